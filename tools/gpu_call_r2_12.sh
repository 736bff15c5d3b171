#!/bin/bash
# Round 2, twelfth GPU call: A/B of the weight-tile bulk-copy issue (elected lane vs thread 0) on the wide GEMM tiles.
set -u
O=gpurun_out/${OUT:-r2p}
mkdir -p $O
for i in 1 2; do
SVANON_TC_B_ELECT=1 timeout 100 python tools/bench_gemm.py > $O/gemm_elect_$i.txt 2>&1
SVANON_TC_B_ELECT=0 timeout 100 python tools/bench_gemm.py > $O/gemm_tid0_$i.txt 2>&1
done
SVANON_TC_B_ELECT=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_elect.json 2>&1
SVANON_TC_B_ELECT=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_tid0.json 2>&1
SVANON_TC_B_ELECT=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_elect2.json 2>&1
grep "M=16384\|M=4096" $O/gemm_*.txt; tail -qn1 $O/batch128_*.json
