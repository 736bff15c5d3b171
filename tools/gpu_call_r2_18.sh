#!/bin/bash
# Round 2, call 18: first run of the pair GEMM kernel (gemm_pair.cu): bitwise check against the single-CTA kernel, timings,
# the 128-stream loop parity test and the 128-stream step with the kernel on / off.
set -u
O=gpurun_out/${OUT:-r2w}
mkdir -p $O
timeout 240 python tools/bench_gemm_pair.py check > $O/pair_check.txt 2>&1; echo "check rc=$?" >> $O/pair_check.txt
tail -60 $O/pair_check.txt
timeout 240 python tools/bench_gemm_pair.py time > $O/pair_time.txt 2>&1; echo "time rc=$?" >> $O/pair_time.txt
cat $O/pair_time.txt
if grep -q "CHECK OK" $O/pair_check.txt; then
  timeout 400 python -m pytest tests/test_gpu_batch.py -x -q -k "128 or nine or encode_batch" > $O/pytest_batch.txt 2>&1; tail -5 $O/pytest_batch.txt
  SVANON_GEMM_PAIR=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_pair0.json 2> $O/batch128_pair0.err; tail -2 $O/batch128_pair0.json
  SVANON_GEMM_PAIR=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_pair1.json 2> $O/batch128_pair1.err; tail -2 $O/batch128_pair1.json
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
