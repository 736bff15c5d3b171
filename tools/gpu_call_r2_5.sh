#!/bin/bash
# Round 2, fifth GPU call: cohort merging, config-5 leg, deeper small-tile prefetch: gating suite + bench line.
set -u
O=gpurun_out/r2e
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 ) > $O/pytest_gpu.txt 2>&1
timeout 120 python tools/bench_launch_overhead.py 100 > $O/single.json 2>&1
SVANON_TC_TMA_WEIGHTS=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_no_tma_weights.json 2>&1
timeout 200 python tools/bench_gemm.py 2 > $O/gemm_regpath.txt 2>&1
timeout 200 python tools/bench_gemm.py 2 --static > $O/gemm_tma_weights.txt 2>&1
timeout 200 python tools/bench_gemm.py 2 --half > $O/gemm_half_regpath.txt 2>&1
timeout 200 python tools/bench_gemm.py 2 --static --half > $O/gemm_half_tma_weights.txt 2>&1
SVANON_TC_TMA_WEIGHTS=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_no_tma_weights.json 2>&1
timeout 200 python tools/bench_batch.py 128 > $O/batch128_tma_weights.json 2>&1
( time timeout 1200 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
tail -5 $O/pytest_gpu.txt; cat $O/single.json $O/single_no_tma_weights.json $O/bench_time.txt; tail -3 $O/bench.err
paste -d'\n' $O/gemm_regpath.txt $O/gemm_tma_weights.txt | tail -12; tail -1 $O/batch128_no_tma_weights.json; tail -1 $O/batch128_tma_weights.json
