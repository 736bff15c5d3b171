#!/bin/bash
# Round 2, sixth GPU call: persistent wide-tile GEMM kernel (two TMEM accumulators): unit tests, loop parity, micro-benchmark, batch step.
set -u
O=gpurun_out/r2f
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -8 ) > $O/pytest_gemm.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_stateful.py -x -q 2>&1 | tail -8 ) > $O/pytest_loops.txt 2>&1
timeout 200 python tools/bench_gemm.py 2 --static > $O/gemm_persist.txt 2>&1
SVANON_TC_PERSIST=0 timeout 200 python tools/bench_gemm.py 2 --static > $O/gemm_no_persist.txt 2>&1
timeout 200 python tools/bench_gemm.py 2 --static --half > $O/gemm_half_persist.txt 2>&1
SVANON_TC_PERSIST=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_no_persist.json 2>&1
timeout 200 python tools/bench_batch.py 128 160 > $O/batch128_persist.json 2>&1
SVANON_PRECISION=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_persist_half.json 2>&1
tail -4 $O/pytest_gemm.txt; tail -4 $O/pytest_loops.txt; tail -5 $O/gemm_persist.txt; tail -5 $O/gemm_no_persist.txt; tail -3 $O/gemm_half_persist.txt
tail -1 $O/batch128_no_persist.json; tail -2 $O/batch128_persist.json; tail -1 $O/batch128_persist_half.json
