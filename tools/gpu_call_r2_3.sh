#!/bin/bash
# Round 2, third GPU call: perf mode (fp16 single-pass GEMMs) -- kernel test, fidelity evaluation -- and the full bench line.
set -u
O=gpurun_out/r2c
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s -k "perf_mode or tcgen05" 2>&1 | tail -30 ) > $O/pytest_gemm.txt 2>&1
timeout 300 python tools/eval_perf_mode.py 32 40 > $O/eval_perf_mode.json 2> $O/eval_perf_mode.err
( time timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
tail -5 $O/pytest_gemm.txt; cat $O/eval_perf_mode.json; tail -3 $O/eval_perf_mode.err; cat $O/bench_time.txt; tail -3 $O/bench.err
