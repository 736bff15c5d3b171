#!/bin/bash
# Round 2, eleventh GPU call: tcgen05.mma issued by an elected lane of a warp-uniform control warp (descriptors in uniform
# registers) in gemm_tc.cu and chain.cu: full GPU suite, single-stream timing (chain on / off), 128-stream step, GEMM bench.
set -u
O=gpurun_out/${OUT:-r2n}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/pytest_gpu.txt 2>&1
tail -4 $O/pytest_gpu.txt
( SVANON_CHAIN=1 timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -4 ) > $O/pytest_chain.txt 2>&1
tail -2 $O/pytest_chain.txt
SVANON_CHAIN=1 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain.json 2>&1
SVANON_CHAIN=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_no_chain.json 2>&1
SVANON_CHAIN=1 SVANON_CHAIN_PROF=60 SVANON_CHAIN_TRACE=1 timeout 120 python tools/bench_launch_overhead.py 70 > /dev/null 2> $O/chain_prof.txt
tail -1 $O/single_chain.json; tail -1 $O/single_no_chain.json
timeout 200 python tools/bench_batch.py 128 > $O/batch128.json 2>&1
tail -1 $O/batch128.json
timeout 100 python tools/bench_gemm.py > $O/gemm.txt 2>&1
tail -12 $O/gemm.txt
