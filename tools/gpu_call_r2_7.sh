#!/bin/bash
# Round 2, seventh GPU call: thin 128 x N tensor-core tiles (N = 16 / 32) for the two lowest HiFi-GAN levels at many streams:
# GEMM unit tests, batch-equals-single parity, 128-stream step A/B against the direct conv kernel.
set -u
O=gpurun_out/r2h
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -8 ) > $O/pytest_gemm.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -8 ) > $O/pytest_batch.txt 2>&1
timeout 200 python tools/bench_batch.py 128 > $O/batch128_thin_tc.json 2>&1
SVANON_CONV_SMALL_MAX_M=1000000000 timeout 200 python tools/bench_batch.py 128 > $O/batch128_conv_small.json 2>&1
tail -4 $O/pytest_gemm.txt; tail -4 $O/pytest_batch.txt; tail -1 $O/batch128_thin_tc.json; tail -1 $O/batch128_conv_small.json
