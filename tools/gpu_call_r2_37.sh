#!/bin/bash
# Round 2, call 37: window attention with two CTAs per (stream, head).
set -u
O=gpurun_out/r2zq
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -k "nine or many_rows or 128 or reference_fixture" > $O/pytest_batch.txt 2>&1; tail -3 $O/pytest_batch.txt
timeout 200 python tools/bench_batch.py 128 > $O/batch128.json 2> $O/batch128.err; tail -1 $O/batch128.json
timeout 200 python tools/bench_batch.py 128 > $O/batch128_b.json 2> $O/batch128_b.err; tail -1 $O/batch128_b.json
