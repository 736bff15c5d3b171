#!/bin/bash
# Round 2, eighth GPU call: persistent chain kernel (encoder transformer half): unit tests, loop parity, single-stream timing A/B.
set -u
O=gpurun_out/r2i
mkdir -p $O
( time timeout 500 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -30 ) > $O/pytest_chain.txt 2>&1
tail -12 $O/pytest_chain.txt
timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain.json 2>&1
SVANON_CHAIN=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_no_chain.json 2>&1
tail -2 $O/single_chain.json; tail -2 $O/single_no_chain.json
