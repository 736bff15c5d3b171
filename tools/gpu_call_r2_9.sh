#!/bin/bash
# Round 2, ninth GPU call: chain kernel after the phase-overhead pass (row-per-CTA norm, 8-query attention jobs, weight-job
# descriptors, double-buffered op descriptors): unit tests, loop parity, single-stream timing A/B, phase profile.
set -u
O=gpurun_out/${OUT:-r2j}
mkdir -p $O
( time SVANON_CHAIN=1 timeout 500 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -30 ) > $O/pytest_chain.txt 2>&1
tail -6 $O/pytest_chain.txt
SVANON_CHAIN=1 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain.json 2>&1
SVANON_CHAIN=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_no_chain.json 2>&1
SVANON_CHAIN=1 SVANON_CHAIN_PROF=60 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain_prof.json 2> $O/chain_prof.txt
tail -2 $O/single_chain.json; tail -2 $O/single_no_chain.json
SVANON_CHAIN=1 SVANON_CHAIN_FUSE=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain_nofuse.json 2>&1
tail -1 $O/single_chain_nofuse.json
