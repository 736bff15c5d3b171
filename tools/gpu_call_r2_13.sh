#!/bin/bash
# Round 2, thirteenth GPU call: conv stack inside the chain kernel: chain tests (loop parity against the per-op path), the
# reference-fixture parity file, single-stream timing with / without the conv half, phase profile.
set -u
O=gpurun_out/${OUT:-r2q}
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -x -q 2>&1 | tail -30 ) > $O/pytest.txt 2>&1
tail -5 $O/pytest.txt
SVANON_CHAIN=3 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain_conv.json 2>&1
SVANON_CHAIN=1 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_chain_noconv.json 2>&1
SVANON_CHAIN=0 timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_no_chain.json 2>&1
SVANON_CHAIN=3 SVANON_CHAIN_PROF=60 timeout 120 python tools/bench_launch_overhead.py 70 > /dev/null 2> $O/chain_prof.txt
tail -qn1 $O/single_chain_conv.json $O/single_chain_noconv.json $O/single_no_chain.json
