#!/bin/bash
# Round 2, tenth GPU call: per-slab pipeline trace of one GEMM phase of the chain kernel.
set -u
O=gpurun_out/${OUT:-r2m}
mkdir -p $O
for op in 1 5 6; do
SVANON_CHAIN=1 SVANON_CHAIN_PROF=60 SVANON_CHAIN_TRACE=$op timeout 120 python tools/bench_launch_overhead.py 70 2>&1 >/dev/null | head -20 > $O/trace_op$op.txt
done
cat $O/trace_op1.txt | head -20
