#!/bin/bash
# Round 2, call 38: smoke() of __graft_entry__ (streaming session vs the oracle + the pair GEMM kernel vs fp64).
set -u
O=gpurun_out/r2zr
mkdir -p $O
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.txt 2>&1; tail -6 $O/smoke.txt
