#!/bin/bash
# Round 2, call 51: the GPU suite and smoke() at the last commit of the round.
set -u
O=gpurun_out/r2zzg
mkdir -p $O
( time timeout 400 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt | cut -c1-300
