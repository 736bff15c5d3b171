#!/bin/bash
# Round 2, call 25: GPU suite after allowing the ring-based window-start pass for a single stream in encoder mode 2.
set -u
O=gpurun_out/${OUT:-r2zd}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
