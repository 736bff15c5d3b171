#!/bin/bash
# Round 2, call 36: the long-run cohort-merge test (rings re-filled after a merge).
set -u
O=gpurun_out/r2zp
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -k "merge_then_long_run" ) > $O/pytest_merge.txt 2>&1; tail -15 $O/pytest_merge.txt
