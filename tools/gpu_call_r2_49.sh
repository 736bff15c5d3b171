#!/bin/bash
# Round 2, call 49: svanon_batch_select (cohort compaction of StreamPool): the new tests and the pool / merge tests, then the whole GPU suite and smoke().
set -u
O=gpurun_out/r2zze
mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_batch.py -q -k "pool" > $O/pytest_pool.txt 2>&1; tail -25 $O/pytest_pool.txt | cut -c1-300
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt | cut -c1-300
