#!/bin/bash
# Round 2, fifth GPU call: cohort merging, config-5 leg, deeper small-tile prefetch: gating suite + bench line.
set -u
O=gpurun_out/r2e
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 ) > $O/pytest_gpu.txt 2>&1
timeout 120 python tools/bench_launch_overhead.py 100 > $O/single.json 2>&1
( time timeout 1200 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
tail -5 $O/pytest_gpu.txt; cat $O/single.json $O/bench_time.txt; tail -3 $O/bench.err
