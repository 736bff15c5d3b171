#!/bin/bash
# Round 2, call 34: persistent conv_small (weights staged once per CTA, next tile's rows prefetched through registers).
set -u
O=gpurun_out/r2zn
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_stateful.py -x -q -k "vocoder or 128 or nine or stream_loop or voc or push" > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
timeout 200 python tools/bench_batch.py 128 > $O/batch128.json 2> $O/batch128.err; tail -1 $O/batch128.json; tail -2 $O/batch128.err
SVANON_CONV_SMALL_BIG=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_small_tiles.json 2> $O/batch128_small_tiles.err; tail -1 $O/batch128_small_tiles.json
timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
