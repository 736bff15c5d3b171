#!/bin/bash
# Round 2, call 29: conv_small with padded input rows (bank conflicts) and optionally twice the rows per CTA.
set -u
O=gpurun_out/r2zi
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -k "vocoder or 128 or nine or stream_loop" > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
timeout 200 python tools/bench_batch.py 128 > $O/batch128_pad.json 2> $O/batch128_pad.err; tail -1 $O/batch128_pad.json
SVANON_CONV_SMALL_BIG=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_pad_big.json 2> $O/batch128_pad_big.err; tail -1 $O/batch128_pad_big.json
timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
