#!/bin/bash
# Round 2, call 23: window-start pass from the rings of steady-state layer inputs (enc_conv_stack_head): parity and step time.
set -u
O=gpurun_out/${OUT:-r2zb}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -k "rings or 128 or nine or reference_fixture or merges" > $O/pytest_batch.txt 2>&1; tail -15 $O/pytest_batch.txt
SVANON_ENC_HEAD_TRI=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_tri1.json 2> $O/batch128_tri1.err; tail -1 $O/batch128_tri1.json; tail -3 $O/batch128_tri1.err
SVANON_ENC_HEAD_TRI=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_tri0.json 2> $O/batch128_tri0.err; tail -1 $O/batch128_tri0.json
