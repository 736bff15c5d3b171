#!/bin/bash
# Round 2, call 33: window-start pass and newest-frames pass merged into the same launches: parity and step time.
set -u
O=gpurun_out/r2zm
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -x -q -k "rings or 128 or nine or reference_fixture or merges or ring_buffer" > $O/pytest_batch.txt 2>&1; tail -15 $O/pytest_batch.txt
SVANON_ENC_MERGED=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_merged1.json 2> $O/batch128_merged1.err; tail -1 $O/batch128_merged1.json; tail -3 $O/batch128_merged1.err
SVANON_ENC_MERGED=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_merged0.json 2> $O/batch128_merged0.err; tail -1 $O/batch128_merged0.json
