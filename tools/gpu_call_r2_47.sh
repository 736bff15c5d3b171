#!/bin/bash
# Round 2, call 47: conv form with 16 channels (64-byte swizzle rows): unit tests, parity, step time.
set -u
O=gpurun_out/r2zza
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k "taps_bitwise" > $O/pytest_taps.txt 2>&1; tail -12 $O/pytest_taps.txt | cut -c1-200
if ! grep -q "failed" $O/pytest_taps.txt; then
  timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_stateful.py -x -q -k "128 or nine or vocoder or voc or push or reference_fixture" > $O/pytest_batch.txt 2>&1; tail -4 $O/pytest_batch.txt
  SVANON_CONV16_PAIR=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_c16pair1.json 2> $O/batch128_c16pair1.err; tail -1 $O/batch128_c16pair1.json; tail -2 $O/batch128_c16pair1.err
  SVANON_CONV16_PAIR=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_c16pair0.json 2> $O/batch128_c16pair0.err; tail -1 $O/batch128_c16pair0.json
fi
