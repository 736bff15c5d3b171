#!/bin/bash
# Round 2, call 39: conv form of the pair kernel (taps through 3-D tensor maps, SiLU'd hi / lo operand arrays): unit test, parity, time.
set -u
O=gpurun_out/r2zs
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "taps_bitwise" > $O/pytest_taps.txt 2>&1; tail -12 $O/pytest_taps.txt
if grep -q "passed" $O/pytest_taps.txt && ! grep -q "failed" $O/pytest_taps.txt; then
  timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_stateful.py -x -q -k "128 or nine or vocoder or voc or push or reference_fixture" > $O/pytest_batch.txt 2>&1; tail -6 $O/pytest_batch.txt
  SVANON_GEMM_PAIR_TAPS=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_taps1.json 2> $O/batch128_taps1.err; tail -1 $O/batch128_taps1.json; tail -2 $O/batch128_taps1.err
  SVANON_GEMM_PAIR_TAPS=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_taps0.json 2> $O/batch128_taps0.err; tail -1 $O/batch128_taps0.json
fi
