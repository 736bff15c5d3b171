#!/bin/bash
# Round 2, call 45: conv operand passes of a launch bundled into one kernel: tests and step time.
set -u
O=gpurun_out/r2zy
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "pair" > $O/pytest_pair.txt 2>&1; tail -3 $O/pytest_pair.txt
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_stateful.py -x -q -k "128 or nine or vocoder or voc or push or reference_fixture" > $O/pytest_batch.txt 2>&1; tail -3 $O/pytest_batch.txt
timeout 200 python tools/bench_batch.py 128 > $O/batch128.json 2> $O/batch128.err; tail -1 $O/batch128.json
