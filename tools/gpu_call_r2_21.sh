#!/bin/bash
# Round 2, call 21: fused forms of the pair kernel (SwiGLU gate, RoPE): unit tests, the 128-stream loop parity test, step time.
set -u
O=gpurun_out/${OUT:-r2z}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "pair" > $O/pytest_pair.txt 2>&1; tail -15 $O/pytest_pair.txt
timeout 400 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -x -q -k "128 or nine or encode" > $O/pytest_batch.txt 2>&1; tail -5 $O/pytest_batch.txt
SVANON_GEMM_PAIR=1 timeout 200 python tools/bench_batch.py 128 160 > $O/batch_pair1.json 2> $O/batch_pair1.err; tail -2 $O/batch_pair1.json
SVANON_GEMM_PAIR=0 timeout 200 python tools/bench_batch.py 128 > $O/batch_pair0.json 2> $O/batch_pair0.err; tail -1 $O/batch_pair0.json
