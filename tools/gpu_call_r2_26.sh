#!/bin/bash
# Round 2, call 26: thread-per-query window attention for many streams (attention_rowq_kernel): parity and step time.
set -u
O=gpurun_out/${OUT:-r2ze}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_stateful.py -x -q > $O/pytest_batch.txt 2>&1; tail -8 $O/pytest_batch.txt
SVANON_ATTN_ROWQ=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_rowq1.json 2> $O/batch128_rowq1.err; tail -1 $O/batch128_rowq1.json; tail -3 $O/batch128_rowq1.err
SVANON_ATTN_ROWQ=0 timeout 200 python tools/bench_batch.py 128 > $O/batch128_rowq0.json 2> $O/batch128_rowq0.err; tail -1 $O/batch128_rowq0.json
