#!/bin/bash
# Round 2, call 28: window attention for many streams with two threads per query: parity, step time, ncu.
set -u
O=gpurun_out/${OUT:-r2zg}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q > $O/pytest_batch.txt 2>&1; tail -4 $O/pytest_batch.txt
SVANON_ATTN_ROWQ=1 timeout 200 python tools/bench_batch.py 128 > $O/batch128_rowq1.json 2> $O/batch128_rowq1.err; tail -1 $O/batch128_rowq1.json; tail -3 $O/batch128_rowq1.err
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_rowq -c 1 \
    -o $O/attn_rowq python tools/profile_batch.py 128 > $O/ncu_attn.log 2>&1
ncu -i $O/attn_rowq.ncu-rep --page raw --csv > $O/attn_rowq_raw.csv 2>/dev/null
