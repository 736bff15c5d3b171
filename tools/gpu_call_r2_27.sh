#!/bin/bash
# Round 2, call 27: ncu of the thread-per-query attention kernel inside the 128-stream step.
set -u
O=gpurun_out/${OUT:-r2zf}
mkdir -p $O
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_rowq -c 2 \
    -o $O/attn_rowq python tools/profile_batch.py 128 > $O/ncu_attn.log 2>&1
ncu -i $O/attn_rowq.ncu-rep --page details --csv > $O/attn_rowq_details.csv 2>/dev/null
ncu -i $O/attn_rowq.ncu-rep --page raw --csv > $O/attn_rowq_raw.csv 2>/dev/null
ls -la $O; tail -3 $O/ncu_attn.log
