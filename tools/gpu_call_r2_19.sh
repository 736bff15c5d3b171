#!/bin/bash
# Round 2, call 19: where the pair GEMM kernel's time goes: per-launch durations (split pass vs GEMM), ncu --set full, and the
# 128-stream launch list with the kernel on.
set -u
O=gpurun_out/${OUT:-r2x}
mkdir -p $O
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/pair_few_launches.csv \
    python tools/bench_gemm_pair.py few > $O/ncu_few.log 2>&1
grep -E "gemm_pair|split_lo|gemm_tc" $O/pair_few_launches.csv | awk -F'","' '{print $5, $(NF)}' | head -40
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -c 6 -o $O/gemm_pair \
    python tools/bench_gemm_pair.py few > $O/ncu_pair_full.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
ls -la $O
