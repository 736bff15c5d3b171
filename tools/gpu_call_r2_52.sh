#!/bin/bash
# Round 2, call 52: ncu launch lists (gpu__time_duration.sum) of the 128-stream step and of the single-stream chunk at the final state.
set -u
O=gpurun_out/r2zzh
mkdir -p $O
timeout 110 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1; echo "batch rc=$?"; tail -1 $O/ncu_batch.log | cut -c1-300
timeout 80 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/single_launches.csv python tools/profile_single.py 2 > $O/ncu_single.log 2>&1; echo "single rc=$?"; tail -1 $O/ncu_single.log | cut -c1-300
wc -l $O/*.csv
