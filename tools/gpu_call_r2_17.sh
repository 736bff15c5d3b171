#!/bin/bash
# Round 2, seventeenth GPU call: launch lists and ncu --set full captures at the end-of-round state (chain kernel default,
# tcgen05.mma issued from uniform registers).
set -u
O=gpurun_out/${OUT:-r2v}
mkdir -p $O
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/single_launches.csv python tools/profile_single.py 2 > $O/ncu_single.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 8 \
    -o $O/gemm_tc_single python tools/profile_single.py 1 > $O/ncu_gemm.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 \
    -o $O/chain_single python tools/profile_single.py 1 > $O/ncu_chain.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ar_decode_staged -c 1 \
    -o $O/ar_decode_single python tools/profile_single.py 1 > $O/ncu_ar.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 12 \
    -o $O/gemm_tc_batch128 python tools/profile_batch.py 128 > $O/ncu_gemm_batch.log 2>&1
ls -la $O; tail -3 $O/ncu_chain.log
