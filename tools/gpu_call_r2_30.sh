#!/bin/bash
# Round 2, call 30: shared-memory bank conflicts and wavefront excess per kernel, single-stream chunk and 128-stream step.
set -u
O=gpurun_out/r2zj
mkdir -p $O
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/single_conflicts.csv python tools/profile_single.py 1 > $O/ncu_single.log 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/batch128_conflicts.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
ls -la $O
