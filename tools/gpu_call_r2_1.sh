#!/bin/bash
# Round 2, first GPU call: the gating test suite, the reworked bench line, launch lists and ncu captures of the dominant
# kernels.  Everything lands under gpurun_out/r2a/.
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -40 ) > $O/pytest_gpu.txt 2>&1
( time timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
timeout 120 python tools/bench_launch_overhead.py 100 > $O/launch_overhead.json 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/single_launches.csv python tools/profile_single.py 2 > $O/ncu_single.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 8 \
    -o $O/gemm_tc_single python tools/profile_single.py 1 > $O/ncu_gemm.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ar_decode_staged -c 1 \
    -o $O/ar_decode_single python tools/profile_single.py 1 > $O/ncu_ar.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:arb_attn_slow -c 2 \
    -o $O/attn_slow_128 python tools/profile_batch.py 128 > $O/ncu_attn.log 2>&1
ls -la $O
tail -5 $O/pytest_gpu.txt
