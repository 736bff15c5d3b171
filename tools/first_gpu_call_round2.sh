#!/usr/bin/env bash
# First GPU call of the next round (NEXT.md section 0): everything that was written after round 1's GPU minutes ran out,
# in one gpurun call, results under gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/first_gpu_call_round2.sh'
set -u
mkdir -p gpurun_out
T=r3a
# 1. the unrun parity tests (non-gating markers: read the XPASS / XFAIL lines)
timeout 600 python -m pytest tests/test_zz_gpu_gemm_descriptors.py tests/test_zz_gpu_speaker.py tests/test_zz_gpu_generate_kwargs.py -q -rxXs \
  > gpurun_out/${T}_unrun_tests.log 2>&1
echo "unrun tests rc=$?" >> gpurun_out/${T}_unrun_tests.log
# 2. setup-path timing per step (resample, style, timbre, mix, codec ids, content ids, whole calculate_prompt)
timeout 300 python tools/bench_prompt.py 5 15 > gpurun_out/${T}_prompt.jsonl 2> gpurun_out/${T}_prompt.err
# 3. host issue time vs device time of the single-stream loop (graphs or persistent kernels?)
timeout 200 python tools/bench_launch_overhead.py 200 > gpurun_out/${T}_launch_overhead.json 2> gpurun_out/${T}_launch_overhead.err
# 3a. the user-facing path from state dicts: InferenceWrapper.from_state_dicts -> stream_infer (config 5 shape) -> infer
timeout 300 python tools/demo_stream_infer.py 2 0.7 > gpurun_out/${T}_demo.json 2> gpurun_out/${T}_demo.err
# 3b. BASELINE config 3: 64 offline conversions in lock-step vs sequential infer
timeout 400 python tools/bench_offline_batch.py 64 4 > gpurun_out/${T}_offline_batch.json 2> gpurun_out/${T}_offline_batch.err
# 4. the gating suite and the bench line with the new library
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
# 5. launch list of one calculate_prompt (per-kernel times are serialised under ncu: shares only)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_prompt_launches.csv \
  python tools/bench_prompt.py 5 > gpurun_out/${T}_ncu_prompt.log 2>&1
tail -5 gpurun_out/${T}_unrun_tests.log gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_prompt.jsonl gpurun_out/${T}_launch_overhead.json gpurun_out/${T}_demo.json gpurun_out/${T}_offline_batch.json
