#!/bin/bash
# Round 2, call 24: whole GPU suite, default bench line and the 128-stream launch list at the commit with the pair kernel's fused
# forms and the ring-based window-start pass.
set -u
O=gpurun_out/${OUT:-r2zc}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
python - <<'P'
import json,sys,os
try:
    d=json.loads(open(os.environ.get('O','gpurun_out/r2zc')+'/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','stage_ms_median','gpu_launches') if k in d})
    print('e2e', d.get('e2e'))
    r=d.get('roofline_gemm_many_streams',{}); print('gemm', {k:r.get(k) for k in ('launch_us','with_split_pass_us','single_cta_us','frac')})
    c=d.get('concurrent_streams',{}); print('config4', {k:c.get(k) for k in ('streams_per_gpu','frames_per_s_all_gpus','ms_per_step_mean_max_over_ranks','ms_per_step_p99_max_over_ranks','max_streams_per_gpu_p99_lt_frame_period')})
    for l in c.get('ladder',[]): print('  ladder', {k:l.get(k) for k in ('streams','ms_per_step_mean','ms_per_step_p99','stage_ms')})
    c5=d.get('config5',{}); print('config5', {k:c5.get(k) for k in ('ms_per_step_mean','ms_per_step_p99','rtf_p99')})
except Exception as e:
    print('parse failed', e)
P
