#!/bin/bash
# Round 2, call 20: the whole GPU suite and the default bench line with the pair GEMM kernel in the product.
set -u
O=gpurun_out/${OUT:-r2y}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.err | tail -5
python - <<'P'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2y/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','stage_ms_median','gpu_launches') if k in d})
    print('e2e', d.get('e2e'))
    print('roofline', {k:v for k,v in d.get('roofline',{}).items() if k in ('achieved','frac')})
    r=d.get('roofline_gemm_many_streams',{}); print('gemm', {k:r.get(k) for k in ('launch_us','with_split_pass_us','single_cta_us','frac','pair_kernel_launches_timed')})
    c=d.get('concurrent_streams',{}); print('config4', {k:c.get(k) for k in ('streams_per_gpu','frames_per_s_all_gpus','ms_per_step_mean_max_over_ranks','ms_per_step_p99_max_over_ranks','max_streams_per_gpu_p99_lt_frame_period')})
    print('ladder', c.get('ladder'))
    c5=d.get('config5',{}); print('config5', {k:c5.get(k) for k in ('ms_per_step_mean','ms_per_step_p99','rtf_p99')})
except Exception as e:
    print('parse failed', e)
P
