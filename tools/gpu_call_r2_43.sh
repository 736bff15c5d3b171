#!/bin/bash
# Round 2, call 40: evidence at the final state (conv form of the pair kernel in the product): GPU suite, bench line, reference arm,
# 128-stream launch list, memcheck of the pair kernel unit tests.
set -u
O=gpurun_out/r2zw
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
( time timeout 1200 python bench.py ) > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
( time timeout 600 python bench.py --impl reference ) > $O/bench_reference.json 2> $O/bench_reference.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/batch128_launches.csv python tools/profile_batch.py 128 > $O/ncu_batch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py -x -q -k "taps_bitwise or fused or (pair_gemm_bitwise and 12500)" > $O/sanitizer_memcheck_pair.txt 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck_pair.txt
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2zw/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','stage_ms_median','gpu_launches') if k in d})
    print('e2e', d.get('e2e'))
    print('roofline', {k:v for k,v in d.get('roofline',{}).items() if k in ('achieved','frac','share_of_step','traffic_source')})
    r=d.get('roofline_gemm_many_streams',{}); print('gemm', {k:r.get(k) for k in ('launch_us','with_split_pass_us','single_cta_us','frac')})
    c=d.get('concurrent_streams',{}); print('config4', {k:c.get(k) for k in ('streams_per_gpu','frames_per_s_all_gpus','ms_per_step_mean_max_over_ranks','ms_per_step_p99_max_over_ranks','max_streams_per_gpu_p99_lt_frame_period')})
    for l in c.get('ladder',[]): print('  ladder', {k:l.get(k) for k in ('streams','ms_per_step_mean','ms_per_step_p99','stage_ms')})
    c5=d.get('config5',{}); print('config5', {k:c5.get(k) for k in ('ms_per_step_mean','ms_per_step_p99','rtf_p99')})
    s=d['concurrent_streams_stateful_encoder']
    for l in s['ladder']: print('  stateful', {x:l.get(x) for x in ('streams','ms_per_step_mean','ms_per_step_p99','stage_ms')})
    p=d['perf_mode']
    for k in ('window_encoder','stateful_encoder'):
        for l in p.get(k,[]): print('  perf',k,{x:l.get(x) for x in ('streams','ms_per_step_mean','ms_per_step_p99')})
    for s in d.get('stage_compute',[]): print('stage_compute', s['streams'], {k:(round(s[k]['achieved_tflops'],1), round(s[k]['frac'],3)) for k in ('E','V')})
    r=json.loads(open('gpurun_out/r2zw/bench_reference.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['ms_per_step'])
except Exception as e:
    print('parse failed', e)
P
