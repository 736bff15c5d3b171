#!/bin/bash
# Round 2, call 41: the default bench line with the longer opt-in ladders (last line of the round).
set -u
O=gpurun_out/r2zu
mkdir -p $O
( time timeout 1500 python bench.py ) > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2zu/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','stage_ms_median') if k in d}); print('e2e', d.get('e2e'))
    c=d.get('concurrent_streams',{}); print('config4', {k:c.get(k) for k in ('ms_per_step_mean_max_over_ranks','ms_per_step_p99_max_over_ranks','max_streams_per_gpu_p99_lt_frame_period')})
    for l in c.get('ladder',[]): print('  ladder', {k:l.get(k) for k in ('streams','ms_per_step_mean','ms_per_step_p99')})
    s=d['concurrent_streams_stateful_encoder']; print('stateful max', s.get('max_streams_per_gpu_p99_lt_frame_period'))
    for l in s['ladder']: print('  stateful', {x:l.get(x) for x in ('streams','ms_per_step_mean','ms_per_step_p99','error')})
    p=d['perf_mode']
    for k in ('window_encoder','stateful_encoder'):
        print('perf max', k, p.get('max_streams_per_gpu_p99_lt_frame_period_'+k))
        for l in p.get(k,[]): print('  perf',k,{x:l.get(x) for x in ('streams','ms_per_step_mean','ms_per_step_p99','error')})
except Exception as e:
    print('parse failed', e)
P
