#!/bin/bash
# Round 2, call 32 (2 GPUs): the driver's multi-rank launch of bench.py at the end-of-round state.
set -u
O=gpurun_out/r2zx
mkdir -p $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 600 $O/bench_2gpu.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2zx/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling') if k in d})
    c=d.get('concurrent_streams',{}); print('config4', {k:c.get(k) for k in ('streams_per_gpu','total_streams','frames_per_s_all_gpus','ms_per_step_mean_max_over_ranks','ms_per_step_p99_max_over_ranks','host_issue_ms_per_step_max_over_ranks')})
    c5=d.get('config5',{}); print('config5', {k:c5.get(k) for k in ('ms_per_step_mean','ms_per_step_p99','rtf_p99','frames_per_s_all_gpus')})
except Exception as e:
    print('parse failed', e)
P
