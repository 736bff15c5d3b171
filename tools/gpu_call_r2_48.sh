#!/bin/bash
# Round 2, call 48: the conv-form launches of the pair kernel inside the 128-stream step: duration, tensor pipe, L2 and DRAM per launch.
set -u
O=gpurun_out/r2zzc
mkdir -p $O
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 400 ncu --profile-from-start off --kernel-name-base demangled --metrics $M --clock-control none -k regex:gemm_pair_kernel --csv --log-file $O/pair_launches_metrics.csv python tools/profile_batch.py 128 > $O/ncu.log 2>&1
python - <<'P'
import csv,collections,re
rows=collections.OrderedDict()
with open('gpurun_out/r2zzc/pair_launches_metrics.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    d=rows.setdefault(r['ID'],{'name':re.search(r'gemm_pair_kernel<[^>]*>',r['Kernel Name']).group(0)})
    d[r['Metric Name']]=(r['Metric Value'],r['Metric Unit'])
agg=collections.defaultdict(list)
for d in rows.values(): agg[d['name']].append(d)
for k,v in agg.items():
    def avg(m): return sum(float(x[m][0].replace(',','')) for x in v)/len(v)
    print(k, len(v), 'us', round(avg('gpu__time_duration.sum')/ (1000 if v[0]['gpu__time_duration.sum'][1].startswith('n') else 1),1), 'tensor%', round(avg('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),1), 'L2%', round(avg('lts__throughput.avg.pct_of_peak_sustained_elapsed'),1), 'L1%', round(avg('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),1))
P
