#!/bin/bash
# Round 2, fourteenth GPU call: chain phases on warm operands (micro-benchmark), split vs single TMEM accumulators.
set -u
O=gpurun_out/${OUT:-r2r}
mkdir -p $O
for acc in 0 1; do
for shp in "128 512 1536" "128 1536 512" "340 1536 384"; do
  n=$(echo $shp | tr ' ' 'x')
  SVANON_CHAIN_ACC_SPLIT=$acc SVANON_CHAIN_PROF=3 timeout 100 python tools/bench_chain_phases.py $shp 3 2> $O/warm_${n}_acc$acc.txt >/dev/null
  echo "== acc_split=$acc $shp"; sed -n 3,4p $O/warm_${n}_acc$acc.txt | cut -c1-250
done
done
SVANON_CHAIN_ACC_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
SVANON_CHAIN_ACC_SPLIT=1 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
SVANON_CHAIN_ACC_SPLIT=0 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
