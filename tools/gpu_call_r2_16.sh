#!/bin/bash
# Round 2, sixteenth GPU call: chain GEMM phases with the epilogue staged through shared memory (whole row segments per store).
set -u
O=gpurun_out/${OUT:-r2u}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for shp in "128 512 1536" "128 1536 512" "340 1536 384"; do
  n=$(echo $shp | tr ' ' 'x')
  SVANON_CHAIN_PROF=3 timeout 100 python tools/bench_chain_phases.py $shp 3 2> $O/warm_${n}.txt >/dev/null
  echo "== $shp"; sed -n 3,5p $O/warm_${n}.txt | cut -c1-250
done
SVANON_CHAIN=1 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
SVANON_CHAIN=3 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
SVANON_CHAIN=0 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
