#!/bin/bash
# Round 2, fourth GPU call: batched re-prompt + weight-slice L2 prefetch (full gating suite, A/B), bench line, sanitizer passes.
set -u
O=gpurun_out/r2d
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 ) > $O/pytest_gpu.txt 2>&1
for w in 0 1; do
  SVANON_TC_WPREFETCH=$w timeout 120 python tools/bench_launch_overhead.py 100 > $O/single_wprefetch$w.json 2>&1
  SVANON_TC_WPREFETCH=$w timeout 200 python tools/bench_batch.py 128 > $O/batch128_wprefetch$w.json 2>&1
done
( time timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "ar_streaming_codes or stream_loop_vs_reference or encoder_vs_reference or vocoder_vs_reference" 2>&1 | tail -25 ) > $O/sanitizer_memcheck.txt 2>&1
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "ar_streaming_codes" 2>&1 | tail -25 ) > $O/sanitizer_racecheck.txt 2>&1
tail -4 $O/pytest_gpu.txt; cat $O/single_wprefetch0.json $O/single_wprefetch1.json; tail -1 $O/batch128_wprefetch0.json; tail -1 $O/batch128_wprefetch1.json
cat $O/bench_time.txt; tail -6 $O/sanitizer_memcheck.txt; tail -6 $O/sanitizer_racecheck.txt
