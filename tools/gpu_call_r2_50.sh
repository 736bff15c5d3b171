#!/bin/bash
# Round 2, call 50: memcheck over the compaction tests (svanon_batch_select), the argument-error test, and a short bench line at the final state.
set -u
O=gpurun_out/r2zzf
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_batch.py -q -k "select_argument" > $O/pytest_select_args.txt 2>&1; tail -5 $O/pytest_select_args.txt | cut -c1-300
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_batch.py -x -q -k "drops_streams or select_argument" > $O/sanitizer_memcheck_select.txt 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck_select.txt | cut -c1-300
timeout 200 python bench.py --steps 100 --warmup 5 --concurrent 128 --concurrent-chunks 720 --stateful '' --config5 0 --perf '' --perf-stateful '' --no-prompt-path --cpu-sample 4 > $O/bench_short.json 2> $O/bench_short.err; tail -c 600 $O/bench_short.json; tail -2 $O/bench_short.err
