#!/bin/bash
# Round 2, fifteenth GPU call: full gating suite, default bench line, reference arm, memcheck over the chain-kernel tests.
set -u
O=gpurun_out/${OUT:-r2t}
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 ) > $O/pytest_gpu.txt 2>&1
tail -4 $O/pytest_gpu.txt
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) > $O/bench_time.txt 2>&1
cat $O/bench_time.txt; tail -3 $O/bench.err
( time timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) > $O/bench_reference_time.txt 2>&1
tail -c 600 $O/bench_reference.json
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_chain.py -x -q \
    -k "stream_loop or encoder_transformer" 2>&1 | tail -12 ) > $O/sanitizer_memcheck_chain.txt 2>&1
tail -5 $O/sanitizer_memcheck_chain.txt
