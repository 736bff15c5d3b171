#!/bin/bash
# Round 2, call 22: RoPE epilogue with explicit rounding points (bitwise test), split-K 8 A/B on the many-stream decode,
# memcheck of the pair kernel tests.
set -u
O=gpurun_out/${OUT:-r2za}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "pair" > $O/pytest_pair.txt 2>&1; tail -3 $O/pytest_pair.txt
timeout 400 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -x -q -k "128 or nine or encode" > $O/pytest_batch.txt 2>&1; tail -2 $O/pytest_batch.txt
timeout 200 python tools/bench_batch.py 128 > $O/batch128_split4.json 2> $O/batch128_split4.err; tail -1 $O/batch128_split4.json
SVANON_TC_MAX_SPLIT=8 timeout 200 python tools/bench_batch.py 128 > $O/batch128_split8.json 2> $O/batch128_split8.err; tail -1 $O/batch128_split8.json
SVANON_TC_MAX_SPLIT=8 timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
timeout 120 python tools/bench_launch_overhead.py 100 | tail -1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py -x -q -k "pair_gemm_bitwise and 12500 or fused" > $O/sanitizer_memcheck_pair.txt 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck_pair.txt
