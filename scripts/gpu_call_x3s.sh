#!/bin/bash
# straddling build of the split-operand kernel (libturboae_b200_s.so): small run, tests, bench
mkdir -p gpurun_out
export TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_s.so
echo "== small"; timeout 120 python scripts/x3_small.py 5 2>&1 | tail -3
echo "== small 12"; timeout 120 python scripts/x3_small.py 12 2>&1 | tail -3
echo "== x3 tests"; timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r02_x3_s_tests.log
echo "== bench"; timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep "x3\|rror" | tee gpurun_out/r02_x3_s_bench.json
