#!/bin/bash
mkdir -p gpurun_out
export TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_e.so
echo "== e: x3 tests"; timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_x3_e_tests.log
echo "== e bench"; timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep "x3\|rror" | tee gpurun_out/r02_x3_e_bench.json
echo "== e bench again"; timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep "enc_f16x3\|rror"
