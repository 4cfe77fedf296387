#!/bin/bash
# CTA-pair build of the split-operand kernel (libturboae_b200_pair.so): tests, accuracy, bench; under a short timeout each
mkdir -p gpurun_out
L=$PWD/turboae_b200/lib
export TURBOAE_B200_LIB=$L/libturboae_b200_pair.so
echo "== small"; timeout 120 python scripts/x3_small.py 5 2>&1 | tail -3
echo "== pair: x3 tests"; timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_x3_pair_tests.log
echo "== pair bench"; timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep "x3\|rror" | tee gpurun_out/r02_x3_pair_bench.json
echo "== pair accuracy"; timeout 600 python scripts/x3_accuracy.py > gpurun_out/r02_x3_accuracy_pair.json 2>gpurun_out/acc_pair.err; grep -A3 "x3\"" gpurun_out/r02_x3_accuracy_pair.json | grep "y_max\|x3\|p9999" | tr -d '\n' | sed 's/"c/\n"c/g'; echo
