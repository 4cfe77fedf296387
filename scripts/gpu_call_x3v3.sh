#!/bin/bash
# pipelined x3 kernel with two issuer warps: bf16 split (v3), fp16 split with scaled operands (v3f), fp16 unscaled (v3g): accuracy, tests, bench
mkdir -p gpurun_out
L=$PWD/turboae_b200/lib
for v in v3 v3f v3g; do
echo "== $v: accuracy"; TURBOAE_B200_LIB=$L/libturboae_b200_$v.so timeout 600 python scripts/x3_accuracy.py > gpurun_out/r02_x3_accuracy_$v.json 2>gpurun_out/acc_$v.err; grep -A3 "x3\"" gpurun_out/r02_x3_accuracy_$v.json | grep "y_max\|x3\|p9999" | tr -d '\n' | sed 's/"c/\n"c/g'; echo
echo "== $v: x3 tests"; TURBOAE_B200_LIB=$L/libturboae_b200_$v.so timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_x3_${v}_tests.log
echo "== $v bench"; TURBOAE_B200_LIB=$L/libturboae_b200_$v.so timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep "x3\|rror" | tee gpurun_out/r02_x3_${v}_bench.json
done
