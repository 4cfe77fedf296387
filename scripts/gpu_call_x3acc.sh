#!/bin/bash
# accuracy of the split-operand path with bf16 and with fp16 operand terms (two builds of the library) + x3 tests on both
mkdir -p gpurun_out
L=$PWD/turboae_b200/lib
echo "== bf16 split"; timeout 900 python scripts/x3_accuracy.py > gpurun_out/r02_x3_accuracy_bf16.json 2>gpurun_out/acc_bf16.err; grep -A5 "x3\"" gpurun_out/r02_x3_accuracy_bf16.json | grep "y_max\|x3\|gt"
echo "== fp16 split"; TURBOAE_B200_LIB=$L/libturboae_b200_f16.so timeout 900 python scripts/x3_accuracy.py > gpurun_out/r02_x3_accuracy_fp16.json 2>gpurun_out/acc_fp16.err;  grep -A5 "x3\"" gpurun_out/r02_x3_accuracy_fp16.json | grep "y_max\|x3\|gt"
echo "== fp16: x3 tests"; TURBOAE_B200_LIB=$L/libturboae_b200_f16.so timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_x3_f16_tests.log
echo "== fp16 bench"; TURBOAE_B200_LIB=$L/libturboae_b200_f16.so timeout 300 python scripts/x3_bench.py 50000 2>&1 | grep x3 | tee gpurun_out/r02_x3_f16_bench.json
