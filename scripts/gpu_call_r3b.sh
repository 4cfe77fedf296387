#!/bin/bash
# host-side diet of the training step: persistent flat image, cached parameter order, power_constraint autograd kernels, fused Adam
mkdir -p gpurun_out
echo "== tests (training, host, drop-in)"; timeout 1200 python -m pytest tests/test_gpu_train_tc.py tests/test_reference_dropin.py tests/test_gpu_parity.py -m gpu -q -x -k "train or power or dropin or reference or dense or grad or loop" 2>&1 | tail -6 | tee gpurun_out/r3b_tests.log
echo "== train bench (fused Adam)"; timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | tee gpurun_out/r3b_train_fused.json
echo "== train bench (foreach Adam)"; TRAIN_FUSED_ADAM=0 timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | tee gpurun_out/r3b_train_foreach.json
echo "== cpuprof"; timeout 300 python scripts/train_cpuprof.py > gpurun_out/r3b_cpuprof.log 2>&1; head -40 gpurun_out/r3b_cpuprof.log
