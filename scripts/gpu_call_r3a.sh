#!/bin/bash
# dense encoder parity test + host-side profile of the eager training step
mkdir -p gpurun_out
echo "== dense tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense" 2>&1 | tail -5 | tee gpurun_out/r3a_dense_tests.log
echo "== cpuprof"; timeout 300 python scripts/train_cpuprof.py > gpurun_out/r3a_cpuprof.log 2>&1; head -75 gpurun_out/r3a_cpuprof.log
