#!/bin/bash
# ncu evidence for the tensor-core training kernels: launch list of two training steps + full captures of the
# weight-gradient kernel and the stack-backward kernel (dec_pair_kernel<1>).
TAG=${1:-r01_train}
mkdir -p gpurun_out
export TRAIN_STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/train_bench.py > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 -o gpurun_out/${TAG}_wgrad -f \
    python scripts/train_bench.py > gpurun_out/${TAG}_wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"dec_pair_kernel<\(int\)1>" -s 20 -c 1 -o gpurun_out/${TAG}_bwd -f \
    python scripts/train_bench.py > gpurun_out/${TAG}_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"dec_pair_kernel<\(int\)2>" -s 5 -c 1 -o gpurun_out/${TAG}_fwd -f \
    python scripts/train_bench.py > gpurun_out/${TAG}_fwd.log 2>&1
tail -2 gpurun_out/${TAG}_bwd.log gpurun_out/${TAG}_wgrad.log gpurun_out/${TAG}_fwd.log
ls -la gpurun_out | tail -8
