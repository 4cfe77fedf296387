#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (own arm), GRU and training side benches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/tests.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== rnn" ; RNN_B=18944 timeout 600 python scripts/rnn_bench.py 2>gpurun_out/rnn.err | tee gpurun_out/rnn.json
tail -3 gpurun_out/rnn.err
echo "== train" ; timeout 600 python scripts/train_bench.py 2>gpurun_out/train.err | tee gpurun_out/train.json
tail -3 gpurun_out/train.err
echo "== next-round checks" ; timeout 600 python scripts/gpu_next_checks.py 2>&1 | tail -5 | tee gpurun_out/next_checks.log
