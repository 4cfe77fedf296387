#!/bin/bash
# Run on the GPU box (under gpurun): parity tests, smoke, a short bench.
# Every stage has its own timeout so that a hung kernel cannot eat the whole call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe" ; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "umma" 2>&1 | tail -25 | tee gpurun_out/probe.log
echo "== decoder" ; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decoder_bf16" 2>&1 | tail -40 | tee gpurun_out/dec.log
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/tests.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
