#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for ov in 1 0; do
  echo "== eager train bench, overlap=$ov (rep $rep)"; TRAIN_STEPS=60 TURBOAE_B200_WGRAD_OVERLAP=$ov timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | cut -c100-420
done
done
