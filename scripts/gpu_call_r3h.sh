#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_overlap_ab.py 2>&1 | tail -3 | tee gpurun_out/r3h_overlap_ab.json
TRAIN_B=745 timeout 600 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee -a gpurun_out/r3h_overlap_ab.json
