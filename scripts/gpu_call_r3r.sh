#!/bin/bash
# where the weight-gradient kernel's MMA issuer spends its time (a -DTAE_WGRAD_PROBE=1 build: clock64 around the waits for the stage loads)
mkdir -p gpurun_out
TURBOAE_B200_WGRAD_OVERLAP=0 TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_wgrad_probe.so timeout 300 python scripts/train_small.py 1000 2>&1 | grep "wgrad job\|train small" | tail -40 | tee gpurun_out/r3r_wgrad_probe.log
