#!/bin/bash
mkdir -p gpurun_out
echo "== old"; timeout 300 python scripts/gru_timeline.py 2>&1 | tail -14 | tee gpurun_out/r02_gru_tl_old.log
echo "== new"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_e.so timeout 300 python scripts/gru_timeline.py 2>&1 | tail -14 | tee gpurun_out/r02_gru_tl_new.log
