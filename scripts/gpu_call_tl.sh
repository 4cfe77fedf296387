#!/bin/bash
# timeline only (TAG)
TAG=${1:-tl}
mkdir -p gpurun_out
TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_timeline.so TL_TAG=${TAG}_timeline timeout 300 python scripts/dec_timeline.py > gpurun_out/${TAG}_tl.log 2>&1; tail -40 gpurun_out/${TAG}_tl.log
