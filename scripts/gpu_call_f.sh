#!/bin/bash
TAG=${1:-r02_f}
mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
