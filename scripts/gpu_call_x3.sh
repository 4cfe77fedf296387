#!/bin/bash
TAG=${1:-r02_x3}
mkdir -p gpurun_out
echo "== x3 tests"; timeout 900 python -m pytest tests/test_gpu_x3.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.log
echo "== x3 bench"; timeout 600 python scripts/x3_bench.py 50000 2>&1 | tail -30 | tee gpurun_out/${TAG}_bench.json
