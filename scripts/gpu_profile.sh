#!/bin/bash
# ncu evidence for the decode kernel: launch list of a short bench run + one full capture of the fused kernel.
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dec_ -s 3 -c 1 -o gpurun_out/${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log
ls -la gpurun_out
