#!/bin/bash
# round 2: full GPU suite, 2 bench runs, ncu launch list + full capture of the fused decoder (must complete without a launch failure)
TAG=${1:-r02_e}
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
export BENCH_SKIP_SECONDARY=1
for i in 1 2; do
echo "== bench $i"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench$i.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline']['launch_ms_min'], d['e2e']['value'])
"
done
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1; tail -1 gpurun_out/${TAG}_launches.log | cut -c1-160
echo "== dec full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_pair -s 4 -c 1 -o gpurun_out/${TAG}_dec -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_dec_full.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/${TAG}_dec_full.log | head -5; tail -1 gpurun_out/${TAG}_dec_full.log | cut -c1-160
