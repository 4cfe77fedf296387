#!/bin/bash
# round-2 closing run: full GPU suite, smoke, bench (both arms), ncu launch list of the bench, full captures of x3_kernel and wgrad_kernel
TAG=${1:-r02_final}
mkdir -p gpurun_out
echo "== tests"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('secondary', {}), indent=1))
"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-400
echo "== launch list"; BENCH_SKIP_SECONDARY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1; tail -1 gpurun_out/${TAG}_launches.log | cut -c1-160
echo "== x3 full (decoder launch)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:x3_kernel -s 14 -c 1 -o gpurun_out/${TAG}_x3 -f \
    python scripts/x3_bench.py 50000 > gpurun_out/${TAG}_x3_full.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/${TAG}_x3_full.log | head -5
echo "== wgrad full"; TRAIN_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 -o gpurun_out/${TAG}_wgrad -f \
    python scripts/train_bench.py > gpurun_out/${TAG}_wgrad.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/${TAG}_wgrad.log | head -5
ls -la gpurun_out | grep ${TAG}
