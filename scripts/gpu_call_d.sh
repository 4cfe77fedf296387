#!/bin/bash
# round 2: full GPU suite + smoke + full bench (own arm and reference arm)
TAG=${1:-r02_d}
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>gpurun_out/${TAG}_ref.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-400
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('secondary'), indent=1)); print(d.get('cpu_baseline')); print(d.get('eager_gpu_baseline'))
"
tail -5 gpurun_out/${TAG}_bench.err
