#!/bin/bash
# closing run #2 (after the split backward): full GPU suite, smoke, bench; compute-sanitizer memcheck over small training steps
TAG=${1:-r02_close3}
mkdir -p gpurun_out
echo "== tests"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline'].get('frac_fastest_launch'), d['e2e']['value'])
print({k: v for k, v in d.get('secondary', {}).items() if 'train' in k})
"
for b in 23 745; do
  echo "== memcheck, training step B=$b"; timeout 900 compute-sanitizer --tool memcheck python scripts/train_small.py $b 2>&1 | grep -v "^$" | tail -4 | tee gpurun_out/${TAG}_memcheck_train_b$b.log
done
if [ -n "$2" ]; then
  echo "== ncu wgrad"; bash scripts/gpu_call_r3o.sh 2>&1 | tail -10 | tee gpurun_out/${TAG}_wgrad_ncu.txt
fi
