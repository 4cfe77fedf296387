#!/bin/bash
# closing run of the round: full GPU suite, smoke, bench (both arms), BASELINE config 4 through the reference's own trainer
# (this package's classes vs the reference's own classes on the same GPU)
TAG=${1:-r02_close}
mkdir -p gpurun_out
echo "== tests"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.log
echo "== c4 through the reference trainer (turboae_b200 classes)"
timeout 600 python scripts/run_reference_dropin.py --mode c4 --seed 7 --out gpurun_out/${TAG}_dropin_c4.json 2>&1 | cut -c1-600 | tail -3
echo "== c4 through the reference trainer (reference's own classes, torch eager)"
timeout 900 python scripts/run_reference_dropin.py --mode c4 --seed 7 --stock --out gpurun_out/${TAG}_stock_c4.json 2>&1 | cut -c1-600 | tail -3
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline'].get('frac_fastest_launch'), d['e2e']['value'])
print({k: v for k, v in d.get('secondary', {}).items() if 'train' in k})
"
