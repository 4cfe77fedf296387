#!/bin/bash
# round 2: parity tests + timeline + short bench of the current kernel (TAG = label for the outputs)
TAG=${1:-r02_b}
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
echo "== timeline"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_timeline.so TL_TAG=${TAG}_timeline timeout 300 python scripts/dec_timeline.py > gpurun_out/${TAG}_tl.log 2>&1; head -50 gpurun_out/${TAG}_tl.log; tail -12 gpurun_out/${TAG}_tl.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline']['launch_ms_min'], d['e2e']['value'], d.get('secondary'))
"
tail -3 gpurun_out/${TAG}_bench.err
