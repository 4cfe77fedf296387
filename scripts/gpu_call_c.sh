#!/bin/bash
# round 2: decoder-only checks of a kernel variant: decoder parity tests, timeline, 2 bench runs (TAG)
TAG=${1:-r02_c}
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_tc.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
echo "== timeline"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_timeline.so TL_TAG=${TAG}_timeline timeout 300 python scripts/dec_timeline.py > gpurun_out/${TAG}_tl.log 2>&1; head -6 gpurun_out/${TAG}_tl.log
for i in 1 2; do
echo "== bench $i"; BENCH_SKIP_SECONDARY=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench$i.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline']['launch_ms_min'], d['e2e']['value'])
"
done
tail -3 gpurun_out/${TAG}_bench.err
