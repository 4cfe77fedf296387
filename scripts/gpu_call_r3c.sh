#!/bin/bash
# bench line with the leaner training step + ncu launch list of the training step
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/r3c_bench.err | tee gpurun_out/r3c_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('secondary', {}), indent=1))
"
tail -3 gpurun_out/r3c_bench.err
echo "== train launch list"
TRAIN_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r3c_train_launches.csv \
    python scripts/train_bench.py > gpurun_out/r3c_train_launches.log 2>&1
tail -2 gpurun_out/r3c_train_launches.log
