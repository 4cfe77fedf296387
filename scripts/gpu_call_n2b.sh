#!/bin/bash
TAG=${1:-r02_n2b}
mkdir -p gpurun_out
echo "== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'ber_0db')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('secondary'), indent=1))
"
grep -v "^$" gpurun_out/${TAG}_bench.err | grep -v "Warning\|^\*\|OMP_NUM\|gfields\|\"\"\"" | tail -8
