#!/bin/bash
N=${1:-8}; TAG=${2:-r02_n8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "== bench --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'ber_0db')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps({k: v for k, v in d.get('secondary', {}).items() if 'train' in k}, indent=1))
"
grep -v "^$" gpurun_out/${TAG}_bench.err | grep -v "Warning\|^\*\|OMP_NUM\|gfields\|\"\"\"" | tail -5
