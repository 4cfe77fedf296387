#!/bin/bash
# round 2, 2 GPUs: NCCL test, bench --gpus 2, the unmodified reference's training under torchrun through the launcher
TAG=${1:-r02_n2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== 2-rank test"; timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.log
echo "== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'ber_0db')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('secondary'), indent=1))
"
tail -5 gpurun_out/${TAG}_bench.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
echo "== dropin scratch training, 2 ranks"; timeout 900 python scripts/run_reference_dropin.py --mode scratch --nproc 2 2>&1 | tail -4
grep -n "Epoch\|Test set" gpurun_out/dropin_scratch_n2.log | head -20
