#!/bin/bash
# round 2, call A: timeline of the fused decoder, layer / iteration scan, the unmodified reference through the launcher
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== timeline"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_timeline.so TL_TAG=r02_dec_timeline_base timeout 300 python scripts/dec_timeline.py > gpurun_out/tl.log 2>&1; tail -60 gpurun_out/tl.log
echo "== scan"; SCAN='[(2,6),(3,6),(5,6),(5,3),(5,1)]' timeout 300 python scripts/dec_scan.py 2>&1 | tail -6
echo "== dropin eval (turboae_b200 classes)"; timeout 900 python scripts/run_reference_dropin.py --mode eval --seed 7 --num-block 200000 --batch-size 50000 2>&1 | tail -3
echo "== stock eval (reference classes, torch eager CUDA)"; timeout 900 python scripts/run_reference_dropin.py --mode eval --stock --seed 7 --num-block 200000 --batch-size 50000 2>&1 | tail -3
echo "== dropin train"; timeout 900 python scripts/run_reference_dropin.py --mode train --seed 7 2>&1 | tail -3
echo "== dropin README eval (unseeded, batch 500, 100k blocks)"; timeout 900 python scripts/run_reference_dropin.py --mode eval --num-block 100000 --batch-size 500 --out gpurun_out/dropin_eval_readme.json 2>&1 | tail -3
ls gpurun_out | head -50
