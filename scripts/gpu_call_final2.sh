#!/bin/bash
# closing run with the straddling x3 kernel: full GPU suite, smoke, x3 accuracy + sanitizer + ncu capture (replays), bench (both arms)
TAG=${1:-r02_final2}
mkdir -p gpurun_out
echo "== tests"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.log
echo "== x3 accuracy"; timeout 600 python scripts/x3_accuracy.py > gpurun_out/${TAG}_x3_accuracy.json 2>gpurun_out/${TAG}_acc.err; grep -A3 "x3\"" gpurun_out/${TAG}_x3_accuracy.json | grep "y_max\|x3\|p9999" | tr -d '\n' | sed 's/"c/\n"c/g'; echo
echo "== x3 bench"; timeout 300 python scripts/x3_bench.py 50000 2>&1 | tail -22 | tee gpurun_out/${TAG}_x3_bench.json
bash scripts/gpu_call_sanitize.sh
echo "== x3 ncu (decoder launch)"; timeout 600 ncu --set full --clock-control none --import-source on -k x3_kernel -s 13 -c 1 -o gpurun_out/${TAG}_x3_dec -f \
    python scripts/x3_bench.py 50000 > gpurun_out/${TAG}_x3_dec_full.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/${TAG}_x3_dec_full.log | head -5
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ber_0db', 'clocks')}, d['roofline']['frac'], d['roofline'].get('frac_fastest_launch'), d['e2e']['value'])
print(json.dumps(d.get('secondary', {}), indent=1))
"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
