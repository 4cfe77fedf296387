#!/bin/bash
mkdir -p gpurun_out
echo "== test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_noise or awgn" 2>&1 | tail -4
echo "== c4 through the reference trainer, --device-channel"
timeout 600 python scripts/run_reference_dropin.py --mode c4 --seed 7 --device-channel --out gpurun_out/r3k_dropin_c4_devch.json 2>&1 | cut -c1-400 | tail -3
echo "== c4 through the reference trainer, host channel"
timeout 600 python scripts/run_reference_dropin.py --mode c4 --seed 7 --out gpurun_out/r3k_dropin_c4.json 2>&1 | cut -c1-400 | tail -3
python - <<'PY'
import json
for f in ("r3k_dropin_c4_devch.json", "r3k_dropin_c4.json"):
    d = json.load(open("gpurun_out/" + f))["c4"]
    print(f, [round(x, 4) for x in d["pass_seconds"]], d.get("train_cw_per_s_median"), d["validate_lines"], "total s", round(d["seconds"], 1))
PY
