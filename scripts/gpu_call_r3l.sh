#!/bin/bash
mkdir -p gpurun_out
for opt in "--device-channel" ""; do
  echo "== README evaluation (100000 blocks per point, batch 500) $opt"
  timeout 900 python scripts/run_reference_dropin.py --mode eval --num-block 100000 --batch-size 500 --seed 7 $opt --out gpurun_out/r3l_eval$opt.json 2>&1 | cut -c1-300 | tail -2
done
python - <<'PY'
import json
for f in ("r3l_eval--device-channel.json", "r3l_eval.json"):
    d = json.load(open("gpurun_out/" + f))["eval"]
    print(f, "seconds", round(d["seconds"], 1), "ber_0db", d.get("ber_0db"), "max abs diff vs golden", max(d["abs_diff"]), all(d["within_1e-4"]))
PY
