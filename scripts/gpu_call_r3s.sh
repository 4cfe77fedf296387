#!/bin/bash
for rep in 1 2; do
  TRAIN_STEPS=60 timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | cut -c100-330
done
python - <<'PY'
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
print(d["secondary"]["train_step"], "|", d["secondary"]["train_step_graphed"])
PY
