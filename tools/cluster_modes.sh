#!/bin/bash
# C4 step under the three cluster modes: 1 = never, 0 = default rule (backward role only), 2 = both roles
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/sweep_shapes.py --count 100 --seed 5 2>&1 | tail -1
for c in 1 0 2; do
  PLS_B200_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/clm_$c.json
  python - "$c" /tmp/clm_$c.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); r = d["roofline"]
print("cluster mode", sys.argv[1], round(d["value"], 1), round(d["ms_per_step"], 2), round(r["frac"], 4), {k: round(v["tflops"], 2) for k, v in r["per_role"].items()}, d["config"]["row_chunk"])
PY
done
