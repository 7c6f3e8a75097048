#!/bin/bash
# C4 step with different bounds on the Dc chunk (the only N-sized intermediate): fewer, larger launches per step
for g in 8 16 33; do
  PLS_B200_DC_BUDGET_GIB=$g python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/dc_$g.json
  python - "$g" /tmp/dc_$g.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); r = d["roofline"]
print("dc_budget_gib", sys.argv[1], round(d["value"], 1), round(d["ms_per_step"], 2), round(r["frac"], 4), d["config"]["row_chunk"],
      {k: round(v["tflops"], 2) for k, v in r["per_role"].items()}, round(r["kernel_share_of_step_this_gpu"], 4))
PY
done
