#!/bin/bash
# the whole GPU suite and the randomised shape sweep with the cluster variant forced wherever the shape allows
PLS_B200_CLUSTER=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PLS_B200_CLUSTER=2 timeout 600 python tools/sweep_shapes.py --count 150 --seed 11 2>&1 | tail -2
PLS_B200_CLUSTER=2 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/cl2.json
python - <<'PY'
import json
d = json.load(open("/tmp/cl2.json")); r = d["roofline"]
print("cluster=2 bench", round(d["value"], 1), round(r["frac"], 4), {k: round(v["tflops"], 2) for k, v in r["per_role"].items()})
PY
