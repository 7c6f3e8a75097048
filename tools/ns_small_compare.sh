#!/bin/bash
# C2 / C3 (small M) with and without the tensor-memory accumulator parking: which tile is better when a tile is only 2 - 8 chunks?
for ns in 0 1; do for w in c2 c3; do
  PLS_B200_TILE_NS=$ns python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/ns_$ns_$w.json
  python - "$ns" "$w" /tmp/ns_$ns_$w.json <<'PY'
import json, sys
d = json.load(open(sys.argv[3])); r = d["roofline"]
print("tile_ns_env", sys.argv[1], sys.argv[2], round(d["value"]), round(d["ms_per_step"], 4),
      {k: (round(v["tflops"], 2), round(v["ms_per_launch"], 4)) for k, v in r["per_role"].items()}, round(r["kernel_share_of_step_this_gpu"], 3))
PY
done; done
