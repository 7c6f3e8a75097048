#!/bin/bash
# launch lists (kernel shares) of the C2 and C3 steps, round 2
for w in c2 c3; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/r02_plain_$w.json 2> /dev/null || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${w}_r02.csv $CMD > /dev/null 2>&1
done
