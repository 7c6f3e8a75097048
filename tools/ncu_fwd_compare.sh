#!/bin/bash
# ncu --set full of ONE forward (and one backward) launch of the hot kernel at a C4-shaped slice, with and without the parked
# accumulator set (NS = 2 / NS = 1); reports land in gpurun_out/.
set -x
N=${N:-131072}
for ns in 0 1; do
  for role in forward backward; do
    PLS_B200_TILE_NS=$ns python tools/bench_gen_gemm.py --n $N --reps 1 --roles $role > gpurun_out/plain_${role}_ns$ns.log 2>&1 || exit 1
    PLS_B200_TILE_NS=$ns ncu --set full --clock-control none --import-source on -k regex:gen_gemm --launch-skip 1 -c 1 \
      --metrics sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum \
      -f -o gpurun_out/r2_${role}_ns$ns python tools/bench_gen_gemm.py --n $N --reps 1 --roles $role > gpurun_out/ncu_${role}_ns$ns.log 2>&1
  done
done
cat gpurun_out/plain_*_ns*.log
