#!/bin/bash
# Round-2 ncu evidence for the default bench command (C4, generated Gram, NS = 2 kernels):
#   1. launch list (gpu__time_duration.sum per launch)                       -> gpurun_out/launches_r02.csv
#   2. ncu --set full (+ the DMMA sub-pipe counters) of one forward launch   -> gpurun_out/r02_fwd.ncu-rep
#   3. the same for one backward launch                                       -> gpurun_out/r02_bwd.ncu-rep
# (one Dc chunk per step at C4: two contraction launches per step, so launch 6 / 7 = the forward / backward of the 4th step)
# Each capture runs only after the same command has exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e"
DMMA=sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --metrics $DMMA --clock-control none --import-source on -k regex:gen_gemm_kernel -s 6 -c 1 -f -o gpurun_out/r02_fwd $CMD > gpurun_out/r02_ncu_fwd.log 2>&1
ncu --set full --metrics $DMMA --clock-control none --import-source on -k regex:gen_gemm_kernel -s 7 -c 1 -f -o gpurun_out/r02_bwd $CMD > gpurun_out/r02_ncu_bwd.log 2>&1
tail -2 gpurun_out/r02_ncu_fwd.log gpurun_out/r02_ncu_bwd.log
