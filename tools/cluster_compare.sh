#!/bin/bash
# experiment: pairs of CTAs sharing the Gram generation through distributed shared memory (PLS_B200_CLUSTER=2)
PLS_B200_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "parked" 2>&1 | tail -3
for c in 0 2; do
  echo "cluster $c"
  PLS_B200_CLUSTER=$c timeout 200 python tools/bench_gen_gemm.py --n 262144 --reps 3
done
