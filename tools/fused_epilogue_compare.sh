#!/bin/bash
# fused training epilogue (cost derivative + cost sums from one forward, pls_forward_step_f64) per cost at M = 256
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in poisson bernoulli student_t multimodal; do
  python tools/bench_gen_gemm.py --n 262144 --reps 3 --m 256 --cost $c --roles forward --epilogue 3
done
python tools/bench_gen_gemm.py --n 262144 --reps 5 --roles forward
