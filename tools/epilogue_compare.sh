#!/bin/bash
# forward throughput per cost epilogue at M = 256 and the headline Gaussian forward (regression check)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python tools/bench_gen_gemm.py --n 262144 --reps 5 --roles forward; done
python tools/bench_gen_gemm.py --n 262144 --reps 3 --m 256 --cost poisson --roles forward
python tools/bench_gen_gemm.py --n 262144 --reps 3 --m 256 --cost bernoulli --roles forward
python tools/bench_gen_gemm.py --n 262144 --reps 3 --m 256 --cost student_t --roles forward
tools/ns_small_compare.sh 2>&1 | head -2
