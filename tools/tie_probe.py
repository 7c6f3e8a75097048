"""Shows that the reference selector's choice on EXACT ties is a property of numpy's unstable argsort on the host it runs on:
the README demo (BASELINE config 1) re-selected by the oracle's literal restatement of conditional_variance.py:105-109.
    python tools/tie_probe.py                                   -> equals the committed reference run (AVX512 / AVX2 hosts)
    NPY_DISABLE_CPU_FEATURES="AVX2 FMA3 AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR" python tools/tie_probe.py
                                                                -> different indices (numpy's scalar introsort path)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pls_oracle import RBFScaleKernel, conditional_variance_select, set_seed  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "readme_demo.npz"))
set_seed(0)
x = torch.from_numpy(g["x"])
k = RBFScaleKernel(torch.tensor([float(g["lengthscale"])], dtype=torch.float64), float(g["outputscale"]))
_, idx = conditional_variance_select(x, 10, k)
print("numpy default argsort:", idx.tolist(), "reference run:", g["induce_idx"].tolist(), "equal:", idx.tolist() == g["induce_idx"].tolist())
