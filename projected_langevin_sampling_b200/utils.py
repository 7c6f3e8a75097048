"""set_seed (reference: src/utils.py:8-22) -- the parity protocol relies on it: the selector permutes with numpy's
global generator (conditional_variance.py:60) and the Langevin noise comes from torch's global CPU generator."""
import os
import random

import numpy as np
import torch


def set_seed(seed: int = 42) -> None:
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    os.environ["PYTHONHASHSEED"] = str(seed)
