"""Host-side helpers the model layer imports (reference: utils/misc.py:124-133,162-171).

Only what the hot path touches is mirrored; dataset / plotting / checkpoint-averaging utilities of the
reference's utils/misc.py are host policy and out of scope."""
import random

import numpy as np
import torch


def check_device(use_gpu):
    """Same contract as the reference: CUDA when requested and present, else CPU.  The b200st kernels
    themselves are CUDA-only and raise on CPU tensors — there is no CPU compute path."""
    if use_gpu and torch.cuda.is_available():
        return torch.device('cuda')
    return torch.device('cpu')


def set_global_seeds(i):
    torch.manual_seed(i)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(i)
    np.random.seed(i)
    random.seed(i)
