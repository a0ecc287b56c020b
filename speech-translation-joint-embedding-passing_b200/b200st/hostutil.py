"""Host-side constants and helpers the model layer needs when it runs WITHOUT the reference tree on sys.path.

With the reference present (drop-in use, b200st/dropin.py) `utils.config` / `utils.misc` are the reference's own,
unmodified files; these are the two things of them the hot path touches, restated so that the package is
self-contained on a box that has no reference checkout (reference: utils/config.py:1-7, utils/misc.py:124-133)."""
import torch

PAD_TOKEN, UNK_TOKEN, BOS_TOKEN, EOS_TOKEN, SPC_TOKEN = '<pad>', '<unk>', '<s>', '</s>', '<spc>'
PAD, UNK, BOS, EOS, SPC = 0, 1, 2, 3, 4


def check_device(use_gpu):
    """Same contract as the reference: CUDA when requested and present, else CPU.  The b200st kernels
    themselves are CUDA-only and raise on CPU tensors — there is no CPU compute path."""
    if use_gpu and torch.cuda.is_available():
        return torch.device('cuda')
    return torch.device('cpu')
