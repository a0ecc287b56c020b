"""Run the oracle restatement (oracle/st_oracle.py) with its tensors on cuda:0.  TEST INFRASTRUCTURE ONLY.

The oracle is the reference's algorithm written with the same torch primitives at the same call sites; at BASELINE.json's
full sizes it takes minutes per step on host cores, so the full-size parity tests execute THE SAME oracle code with stock
PyTorch CUDA kernels (cuDNN LSTM, cuBLAS, ATen softmax / LayerNorm / autograd) in strict fp32 (TF32 off for matmul and
cuDNN).  Nothing here touches the product path: the oracle's results are moved to the host before the product runs.

Shims (none touches arithmetic):
  * pack_padded_sequence wants its lengths on the host;
  * cuDNN's LSTM backward needs the training-mode forward (`train=True`; dropout is 0.0 either way);
  * the oracle builds its masks / index tensors on the default device.
"""
import contextlib

import torch


@contextlib.contextmanager
def oracle_on_cuda(device='cuda:0'):
    dev = torch.device(device)
    rnn = torch.nn.utils.rnn
    pps, lstm = rnn.pack_padded_sequence, torch._VF.lstm
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rnn.pack_padded_sequence = lambda x, lens, **kw: pps(x, lens.cpu() if torch.is_tensor(lens) else lens, **kw)
    torch._VF.lstm = lambda *a: lstm(*a[:7], True, *a[8:])
    torch.set_default_device(dev)
    try:
        yield dev
    finally:
        torch.set_default_device('cpu')
        torch._VF.lstm = lstm
        rnn.pack_padded_sequence = pps
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.synchronize()


def params_to(P, dev, requires_grad=True):
    return {k: v.detach().to(dev).clone().requires_grad_(requires_grad) for k, v in P.items()}


def grads_to_host(Pg):
    """{name: grad on the host} for every parameter that received a non-zero gradient."""
    return {k: v.grad.detach().cpu() for k, v in Pg.items() if v.grad is not None and float(v.grad.abs().sum()) > 0}
