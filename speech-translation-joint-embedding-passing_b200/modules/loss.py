"""Loss objects with the protocol the reference trainers use (reference: modules/loss.py:12-132):
reset / eval_batch / eval_batch_with_mask / norm_term / normalise / mul / add / get_loss / backward.
Only NLLLoss is on the ST path (trainer_st.py:235-288); it runs the b200st masked-NLL kernels."""
import torch
import torch.nn as nn

from b200st import functional as BF


class Loss(object):

    def __init__(self, name, criterion):
        self.name = name
        self.criterion = criterion
        if not issubclass(type(self.criterion), nn.modules.loss._Loss):
            raise ValueError("Criterion has to be a subclass of torch.nn._Loss")
        self.acc_loss = 0
        self.norm_term = 1

    def reset(self):
        self.acc_loss = 0
        self.norm_term = 1

    def get_loss(self):
        raise NotImplementedError

    def eval_batch(self, outputs, target):
        raise NotImplementedError

    def cuda(self):
        self.criterion.cuda()

    def backward(self, retain_graph=False):
        if type(self.acc_loss) is int:
            raise ValueError("No loss to back propagate.")
        self.acc_loss.backward(retain_graph=retain_graph)

    def normalise(self):
        self.acc_loss = self.acc_loss / (1.0 * self.norm_term)

    def mul(self, coeff):
        self.acc_loss = self.acc_loss * coeff

    def add(self, loss):
        self.acc_loss = self.acc_loss + loss.acc_loss


class NLLLoss(Loss):
    """Sum of -logp[target] (optionally over a mask); reduction='none' + masked_select + sum in the
    reference (loss.py:116-132).  `weight` is not used by any trainer and is not supported."""

    _NAME = "NLLLoss"

    def __init__(self, weight=None, mask=None, reduction='none'):
        if weight is not None or mask is not None:
            raise NotImplementedError('per-class weights are never used by the trainers')
        self.mask = mask
        super().__init__(self._NAME, nn.NLLLoss(weight=weight, reduction=reduction))

    def get_loss(self):
        if isinstance(self.acc_loss, int):
            return 0
        return self.acc_loss.data.detach().item()

    def eval_batch(self, outputs, target):
        self.acc_loss = self.acc_loss + BF.masked_nll_sum(outputs, target, None)

    def eval_batch_with_mask(self, outputs, target, mask):
        self.acc_loss = self.acc_loss + BF.masked_nll_sum(outputs, target, mask)
