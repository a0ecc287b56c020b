"""`NLLLoss` with the object protocol the reference trainers drive (reference: modules/loss.py:36-132; call sites
trainer_st.py:235-288): reset / eval_batch / eval_batch_with_mask / norm_term / normalise / mul / add / get_loss /
backward / acc_loss.  The summed masked negative log-likelihood and its gradient are b200st kernels
(`b200st_masked_nll_fwd/bwd`) instead of nn.NLLLoss(reduction='none') + masked_select + sum.

Only NLLLoss is on the hot path.  The reference's trainers and translate.py also IMPORT BCELoss, CrossEntropyLoss,
KLDivLoss and MSELoss from this module name (never used by the ST path); with the reference tree on sys.path those
come from the reference's own file through `inherit_shadowed` (b200st/dropin.py), so this file does not restate them."""
from b200st import functional as BF
from b200st.dropin import inherit_shadowed


class NLLLoss(object):
    name = _NAME = 'NLLLoss'

    def __init__(self, weight=None, mask=None, reduction='none'):
        if weight is not None or mask is not None:
            raise NotImplementedError('per-class weights / constructor masks are never used by the trainers')
        self.mask = None
        self.reset()

    # ---- accumulator protocol (loss.py:36-89)
    def reset(self):
        self.acc_loss, self.norm_term = 0, 1

    def normalise(self):
        self.acc_loss = self.acc_loss / (1.0 * self.norm_term)

    def mul(self, coeff):
        self.acc_loss = self.acc_loss * coeff

    def add(self, other):
        self.acc_loss = self.acc_loss + other.acc_loss

    def get_loss(self):
        return 0 if isinstance(self.acc_loss, int) else self.acc_loss.detach().item()

    def backward(self, retain_graph=False):
        if isinstance(self.acc_loss, int):
            raise ValueError('No loss to back propagate.')
        self.acc_loss.backward(retain_graph=retain_graph)

    def cuda(self):                 # the reference moves its criterion module; there is none here
        return self

    # ---- the two evaluators (loss.py:116-132): outputs [rows, V] log-probabilities, target [rows], mask [rows]
    def eval_batch(self, outputs, target):
        self.acc_loss = self.acc_loss + BF.masked_nll_sum(outputs, target, None)

    def eval_batch_with_mask(self, outputs, target, mask):
        self.acc_loss = self.acc_loss + BF.masked_nll_sum(outputs, target, mask)


inherit_shadowed(globals())
