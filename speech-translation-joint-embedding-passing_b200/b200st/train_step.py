"""The training step that wraps the hot path, for every trainer mode of the reference — mirrors of
`Trainer._train_batch` in trainer/trainer_st.py:211-299, trainer_asr_st.py:253-357, trainer_mt.py:199-282 and
trainer_asr.py:199-283, and nothing else from the trainers.

These classes are NOT what makes the repo a drop-in: the reference's own, unmodified trainers run on the modules of
this package as they are (b200st/dropin.py, tests/test_dropin_reference.py).  They exist because the reference's step
reads the loss back to the host once per minibatch (`get_loss()`, trainer_st.py:287), which cannot be captured into a
CUDA graph: `_train_batch_device` is the same arithmetic with the loss kept as a device scalar (what
b200st.graph.GraphedTrainStep captures and bench.py times), plus two options the reference has no room for — the
fused softmax + NLL kernel (`fused_loss`) and a data-parallel gradient reducer (`reducer`, b200st/dp.py).

Out of scope: epoch loop, rollback / LR-halving / early-stop policy, BLEU evaluation,
tensorboard, checkpoint I/O — with the drop-in those stay the reference's own host code.

Loss assembly per minibatch (identical in all four trainers): sum of -logp over the non-PAD targets, divided by that
minibatch's #non-PAD (`normalise_loss`), times the mode's `loss_coeff`, divided by n_minibatch; gradients accumulate
over the minibatches; then optimizer.step() / zero_grad().
"""
import torch

from b200st import runtime as rt
from b200st.hostutil import PAD, check_device
from modules.loss import NLLLoss


class _TrainStep(object):
    MODE = None            # forward_train mode string
    # (out_dict key, which ids are the targets, drop the last position?, loss_coeff key | None, result key)
    LOSSES = ()
    NEEDS_ACOUS = True

    def __init__(self, use_gpu=True, batch_size=64, minibatch_partition=1, eval_with_mask=True,
                 normalise_loss=True, loss_coeff=None, optimizer=None, reducer=None, max_grad_norm=1.0,
                 fused_loss=True):
        self.use_gpu = use_gpu
        self.device = check_device(use_gpu)
        self.batch_size = batch_size
        self.minibatch_partition = minibatch_partition
        self.minibatch_size = int(batch_size / minibatch_partition)       # trainer_base.py:85
        self.eval_with_mask = eval_with_mask
        self.normalise_loss = normalise_loss
        self.loss_coeff = loss_coeff or {'nll_asr': 1.0, 'nll_st': 1.0, 'nll_mt': 1.0}
        self.optimizer = optimizer
        self.reducer = reducer
        self.max_grad_norm = max_grad_norm
        # ST only: fused softmax + masked NLL (+ its gradient) straight from the logits instead of
        # log_softmax -> NLLLoss; same value and gradient, no [B, L, V] log-probability tensor
        self.fused_loss = fused_loss

    def _train_batch(self, model, batch_items, dataset=None, step=0, total_steps=0):
        res = self._train_batch_device(model, batch_items)
        if self.optimizer is not None:
            self.optimizer.step()
            model.zero_grad()
        # one D2H read per batch (the reference reads .item() per minibatch, trainer_st.py:287)
        out = {'nll_loss_de': 0, 'nll_loss_en': 0}
        for (_, _, _, _, name), v in zip(self.LOSSES, res if isinstance(res, tuple) else (res,)):
            out[name] = float(v)
        if len(self.LOSSES) == 1 and self.LOSSES[0][4] == 'nll_loss_de':
            out.pop('nll_loss_en')
        return out

    def _masked_loss(self, logps, ids, n_minibatch, coeff):
        """loss.py:116-132 + trainer_st.py:268-288 for one (log-probabilities, target ids) pair."""
        loss = NLLLoss()
        loss.reset()
        tgt = ids[:, 1:]
        if not self.eval_with_mask:
            loss.eval_batch(logps.reshape(-1, logps.size(-1)), tgt.reshape(-1))
            loss.norm_term = 1.0 * ids.size(0) * tgt.size(1)
        else:
            keep = ids.data.ne(PAD)[:, 1:]
            loss.eval_batch_with_mask(logps.reshape(-1, logps.size(-1)), tgt.reshape(-1), keep.reshape(-1))
            loss.norm_term = 1.0 * torch.sum(keep)
        if self.normalise_loss:
            loss.normalise()
        loss.acc_loss = loss.acc_loss * coeff / n_minibatch
        return loss

    def _train_batch_device(self, model, batch_items):
        """The step without any host synchronisation (losses stay device scalars), so it can be captured into a
        CUDA graph (b200st.graph.GraphedTrainStep).  Returns one scalar per entry of LOSSES (a bare scalar if one)."""
        batch_src_ids = batch_items['srcid'][0]
        batch_tgt_ids = batch_items['tgtid'][0] if 'tgtid' in batch_items else None
        batch_acous_feats = batch_items['acous_feat'][0] if self.NEEDS_ACOUS else None
        batch_acous_lengths = batch_items['acouslen'] if self.NEEDS_ACOUS else None
        batch_size = batch_src_ids.size(0)
        n_minibatch = int(batch_size / self.minibatch_size)
        n_minibatch += int(batch_size % self.minibatch_size > 0)
        totals = [0] * len(self.LOSSES)
        for bidx in range(n_minibatch):
            if self.reducer is not None:     # reduce the ACCUMULATED gradient once, from the last minibatch's backward
                self.reducer.arm(bidx == n_minibatch - 1)
            i_start = bidx * self.minibatch_size
            i_end = min(i_start + self.minibatch_size, batch_size)
            kw = {}
            ids = {'src': batch_src_ids[i_start:i_end].to(device=self.device, non_blocking=True)}
            if batch_tgt_ids is not None and self.MODE != 'ASR':
                ids['tgt'] = batch_tgt_ids[i_start:i_end].to(device=self.device, non_blocking=True)
                kw['tgt'] = ids['tgt']
            if self.NEEDS_ACOUS:
                acous_lengths = batch_acous_lengths[i_start:i_end]
                if torch.is_tensor(acous_lengths) and acous_lengths.is_cuda:     # device-resident lengths: no sync
                    acous_len = batch_acous_feats.size(1)
                else:
                    acous_len = max(int(n) for n in acous_lengths)
                    acous_len = acous_len + 8 - acous_len % 8                    # trainer_st.py:252
                kw['acous_feats'] = batch_acous_feats[i_start:i_end, :acous_len].to(device=self.device,
                                                                                   non_blocking=True)
                kw['acous_lens'] = acous_lengths
            if self.MODE == 'ST' and self.fused_loss and self.eval_with_mask and self.normalise_loss:
                # same loss (trainer_st.py:268-288), formed by the fused softmax + NLL kernel from the logits
                keep = ids['tgt'].data.ne(PAD)[:, 1:]
                scale = (self.loss_coeff['nll_st'] / n_minibatch) / torch.sum(keep).to(torch.float32)
                out_dict = model.forward_train(ids['src'], mode='ST', use_gpu=self.use_gpu,
                                               st_loss_scale=scale.reshape(1), **kw)
                loss = out_dict['loss_st']
                loss.backward()
                totals[0] = totals[0] + loss.detach()
                continue
            out_dict = model.forward_train(ids['src'], mode=self.MODE, use_gpu=self.use_gpu, **kw)
            total = None
            for j, (key, which, drop_last, coeff, _) in enumerate(self.LOSSES):
                logps = out_dict[key][:, :-1, :] if drop_last else out_dict[key]
                part = self._masked_loss(logps, ids[which], n_minibatch, 1.0 if coeff is None else self.loss_coeff[coeff])
                totals[j] = totals[j] + part.acc_loss.detach()
                if total is None:
                    total = part
                else:
                    total.add(part)                                          # trainer_asr_st.py:345-346
            total.backward()
        rt.join_deferred()          # weight-gradient GEMMs forked onto side streams during backward
        if self.reducer is not None:
            self.reducer.finish()
        return totals[0] if len(totals) == 1 else tuple(totals)


class Trainer_ST(_TrainStep):
    """trainer/trainer_st.py:211-299: free-running LAS -> mix -> Transformer; NLL of logps_st[:, :-1] vs tgt[:, 1:]."""
    MODE = 'ST'
    LOSSES = (('logps_st', 'tgt', True, 'nll_st', 'nll_loss_de'),)


class Trainer_ASR_ST(_TrainStep):
    """trainer/trainer_asr_st.py:253-357: teacher-forced LAS with SpecAug feeds the ST branch; the ST loss and the ASR
    loss (logps_asr vs src[:, 1:]) are each normalised by their own #non-PAD and their SUM is back-propagated."""
    MODE = 'ASR_ST'
    LOSSES = (('logps_st', 'tgt', True, 'nll_st', 'nll_loss_de'), ('logps_asr', 'src', False, 'nll_asr', 'nll_loss_en'))


class Trainer_MT(_TrainStep):
    """trainer/trainer_mt.py:199-282: static embedding + the constant average dynamic embedding; no acoustics."""
    MODE = 'MT'
    LOSSES = (('logps_mt', 'tgt', True, 'nll_mt', 'nll_loss_de'),)
    NEEDS_ACOUS = False


class Trainer_ASR(_TrainStep):
    """trainer/trainer_asr.py:199-283: teacher-forced LAS with SpecAug alone; no loss coefficient."""
    MODE = 'ASR'
    LOSSES = (('logps_asr', 'src', False, None, 'nll_loss_en'),)
