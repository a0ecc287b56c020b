"""The ST training step that wraps the hot path — mirror of Trainer_ST._train_batch
(reference: trainer/trainer_st.py:211-299), and nothing else from the trainer.

Out of scope: epoch loop, rollback / LR-halving / early-stop policy, BLEU
evaluation, tensorboard, checkpoint I/O.  `Trainer_ST` here keeps the constructor arguments the step
depends on and the exact loss assembly: per minibatch, sum of -logp over non-PAD targets of
`logps_st[:, :-1]` vs `tgt[:, 1:]`, divided by that minibatch's #non-PAD, times loss_coeff['nll_st'],
divided by n_minibatch; gradients accumulate over minibatches; then optimizer.step() / zero_grad().
With `reducer` set (b200st.dp.GradAllReducer) the accumulated gradient is mean-all-reduced across
data-parallel ranks before the optimizer step.
"""
import torch

from b200st import runtime as rt
from modules.loss import NLLLoss
from utils.config import PAD
from utils.misc import check_device


class Trainer_ST(object):

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
        self.loss_coeff = loss_coeff or {'nll_st': 1.0}
        self.optimizer = optimizer
        self.reducer = reducer
        self.max_grad_norm = max_grad_norm
        # fused softmax + masked NLL (+ its gradient) straight from the logits instead of log_softmax -> NLLLoss;
        # same value and gradient, no [B, L, V] log-probability tensor
        self.fused_loss = fused_loss

    def _train_batch(self, model, batch_items, dataset=None, step=0, total_steps=0):
        resloss_de = self._train_batch_device(model, batch_items)
        if self.optimizer is not None:
            self.optimizer.step()
            model.zero_grad()
        # one D2H read per batch (the reference reads .item() per minibatch, trainer_st.py:287)
        return {'nll_loss_de': float(resloss_de)}

    def _train_batch_device(self, model, batch_items):
        """The same step without any host synchronisation (returns the loss as a device scalar), so it
        can be captured into a CUDA graph (b200st.graph.GraphedTrainStep)."""
        batch_src_ids = batch_items['srcid'][0]
        batch_tgt_ids = batch_items['tgtid'][0]
        batch_acous_feats = batch_items['acous_feat'][0]
        batch_acous_lengths = batch_items['acouslen']
        batch_size = batch_src_ids.size(0)
        n_minibatch = int(batch_size / self.minibatch_size)
        n_minibatch += int(batch_size % self.minibatch_size > 0)
        resloss_de = 0
        for bidx in range(n_minibatch):
            loss_de = NLLLoss()
            loss_de.reset()
            i_start = bidx * self.minibatch_size
            i_end = min(i_start + self.minibatch_size, batch_size)
            acous_lengths = batch_acous_lengths[i_start:i_end]
            if torch.is_tensor(acous_lengths) and acous_lengths.is_cuda:     # device-resident lengths: no sync
                acous_len = batch_acous_feats.size(1)
            else:
                acous_len = max(int(n) for n in acous_lengths)
                acous_len = acous_len + 8 - acous_len % 8                    # trainer_st.py:252
            src_ids = batch_src_ids[i_start:i_end].to(device=self.device, non_blocking=True)
            tgt_ids = batch_tgt_ids[i_start:i_end].to(device=self.device, non_blocking=True)
            acous_feats = batch_acous_feats[i_start:i_end, :acous_len].to(device=self.device,
                                                                          non_blocking=True)
            non_padding_mask_tgt = tgt_ids.data.ne(PAD)
            if self.fused_loss and self.eval_with_mask and self.normalise_loss:
                # same loss (trainer_st.py:268-288), formed by the fused softmax + NLL kernel from the logits
                scale = (self.loss_coeff['nll_st'] / n_minibatch) / torch.sum(non_padding_mask_tgt[:, 1:]).to(torch.float32)
                out_dict = model.forward_train(src_ids, tgt=tgt_ids, acous_feats=acous_feats,
                                               acous_lens=acous_lengths, mode='ST', use_gpu=self.use_gpu,
                                               st_loss_scale=scale.reshape(1))
                loss = out_dict['loss_st']
                loss.backward()
                resloss_de = resloss_de + loss.detach()
                continue
            out_dict = model.forward_train(src_ids, tgt=tgt_ids, acous_feats=acous_feats,
                                           acous_lens=acous_lengths, mode='ST', use_gpu=self.use_gpu)
            logps_de = out_dict['logps_st'][:, :-1, :]
            if not self.eval_with_mask:
                loss_de.eval_batch(logps_de.reshape(-1, logps_de.size(-1)), tgt_ids[:, 1:].reshape(-1))
                loss_de.norm_term = 1.0 * tgt_ids.size(0) * tgt_ids[:, 1:].size(1)
            else:
                loss_de.eval_batch_with_mask(logps_de.reshape(-1, logps_de.size(-1)),
                                             tgt_ids[:, 1:].reshape(-1),
                                             non_padding_mask_tgt[:, 1:].reshape(-1))
                loss_de.norm_term = 1.0 * torch.sum(non_padding_mask_tgt[:, 1:])
            if self.normalise_loss:
                loss_de.normalise()
            loss_de.acc_loss = loss_de.acc_loss * self.loss_coeff['nll_st']
            loss_de.acc_loss = loss_de.acc_loss / n_minibatch
            loss_de.backward()
            resloss_de = resloss_de + loss_de.acc_loss.detach()
        rt.join_deferred()          # weight-gradient GEMMs forked onto side streams during backward
        if self.reducer is not None:
            self.reducer.finish()
        return resloss_de
