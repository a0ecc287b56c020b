"""The joint ASR + ST training step — mirror of Trainer._train_batch of the reference's trainer/trainer_asr_st.py:253-357
(SURVEY.md §8 f-4), and nothing else from that trainer (epoch loop, evaluation, checkpointing are host policy).

Per minibatch: `forward_train(mode='ASR_ST')` = teacher-forced LAS with SpecAug (the ASR branch, Seq2seq.py:422-436) whose
cell values feed the mix + Transformer (the ST branch); two masked NLL losses — `logps_asr` against `src[:, 1:]` and
`logps_st[:, :-1]` against `tgt[:, 1:]` — each normalised by its own #non-PAD, scaled by `loss_coeff` and divided by
n_minibatch; their SUM is back-propagated (trainer_asr_st.py:341-346); then optimizer.step() / zero_grad().
"""
import torch

from b200st import runtime as rt
from modules.loss import NLLLoss
from utils.config import PAD
from utils.misc import check_device


class Trainer_ASR_ST(object):

    def __init__(self, use_gpu=True, batch_size=64, minibatch_partition=1, eval_with_mask=True,
                 normalise_loss=True, loss_coeff=None, optimizer=None, reducer=None, max_grad_norm=1.0):
        self.use_gpu = use_gpu
        self.device = check_device(use_gpu)
        self.batch_size = batch_size
        self.minibatch_partition = minibatch_partition
        self.minibatch_size = int(batch_size / minibatch_partition)
        self.eval_with_mask = eval_with_mask
        self.normalise_loss = normalise_loss
        self.loss_coeff = loss_coeff or {'nll_asr': 1.0, 'nll_st': 1.0}
        self.optimizer = optimizer
        self.reducer = reducer
        self.max_grad_norm = max_grad_norm

    def _train_batch(self, model, batch_items, dataset=None, step=0, total_steps=0):
        de, en = self._train_batch_device(model, batch_items)
        if self.optimizer is not None:
            self.optimizer.step()
            model.zero_grad()
        return {'nll_loss_de': float(de), 'nll_loss_en': float(en)}

    def _train_batch_device(self, model, batch_items):
        batch_src_ids = batch_items['srcid'][0]
        batch_tgt_ids = batch_items['tgtid'][0]
        batch_acous_feats = batch_items['acous_feat'][0]
        batch_acous_lengths = batch_items['acouslen']
        batch_size = batch_src_ids.size(0)
        n_minibatch = int(batch_size / self.minibatch_size)
        n_minibatch += int(batch_size % self.minibatch_size > 0)
        resloss_de = resloss_en = 0
        for bidx in range(n_minibatch):
            loss_de, loss_en = NLLLoss(), NLLLoss()
            loss_de.reset()
            loss_en.reset()
            i_start = bidx * self.minibatch_size
            i_end = min(i_start + self.minibatch_size, batch_size)
            acous_lengths = batch_acous_lengths[i_start:i_end]
            if torch.is_tensor(acous_lengths) and acous_lengths.is_cuda:
                acous_len = batch_acous_feats.size(1)
            else:
                acous_len = max(int(n) for n in acous_lengths)
                acous_len = acous_len + 8 - acous_len % 8                    # trainer_asr_st.py:300
            src_ids = batch_src_ids[i_start:i_end].to(device=self.device, non_blocking=True)
            tgt_ids = batch_tgt_ids[i_start:i_end].to(device=self.device, non_blocking=True)
            acous_feats = batch_acous_feats[i_start:i_end, :acous_len].to(device=self.device, non_blocking=True)
            non_padding_mask_src = src_ids.data.ne(PAD)
            non_padding_mask_tgt = tgt_ids.data.ne(PAD)
            out_dict = model.forward_train(src_ids, tgt=tgt_ids, acous_feats=acous_feats, acous_lens=acous_lengths,
                                           mode='ASR_ST', use_gpu=self.use_gpu)
            logps_de = out_dict['logps_st'][:, :-1, :]
            logps_en = out_dict['logps_asr']
            if not self.eval_with_mask:
                loss_de.eval_batch(logps_de.reshape(-1, logps_de.size(-1)), tgt_ids[:, 1:].reshape(-1))
                loss_de.norm_term = 1.0 * tgt_ids.size(0) * tgt_ids[:, 1:].size(1)
                loss_en.eval_batch(logps_en.reshape(-1, logps_en.size(-1)), src_ids[:, 1:].reshape(-1))
                loss_en.norm_term = 1.0 * src_ids.size(0) * src_ids[:, 1:].size(1)
            else:
                loss_de.eval_batch_with_mask(logps_de.reshape(-1, logps_de.size(-1)), tgt_ids[:, 1:].reshape(-1),
                                             non_padding_mask_tgt[:, 1:].reshape(-1))
                loss_de.norm_term = 1.0 * torch.sum(non_padding_mask_tgt[:, 1:])
                loss_en.eval_batch_with_mask(logps_en.reshape(-1, logps_en.size(-1)), src_ids[:, 1:].reshape(-1),
                                             non_padding_mask_src[:, 1:].reshape(-1))
                loss_en.norm_term = 1.0 * torch.sum(non_padding_mask_src[:, 1:])
            if self.normalise_loss:
                loss_de.normalise()
                loss_en.normalise()
            loss_de.acc_loss = loss_de.acc_loss * self.loss_coeff['nll_st'] / n_minibatch
            resloss_de = resloss_de + loss_de.acc_loss.detach()
            loss_en.acc_loss = loss_en.acc_loss * self.loss_coeff['nll_asr'] / n_minibatch
            resloss_en = resloss_en + loss_en.acc_loss.detach()
            loss_en.add(loss_de)                                             # trainer_asr_st.py:345-346
            loss_en.backward()
        rt.join_deferred()
        if self.reducer is not None:
            self.reducer.finish()
        return resloss_de, resloss_en
