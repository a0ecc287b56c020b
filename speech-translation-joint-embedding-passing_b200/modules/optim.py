"""`Optimizer` — mirror of the reference wrapper (modules/optim.py:6-56): same constructor, `set_scheduler`,
`step`, `update`, and the `.optimizer` / `.max_grad_norm` attributes the trainers touch
(trainer_base.py:197-200,422-426; trainer_st.py:325-328).  `step()` = gradient-norm clip + Adam, executed by the
fused multi-tensor kernels of b200st.optim instead of ~270 x several PyTorch launches; `update()` is the
reference's scheduler hook unchanged.  `lr_scheduler()` restates the warm-up / inverse-square-root rule
(trainer_base.py:135-154)."""
import torch

from b200st.optim import FusedClipAdam


class Optimizer(object):

    _ARG_MAX_GRAD_NORM = 'max_grad_norm'
    _fused = None          # class default: instances unpickled from a checkpoint the REFERENCE wrote have no such attribute

    def __init__(self, optim, max_grad_norm=0):
        self.optimizer = optim
        self.scheduler = None
        self.max_grad_norm = max_grad_norm
        self._fused = None

    def set_scheduler(self, scheduler):
        self.scheduler = scheduler

    def _engine(self):
        # trainers swap `.optimizer` for a fresh instance on resume (trainer_base.py:197-200): follow it
        if self._fused is None or self._fused.adam is not self.optimizer:
            self._fused = FusedClipAdam(self.optimizer, self.max_grad_norm)
        self._fused.max_grad_norm = float(self.max_grad_norm or 0.0)
        return self._fused

    def step(self):
        """clip_grad_norm_(all params, max_grad_norm) if max_grad_norm > 0, then Adam (modules/optim.py:31-36)."""
        self._engine().step()

    def set_lr(self, lr):
        for group in self.optimizer.param_groups:
            group['lr'] = lr
        self._engine().set_lr(lr)

    def update(self, loss, epoch):
        """Scheduler hook of the reference (modules/optim.py:38-52): plateau schedulers are fed the loss, any other
        scheduler is simply stepped; no scheduler, no effect.  `epoch` is accepted for signature compatibility."""
        sched = self.scheduler
        if sched is not None:
            args = (loss,) if isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau) else ()
            sched.step(*args)


def lr_at(step, init_lr=0.00001, peak_lr=0.0005, warmup_steps=16000):
    """The reference's learning-rate rule (trainer_base.py:145-149): linear warm-up from init_lr to peak_lr over
    warmup_steps, then peak_lr * step^-0.5 * warmup_steps^0.5 (same floating-point expression order as upstream)."""
    if step <= warmup_steps:
        return step * 1. * (peak_lr - init_lr) / warmup_steps + init_lr
    return peak_lr * (step ** (-0.5)) * (warmup_steps ** 0.5)


def lr_scheduler(optimizer, step, init_lr=0.00001, peak_lr=0.0005, warmup_steps=16000):
    """`Trainer.lr_scheduler` (trainer_base.py:135-154): writes lr_at(step) into every parameter group of the wrapped
    torch optimizer and returns it; warmup_steps <= 0 switches the rule off."""
    if warmup_steps > 0:
        lr = lr_at(step, init_lr, peak_lr, warmup_steps)
        for group in optimizer.param_groups:
            group['lr'] = lr
    return optimizer
