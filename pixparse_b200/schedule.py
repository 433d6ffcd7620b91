"""Cosine learning-rate schedule with linear warm-up, stepped per optimizer update.

Behavioural mirror of what the reference obtains from ``timm.scheduler.create_scheduler_v2(optimizer, 'cosine',
warmup_lr=..., warmup_epochs=num_warmup_intervals, num_epochs=num_intervals, step_on_epochs=False,
updates_per_epoch=...)`` (/root/reference/src/pixparse/task/task_cruller_pretrain.py:215-224, :294): the cosine is
NOT shifted by the warm-up length, lr_min = 0, one cycle, and a group's ``lr_scale`` multiplies its lr.
"""
import math


class CosineSchedule:
    def __init__(self, optimizer, total_updates, warmup_updates=0, warmup_lr=0.0, lr_min=0.0):
        self.optimizer = optimizer
        self.total_updates = max(1, int(total_updates))
        self.warmup_updates = int(warmup_updates)
        self.warmup_lr = float(warmup_lr)
        self.lr_min = float(lr_min)
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self.base_lrs = [g['initial_lr'] for g in optimizer.param_groups]
        self._apply([self.warmup_lr] * len(self.base_lrs) if self.warmup_updates else self.base_lrs)

    def lr_at(self, t):
        out = []
        for base in self.base_lrs:
            if t < self.warmup_updates:
                out.append(self.warmup_lr + t * (base - self.warmup_lr) / self.warmup_updates)
            elif t < self.total_updates:
                out.append(self.lr_min + 0.5 * (base - self.lr_min) * (1.0 + math.cos(math.pi * t / self.total_updates)))
            else:
                out.append(self.lr_min)
        return out

    def _apply(self, lrs):
        for g, lr in zip(self.optimizer.param_groups, lrs):
            g['lr'] = lr * g['lr_scale'] if 'lr_scale' in g else lr

    def step_update(self, num_updates, metric=None):
        self._apply(self.lr_at(num_updates))

    def step(self, epoch, metric=None):      # schedules are stepped on updates, not on intervals
        pass

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != 'optimizer'}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def create_scheduler(optimizer, sched='cosine', warmup_lr=0.0, warmup_intervals=0, num_intervals=1,
                     updates_per_interval=1):
    if sched != 'cosine':
        raise ValueError(f"scheduler {sched!r} is not on the Cruller hot path (only 'cosine', framework/config.py:9)")
    s = CosineSchedule(optimizer, total_updates=num_intervals * updates_per_interval,
                       warmup_updates=warmup_intervals * updates_per_interval, warmup_lr=warmup_lr)
    return s, num_intervals
