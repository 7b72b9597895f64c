"""Learning-rate schedules: host-side scalar arithmetic on ``optimizer.lr`` (reference optim/lr_scheduler.py:16-160; SURVEY.md §2
row 17 marks it out of the tensor hot path — kept so that training scripts run unchanged).

Results follow the reference, quirks included (pinned by tests/golden/lr_schedulers.json, generated from the unmodified reference):
* every schedule is CHAINED on the optimizer's current lr, and ``ExponentialLR`` / ``StepLR`` multiply it by ``gamma**last_epoch`` /
  ``gamma**(last_epoch // step_size)`` at EVERY step (lr_scheduler.py:101, 117-118) — so ExponentialLR decays like
  ``gamma**(t(t+1)/2)``, not ``gamma**t``;
* ``get_last_lr()`` returns the lr that was in force BEFORE the latest ``step()`` (``_last_lr`` is taken before the update, :80-81);
* the constructor records ``optimizer.initial_lr``, counts ``optimizer.step()`` calls in ``optimizer._step_count`` and performs the
  first ``step()`` itself (:19-24, :56-60)."""
import math
from collections import Counter
from functools import wraps


class _LRScheduler:

    def __init__(self, optimizer, last_epoch: int = -1) -> None:
        self.optimizer = optimizer
        self.last_epoch = last_epoch
        if last_epoch == -1:
            optimizer.initial_lr = optimizer.lr
        else:
            assert hasattr(optimizer, "initial_lr"), "last_epoch=1 but no 'initial_lr' attribute in optimizer!"
        if not getattr(optimizer.step, "_with_counter", False):
            inner = optimizer.step

            @wraps(inner)
            def counted(*args, **kwargs):
                optimizer._step_count += 1
                return inner(*args, **kwargs)

            counted._with_counter = True
            optimizer.step = counted
        optimizer._step_count = 0
        self._step_count = 0
        self.step()

    def get_lr(self) -> float:
        raise NotImplementedError

    def step(self):
        self._step_count += 1
        self.last_epoch += 1
        lr = self.get_lr()
        self._last_lr = self.optimizer.lr  # the value BEFORE this update (reference :80)
        self.optimizer.lr = lr

    def get_last_lr(self):
        return self._last_lr


class ExponentialLR(_LRScheduler):

    def __init__(self, optimizer, gamma: float = 0.1, last_epoch: int = -1) -> None:
        self.gamma = gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.optimizer.lr * self.gamma**self.last_epoch


class StepLR(_LRScheduler):

    def __init__(self, optimizer, step_size: int, gamma=0.1, last_epoch: int = -1) -> None:
        self.step_size, self.gamma = step_size, gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.optimizer.lr * self.gamma**(self.last_epoch // self.step_size)


class MultiStepLR(_LRScheduler):

    def __init__(self, optimizer, milestones, gamma=0.1, last_epoch: int = -1) -> None:
        self.milestones, self.gamma = Counter(milestones), gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        if self.last_epoch not in self.milestones:
            return self.optimizer.lr
        return self.optimizer.lr * self.gamma**self.milestones[self.last_epoch]


class CosineAnnealingLR(_LRScheduler):

    def __init__(self, optimizer, T_max: int, eta_min: float = 0, last_epoch: int = -1) -> None:
        self.T_max, self.eta_min = T_max, eta_min
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        base, t, T = self.optimizer.initial_lr, self.last_epoch, self.T_max
        if t == 0:
            return base
        if (t - 1 - T) % (2 * T) == 0:
            return self.get_last_lr() + (base - self.eta_min) * (1 - math.cos(math.pi / T)) / 2
        return (1 + math.cos(math.pi * t / T)) / (1 + math.cos(math.pi * (t - 1) / T)) * (self.get_last_lr() - self.eta_min) + self.eta_min
