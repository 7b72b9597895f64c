"""Learning-rate schedules: host-side scalar arithmetic on ``optimizer.lr`` (surface of reference
optim/lr_scheduler.py:16-160; SURVEY.md §2 row 17 marks it out of the tensor hot path — kept for API completeness)."""
import math
from bisect import bisect_right


class _LRScheduler:

    def __init__(self, optimizer, last_epoch: int = -1) -> None:
        self.optimizer = optimizer
        self.initial_lr = optimizer.lr if last_epoch == -1 else getattr(optimizer, "initial_lr", optimizer.lr)
        self.last_epoch = last_epoch
        self.step()

    def get_lr(self) -> float:
        raise NotImplementedError

    def step(self):
        self.last_epoch += 1
        self.optimizer.lr = self.get_lr()
        return self.optimizer.lr


class ExponentialLR(_LRScheduler):

    def __init__(self, optimizer, gamma: float = 0.1, last_epoch: int = -1) -> None:
        self.gamma = gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.initial_lr * self.gamma**self.last_epoch


class StepLR(_LRScheduler):

    def __init__(self, optimizer, step_size: int, gamma: float = 0.1, last_epoch: int = -1) -> None:
        self.step_size, self.gamma = step_size, gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.initial_lr * self.gamma**(self.last_epoch // self.step_size)


class MultiStepLR(_LRScheduler):

    def __init__(self, optimizer, milestones, gamma: float = 0.1, last_epoch: int = -1) -> None:
        self.milestones, self.gamma = sorted(milestones), gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.initial_lr * self.gamma**bisect_right(self.milestones, self.last_epoch)


class CosineAnnealingLR(_LRScheduler):

    def __init__(self, optimizer, T_max: int, eta_min: float = 0., last_epoch: int = -1) -> None:
        self.T_max, self.eta_min = T_max, eta_min
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return self.eta_min + (self.initial_lr - self.eta_min) * (1 + math.cos(math.pi * self.last_epoch / self.T_max)) / 2
