"""Generates tests/golden/lr_schedulers.json from the UNMODIFIED reference's optim/lr_scheduler.py: the lr the optimizer holds, and
get_last_lr(), after the constructor and after each of 14 further scheduler steps (an optimizer.step() before each, as in a
training loop), for every schedule."""
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import pydynet as pdn  # noqa: E402
from pydynet.optim import SGD  # noqa: E402
from pydynet.optim.lr_scheduler import CosineAnnealingLR, ExponentialLR, MultiStepLR, StepLR  # noqa: E402

SPECS = [("ExponentialLR", dict(gamma=0.9)), ("ExponentialLR", dict(gamma=0.5)), ("StepLR", dict(step_size=3, gamma=0.5)),
         ("StepLR", dict(step_size=1, gamma=0.9)), ("MultiStepLR", dict(milestones=[2, 5, 5, 9], gamma=0.1)),
         ("MultiStepLR", dict(milestones=[4], gamma=0.5)), ("CosineAnnealingLR", dict(T_max=5, eta_min=0.001)),
         ("CosineAnnealingLR", dict(T_max=4))]
CLS = dict(ExponentialLR=ExponentialLR, StepLR=StepLR, MultiStepLR=MultiStepLR, CosineAnnealingLR=CosineAnnealingLR)
out = []
for name, kw in SPECS:
    w = pdn.Tensor(np.ones(3), dtype=np.float32, requires_grad=True)
    opt = SGD([w], lr=0.2)
    sch = CLS[name](opt, **kw)
    lrs, last = [opt.lr], [sch.get_last_lr()]
    for _ in range(14):
        (w * w).sum().backward()
        opt.step()
        opt.zero_grad()
        sch.step()
        lrs.append(opt.lr)
        last.append(sch.get_last_lr())
    out.append({"name": name, "kwargs": kw, "lr": lrs, "last_lr": last, "initial_lr": opt.initial_lr, "opt_step_count": opt._step_count,
                "sched_step_count": sch._step_count, "last_epoch": sch.last_epoch})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lr_schedulers.json"), "w"))
print(len(out), "schedules;", out[0]["lr"][:4])
