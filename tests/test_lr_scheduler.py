"""Learning-rate schedules against fixtures produced by the unmodified reference (tests/golden/make_golden_sched.py): the lr held by
the optimizer and get_last_lr() after every step, the recorded initial_lr and the step counters — including the reference's
compounding ExponentialLR / StepLR (optim/lr_scheduler.py:101, 117-118)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_schedulers.json")


def test_schedules_match_reference():
    import pydynet_b200 as pdn
    from pydynet_b200.optim import SGD, lr_scheduler
    cases = json.load(open(GOLD))
    assert len(cases) == 8
    for c in cases:
        w = pdn.Tensor(np.ones(3), dtype=np.float32, requires_grad=True)
        opt = SGD([w], lr=0.2)
        sch = getattr(lr_scheduler, c["name"])(opt, **c["kwargs"])
        lrs, last = [opt.lr], [sch.get_last_lr()]
        for _ in range(14):
            (w * w).sum().backward()
            opt.step()
            opt.zero_grad()
            sch.step()
            lrs.append(opt.lr)
            last.append(sch.get_last_lr())
        np.testing.assert_allclose(lrs, c["lr"], rtol=1e-12, atol=0, err_msg=str(c["name"]))
        np.testing.assert_allclose(last, c["last_lr"], rtol=1e-12, atol=0, err_msg=str(c["name"]))
        assert opt.initial_lr == c["initial_lr"] and opt._step_count == c["opt_step_count"]
        assert sch._step_count == c["sched_step_count"] and sch.last_epoch == c["last_epoch"]
