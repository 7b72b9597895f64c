"""Host side of the CUDA path, checked WITHOUT a GPU through a stubbed C ABI (tests/stub_abi.py): steady-state training steps
must not contain synchronising host<->device copies (every `pdn_memcpy_h2d` / `_d2h` drains the stream — a regression here is how a
launch-bound step loses its host run-ahead), must issue the same calls every step, and stay within a launch budget."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SYNCING = ("pdn_memcpy_h2d", "pdn_memcpy_d2h", "pdn_sync")


def _measure(name):
    r = subprocess.run([sys.executable, os.path.join(HERE, "stub_abi.py"), name], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("name,budget", [("lenet", 100), ("encoder", 260), ("matmul", 30)])
def test_steady_state_step_has_no_synchronising_copies(name, budget):
    steps = _measure(name)
    assert steps[0] == steps[1] == steps[2], "a training step must issue the same C-ABI calls every time"
    for s in steps:
        assert not any(s.get(k, 0) for k in SYNCING), {k: s.get(k, 0) for k in SYNCING}
        launches = sum(v for k, v in s.items() if k not in ("pdn_malloc", "pdn_free", "pdn_get_device", "pdn_set_device"))
        assert launches <= budget, (launches, s)
    if name == "matmul":  # x @ w forward + two gradient products: three GEMMs through the plane-caching entry point
        assert steps[0].get("pdn_gemm_cached", 0) == 3 and steps[0].get("pdn_gemm", 0) == 0


def test_graphed_step_replays_one_graph_launch_per_step():
    """pydynet_b200.cuda.graphed_step: after 2 eager calls and the recording, a training step is the two input copies + ONE graph
    launch — no kernel launch, allocation or synchronising copy issued from Python — and the optimizer's host step counter keeps
    counting (it starts at 1; 4 warm + 3 measured steps have run)."""
    *steps, extra = _measure("lenet_graphed")
    assert steps[0] == steps[1] == steps[2]
    s = steps[0]
    assert s.get("pdn_graph_launch") == 1, s
    assert not any(s.get(k, 0) for k in SYNCING), s
    others = {k: v for k, v in s.items() if k not in ("pdn_graph_launch", "pdn_get_device", "pdn_set_device")}
    assert sum(others.values()) <= 4, others  # copies of X and y into the recorded input buffers
    assert extra["adam_t"] == 1 + 7
