"""Data-parallel host logic on CPU: world_size-2 gloo processes (cpu-device tensors) must reproduce the single-process
result on the concatenated batch — gradients and post-Adam parameters — including the batch-coupled "LayerNorm" of the
encoder, whose statistics are reduced across ranks (pydynet_b200/distributed.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT)
import pydynet_b200 as pdn, pydynet_b200.nn as nn, pydynet_b200.nn.functional as F
from pydynet_b200 import distributed as dist
from pydynet_b200.optim import Adam

def build():
    np.random.seed(7)
    return nn.Sequential(nn.Linear(6, 8, bias=False, dtype=np.float32), nn.LayerNorm(8, dtype=np.float32), nn.ReLU(), nn.Linear(8, 3, dtype=np.float32))

def run(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    rng = np.random.default_rng(0)
    X = rng.standard_normal((8, 6)).astype(np.float32); y = rng.integers(0, 3, 8)
    net = build()
    opt = Adam(net.parameters(), lr=1e-2)
    if world > 1:
        dist.init_process_group("gloo", rank, world)
        dist.sync_batch_stats(True)
        ddp = dist.DataParallel(net, opt)
        Xs, ys = dist.shard(X), dist.shard(y)
    else:
        ddp, Xs, ys = None, X, y
    losses = []
    for step in range(3):
        loss = F.cross_entropy_loss(net(pdn.Tensor(Xs, dtype=np.float32)), pdn.Tensor(ys))
        opt.zero_grad(); loss.backward()
        if step == 0:
            g0 = {k: np.array(p.grad) for k, p in net._parameters.items() if p.requires_grad}
        (ddp.step() if ddp else opt.step())
        losses.append(float(loss.item()))
    if world > 1:
        # gradients after sync are the global-batch gradients
        import torch.distributed as td
        td.barrier()
    res = {"p." + k: p.data for k, p in net._parameters.items()}
    res["losses"] = np.array(losses)
    np.savez(out, **res)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()

if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
'''


def test_two_rank_gloo_equals_single_process(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    port = 29500 + (os.getpid() % 500)
    single = tmp_path / "single.npz"
    subprocess.run([sys.executable, str(script), "0", "1", str(port), str(single)], check=True, timeout=120)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(port), str(tmp_path / f"r{r}.npz")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    ref = np.load(single)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ref.files:
        if k == "losses":
            # the global loss is the mean of the two shard losses
            np.testing.assert_allclose((r0[k] + r1[k]) / 2, ref[k], rtol=2e-5)
            continue
        np.testing.assert_allclose(r0[k], r1[k], rtol=0, atol=0, err_msg=f"replicas diverged: {k}")
        # (the Linear in front of the batch-statistic norm has no bias on purpose: such a bias has a mathematically ZERO
        # gradient — the norm subtracts the per-feature batch mean — so Adam would amplify pure rounding noise)
        np.testing.assert_allclose(r0[k], ref[k], rtol=2e-4, atol=2e-6, err_msg=k)


def test_shard_and_world_defaults():
    from pydynet_b200 import distributed as dist
    assert dist.get_world_size() == 1 and dist.get_rank() == 0 and not dist.is_initialized()
    a = np.arange(12).reshape(4, 3)
    np.testing.assert_array_equal(dist.shard(a), a)
    assert not dist.sync_stats_enabled()


def test_backward_reports_each_pinned_leaf_after_its_last_consumer():
    """The overlap hook of the tape (core/tensor.py, _LEAF_READY_HOOK): every leaf whose gradient lives in a flat bucket is reported
    exactly once, only after the LAST node that consumes it has run (its gradient is final), newest consumers first, and the end of
    the sweep is signalled — the contract DataParallel(overlap=True) launches its bucket all-reduces on."""
    import pydynet_b200 as pdn
    from pydynet_b200.core import tensor as T
    rng = np.random.default_rng(0)
    mk = lambda *s: pdn.Tensor(rng.standard_normal(s), dtype=np.float64, requires_grad=True)
    w1, w2, w3, x = mk(4, 4), mk(4, 4), mk(4, 4), mk(2, 4)
    for p in (w1, w2, w3):  # what FlatAdamState does: gradient storage pinned to a preallocated view
        p._grad = np.zeros(p.shape)
        p._pinned_grad = True
        p._grad_stale = True
    events, snap = [], {}

    def hook(leaf):
        events.append(None if leaf is None else id(leaf))
        if leaf is not None:
            snap[id(leaf)] = np.array(leaf._grad)

    h1 = x @ w1
    h2 = (h1 @ w2) * 2.0 + h1 @ w1  # w1 is consumed twice: by the first and by a late node
    loss = ((h2 @ w3) ** 2).sum()
    T._LEAF_READY_HOOK[0] = hook
    try:
        loss.backward()
    finally:
        T._LEAF_READY_HOOK[0] = None
    assert events == [id(w3), id(w2), id(w1), None]  # w1 only after its EARLIEST consumer (last in the reverse sweep)
    for p in (w1, w2, w3):
        np.testing.assert_array_equal(snap[id(p)], p.grad)  # the gradient was final when it was reported
    assert not x._pinned_grad and id(x) not in snap


def test_bucket_ranges_cover_the_flat_buffer_in_order():
    from pydynet_b200.distributed import DataParallel

    class P:
        def __init__(self, n): self.size = n

    class Flat:
        params = [P(n) for n in (1000, 64, 5000, 64, 3000, 10, 2000)]
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.size + 63) // 64 * 64

    dp = DataParallel.__new__(DataParallel)
    dp._flat = Flat
    dp._make_buckets(3)
    assert 1 < len(dp._ranges) <= 3
    pos, members = 0, []
    for off, n, mem in dp._ranges:
        assert off == pos and n > 0
        pos += n
        members += mem
    assert pos == Flat.total and members == list(range(len(Flat.params)))
    assert set(dp._owner.values()) == set(range(len(dp._ranges)))


def test_nccl_id_rendezvous_over_tcp_leaves_nothing_behind(tmp_path, monkeypatch):
    """Without torch.distributed the 128-byte id travels over a one-shot TCP socket (no predictable /tmp file that a second job or
    another user could read or plant)."""
    import threading
    from pydynet_b200 import distributed as dist
    port = 31000 + (os.getpid() % 2000)
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("PDN_ID_PORT", str(port))
    payload = bytes(range(128))
    got = {}
    th = [threading.Thread(target=lambda r=r: got.__setitem__(r, dist._exchange_id(r, 3, lambda: payload))) for r in range(3)]
    for t in th[1:]:
        t.start()
    th[0].start()
    for t in th:
        t.join(timeout=60)
    assert got == {0: payload, 1: payload, 2: payload}
    assert not [f for f in os.listdir("/tmp") if f.startswith("pdn_nccl_id_")]
