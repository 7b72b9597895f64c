"""Multi-GPU data-parallel check (run under torchrun on N GPUs, or imported by bench.py): N ranks training the Transformer
encoder on shards of a global batch must reach what a single GPU reaches on the whole batch — NCCL all-reduce of the bucketed
gradient buffer overlapped with backward (pydynet_b200/distributed.py, csrc/comm.cu) + cross-rank batch statistics of the
batch-coupled norms.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

What is compared: the all-reduced, 1/W-scaled gradient of step 0 against the single-GPU gradient of the global
batch (every parameter whose gradient is not mathematically zero), the loss of every step (mean of the shard losses), and the
parameters after 3 Adam steps on the well-conditioned tensors (the matrices; Adam turns the rounding noise of near-zero gradient
entries into lr-sized steps, so biases in front of a batch-statistic norm are excluded, as for the single-GPU goldens). Token ids
are unique across the global batch: the reference's embedding backward is last-write-wins inside a process (tensor.py:937-940),
which has no cross-rank definition for repeated tokens.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

f32 = np.float32
SKIP = ("feed_forward.2.bias", "shift", "running_")  # zero-gradient directions / statistics


def _build(dev, V):
    import pydynet_b200 as pdn  # noqa: F401
    from workloads.encoder import Transformer
    np.random.seed(0)
    net = Transformer(128, 1, 4, 3, 0.05, V, 32)
    net.word_embedding.reset_parameters()
    return net.to(dev)


def _run(dev, X, y, steps, ddp, V):
    import pydynet_b200 as pdn
    from pydynet_b200 import distributed as dist
    from pydynet_b200.optim import Adam
    from workloads.encoder import logistic_loss
    net = _build(dev, V)
    opt = Adam(net.parameters(), lr=5e-4)
    wrap = dist.DataParallel(net, opt) if ddp else None
    Xs, ys = (dist.shard(X), dist.shard(y)) if ddp else (X, y)
    net.train()
    losses, g0 = [], None
    tX, ty = pdn.Tensor(Xs, device=dev), pdn.Tensor(ys, dtype=f32, device=dev)
    for s in range(steps):
        loss = logistic_loss(net(tX, None), ty)
        opt.zero_grad()
        loss.backward()
        if wrap:
            wrap.sync_gradients()
        if s == 0:
            scale = 1.0 / dist.get_world_size() if ddp else 1.0
            g0 = {k: np.asarray(p.grad.get()) * scale for k, p in net._parameters.items() if p.requires_grad}
        if wrap:
            opt.step()
            wrap._arm()
        else:
            opt.step()
        losses.append(float(loss.item()))
    log = list(wrap.launch_log) if wrap else []
    return {k: p.numpy() for k, p in net._parameters.items()}, g0, losses, log


def parity(dev, rank, world):
    """Returns a dict of relative errors (this rank's view). The NCCL process group must be up."""
    from pydynet_b200 import distributed as dist
    per = 16
    V = per * world * 32 + 1
    rng = np.random.default_rng(1)
    X = (rng.permutation(V - 1)[:per * world * 32] + 1).reshape(per * world, 32)
    y = rng.choice([-1, 1], per * world).astype(f32)
    was = dist.sync_stats_enabled()
    saved = dist._S.copy()
    dist._S.update(world=1, rank=0)  # single-GPU run on the GLOBAL batch, on every rank
    try:
        ref_p, ref_g, ref_losses, _ = _run(dev, X, y, 3, False, V)
    finally:
        dist._S.update(saved)
    dist.sync_batch_stats(True)
    p, g, losses, log = _run(dev, X, y, 3, True, V)
    dist.sync_batch_stats(was)
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    g_err = {k: rel(g[k], v) for k, v in ref_g.items() if not any(s in k for s in SKIP)}
    p_err = {k: rel(p[k], v) for k, v in ref_p.items() if not any(s in k for s in SKIP) and v.ndim >= 2}
    import torch
    import torch.distributed as td
    lt = torch.tensor(losses, dtype=torch.float64)
    if td.is_initialized() and world > 1:
        td.all_reduce(lt)
        lt /= world
    loss_err = float(np.max(np.abs(lt.numpy() - np.array(ref_losses)) / np.abs(ref_losses)))
    return {"world": world, "grad_rel_err_max": max(g_err.values()), "grad_worst": max(g_err, key=g_err.get),
            "param_rel_err_max_after_3_adam_steps": max(p_err.values()), "loss_rel_err_max": loss_err,
            "buckets_queued_during_backward": sum(1 for (_, sweeps) in log[:8] if sweeps == 0), "buckets": len({b for b, _ in log}),
            "tolerance": 1e-4, "ok": bool(max(g_err.values()) < 1e-4 and max(p_err.values()) < 1e-4 and loss_err < 1e-4)}


def main():
    import torch.distributed as td
    from pydynet_b200 import distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group("gloo", rank=rank, world_size=world)
    dist.init_process_group("nccl")
    res = parity(f"cuda:{local}", rank, world)
    if rank == 0:
        print(json.dumps(res), flush=True)
    td.barrier()
    dist.destroy_process_group()
    td.destroy_process_group()
    assert res["ok"], res


if __name__ == "__main__":
    main()
