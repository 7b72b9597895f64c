"""Multi-GPU data-parallel check (run under torchrun on N GPUs): N ranks training the Transformer encoder / LeNet on
shards of a global batch must end with the parameters a single GPU reaches on the whole batch (NCCL gradient-bucket
all-reduce + cross-rank batch statistics), and reports the step time.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydynet_b200 as pdn  # noqa: E402
import pydynet_b200.nn.functional as F  # noqa: E402
from pydynet_b200 import distributed as dist  # noqa: E402
from pydynet_b200.backend import lib  # noqa: E402
from pydynet_b200.optim import Adam  # noqa: E402
from workloads.encoder import Transformer, logistic_loss  # noqa: E402

f32 = np.float32


def build(dev):
    np.random.seed(0)
    net = Transformer(128, 1, 4, 3, 0.05, 500, 32)
    net.word_embedding.reset_parameters()
    net.layers[0].feed_forward.module_list[2].bias.requires_grad = True
    return net.to(dev)


def run(dev, X, y, steps, ddp):
    net = build(dev)
    opt = Adam(net.parameters(), lr=5e-4)
    wrap = dist.DataParallel(net, opt) if ddp else None
    Xs, ys = (dist.shard(X), dist.shard(y)) if ddp else (X, y)
    net.train()
    losses = []
    tX, ty = pdn.Tensor(Xs, device=dev), pdn.Tensor(ys, dtype=f32, device=dev)
    for _ in range(steps):
        loss = logistic_loss(net(tX, None), ty)
        opt.zero_grad()
        loss.backward()
        (wrap.step() if wrap else opt.step())
        losses.append(float(loss.item()))
    return {k: p.numpy() for k, p in net._parameters.items()}, losses


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = f"cuda:{local}"
    rng = np.random.default_rng(1)
    X = rng.integers(1, 500, (16 * world, 32))
    y = rng.choice([-1, 1], 16 * world).astype(f32)
    ref_params, ref_losses = run(dev, X, y, 3, ddp=False)  # every rank: single-GPU result on the GLOBAL batch
    dist.init_process_group("nccl")
    dist.sync_batch_stats(True)
    params, losses = run(dev, X, y, 3, ddp=True)
    worst = 0.0
    for k, v in ref_params.items():
        if "feed_forward.2.bias" in k or "shift" in k:  # zero-gradient directions amplified by Adam
            continue
        err = np.linalg.norm(params[k] - v) / max(np.linalg.norm(v), 1e-30)
        worst = max(worst, err)
    # timing of the DP step at BASELINE config 4 size per GPU (weak scaling)
    net = Transformer(512, 1, 8, 3, 0.05, 8192, 512)
    np.random.seed(0)
    net.word_embedding.reset_parameters()
    net.to(dev)
    opt = Adam(net.parameters(), lr=5e-4)
    wrap = dist.DataParallel(net, opt)
    B = int(os.environ.get("PDN_DP_BATCH", 32))
    tX = pdn.Tensor(rng.integers(1, 8192, (B, 512)), device=dev)
    ty = pdn.Tensor(rng.choice([-1, 1], B).astype(f32), device=dev)
    net.train()

    def step():
        loss = logistic_loss(net(tX, None), ty)
        opt.zero_grad()
        loss.backward()
        wrap.step()

    for _ in range(3):
        step()
    pdn.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    pdn.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    res = {"rank": rank, "world": world, "dp_vs_single_gpu_param_rel_err_max": float(worst), "losses_dp_rank": losses, "losses_single": ref_losses,
           "c4_dp_step_ms_per_gpu_batch%d" % B: dt * 1e3, "tokens_per_s_all_ranks": world * B * 512 / dt}
    print(json.dumps(res), flush=True)
    assert worst < 2e-3, worst
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
