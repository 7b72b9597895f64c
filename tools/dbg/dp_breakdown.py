"""Where the data-parallel overhead of the encoder step goes (torchrun --nproc-per-node N): the step of bench.py's dp_train with
(a) everything on, (b) per-rank batch statistics, (c) all-reduce after backward instead of overlapped, (d) no gradient exchange."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
import pydynet_b200 as pdn
from pydynet_b200 import distributed as pd
from pydynet_b200.backend import lib
from pydynet_b200.optim import Adam
from baseline import refload
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
Transformer = refload.dropin_model("examples/pydynet/transformer.py", lines=(52, 192), extra=refload.dropin_extra())["Transformer"]
device = f"cuda:{local}"
pd.init_process_group("nccl", rank, world)
C4 = bench.C4
E, S, V = C4["E"], C4["S"], C4["V"]


def run(tag, sync, overlap, exchange):
    pd.sync_batch_stats(sync)
    np.random.seed(0)
    net = Transformer(E, 1, C4["H"], C4["FFX"], 0.05, V, S)
    net.word_embedding.reset_parameters()
    net.to(device)
    opt = Adam(net.parameters(), lr=5e-4)
    ddp = pd.DataParallel(net, opt, buckets=4, overlap=overlap) if exchange else None
    rng = np.random.default_rng(10 + rank)
    X = pdn.Tensor(rng.integers(1, V, (C4["B"], S)), device=device)
    y = pdn.Tensor(rng.choice([-1, 1], C4["B"]).astype(np.float32), device=device)
    net.train()

    def step():
        loss = pdn.log(1 + pdn.exp(-y * pdn.squeeze(net(X, None)))).mean()
        opt.zero_grad()
        loss.backward()
        (ddp or opt).step()

    for _ in range(3):
        step()
    pdn.cuda.synchronize()
    ev0, ev1 = bench._event_pair(lib)
    lib.call("pdn_event_record", ev0)
    for _ in range(10):
        step()
    lib.call("pdn_event_record", ev1)
    pdn.cuda.synchronize()
    sec = bench._elapsed_s(lib, ev0, ev1) / 10
    if rank == 0:
        print(f"{tag:42s} {sec * 1e3:.3f} ms/step", flush=True)


run("(d) no gradient exchange, local stats", False, False, False)
run("(b) exchange overlapped, local stats", False, True, True)
run("(c) exchange after backward, global stats", True, False, True)
run("(a) exchange overlapped, global stats", True, True, True)
pd.sync_batch_stats(False)
pd.destroy_process_group()
