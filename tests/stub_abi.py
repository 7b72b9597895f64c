"""A stubbed C ABI for exercising the CUDA code path of the Python layer WITHOUT a GPU: `lib.call` is replaced by a recorder that
hands out deterministic fake device pointers, treats kernels as no-ops and zero-fills D2H copies. What it checks is everything that
happens on the host — shape / stride / dtype logic, which entry points are called with which arguments, how many synchronising
copies a step performs — not numerics (those are the -m gpu tests). Run as a script it prints the per-step call counts:
    python tests/stub_abi.py lenet|encoder|matmul
Must be installed before pydynet_b200 creates any device array (the test runs it in a subprocess)."""
import collections
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

trace = []
_next = [1 << 30]


def _norm(a):
    if isinstance(a, (int, float, str, bytes, type(None))):
        return a
    if isinstance(a, C.c_void_p):
        return ("vp", a.value)
    if isinstance(a, C.Array):
        return ("arr", tuple(a))
    if hasattr(a, "_obj"):
        return ("byref", )
    if hasattr(a, "value"):
        return ("c", a.value)
    return ("obj", type(a).__name__)


def install():
    from pydynet_b200.backend import lib as L

    def fake_call(name, *args):
        if name in ("pdn_malloc", "pdn_malloc_host"):
            size = int(args[1])
            args[0]._obj.value = _next[0]
            _next[0] += (size + 511) // 512 * 512
            trace.append((name, size))
            return
        if name == "pdn_get_device":
            args[0]._obj.value = 0
        elif name == "pdn_memcpy_d2h" and isinstance(args[0], int):
            C.memset(args[0], 0, int(args[2]))
        elif name in ("pdn_event_create", "pdn_gemm_prepack", "pdn_graph_end"):
            args[-1]._obj.value = 1234
        elif name == "pdn_decoder_create":
            args[0]._obj.value = 4321
        vals = [_norm(a) for a in args]
        if name == "pdn_memcpy_h2d":
            vals[1] = "host"
        if name == "pdn_memcpy_d2h":
            vals[0] = "host"
        trace.append((name, ) + tuple(vals))

    class FakeLib:

        def pdn_free(self, p):
            trace.append(("pdn_free", p))
            return 0

        def __getattr__(self, n):
            return lambda *a: 0

    L.call, L._device_count, L._lib = fake_call, 1, FakeLib()
    L.load = lambda: L._lib


def workload(name):
    import numpy as np
    import pydynet_b200 as pdn
    from pydynet_b200.optim import Adam
    dev = "cuda:0"
    if name == "lenet":
        from workloads.lenet import ConvNet, train_step
        np.random.seed(42)
        net = ConvNet().to(dev)
        opt = Adam(net.parameters(), lr=1e-4)
        X, y = pdn.Tensor(np.random.rand(64, 1, 28, 28).astype(np.float32), device=dev), pdn.Tensor(np.random.randint(0, 10, 64), device=dev)
        net.train()
        return lambda: train_step(net, opt, X, y)
    if name == "lenet_graphed":
        # the same step recorded into a CUDA graph by pydynet_b200.cuda.graphed_step (2 eager warm-up calls, 1 recording, replays)
        from workloads.lenet import ConvNet, train_step
        np.random.seed(42)
        net = ConvNet().to(dev)
        opt = Adam(net.parameters(), lr=1e-4)
        X, y = pdn.Tensor(np.random.rand(64, 1, 28, 28).astype(np.float32), device=dev), pdn.Tensor(np.random.randint(0, 10, 64), device=dev)
        net.train()
        gs = pdn.cuda.graphed_step(lambda a, b: train_step(net, opt, a, b), optimizers=[opt])
        step = lambda: gs(X, y)
        step.optimizer = opt
        return step
    if name == "encoder":
        from workloads.encoder import Transformer, train_step
        np.random.seed(0)
        net = Transformer(64, 1, 4, 3, 0.05, 100, 32)
        net.word_embedding.reset_parameters()
        net.to(dev)
        opt = Adam(net.parameters(), lr=5e-4)
        X, y = pdn.Tensor(np.random.randint(1, 100, (8, 32)), device=dev), pdn.Tensor(np.random.choice([-1, 1], 8).astype(np.float32), device=dev)
        net.train()
        return lambda: train_step(net, opt, X, y, None)
    if name == "matmul":
        A = pdn.Tensor(np.random.rand(96, 80).astype(np.float32), device=dev, requires_grad=True)
        B = pdn.Tensor(np.random.rand(80, 72).astype(np.float32), device=dev, requires_grad=True)

        def step():
            A.zero_grad()
            B.zero_grad()
            pdn.matmul(A, B).sum().backward()

        return step
    if name in ("llama", "llama_ref", "llama_ref_b1"):
        # one greedy decode step of a Llama-style decoder at batch 48 through Module.__call__ -> inference plan (nn/_plans.py);
        # llama_ref: the reference's OWN llm/llama/model.py exec'd unchanged on top of pydynet_b200
        B = 1 if name.endswith("_b1") else 48
        if name.startswith("llama_ref"):
            from baseline import refload
            Llama = refload.dropin_model("llm/llama/model.py")["Llama"]
        else:
            from workloads.llama import Llama
        net = Llama(512, 128, 4, 256, 64, B, 3, np.float32).to(dev)
        net.eval()
        gen = net.generate(np.random.randint(1, 512, (B, 4)), 64)
        pdn.autograd.set_grad_enabled(False)
        next(gen)  # prefill

        def step():
            return next(gen)

        return step
    raise SystemExit(f"unknown workload {name}")


def measure(name, warm=2, steps=3):
    install()
    step = workload(name)
    if name.endswith("_graphed"):
        warm = 4  # 2 eager calls + the recording + its first replay
    for _ in range(warm):
        step()
    per_step = []
    for _ in range(steps):
        trace.clear()
        step()
        per_step.append(collections.Counter(ev[0] for ev in trace))
    out = [dict(c) for c in per_step]
    if hasattr(step, "optimizer"):
        out.append({"adam_t": step.optimizer.t})
    return out


if __name__ == "__main__":
    print(json.dumps(measure(sys.argv[1] if len(sys.argv) > 1 else "lenet")))
