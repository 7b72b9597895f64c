import sys, os, time, cProfile, pstats
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import stub_abi
stub_abi.install()
stub_abi.trace = type("NullList", (list,), {"append": lambda self, x: None})()
import numpy as np
import pydynet_b200 as pdn
from pydynet_b200.optim import Adam
which = sys.argv[1]
dev = "cuda:0"
if which == "matmul":
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((512, 512)).astype(np.float32), rng.standard_normal((512, 512)).astype(np.float32)
    x, w = pdn.Tensor(A, dtype=A.dtype, device=dev, requires_grad=True), pdn.Tensor(B, dtype=B.dtype, device=dev, requires_grad=True)
    def step():
        x.data.buf.version += 1; w.data.buf.version += 1
        x.zero_grad(); w.zero_grad()
        pdn.matmul(x, w).sum().backward()
else:
    from workloads.lenet import ConvNet, train_step
    np.random.seed(42)
    net = ConvNet().to(dev)
    opt = Adam(net.parameters(), lr=1e-4)
    X = pdn.Tensor(np.random.rand(256, 1, 28, 28).astype(np.float32), dtype=np.float32, device=dev)
    y = pdn.Tensor(np.random.randint(0, 10, 256), device=dev)
    net.train()
    def step():
        train_step(net, opt, X, y)
for _ in range(20): step()
t0 = time.perf_counter()
N = 300
for _ in range(N): step()
print(f"{which}: host {1e6 * (time.perf_counter() - t0) / N:.1f} us/step (stubbed C ABI)")
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
