"""Time forward + backward of nn.RNN(512, 512) over T steps at batch B on cuda:0 (PDN_GRU_PERSIST=0 for the per-step host loop)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
import pydynet_b200.nn as nn
from pydynet_b200.backend import lib
T = int(os.environ.get("T", 256)); B = int(os.environ.get("B", 32)); H = int(os.environ.get("H", 512))
np.random.seed(0)
net = nn.RNN(H, H, dtype=np.float32).to("cuda:0")
X = pdn.Tensor(np.random.randn(T, B, H).astype(np.float32), dtype=np.float32, device="cuda:0")


def step():
    out, hn = net(X)
    loss = (out * out).sum()
    for p in net.parameters():
        p.zero_grad()
    loss.backward()


for _ in range(3):
    step()
pdn.cuda.synchronize()
lib.reset_launch_count()
t0 = time.perf_counter()
for _ in range(5):
    step()
pdn.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"RNN H{H} T{T} B{B} persist={os.environ.get('PDN_GRU_PERSIST', '1')}: {dt * 1e3:.2f} ms / fwd+bwd ({dt / T * 1e6:.1f} us per time step), {lib.launch_count() // 5} launches")
