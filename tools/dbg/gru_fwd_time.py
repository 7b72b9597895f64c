import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn, pydynet_b200.nn as nn
dev = "cuda:0"
np.random.seed(0)
T = int(os.environ.get("T", 1024)); B = int(os.environ.get("B", 256))
rnn = nn.GRU(512, 512, 1, batch_first=True, dtype=np.float32).to(dev)
X = pdn.Tensor(np.random.randn(B, T, 512).astype(np.float32), dtype=np.float32, device=dev)
with pdn.no_grad():
    for _ in range(2): rnn(X, None)
    pdn.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): rnn(X, None)
    pdn.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
print(f"GRU forward T={T} B={B}: {dt*1e3:.2f} ms = {dt/T*1e6:.2f} us per time step (PDN_GRU_PERSIST={os.environ.get('PDN_GRU_PERSIST','1')})")
