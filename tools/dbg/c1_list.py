import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
rng = np.random.default_rng(0)
A, B = (rng.standard_normal((512, 512)).astype(np.float32) for _ in range(2))
x = pdn.Tensor(A, dtype=np.float32, device="cuda:0", requires_grad=True); w = pdn.Tensor(B, dtype=np.float32, device="cuda:0", requires_grad=True)
def step():
    x.zero_grad(); w.zero_grad()
    pdn.matmul(x, w).sum().backward()
for _ in range(3): step()
pdn.cuda.synchronize()
step(); pdn.cuda.synchronize()
