import os, sys, ctypes as C, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.backend import lib
from pydynet_b200.optim import Adam
from workloads.lenet import ConvNet, train_step
dev = "cuda:0"
np.random.seed(42)
net = ConvNet().to(dev); opt = Adam(net.parameters(), lr=1e-3); net.train()
X = pdn.Tensor(np.random.rand(32, 1, 28, 28).astype(np.float32), dtype=np.float32, device=dev); y = pdn.Tensor(np.random.randint(0, 10, 32), device=dev)
for _ in range(2): train_step(net, opt, X, y)
opt._flat.sync_device_state(opt.t, opt.lr)
orig = lib.call
bad = [None]
def call(name, *a):
    orig(name, *a)
    if bad[0] is None and name not in ("pdn_graph_status", "pdn_graph_end"):
        st = C.c_int(0); orig("pdn_graph_status", C.byref(st))
        if st.value == 2:
            bad[0] = name; print("FIRST INVALIDATING CALL:", name, flush=True)
lib.call = call
g = pdn.cuda.Graph(); g.begin()
try:
    train_step(net, opt, X, y)
finally:
    try: g.end()
    except Exception as e: print("end:", str(e)[:200])
