import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.optim import Adam
from workloads.gru import GRURegressor, train_step
dev = "cuda:0"
np.random.seed(0)
T = int(os.environ.get("T", 128)); B = int(os.environ.get("B", 256))
net = GRURegressor(512, 512).to(dev)
opt = Adam(net.parameters(), lr=0.01)
X = pdn.Tensor(np.random.randn(B, T, 512).astype(np.float32), dtype=np.float32, device=dev)
Y = pdn.Tensor(np.random.randn(B, 1).astype(np.float32), dtype=np.float32, device=dev)
for _ in range(2):
    train_step(net, opt, X, Y)
pdn.cuda.synchronize()
