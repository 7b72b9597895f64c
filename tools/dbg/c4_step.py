import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.optim import Adam
from baseline import refload
Transformer = refload.dropin_model("examples/pydynet/transformer.py", lines=(52, 192), extra=refload.dropin_extra())["Transformer"]
dev = "cuda:0"
np.random.seed(0)
B, S, V = 128, 512, 8192
net = Transformer(512, 1, 8, 3, 0.05, V, S); net.word_embedding.reset_parameters(); net.to(dev)
opt = Adam(net.parameters(), lr=5e-4)
X = pdn.Tensor(np.random.randint(1, V, (B, S)), device=dev); y = pdn.Tensor(np.random.choice([-1, 1], B).astype(np.float32), device=dev)
net.train()
for _ in range(3):
    loss = pdn.log(1 + pdn.exp(-y * pdn.squeeze(net(X, None)))).mean()
    opt.zero_grad(); loss.backward(); opt.step()
pdn.cuda.synchronize()
