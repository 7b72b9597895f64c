import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from oracle import pdn_oracle as O
from baseline import refload
V, D, H, FF, S, L = 32000, 288, 6, 768, 1024, 6
B = int(os.environ.get("B", 1)); TOTAL = int(os.environ.get("TOTAL", 40))
Llama = refload.dropin_model("llm/llama/model.py")["Llama"]
params = O.synthetic_llama_params(V, D, H, FF, L, seed=0, std=0.05)
net = Llama(V, D, H, FF, S, B, L, np.float32).to("cuda:0")
for name, p in net._parameters.items():
    if name in params:
        with p.device: p.data[...] = params[name]
net.eval()
prompt = np.random.default_rng(100).integers(1, V, (B, 4))
with pdn.no_grad():
    pd = pdn.Tensor(prompt, device="cuda:0")
    for _ in range(2): toks = [t for t in net.generate(pd, TOTAL)]
    pdn.cuda.synchronize()
