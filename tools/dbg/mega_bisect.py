import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.nn import _plans
from oracle import pdn_oracle as O
from workloads.llama import Llama
_plans.VERIFY = False
def run(V, D, H, FF, S, L, B=1, plen=1):
    params = O.synthetic_llama_params(V, D, H, FF, L, seed=0, std=0.05)
    net = Llama(V, D, H, FF, S, B, L, np.float32).to("cuda:0")
    for name, p in net._parameters.items():
        if name in params:
            with p.device: p.data[...] = params[name]
    net.eval()
    prompt = np.random.default_rng(1).integers(1, V, (B, plen))
    ref = O.LlamaOracle(params, H, S, B, L).step(prompt, 0)
    with pdn.no_grad():
        got = net(pdn.Tensor(prompt, device="cuda:0"), 0).numpy()
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"V{V} D{D} H{H} FF{FF} S{S} L{L} B{B} plen{plen}: rel err {err:.3e}  argmax {got.argmax(-1).ravel()[:3]} vs {ref.argmax(-1).ravel()[:3]}", flush=True)
run(512, 96, 2, 256, 64, 1)
run(512, 288, 6, 768, 64, 1)
run(512, 288, 6, 256, 64, 1)
run(512, 96, 2, 768, 64, 1)
run(32000, 96, 2, 256, 64, 1)
run(32000, 288, 6, 768, 1024, 1)
run(32000, 288, 6, 768, 1024, 6)
run(32000, 288, 6, 768, 1024, 6, plen=4)
