import os, sys, time, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.backend import lib
from oracle import pdn_oracle as O
from baseline import refload
V, D, H, FF, S, L = 32000, 288, 6, 768, 1024, 6
B = int(os.environ.get("B", 1)); TOTAL = 256
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
    for _ in range(3): toks = [t for t in net.generate(pd, TOTAL)]
    pdn.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        toks = [t for t in net.generate(pd, TOTAL)]
        t1 = time.perf_counter()
        pdn.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"resident: host {1e6*(t1-t0)/TOTAL:.1f} us/tok, total {1e6*(t2-t0)/TOTAL:.1f} us/tok -> {B*TOTAL/(t2-t0):.0f} tok/s", flush=True)
    for rep in range(3):
        t0 = time.perf_counter()
        out = [t[0].numpy().tolist() for t in net.generate(prompt, TOTAL)]
        t2 = time.perf_counter()
        print(f"e2e (infer.py loop): {1e6*(t2-t0)/TOTAL:.1f} us/tok -> {B*TOTAL/(t2-t0):.0f} tok/s", flush=True)
    got = np.concatenate([t.numpy() for t in net.generate(pd, TOTAL)], axis=1)
    ref, mar = O.LlamaOracle(params, H, S, B, L).generate_with_margins(prompt, TOTAL)
    print("oracle check", O.check_greedy_tokens(got, ref, mar), "distinct", len(set(got.ravel().tolist())))
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    toks = [t for t in net.generate(pd, TOTAL)]
    pr.disable(); pdn.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
