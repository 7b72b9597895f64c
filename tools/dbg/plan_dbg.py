import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.nn import _plans
from oracle import pdn_oracle as O
from workloads.llama import Llama
sys.path.insert(0, "tests")
from test_llama_fast_gpu import _build, _generate
cfg = (V, D, H, FF, S, L) = (512, 96, 4, 256, 64, 3)
for B, total in [(48, 40), (1, 40), (3, 24)]:
    params = O.synthetic_llama_params(V, D, H, FF, L, seed=3, std=0.08)
    prompt = np.random.default_rng(5).integers(1, V, (B, 4))
    ref_toks, margins = O.LlamaOracle(params, H, S, B, L).generate_with_margins(prompt, total)
    net = _build(Llama, B, cfg, params)
    def cmp(name, t):
        bad = (t != ref_toks)
        first = [int(np.argmax(r)) if r.any() else -1 for r in bad]
        print(B, name, "mismatch", bad.sum(), "first-bad-steps", sorted(set(first))[:8], flush=True)
    toks = _generate(net, prompt, total); cmp("graph", toks)
    toks2 = _generate(net, prompt, total); cmp("graph2", toks2)
    os.environ["PDN_DECODE_GRAPH"] = "0"
    cmp("nograph", _generate(net, prompt, total))
    os.environ.pop("PDN_DECODE_GRAPH")
    _plans.ENABLED = False
    cmp("eager", _generate(net, prompt, total))
    _plans.ENABLED = True
