"""Generates tests/golden/llama_app/* by running the UNMODIFIED reference's llm/llama/{io,tokenizer,model,finetune}.py
(/root/reference) on a small synthetic checkpoint and tokenizer. Run in the build container only
(`python tests/golden/make_golden_llama_app.py`); the fixtures are committed and the tests read nothing else."""
import io as _io
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import pydynet as pdn  # noqa: E402  (the reference)
import pydynet.optim as optim  # noqa: E402
from llm.llama.io import load_finetuned_parameters, load_model, save_finetuned_parameters  # noqa: E402
from llm.llama.model import Llama  # noqa: E402
from llm.llama.tokenizer import Tokenizer  # noqa: E402
from llm.llama.finetune import build_causal_training_pair  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "llama_app")
os.makedirs(OUT, exist_ok=True)
V, D, H, FF, S, B, L = 96, 24, 2, 48, 64, 1, 2
rng = np.random.default_rng(2024)

# ---- tokenizer fixture: specials, single characters, a few merges (with a duplicated string and tied scores)
chars = list(" abcdefghijklmnopqrstuvwxyzABT.,!'")
merges = ["th", "he", "the", " t", " the", "re", "er", "as", "wa", "was", " a", "bo", "oy", "boy", "in", "ing", "s ", "ll", "he", "st", "or", "ies"]
tokens = ["<unk>", "<s>", "</s>"] + chars + merges
tokens += [f"<pad{i}>" for i in range(V - len(tokens))]
assert len(tokens) == V, len(tokens)
scores = [0.0, 0.0, 0.0] + [-1e3] * len(chars) + [float(-i) for i in range(len(merges))] + [-1e9] * (V - 3 - len(chars) - len(merges))
scores[tokens.index("re")] = scores[tokens.index("er")]  # a tie: the leftmost pair must win
with open(os.path.join(OUT, "tokenizer.model.np"), "w", encoding="utf-8") as f:
    json.dump({"tokens": tokens, "scores": scores}, f)
tok = Tokenizer(os.path.join(OUT, "tokenizer.model.np"))
texts = ["There was a boy", "the stories", "was there ?", "", "sss", "B0b's erer the", "</s>hello<s>", "Tall boys sing."]
cases = []
for t in texts:
    for bos, eos in ((True, False), (True, True), (False, False)):
        ids = tok.encode(t, add_bos=bos, add_eos=eos)
        cases.append({"text": t, "bos": bos, "eos": eos, "ids": ids, "decoded": tok.decode(ids)})
json.dump(cases, open(os.path.join(OUT, "tokenizer_cases.json"), "w"))

# ---- checkpoint fixture (HuggingFace names, projections stored [out, in])
ck = {"model.embed_tokens.weight": rng.standard_normal((V, D)) * 0.08, "lm_head.weight": rng.standard_normal((V, D)) * 0.08,
      "model.norm.weight": 1 + 0.1 * rng.standard_normal(D)}
for i in range(L):
    p = f"model.layers.{i}."
    for n in "qkvo":
        ck[p + f"self_attn.{n}_proj.weight"] = rng.standard_normal((D, D)) * 0.08
    ck[p + "mlp.up_proj.weight"] = rng.standard_normal((FF, D)) * 0.08
    ck[p + "mlp.gate_proj.weight"] = rng.standard_normal((FF, D)) * 0.08
    ck[p + "mlp.down_proj.weight"] = rng.standard_normal((D, FF)) * 0.08
    ck[p + "input_layernorm.weight"] = 1 + 0.1 * rng.standard_normal(D)
    ck[p + "post_attention_layernorm.weight"] = 1 + 0.1 * rng.standard_normal(D)
ck = {k: v.astype(np.float32) for k, v in ck.items()}
np.savez(os.path.join(OUT, "checkpoint.model.npz"), **ck)

np.random.seed(7)
model = load_model(Llama(V, D, H, FF, S, B, L, dtype=np.float32), os.path.join(OUT, "checkpoint.model.npz"))
gold = {"cfg": np.array([V, D, H, FF, S, B, L])}
gold.update({"p." + k: v.data.copy() for k, v in model._parameters.items()})  # includes the (random, unloaded) lm_head.bias
model.eval()
prompt = np.array([tok.encode("There was a boy")])
with pdn.no_grad():
    gold["gen.prompt"] = prompt
    gold["gen.tokens"] = np.concatenate([t.numpy() for t in model.generate(prompt, 40)], axis=1)
pdn.autograd.set_grad_enabled(True)

# ---- fine-tune: 3 Adam steps on lm_head, save, reload into a fresh model
model.train()
n_train, n_frozen = model.set_trainable_parameters(("lm_head", ))
gold["ft.counts"] = np.array([n_train, n_frozen])
opt = optim.Adam(model.parameters(), lr=1e-3)
x, y = build_causal_training_pair(tok, "the boy was there", S)
gold["ft.input_ids"], gold["ft.target_ids"] = x, y
gold["ft.losses"] = np.array([model.finetune_step(x, y, opt) for _ in range(3)])
save_finetuned_parameters(model, os.path.join(OUT, "finetuned_ref.npz"))
saved = np.load(os.path.join(OUT, "finetuned_ref.npz"))
gold["ft.saved_keys"] = np.array(sorted(saved.files))
for k in saved.files:
    gold["ft." + k] = saved[k]
np.savez_compressed(os.path.join(OUT, "llama_app.npz"), **gold)
print("wrote", sorted(os.listdir(OUT)))
