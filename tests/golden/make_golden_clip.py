"""Generates tests/golden/clip.npz by running the UNMODIFIED reference's llm/clip/model.py (/root/reference, NumPy path) on a small
synthetic configuration: forward logits, the cross-entropy loss of one (image, texts, target) triple and its gradients w.r.t. the
text encoder."""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import pydynet as pdn  # noqa: E402
import pydynet.nn as nn  # noqa: E402
from llm.clip.model import CLIP  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
cfg = dict(image_dim=32, image_heads=4, image_mlp_dim=64, image_patch=4, image_layers=2, text_dim=24, text_heads=3, text_mlp_dim=48,
           text_layers=2, final_dim=16, vocab_size=50, vision_tokens=5, text_tokens=7)
np.random.seed(3)
net = CLIP(**cfg)
rng = np.random.default_rng(1)
for name, p in net._parameters.items():
    # the randn embeddings / patch kernel are tamed so that the logits are not saturated; the token table is UNINITIALISED memory in
    # the reference (Embedding's constructor only allocates it) and gets defined values here
    if name in ("class_embed", "v_pos_emb", "t_pos_emb", "image_encoder.kernel", "text_encoder.token_embed.weight"):
        p.data[...] = (rng.standard_normal(p.shape) * 0.1).astype(np.float32)
img = rng.standard_normal((1, 3, 8, 8)).astype(np.float32)  # ONE image (the reference concatenates a (1,1,D) class token: batch 1 only)
idx = rng.integers(1, 49, (4, 7))
idx[np.arange(4), [6, 3, 5, 2]] = 49  # the end-of-text token (highest id) marks the pooled position
g = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array(list(cfg.values())), "img": img, "idx": idx}
g.update({"p." + k: v.data.copy() for k, v in net._parameters.items()})
net.eval()
with pdn.no_grad():
    g["logits"] = net(pdn.Tensor(img), idx).numpy()
pdn.autograd.set_grad_enabled(True)
net.train()
g["counts"] = np.array(net.set_trainable_parameters(("text_encoder", )))
# NOTE: the reference's CLIP.finetune_step cannot run: Adam(model.parameters()).zero_grad() hits the grad-less running statistics
# of CLIPLayerNorm (optimizer.py:29 -> tensor.py:383, TypeError). The fixture pins forward + backward of the same loss instead.
targets = np.array([2])
g["targets"] = targets
logits = net(pdn.Tensor(img), idx)
loss = nn.CrossEntropyLoss()(logits.reshape(1, 4), pdn.Tensor(targets, dtype=np.int64))
g["loss"] = np.array(loss.item())
loss.backward()
g.update({"g." + k: np.array(v.grad, copy=True) for k, v in net._parameters.items() if v.requires_grad and v.grad is not None})  # (running statistics: flagged trainable by the prefix match, but they never get a grad buffer)
assert all(np.isfinite(v).all() for k, v in g.items() if v.dtype.kind == "f"), "non-finite values in the fixture"
np.savez_compressed(os.path.join(HERE, "clip.npz"), **g)
print("clip golden:", g["logits"], float(g["loss"]), len([k for k in g if k.startswith("g.")]), os.path.getsize(os.path.join(HERE, "clip.npz")))
