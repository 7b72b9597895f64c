"""Checkpoint IO of the Llama demo (reference llm/llama/io.py:8-57).

``load_model`` reads a HuggingFace-named ``.npz`` (``model.layers.{i}.self_attn.q_proj.weight`` …, projection matrices stored
[out, in]) into a ``workloads.llama.Llama`` whose Linear weights are [in, out]; the fine-tune helpers write / read the
grad-requiring parameters under their ``_parameters`` names. Parameters may already live on a cuda device: the copy goes through
``param.data[...] =`` inside the parameter's device scope (one H2D per tensor, the cached GEMM operand planes of that weight are
invalidated by the write counter)."""
import numpy as np

import pydynet_b200 as pdn

# reference parameter name (relative to a layer) -> (checkpoint key template, stored transposed?)
_LAYER_KEYS = {
    "attention.Q.weight": ("self_attn.q_proj.weight", True),
    "attention.K.weight": ("self_attn.k_proj.weight", True),
    "attention.V.weight": ("self_attn.v_proj.weight", True),
    "attention.O.weight": ("self_attn.o_proj.weight", True),
    "ffn.up.weight": ("mlp.up_proj.weight", True),
    "ffn.gate.weight": ("mlp.gate_proj.weight", True),
    "ffn.down.weight": ("mlp.down_proj.weight", True),
    "input_norm.weight": ("input_layernorm.weight", False),
    "post_attn_norm.weight": ("post_attention_layernorm.weight", False),
}


def checkpoint_key_map(n_layers: int) -> dict:
    """``_parameters`` name -> (checkpoint key, transposed) for every tensor the reference loads (io.py:12-38). ``lm_head.bias``
    is NOT in the checkpoint: it keeps its random initial value."""
    keys = {"tok_embedding.weight": ("model.embed_tokens.weight", False), "lm_head.weight": ("lm_head.weight", True),
            "norm.weight": ("model.norm.weight", False)}
    for i in range(n_layers):
        for ours, (theirs, tr) in _LAYER_KEYS.items():
            keys[f"layers.{i}.{ours}"] = (f"model.layers.{i}.{theirs}", tr)
    return keys


def _assign(param, value):
    value = np.asarray(value)
    if tuple(value.shape) != tuple(param.shape):
        raise ValueError(f"could not broadcast input array from shape {value.shape} into shape {tuple(param.shape)}")
    with param.device:
        param.data[...] = value.astype(param.dtype, copy=False)


@pdn.no_grad()
def load_model(llama, model_path: str):
    weights = np.load(model_path)
    params = llama._parameters
    for name, (key, transposed) in checkpoint_key_map(llama.n_layers).items():
        w = weights[key]
        _assign(params[name], w.T if transposed else w)
    return llama


@pdn.no_grad()
def save_finetuned_parameters(model, output_path: str):
    np.savez(output_path, **{name: p.numpy() for name, p in model._parameters.items() if p.requires_grad})


@pdn.no_grad()
def load_finetuned_parameters(model, finetuned_path: str):
    weights = np.load(finetuned_path)
    for name, p in model._parameters.items():
        if name in weights:
            _assign(p, weights[name])
    return model
