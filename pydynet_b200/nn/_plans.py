"""Inference plans: fused execution of a whole eval-mode forward BEHIND the unchanged ``nn.Module`` surface.

The reference runs a model as a chain of ~490 eager array expressions per decoded token (SURVEY.md §3.4,
llm/llama/model.py:95-121, 142-150, 192-207, 254-256).  A model file written against the reference — its own
``llm/llama/model.py``, exec'd unchanged with ``import pydynet_b200 as pydynet`` — calls ``model(input_ids, start_pos)``;
``Module.__call__`` (nn/modules/module.py) asks this module for a plan before it falls through to ``forward``:

* ``DecoderPlan.match`` recognises the Llama-style decoder STRUCTURALLY (token embedding, RoPE tables, blocks of
  {RMSNorm, bias-free Q/K/V/O, KV-cache parameters, RMSNorm, gate/up/down SwiGLU}, final RMSNorm, lm_head) — attribute
  names and parameter shapes, nothing about the class identity;
* the FIRST call of each kind (prefill, decode) is verified against the model's own eager ``forward`` on the same
  inputs (normwise 1e-3); a mismatch retires the plan for good and the model keeps running eagerly, so a class that
  merely looks like the reference's but computes something else is never mis-served;
* afterwards a call costs: rows >= 32 (batched decode / prefill) — 9 launches per block on the tcgen05 GEMM path
  (RMSNorm -> operand planes, fused QKV GEMM, RoPE + cache append, cached attention -> planes, O-proj accumulating onto
  the residual, RMSNorm -> planes, fused gate|up GEMM, SwiGLU -> planes, down-proj accumulating onto the residual), the
  decode step recorded once per batch size as a CUDA graph with position and token ids in device memory and replayed per
  token; rows < 32 (the reference's own B = 1 loop) — ONE cooperative persistent kernel per token
  (csrc/decode_mega.cu).

Results are those of the eager chain (fp32; tested token-exact against the oracle).  A plan only ever serves
``not model._train and not grad-mode`` calls on a cuda device; everything else falls through.  Validity is tracked by two
epochs: module structure / parameter rebinding (``note_structure_change``) and user-level in-place writes to device
buffers (``backend.array.WRITE_EPOCH``); a change triggers a re-check of every weight pointer and write counter and
drops recorded graphs and packed weights that no longer match.

``logits[:, -1, :].argmax(-1, True)`` (reference model.py:268) on a plan's logits returns the argmax the lm_head GEMM
epilogue / the decode kernel already produced (``Tensor._pdn_hint``): same values, no second pass over [B, 32000].
"""
import ctypes
import math
import os
import sys
import warnings

import numpy as np

from ..autograd import is_grad_enable
from ..backend.array import WRITE_EPOCH
from ..core.tensor import Tensor, _result
from . import _fused
from ._fused import Planes, DevicePos, _call, _empty, _i64, _bhl_strides, _PackedWeight, _c
from ..backend.array import ndarray

ENABLED = os.environ.get("PDN_PLANS", "1") != "0"
VERIFY = os.environ.get("PDN_PLAN_VERIFY", "1") != "0"
STRUCT_EPOCH = [0]
F32 = np.dtype(np.float32)
I64 = np.dtype(np.int64)
MEGA_MAX_ROWS = 8  # batch rows one decode_mega launch serves (register accumulators per warp)


def decode_branches(B):
    """Concurrent batch slices of the recorded decode step: 1 (one launch chain) unless PDN_DECODE_BRANCHES=2|4 asks for a fork.
    Measured at batch 1024 on B200: 1.455 M tokens/s unforked, 1.354 M with 2 slices, 1.235 M with 4 — the
    KV-cache attention and the lm_head GEMM lose more at half / quarter batch (partial last wave; the 37 MB of lm_head weight
    planes re-read per slice) than the overlapped launch chains win, so the fork stays an opt-in experiment."""
    nb = int(os.environ.get("PDN_DECODE_BRANCHES", "1"))
    nb = max(1, min(4, nb))
    while nb > 1 and (B % nb or B // nb < 32):
        nb //= 2
    return nb


def note_structure_change():
    """A Module gained / lost a Parameter or sub-Module, a Parameter was re-bound or moved: every plan re-validates."""
    STRUCT_EPOCH[0] += 1


def _write_epoch():
    return WRITE_EPOCH[0]


def attach(module):
    """Called once per Module instance from ``Module.__call__``; returns a plan or False."""
    plan = False
    if ENABLED:
        try:
            plan = DecoderPlan.match(module) or AttentionPlan.match(module) or False
        except Exception:  # a structural probe must never break a forward
            plan = False
    object.__setattr__(module, "_pdn_plan", plan)
    return plan


class _Hint:
    """Memoised ``argmax over the vocabulary`` of a plan's logits, valid while the producing step is the latest one."""
    __slots__ = ("state", "step_no", "slot", "sliced")

    def __init__(self, state, step_no, slot, sliced=False):
        self.state, self.step_no, self.slot, self.sliced = state, step_no, slot, sliced


class _LogitsCore:
    """The lm_head input of one batched decode step (final-RMSNorm rows as GEMM operand planes) + where its argmax went.
    The step itself ran the lm_head GEMM with the argmax epilogue only — greedy decoding (reference model.py:268) needs
    nothing else and the [B, 32000] logits never reach HBM.  Anything that reads the logits' values materialises them here
    with the same GEMM (+bias) from the kept planes; a recorded step that is about to overwrite the planes of a core that is
    still alive materialises it first (``_DecodeState.release_slot``), so a caller can never observe later data."""
    __slots__ = ("plan", "planes", "buf", "__weakref__")

    def __init__(self, plan, planes):
        self.plan, self.planes, self.buf = plan, planes, None

    def materialise(self):
        if self.buf is None:
            plan, pl = self.plan, self.planes
            m = plan.model
            V = m.lm_head.weight.shape[1]
            bias = m.lm_head.bias
            parts = pl if isinstance(pl, list) else [pl]  # a forked step leaves one set of planes per batch slice
            with plan.device:
                out = _empty((sum(q.M for q in parts), V))
                row = 0
                for q in parts:
                    _call("pdn_gemm_prepacked_planes", q.ptr, q.M, q.Kp, _fused._packed(m.lm_head.weight).handle, out.ptr + row * V * 4, V,
                          _c(bias.data).ptr if bias is not None else None, 0)
                    row += q.M
            self.buf, self.planes = out, None
        return self.buf


class _LazyLogits(Tensor):
    """``Tensor`` whose array ([B, 1, V], or the [B, V] slice of it) is produced on first access of ``.data``."""

    @classmethod
    def make(cls, core, shape, device, hint):
        out = Tensor.__new__(cls)
        out.device = device
        out._core, out._shape, out._data = core, shape, None
        out._grad = None
        out._pinned_grad = out._grad_stale = False
        out._node = None
        out.requires_grad = False
        out._pdn_hint = hint
        return out

    @property
    def data(self):
        d = self._data
        if d is None:
            d = self._data = self._core.materialise().reshape(self._shape)
        return d

    @data.setter
    def data(self, value):
        self._data = value
        self._pdn_hint = None  # re-bound storage: the memo no longer describes this tensor

    shape = property(lambda self: self._shape if self._data is None else self._data.shape)
    ndim = property(lambda self: len(self.shape))
    dtype = property(lambda self: F32 if self._data is None else self._data.dtype)
    size = property(lambda self: int(np.prod(self.shape)))


_ALL = slice(None)


def hint_getitem(x, key):
    """``logits[:, -1, :]`` on a plan's [B, 1, V] logits (reference model.py:268): the [B, V] view, built directly (no generic
    index parsing) and still carrying the argmax memo. Any other key: None (the generic path serves it, memo dropped)."""
    h = x._pdn_hint
    if (not h.sliced and type(key) is tuple and len(key) == 3 and key[0] == _ALL and key[2] == _ALL
            and isinstance(key[1], (int, np.integer)) and key[1] in (-1, 0)):
        hint = _Hint(h.state, h.step_no, h.slot, True)
        if type(x) is _LazyLogits and x._data is None:
            return _LazyLogits.make(x._core, (x._shape[0], x._shape[2]), x.device, hint)
        d = x.data
        out = _result(d._view((d.shape[0], d.shape[2]), (d.estrides[0], d.estrides[2])), x.device, (), None, "_get_slice")
        out._pdn_hint = hint
        return out
    return None


def hint_argmax(x, axis, keepdims):
    """Returns the memoised ids Tensor for ``x.argmax(-1, True)`` or None when the hint does not apply (any more)."""
    h = x._pdn_hint
    if not h.sliced or not keepdims or axis not in (-1, 1) or x.ndim != 2:
        return None
    return h.state.take_ids(h)


# ------------------------------------------------------------------------------------------- decode state -----
class _DecodeState:
    """Per (plan, batch size) buffers of the recorded decode step: ping-pong ids buffers and lm_head operand planes, the
    device position, the two CUDA graphs (one per slot)."""

    def __init__(self, plan, B):
        self.plan, self.B = plan, B
        dev = plan.device
        with dev:
            self.ids = [ndarray.empty((B, 1), I64) for _ in range(2)]
            self.pos = DevicePos(Tensor(np.zeros(1, dtype=np.int64), device=dev))
        self.planes = [None, None]  # lm_head operand planes written by the recorded step of each slot
        self.live = [None, None]  # weakref to the _LogitsCore that still points at planes[slot]
        self.graphs = [None, None]
        self.slot = 0  # slot the NEXT step writes
        self.step_no = 0
        self.dev_pos = None  # value currently held by the device-side position (None: unknown)
        self.last_out = None  # (ids Tensor handed to the caller, version of its buffer, slot whose device copy equals it)
        self.steps_run = 0

    def destroy(self):
        for slot in (0, 1):
            self.release_slot(slot)
        for g in self.graphs:
            if g is not None:
                with self.plan.device:
                    g.destroy()
        self.graphs = [None, None]

    def release_slot(self, slot):
        """The planes of ``slot`` are about to be overwritten (or freed): a logits tensor that still refers to them gets its
        values now."""
        ref = self.live[slot]
        if ref is not None:
            core = ref()
            if core is not None and core.buf is None and core.planes is not None:
                core.materialise()
            self.live[slot] = None

    def take_ids(self, hint):
        if hint.step_no != self.step_no:
            return None  # a later step has overwritten the slot: the caller's argmax runs on the logits it holds
        with self.plan.device:
            out = ndarray.empty((self.B, 1), I64)
            _call("pdn_memcpy_d2d", out.ptr, self.ids[hint.slot].ptr, self.B * 8)
        t = _result(out, self.plan.device, (), None, "argmax")
        self.last_out = (t, out.buf.version, hint.slot)
        return t


# ------------------------------------------------------------------------------------------- the plan ---------
def _is_linear(m, bias):
    w = getattr(m, "weight", None)
    if not isinstance(w, Tensor) or w.ndim != 2:
        return False
    b = getattr(m, "bias", None)
    return (b is None) if bias is False else True


def _is_rmsnorm(m, dim):
    w = getattr(m, "weight", None)
    return isinstance(w, Tensor) and w.shape == (dim, ) and hasattr(m, "eps") and type(m).__name__ == "RMSNorm"


class DecoderPlan:

    @classmethod
    def match(cls, m):
        from .modules.module import Module, ModuleList
        from .modules.layers import Embedding, Linear
        emb, layers = getattr(m, "tok_embedding", None), getattr(m, "layers", None)
        if not isinstance(emb, Embedding) or emb.padding_idx is not None or not isinstance(layers, ModuleList) or len(layers) == 0:
            return None
        cos, sin = getattr(m, "freqs_cos", None), getattr(m, "freqs_sin", None)
        head, norm = getattr(m, "lm_head", None), getattr(m, "norm", None)
        if not (isinstance(cos, Tensor) and isinstance(sin, Tensor) and isinstance(head, Linear)):
            return None
        V, D = emb.weight.shape
        if not _is_rmsnorm(norm, D) or head.weight.shape != (D, V) or cos.ndim != 2 or cos.shape != sin.shape:
            return None
        blocks = []
        for blk in layers:
            att, ffn = getattr(blk, "attention", None), getattr(blk, "ffn", None)
            n1, n2 = getattr(blk, "input_norm", None), getattr(blk, "post_attn_norm", None)
            if not (isinstance(att, Module) and isinstance(ffn, Module) and _is_rmsnorm(n1, D) and _is_rmsnorm(n2, D)):
                return None
            for nm in "QKVO":
                lin = getattr(att, nm, None)
                if not isinstance(lin, Linear) or lin.bias is not None or lin.weight.shape != (D, D):
                    return None
            ck, cv = getattr(att, "cache_k", None), getattr(att, "cache_v", None)
            H, hd = getattr(att, "n_heads", None), getattr(att, "head_dim", None)
            if not (isinstance(ck, Tensor) and isinstance(cv, Tensor) and isinstance(H, int) and isinstance(hd, int)):
                return None
            if H * hd != D or ck.ndim != 4 or ck.shape != cv.shape or ck.shape[2:] != (H, hd) or cos.shape[1] * 2 != hd:
                return None
            up, gate, down = (getattr(ffn, nm, None) for nm in ("up", "gate", "down"))
            if not all(isinstance(x, Linear) and x.bias is None for x in (up, gate, down)):
                return None
            FF = up.weight.shape[1]
            if up.weight.shape != (D, FF) or gate.weight.shape != (D, FF) or down.weight.shape != (FF, D):
                return None
            blocks.append((blk, att, ffn, n1, n2))
        if hd > 128 or hd % 4 != 0 or D % 8 != 0:
            return None
        return cls(m, blocks)

    def __init__(self, model, blocks):
        self.model, self.blocks = model, blocks
        self.dead = False
        self.verified = set()
        self.device = None
        self._epoch = None
        self._sig = None
        self._cat = {}  # fused [W0 | W1 | ...] operand planes per (layer, kind)
        self._dec = {}  # batch size -> _DecodeState
        self._mega = None
        self._mega_failed = False  # the decode kernel refused this model (too wide for one task per warp on this GPU)
        self._masks = {}

    # ---------------------------------------------------------------- validity -----------------
    def _weights(self):
        m = self.model
        ws = [m.tok_embedding.weight, m.freqs_cos, m.freqs_sin, m.norm.weight, m.lm_head.weight]
        if m.lm_head.bias is not None:
            ws.append(m.lm_head.bias)
        caches = []
        for blk, att, ffn, n1, n2 in self.blocks:
            ws += [att.Q.weight, att.K.weight, att.V.weight, att.O.weight, ffn.gate.weight, ffn.up.weight, ffn.down.weight, n1.weight,
                   n2.weight]
            caches += [att.cache_k, att.cache_v]
        return ws, caches

    def _revalidate(self) -> bool:
        """Re-derives the weight signature; drops recorded graphs / packed operands when a weight buffer moved or was
        written. False: the model cannot be served right now (not on one cuda device in fp32)."""
        ws, caches = self._weights()
        dev = ws[0].device
        if not dev.is_cuda:
            return False
        for t in ws + caches:
            if t.device != dev or t.data.dtype != F32 or not t.data.is_contiguous:
                return False
        sig = tuple((t.data.ptr, t.data.buf.version, t.data.shape) for t in ws) + tuple((t.data.ptr, t.data.shape) for t in caches)
        if sig != self._sig:
            self._drop_recorded()
            self._sig, self.device = sig, dev
        return True

    def _drop_recorded(self):
        for st in self._dec.values():
            st.destroy()
        self._dec.clear()
        self._cat.clear()
        self._mega = None

    # ---------------------------------------------------------------- entry --------------------
    def __call__(self, args):
        m = self.model
        if self.dead or not ENABLED or m._train or is_grad_enable() or len(args) != 2:
            return NotImplemented
        ids, pos = args
        if isinstance(pos, (bool, float)) or not isinstance(pos, (int, np.integer)):
            return NotImplemented
        epoch = (STRUCT_EPOCH[0], _write_epoch())
        if epoch != self._epoch:
            if not self._revalidate():
                return NotImplemented
        if isinstance(ids, Tensor):
            if ids.device != self.device or ids.ndim != 2 or ids.data.dtype != I64:
                return NotImplemented
        elif isinstance(ids, np.ndarray) and ids.ndim == 2 and np.issubdtype(ids.dtype, np.integer):
            pass
        else:
            return NotImplemented
        B, L = ids.shape
        pos = int(pos)
        ck = self.blocks[0][1].cache_k
        if B > ck.shape[0] or pos < 0 or pos + L > ck.shape[1] or L == 0:
            return NotImplemented  # the eager path raises the reference's own error for these
        kind = "decode" if L == 1 else "prefill"
        path = "mega" if (B * L < 32 and B <= MEGA_MAX_ROWS and self._mega_ok()) else ("tc" if B * L >= 32 else None)
        if path is None:
            return NotImplemented
        with self.device:
            if VERIFY and (kind, path) not in self.verified:
                out = self._verified_first_call(kind, path, ids, pos, B, L)
            else:
                out = self._run(kind, path, ids, pos, B, L)
        self._epoch = (STRUCT_EPOCH[0], _write_epoch())  # absorbs this call's own internal writes
        return out

    def _verified_first_call(self, kind, path, ids, pos, B, L):
        m = self.model
        want = m.forward(ids, pos)
        got = self._run(kind, path, ids, pos, B, L)
        a, b = want.numpy().astype(np.float64), got.numpy().astype(np.float64)
        err = np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30) if a.shape == b.shape else float("inf")
        if not (err < 1e-3):
            self.dead = True
            self._drop_recorded()
            warnings.warn(f"pydynet_b200: inference plan for {type(m).__name__} disagrees with its eager forward "
                          f"({kind}/{path}: normwise {err:.2e}); the model keeps running eagerly")
            return want
        self.verified.add((kind, path))
        return got

    def _run(self, kind, path, ids, pos, B, L):
        if path == "mega":
            return self._run_mega(ids, pos, B, L)
        if kind == "prefill":
            return self._prefill_tc(ids, pos, B, L)
        return self._decode_tc(ids, pos, B)

    # ---------------------------------------------------------------- rows >= 32: tcgen05 GEMM path ------------
    def _ids_tensor(self, ids):
        return ids if isinstance(ids, Tensor) else Tensor(ids, dtype=np.int64, device=self.device)

    def _packed_cat(self, key, weights):
        sig = tuple((w.data.ptr, w.data.buf.version) for w in weights)
        ent = self._cat.get(key)
        if ent is None or ent[0] != sig:
            from .. import backend as bk
            cat = bk.concatenate([w.data for w in weights], axis=1)
            ent = self._cat[key] = (sig, _PackedWeight(cat), cat.shape[1])
        return ent[1], ent[2]

    def _gemm_cat(self, key, pl, weights):
        pw, N = self._packed_cat(key, weights)
        out = _empty((pl.M, N))
        _call("pdn_gemm_prepacked_planes", pl.ptr, pl.M, pl.Kp, pw.handle, out.ptr, N, None, 0)
        return out

    def _block(self, i, x, start_pos, mask, B, L, b0=0, chain=None):
        """One transformer block in place on the residual stream ``x`` [B, L, dim] (reference model.py:142-150, 95-121);
        ``b0``: first sequence of the KV cache this batch slice belongs to; ``chain``: [token] of the previous attention launch
        of a forked step (the attention launches of all branches run one after the other)."""
        blk, att, ffn, n1, n2 = self.blocks[i]
        m = self.model
        H, D = att.n_heads, att.head_dim
        dim = H * D
        qkv = self._gemm_cat((i, "qkv"), _fused.rmsnorm_planes(x, n1.weight, n1.eps), (att.Q.weight, att.K.weight, att.V.weight))
        q3 = qkv.reshape(B, L, 3, H, D)
        q, k, v = q3[:, :, 0], q3[:, :, 1], q3[:, :, 2]
        ck, cv = att.cache_k.data, att.cache_v.data
        if B != ck.shape[0]:
            ck, cv = ck[b0:b0 + B], cv[b0:b0 + B]  # views: same buffer (and version counter)
        S = ck.shape[1]
        cos, sin = m.freqs_cos.data, m.freqs_sin.data
        pl = Planes((B, L), dim, self.device)
        cstr = _i64((ck.estrides[0], ck.estrides[2], ck.estrides[1]))
        scale = 1.0 / math.sqrt(D)
        ld = 3 * dim
        if isinstance(start_pos, DevicePos):
            pp = start_pos.tensor.data.ptr
            _call("pdn_rope_kv_append_dev", q.ptr, k.ptr, v.ptr, cos.ptr, sin.ptr, ck.ptr, cv.ptr, B, L, H, D, S, pp, ld)
            if chain is not None:
                _call("pdn_branch_wait", chain[0])
            _call("pdn_attention_fwd_dev", q.ptr, ck.ptr, cv.ptr, None, B, H, L, D, _bhl_strides(q), cstr, cstr, scale, pp, L, pl.ptr, pl.Kp)
            if chain is not None:
                tok = ctypes.c_int32(-1)
                _call("pdn_branch_mark", ctypes.byref(tok))
                chain[0] = tok.value
        else:
            _call("pdn_rope_kv_append", q.ptr, k.ptr, v.ptr, cos.ptr, sin.ptr, ck.ptr, cv.ptr, B, L, H, D, S, start_pos, ld)
            Lk = start_pos + L
            keep, mptr, mstr = _fused._mask_args(mask, B, H, L, Lk)
            _call("pdn_attention_fwd", q.ptr, ck.ptr, cv.ptr, mptr, None, None, B, H, L, Lk, D, _bhl_strides(q), cstr, cstr, mstr, scale,
                  pl.ptr, pl.Kp)
            del keep
        ck.buf.version += 1
        cv.buf.version += 1
        z = _fused.linear_residual_(pl, att.O.weight, x)
        gu = self._gemm_cat((i, "gu"), _fused.rmsnorm_planes(z, n2.weight, n2.eps), (ffn.gate.weight, ffn.up.weight))
        FF = ffn.up.weight.shape[1]
        hp = Planes((B, L), FF, self.device)
        _call("pdn_swiglu_rows_planes", gu.ptr, hp.ptr, hp.M, FF, hp.Kp)
        return _fused.linear_residual_(hp, ffn.down.weight, z)

    def _mask(self, L, start_pos):
        """Causal mask over [cached positions | new positions], built on the host like the reference (model.py:199-203)."""
        key = (L, start_pos)
        mk = self._masks.get(key)
        if mk is None:
            if len(self._masks) > 16:
                self._masks.clear()
            host = np.concatenate([np.zeros((L, start_pos)), np.triu(np.full((L, L), float("-inf")), k=1)], axis=1)
            mk = self._masks[key] = Tensor(host, device=self.device, dtype=np.float32)
        return mk

    def _hidden_tc(self, ids_t, start_pos, B, L):
        from . import functional as F
        m = self.model
        h = F.embedding(ids_t, m.tok_embedding.weight, None)  # fresh [B, L, dim] buffer: becomes the residual stream
        mask = self._mask(L, start_pos) if L > 1 else None
        for i in range(len(self.blocks)):
            h = self._block(i, h, start_pos, mask, B, L)
        return h

    def _head_planes(self, h, B, L):
        m = self.model
        last = h if L == 1 else h[:, -1, :]
        return _fused.rmsnorm_planes(last, m.norm.weight, m.norm.eps)

    def _prefill_tc(self, ids, start_pos, B, L):
        m = self.model
        h = self._hidden_tc(self._ids_tensor(ids), start_pos, B, L)
        pl = self._head_planes(h, B, L)
        V = m.lm_head.weight.shape[1]
        out = _empty((B, V))
        bias = m.lm_head.bias
        _call("pdn_gemm_prepacked_planes", pl.ptr, pl.M, pl.Kp, _fused._packed(m.lm_head.weight).handle, out.ptr, V,
              _c(bias.data).ptr if bias is not None else None, 0)
        return _result(out.reshape(B, 1, V), self.device, (), None, "plan_logits")

    def _decode_step_into(self, st, slot_in, slot_out, pos):
        """Launch sequence of one decode step: ids[slot_in] -> lm_head operand planes, ids[slot_out] = argmax of the logits
        (GEMM epilogue); ``pos`` is a host int (eager launches) or the state's DevicePos (graph recording; the recorded step
        also advances it). Returns the planes."""
        m, B = self.model, st.B
        nb = decode_branches(B) if isinstance(pos, DevicePos) else 1
        if nb > 1:
            return self._decode_step_forked(st, slot_in, slot_out, pos, nb)
        ids_t = _result(st.ids[slot_in], self.device, (), None, "ids")
        h = self._hidden_tc(ids_t, pos, B, 1)
        pl = self._head_planes(h, B, 1)
        bias = m.lm_head.bias
        bptr = _c(bias.data).ptr if bias is not None else None
        _call("pdn_gemm_prepacked_planes_argmax", pl.ptr, pl.M, pl.Kp, _fused._packed(m.lm_head.weight).handle, bptr, st.ids[slot_out].ptr)
        if isinstance(pos, DevicePos):
            pos.tensor += 1
        return pl

    def _decode_step_forked(self, st, slot_in, slot_out, pos, nb):
        """The recorded decode step as ``nb`` concurrent branches over batch slices (``pdn_branch_*``): sequences are independent,
        and at batch 1024 half of the step is a chain of ~50 launch-latency-bound kernels (GEMMs of 40-120 CTAs, row kernels)
        between the HBM-bound KV-cache attention launches — the chain of one slice runs in the shadow of another slice's
        attention. Layers are issued round-robin over the branches; the attention launches are ordered one after the other across
        branches (each keeps the whole HBM bandwidth, and its CUDA-event timing stays that of one kernel)."""
        from . import functional as F
        m, B = self.model, st.B
        Bs = B // nb
        bias = m.lm_head.bias
        bptr = _c(bias.data).ptr if bias is not None else None
        _call("pdn_branch_begin", nb)
        try:
            hs = []
            for r in range(nb):
                _call("pdn_branch_select", r)
                ids_t = _result(st.ids[slot_in][r * Bs:(r + 1) * Bs], self.device, (), None, "ids")
                hs.append(F.embedding(ids_t, m.tok_embedding.weight, None))
            chain = [-1] if os.environ.get("PDN_DECODE_ORDER", "1") != "0" else None
            for i in range(len(self.blocks)):
                for r in range(nb):
                    _call("pdn_branch_select", r)
                    hs[r] = self._block(i, hs[r], pos, None, Bs, 1, r * Bs, chain)
            planes = []
            for r in range(nb):
                _call("pdn_branch_select", r)
                pl = self._head_planes(hs[r], Bs, 1)
                _call("pdn_gemm_prepacked_planes_argmax", pl.ptr, pl.M, pl.Kp, _fused._packed(m.lm_head.weight).handle, bptr,
                      st.ids[slot_out].ptr + r * Bs * 8)
                planes.append(pl)
            del hs
        finally:
            _call("pdn_branch_end")
        pos.tensor += 1
        return planes

    def _decode_tc(self, ids, start_pos, B):
        import weakref
        from .. import cuda
        st = self._dec.get(B)
        if st is None:
            st = self._dec[B] = _DecodeState(self, B)
        slot_out, slot_in = st.slot, st.slot ^ 1
        # token ids of this step: already in ids[slot_in] when the caller passes back the argmax Tensor of the previous step
        lo = st.last_out
        if not (lo is not None and ids is lo[0] and lo[2] == slot_in and ids.data.buf.version == lo[1]):
            if isinstance(ids, Tensor):
                _call("pdn_memcpy_d2d", st.ids[slot_in].ptr, _c(ids.data).ptr, B * 8)
            else:
                host = np.ascontiguousarray(ids, dtype=np.int64)
                _call("pdn_memcpy_h2d", st.ids[slot_in].ptr, host.ctypes.data, host.nbytes)
        st.last_out = None
        use_graph = os.environ.get("PDN_DECODE_GRAPH", "1") != "0" and st.steps_run >= 1 and not cuda.is_capturing()
        if not use_graph:
            pl = self._decode_step_into(st, slot_in, slot_out, start_pos)  # fresh planes: the core simply keeps them
            core = _LogitsCore(self, pl)
        else:
            st.release_slot(slot_out)
            if st.dev_pos != start_pos:
                st.pos.tensor.data.fill(start_pos)
            g = st.graphs[slot_out]
            if g is None:
                g = cuda.Graph()
                g.begin()
                try:
                    st.planes[slot_out] = self._decode_step_into(st, slot_in, slot_out, st.pos)
                finally:
                    g.end()
                st.graphs[slot_out] = g
            g.launch()
            st.dev_pos = start_pos + 1  # the recorded step advances it
            for blk, att, ffn, n1, n2 in self.blocks:
                att.cache_k.data.buf.version += 1
                att.cache_v.data.buf.version += 1
            core = _LogitsCore(self, st.planes[slot_out])
            st.live[slot_out] = weakref.ref(core)
        V = self.model.lm_head.weight.shape[1]
        st.steps_run += 1
        st.step_no += 1
        st.slot = slot_in
        return _LazyLogits.make(core, (B, 1, V), self.device, _Hint(st, st.step_no, slot_out))

    # ---------------------------------------------------------------- rows < 32: persistent decode kernel ------
    def _mega_ok(self):
        if os.environ.get("PDN_DECODE_MEGA", "1") == "0":
            return False
        att, ffn = self.blocks[0][1], self.blocks[0][2]
        D, FF = att.n_heads * att.head_dim, ffn.up.weight.shape[1]
        return (not self._mega_failed and att.head_dim in (32, 48, 64) and D % 4 == 0 and FF % 4 == 0 and D <= 1024 and FF <= 1024
                and att.cache_k.shape[1] <= 2048 and all(b[2].up.weight.shape[1] == FF for b in self.blocks))

    def _mega_weights(self):
        """Transposed ([out][in]) copies the persistent kernel streams row by row; rebuilt when a weight buffer changes
        (``_drop_recorded``)."""
        if self._mega is None:
            from .. import backend as bk
            m = self.model
            T = lambda w: w.data.swapaxes(0, 1).copy()
            layers = []
            for blk, att, ffn, n1, n2 in self.blocks:
                wqkv = bk.concatenate([T(att.Q.weight), T(att.K.weight), T(att.V.weight)], axis=0)
                FF, D = ffn.up.weight.shape[1], ffn.up.weight.shape[0]
                wgu = ndarray.empty((FF, 2, D), F32)
                wgu[:, 0, :] = T(ffn.gate.weight)
                wgu[:, 1, :] = T(ffn.up.weight)
                H, hd = att.n_heads, att.head_dim
                wo_h = att.O.weight.data.reshape(H, hd, D).swapaxes(1, 2).copy()  # [H][out][hd]: the head slice of Woᵀ is contiguous
                layers.append((wqkv, wo_h, wgu, T(ffn.down.weight), n1, n2, att))
            self._mega = {"layers": layers, "wlm": T(m.lm_head.weight), "states": {}}
        return self._mega

    def _run_mega(self, ids, pos, B, L):
        mg = self._mega_weights()
        st = mg["states"].get(B)
        if st is None:
            try:
                st = mg["states"][B] = _MegaState(self, mg, B)
            except RuntimeError:
                self._mega_failed = True
                return self.model.forward(ids, pos)
        return st.run(ids, pos, L)


class _MegaState:
    """Handle + output buffers of the persistent decode kernel for one batch size (csrc/decode_mega.cu)."""

    def __init__(self, plan, mg, B):
        import ctypes as C
        self.plan, self.B = plan, B
        m = plan.model
        att, ffn = plan.blocks[0][1], plan.blocks[0][2]
        self.H, self.D, self.FF = att.n_heads, att.n_heads * att.head_dim, ffn.up.weight.shape[1]
        self.V, self.S = m.lm_head.weight.shape[1], att.cache_k.shape[1]
        n = len(mg["layers"])
        ptrs = (C.c_void_p * (8 * n))()
        eps = (C.c_float * (2 * n))()
        for i, (wqkv, wo, wgu, wd, n1, n2, a) in enumerate(mg["layers"]):
            for j, arr in enumerate((wqkv, wo, wgu, wd, n1.weight.data, n2.weight.data, a.cache_k.data, a.cache_v.data)):
                ptrs[8 * i + j] = arr.ptr
            eps[2 * i], eps[2 * i + 1] = n1.eps, n2.eps
        bias = m.lm_head.bias
        self.handle = C.c_void_p()
        _call("pdn_decoder_create", C.byref(self.handle), n, B, self.D, self.H, self.FF, self.V, self.S, ptrs, eps,
              m.tok_embedding.weight.data.ptr, m.freqs_cos.data.ptr, m.freqs_sin.data.ptr, m.norm.weight.data.ptr, m.norm.eps,
              mg["wlm"].ptr, bias.data.ptr if bias is not None else None)
        self._keep = mg  # the transposed weights live as long as the handle
        self.logits = [None, None]
        self.base_ref = [0, 0]
        self.slot = 0
        self.hist = None  # [S + 1, B] int64: the id produced at position p lives in row p (distinct storage per step)
        self.hist_ref = 0
        self.gen = 0
        self.last_row = -1
        self.last_out = None  # (ids Tensor handed out, version, row)
        self.staging = ndarray.empty((B, 1), I64)
        self.caches = [(a.cache_k.data.buf, a.cache_v.data.buf) for (_, _, _, _, _, _, a) in mg["layers"]]

    def __del__(self):
        try:
            from ..backend import lib
            if lib._lib is not None and self.handle:
                lib._lib.pdn_decoder_destroy(self.handle)
        except Exception:
            pass

    def _new_hist(self):
        self.hist = ndarray.empty((self.S + 1, self.B), I64)
        self.hist_ref = sys.getrefcount(self.hist.buf)
        self.gen += 1

    def take_ids(self, hint):
        if hint.step_no != self.gen:
            return None
        row = hint.slot
        view = self.hist._view((self.B, 1), (1, 1), row * self.B)
        t = _result(view, self.plan.device, (), None, "argmax")
        self.last_out = (t, self.hist.buf.version, row)
        return t

    def run(self, ids, pos, L):
        B, V = self.B, self.V
        # where this step's token ids live on the device: the history row when the caller passes back our own argmax Tensor,
        # the caller's own int64 buffer otherwise (no copy), a staging buffer for host ids
        lo = self.last_out
        if lo is not None and ids is lo[0] and L == 1 and self.hist.buf.version == lo[1]:
            ids_ptr, stride = self.hist.ptr + lo[2] * B * 8, 1
        elif isinstance(ids, Tensor):
            d = ids.data
            if d.estrides[1] != 1 and L > 1:
                d = d.copy()
            ids_ptr, stride = d.ptr, d.estrides[0] if B > 1 else L
            keep = d
        else:
            host = np.ascontiguousarray(ids, dtype=np.int64)
            if L > 1:
                keep = ndarray.from_host(host)
                ids_ptr, stride = keep.ptr, L
            else:
                _call("pdn_memcpy_h2d", self.staging.ptr, host.ctypes.data, host.nbytes)
                ids_ptr, stride = self.staging.ptr, 1
        row = pos + L - 1
        if self.hist is None or row <= self.last_row:
            # a new generation (positions restart): id Tensors of the previous one that the caller still holds keep their rows
            if self.hist is None or sys.getrefcount(self.hist.buf) > self.hist_ref:
                self._new_hist()
            else:
                self.gen += 1
        self.last_row = row
        slot = self.slot
        lg = self.logits[slot]
        if lg is None or sys.getrefcount(lg.buf) > self.base_ref[slot]:
            lg = self.logits[slot] = ndarray.empty((B, V), F32)
        self.base_ref[slot] = sys.getrefcount(lg.buf)
        h = self.handle
        for l in range(L - 1):  # prompt positions before the last one: KV cache only
            _call("pdn_decoder_step", h, ids_ptr + 8 * l, stride, pos + l, None, None)
        _call("pdn_decoder_step", h, ids_ptr + 8 * (L - 1), stride, row, lg.ptr, self.hist.ptr + row * B * 8)
        for ck, cv in self.caches:
            ck.version += 1
            cv.version += 1
        self.slot = slot ^ 1
        self.last_out = None
        out = _result(lg.reshape(B, 1, V), self.plan.device, (), None, "plan_logits")
        out._pdn_hint = _Hint(self, self.gen, row)
        return out


# ------------------------------------------------------------------------------------------- attention module ---
class AttentionPlan:
    """A multi-head self-attention MODULE written as an operator chain — the reference's examples/pydynet/transformer.py:53-104:
    bias-free ``Q/K/V/O`` projections of ``values``, head split by reshape / transpose, ``q kT / sqrt(hd)`` (+ additive mask whose
    entries equal to 1 are first turned into -inf IN PLACE) -> softmax -> ``@ v`` -> merge heads -> ``O`` — is served by ONE fused
    attention operator (forward and backward, csrc/attention_tc.cu / attention.cu) between the unchanged projections; the
    [B, H, L, L] score tensor and its five temporaries never exist. Works in training and inference; the first call is checked
    against the module's own ``forward`` (normwise 1e-3), a mismatch retires the plan."""

    @classmethod
    def match(cls, m):
        import inspect
        from .modules.layers import Linear
        E, H, hd = getattr(m, "embed_size", None), getattr(m, "heads", None), getattr(m, "head_dim", None)
        if not (isinstance(E, int) and isinstance(H, int) and isinstance(hd, int) and H * hd == E):
            return None
        for nm in "QKVO":
            lin = getattr(m, nm, None)
            if not isinstance(lin, Linear) or lin.bias is not None or lin.weight.shape != (E, E):
                return None
        try:
            names = tuple(inspect.signature(m.forward).parameters)
        except (TypeError, ValueError):
            return None
        if names != ("values", "keys", "query", "mask"):
            return None
        return cls(m)

    def __init__(self, module):
        self.m, self.dead, self.verified = module, False, False

    def __call__(self, args):
        if self.dead or not ENABLED or len(args) != 4:
            return NotImplemented
        values, keys, query, mask = args
        m = self.m
        if not (isinstance(values, Tensor) and isinstance(keys, Tensor) and isinstance(query, Tensor) and values.ndim == 3
                and values.shape == keys.shape == query.shape and values.shape[2] == m.embed_size):
            return NotImplemented
        if mask is not None and not isinstance(mask, Tensor):
            return NotImplemented
        if not _fused.usable(values, m.Q.weight, m.K.weight, m.V.weight, m.O.weight, op="attention"):
            return NotImplemented
        N, L = values.shape[0], values.shape[1]
        H, D = m.heads, m.head_dim
        if _fused._attention_impl(N, H, L, L, D) == "":
            return NotImplemented
        if VERIFY and not self.verified:
            from ..autograd import no_grad
            with no_grad():
                want = m.forward(values, keys, query, mask).numpy().astype(np.float64)
                got = self._run(values, mask, N, L, H, D).numpy().astype(np.float64)
            err = np.linalg.norm(want - got) / max(np.linalg.norm(want), 1e-30) if want.shape == got.shape else float("inf")
            if not (err < 1e-3):
                self.dead = True
                warnings.warn(f"pydynet_b200: fused attention plan for {type(m).__name__} disagrees with its forward ({err:.2e}); eager path kept")
                return NotImplemented
            self.verified = True
        return self._run(values, mask, N, L, H, D)

    def _run(self, values, mask, N, L, H, D):
        m = self.m
        xq, xk, xv = (proj(values).reshape(N, L, H, D) for proj in (m.Q, m.K, m.V))
        if mask is not None:
            mask[mask.eq(1)] = np.float32("-inf")  # the reference mutates the caller's mask (transformer.py:97)
        return m.O(_fused.attention(xq, xk, xv, mask, 1.0 / D**.5))
