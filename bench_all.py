"""bench_all.py — the other BASELINE configs and the micro-benchmarks of SURVEY.md §8(d), one JSON line each.

(bench.py carries the headline metric; this script produces the per-config numbers quoted in README.md / profiles/.)
  C1 matmul fwd+bwd (512^3 literal + size sweep)      C2 LeNet b256 train step        C4 Transformer encoder train step
  C5 GRU T=1024 train step                            micro: Conv2d fwd+bwd, attention core fwd+bwd, GEMM fwd
Usage: python bench_all.py [--only c1,c2,...] [--steps K] [--small]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import pydynet_b200 as pdn  # noqa: E402
import pydynet_b200.nn as nn  # noqa: E402
import pydynet_b200.nn.functional as F  # noqa: E402
from pydynet_b200.backend import lib  # noqa: E402
from pydynet_b200.optim import Adam  # noqa: E402

DEV = "cuda:0"
f32 = np.float32
ARGS = None
PEAK_BF16 = 1362.4e12
PEAK_HBM = 6552.6e9
try:
    _p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_BF16, PEAK_HBM = _p.get("bf16_tflops_sustained", _p["bf16_tflops"]) * 1e12, _p["hbm_gbs"] * 1e9
except Exception:
    pass


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    pdn.cuda.synchronize()
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.call("pdn_event_create", C.byref(e0))
    lib.call("pdn_event_create", C.byref(e1))
    lib.reset_launch_count()
    lib.call("pdn_event_record", e0)
    for _ in range(steps):
        fn()
    lib.call("pdn_event_record", e1)
    ms = C.c_float()
    lib.load().pdn_event_elapsed_ms(e0, e1, C.byref(ms))
    return ms.value / steps / 1e3, lib.launch_count() // steps


def emit(name, sec, flops=None, nbytes=None, launches=None, **extra):
    r = {"config": name, "ms_per_step": sec * 1e3, "launches_per_step": launches}
    if flops:
        r["tflops_alg"] = flops / sec / 1e12
        r["frac_of_bf16_peak"] = flops / sec / PEAK_BF16
        r["frac_of_bf16x3_ceiling"] = 3 * flops / sec / PEAK_BF16
    if nbytes:
        r["gbs_alg"] = nbytes / sec / 1e9
        r["frac_of_hbm_peak"] = nbytes / sec / PEAK_HBM
    r.update(extra)
    print(json.dumps(r), flush=True)
    return r


def T(a, rg=False):
    return pdn.Tensor(a, dtype=a.dtype, device=DEV, requires_grad=rg)


# ------------------------------------------------------------------------------------------------- the reference beside every number
# The UNMODIFIED reference package (NumPy path) from /root/reference or the staged copy baseline/_ref, timed on this box's host
# cores on a bounded sample of the same workload; None when neither tree is present (--no-cpu skips it).
_REF = {}


def ref_ns():
    if "ns" not in _REF:
        _REF["ns"] = None
        try:
            from baseline import refload
            if refload.available():
                ns = dict(refload.ref_extra())
                ns["refload"] = refload
                try:
                    import threadpoolctl
                    threadpoolctl.threadpool_limits(limits=os.cpu_count())
                    ns["cores"] = max([i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()] or [os.cpu_count()])
                except Exception:
                    ns["cores"] = os.cpu_count()
                _REF["ns"] = ns
        except Exception as e:  # noqa: BLE001
            print(f"note: reference not usable ({e})", file=sys.stderr)
    return _REF["ns"]


def cpu_ref(build, reps, sample, scale=1.0, warm=1):
    """build(ns) -> step callable on the reference; returns the cpu_baseline object (ms_per_step scaled by ``scale`` to the
    full workload, the scaling stated in ``sample``)."""
    ns = ref_ns() if ARGS.cpu else None
    if ns is None:
        return None
    try:
        step = build(ns)
        for _ in range(warm):
            step()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = (time.perf_counter() - t0) / reps
        ns["pdn"].autograd.set_grad_enabled(True) if hasattr(ns["pdn"], "autograd") else None
        return {"ms_per_step": dt * scale * 1e3, "kind": "reference", "cores": int(ns["cores"]),
                "sample": f"{sample}; {reps} rep(s) of {dt * 1e3:.1f} ms after {warm} warm-up" + (f", x{scale:g} to the full workload" if scale != 1.0 else "")}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:200]}


def with_ref(extra, sec, base):
    if base and "ms_per_step" in base:
        extra["cpu_baseline"] = base
        extra["speedup_vs_cpu_reference"] = base["ms_per_step"] / (sec * 1e3)
    elif base:
        extra["cpu_baseline"] = base
    return extra


def c1(args):
    rng = np.random.default_rng(0)
    for n in ([512] if args.small else [512, 1024, 2048, 4096, 8192]):
        A, B = rng.standard_normal((n, n)).astype(f32), rng.standard_normal((n, n)).astype(f32)
        x, w = T(A, True), T(B, True)

        def step():
            # new operand values every step as far as the library can tell (write counters bumped): the operand-plane cache only
            # serves re-uses INSIDE a step (x and w in forward and backward), every step packs x, w and g once
            x.data.buf.version += 1; w.data.buf.version += 1
            x.zero_grad(); w.zero_grad()
            pdn.matmul(x, w).sum().backward()

        sec, nl = timed(step, args.steps if n < 8192 else max(2, args.steps // 2))
        base = None
        if n <= 1024:
            def build(ns, A=A, B=B):
                rx, rw = ns["pdn"].Tensor(A, dtype=f32, requires_grad=True), ns["pdn"].Tensor(B, dtype=f32, requires_grad=True)

                def rstep():
                    rx.zero_grad(); rw.zero_grad()
                    ns["pdn"].matmul(rx, rw).sum().backward()
                return rstep
            base = cpu_ref(build, 5, f"the full workload: matmul(x, w).sum().backward() at {n}^3 through the reference's Tensor/autograd")
        emit(f"C1 matmul fwd+bwd {n}^3 fp32 (matmul(x,w).sum().backward())", sec, flops=6.0 * n**3, nbytes=9 * 4.0 * n * n, launches=nl,
             **with_ref({}, sec, base))
        if n <= 1024:  # the same step recorded once into a CUDA graph (pdn.cuda.graphed_step): the per-op Python cost disappears
            gs = pdn.cuda.graphed_step(step)
            sec_g, nl_g = timed(gs, max(args.steps, 50), warmup=5)
            emit(f"C1 matmul fwd+bwd {n}^3 fp32, step recorded with pdn.cuda.graphed_step (CUDA-graph replay)", sec_g, flops=6.0 * n**3,
                 nbytes=9 * 4.0 * n * n, launches=nl_g, **with_ref({}, sec_g, base))
    for n in ([] if args.small else [4096, 8192]):
        a, b = pdn.backend.array(rng.standard_normal((n, n)).astype(f32)), pdn.backend.array(rng.standard_normal((n, n)).astype(f32))
        out = pdn.backend.empty((n, n), f32)
        def fwd():
            a.buf.version += 1; b.buf.version += 1  # fresh operands: both packs are part of every step
            pdn.backend.gemm_into(out, a, b)

        sec, nl = timed(fwd, args.steps)
        emit(f"micro GEMM fwd {n}^3 fp32 (pack + tcgen05 BF16x3)", sec, flops=2.0 * n**3, launches=nl)


def gemm(args):
    """Single large fp32 GEMM (the ncu --set full target for the tcgen05 kernel)."""
    rng = np.random.default_rng(0)
    n = 2048 if args.small else 8192
    a, b = pdn.backend.array(rng.standard_normal((n, n)).astype(f32)), pdn.backend.array(rng.standard_normal((n, n)).astype(f32))
    out = pdn.backend.empty((n, n), f32)
    def fwd():
        a.buf.version += 1; b.buf.version += 1  # fresh operands: both packs are part of every step
        pdn.backend.gemm_into(out, a, b)

    sec, nl = timed(fwd, args.steps)
    emit(f"micro GEMM fwd {n}^3 fp32 (pack + tcgen05 BF16x3)", sec, flops=2.0 * n**3, launches=nl)


def rows(args):
    """HBM-bound row kernels at encoder-like sizes: softmax, feature norm, Adam (ncu --set full targets)."""
    rng = np.random.default_rng(0)
    R, Cn = (4096, 512) if args.small else (65536, 512)
    x = T(rng.standard_normal((R, Cn)).astype(f32))
    with pdn.no_grad():
        sec, nl = timed(lambda: F.softmax(x, axis=-1), args.steps)
    emit(f"micro softmax fwd [{R},{Cn}] fp32", sec, nbytes=8.0 * R * Cn, launches=nl)
    ln = nn.LayerNorm(Cn, dtype=f32).to(DEV)
    ln.train()
    xg = T(rng.standard_normal((R, Cn)).astype(f32), True)

    def lnstep():
        xg.zero_grad(); ln.scale.zero_grad(); ln.shift.zero_grad()
        ln(xg).sum().backward()

    sec, nl = timed(lnstep, args.steps)
    emit(f"micro batch-statistic LayerNorm fwd+bwd [{R},{Cn}] fp32", sec, nbytes=4.0 * R * Cn * (4 + 5), launches=nl)
    try:  # the same step recorded once and replayed as a CUDA graph: the ~16 eager launches cost more host time than GPU time
        gstep = pdn.cuda.graphed_step(lnstep)
        sec, nl = timed(gstep, args.steps)
        emit(f"micro batch-statistic LayerNorm fwd+bwd [{R},{Cn}] fp32, step recorded with pdn.cuda.graphed_step (CUDA-graph replay)", sec,
             nbytes=4.0 * R * Cn * (4 + 5), launches=nl)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"config": "LayerNorm graphed", "error": repr(e)[:200]}), flush=True)
    n = 1 << 20 if args.small else 1 << 26
    p = T(rng.standard_normal(n).astype(f32), True)
    opt = Adam([p], lr=1e-3)
    p.grad[...] = 0.5

    def adam():
        p._grad_stale = False
        opt.step()

    sec, nl = timed(adam, args.steps)
    emit(f"micro Adam step {n} params fp32 (one fused kernel)", sec, nbytes=28.0 * n, launches=nl)


def c2(args):
    from workloads.lenet import ConvNet, train_step
    np.random.seed(42)
    net = ConvNet().to(DEV)
    opt = Adam(net.parameters(), lr=1e-4)
    B = 256
    X, y = T(np.random.rand(B, 1, 28, 28).astype(f32)), T(np.random.randint(0, 10, B))
    net.train()
    sec, nl = timed(lambda: train_step(net, opt, X, y), args.steps)

    def build(ns):
        ConvNet_r = ns["refload"].ref_model("examples/pydynet/mnist.py", lines=(81, 98), extra=ns["refload"].ref_extra())["ConvNet"]
        np.random.seed(42)
        rnet = ConvNet_r()
        ropt = ns["refload"]._REF_PKG["mods"]["pydynet.optim"].Adam(rnet.parameters(), lr=1e-4)
        rX, ry = ns["pdn"].Tensor(X.numpy(), dtype=f32), ns["pdn"].Tensor(y.numpy())
        rnet.train()

        def rstep():
            loss = ns["F"].cross_entropy_loss(rnet(rX), ry)
            ropt.zero_grad(); loss.backward(); ropt.step()
        return rstep
    base = cpu_ref(build, 2, "the full workload: reference ConvNet (examples/pydynet/mnist.py:82-98) batch 256, CE + backward + Adam")
    emit("C2 LeNet b256 train step (CE, backward, Adam)", sec, flops=4_743_290_880, nbytes=170e6, launches=nl, images_per_s=B / sec,
         **with_ref({}, sec, base))
    gs = pdn.cuda.graphed_step(lambda a, b: train_step(net, opt, a, b), optimizers=[opt])
    sec_g, nl_g = timed(lambda: gs(X, y), max(args.steps, 30), warmup=5)
    emit("C2 LeNet b256 train step, recorded with pdn.cuda.graphed_step (CUDA-graph replay)", sec_g, flops=4_743_290_880, nbytes=170e6,
         launches=nl_g, images_per_s=B / sec_g, **with_ref({}, sec_g, base))


def c4(args):
    from workloads.encoder import Transformer, train_step
    np.random.seed(0)
    B, S, V = (16, 128, 1000) if args.small else (128, 512, 8192)
    net = Transformer(512, 1, 8, 3, 0.05, V, S)
    net.word_embedding.reset_parameters()
    net.to(DEV)
    opt = Adam(net.parameters(), lr=5e-4)
    X, y = T(np.random.randint(1, V, (B, S))), T(np.random.choice([-1, 1], B).astype(f32))
    net.train()
    sec, nl = timed(lambda: train_step(net, opt, X, y, None), max(2, args.steps // 2), warmup=2)
    fl = 3 * (2.0 * B * S * 512 * 512 * 4 + 4.0 * B * 8 * S * S * 64 + 2.0 * B * S * 512 * 1536 * 2)
    Bc = 8 if args.small else 16

    def build(ns):
        Tr = ns["refload"].ref_model("examples/pydynet/transformer.py", lines=(52, 192), extra=ns["refload"].ref_extra())["Transformer"]
        np.random.seed(0)
        rnet = Tr(512, 1, 8, 3, 0.05, V, S)
        rnet.word_embedding.reset_parameters()
        ropt = ns["refload"]._REF_PKG["mods"]["pydynet.optim"].Adam(rnet.parameters(), lr=5e-4)
        rX, ry = ns["pdn"].Tensor(X.numpy()[:Bc]), ns["pdn"].Tensor(y.numpy()[:Bc], dtype=f32)
        rnet.train()

        def rstep():
            loss = ns["pdn"].log(1 + ns["pdn"].exp(-ry * ns["pdn"].squeeze(rnet(rX, None)))).mean()
            ropt.zero_grad(); loss.backward(); ropt.step()
        return rstep
    base = cpu_ref(build, 1, f"reference Transformer (examples/pydynet/transformer.py:53-192) at batch {Bc} of {B} (batch {B} needs 22 GB of host RAM "
                             "for its score temporaries); time scaled linearly in the batch", scale=B / Bc, warm=1)
    emit(f"C4 Transformer encoder d512 h8 S{S} B{B} train step", sec, flops=fl, launches=nl, tokens_per_s=B * S / sec, **with_ref({}, sec, base))


def c5(args):
    from workloads.gru import GRURegressor, train_step
    np.random.seed(0)
    B, Tn = (64, 128) if args.small else (256, 1024)
    net = GRURegressor(512, 512).to(DEV)
    opt = Adam(net.parameters(), lr=0.01)
    X, Y = T(np.random.randn(B, Tn, 512).astype(f32)), T(np.random.randn(B, 1).astype(f32))
    sec, nl = timed(lambda: train_step(net, opt, X, Y), max(2, args.steps // 3), warmup=2)
    Tc = 32 if args.small else 64

    def build(ns):
        np.random.seed(0)
        rg = ns["nn"].GRU(512, 512, 1, batch_first=True, dtype=f32)
        rh = ns["nn"].Linear(512, 1, dtype=f32)
        ropt = ns["refload"]._REF_PKG["mods"]["pydynet.optim"].Adam(list(rg.parameters()) + list(rh.parameters()), lr=0.01)
        rX, rY = ns["pdn"].Tensor(X.numpy()[:, :Tc], dtype=f32), ns["pdn"].Tensor(Y.numpy(), dtype=f32)

        def rstep():
            _, h = rg(rX, None)
            loss = ns["F"].mse_loss(rh(h[:, 0, :]), rY)
            ropt.zero_grad(); loss.backward(); ropt.step()
        return rstep
    base = cpu_ref(build, 1, f"reference nn.GRU at T = {Tc} of {Tn} time steps, batch {B} (the reference's step time grows faster than linearly in T: "
                             "92 s at T = 1024 on 8 cores, SURVEY.md 6); time scaled LINEARLY in T, which favours the CPU", scale=Tn / Tc, warm=0)
    emit(f"C5 GRU in512 h512 T{Tn} B{B} train step", sec, flops=2_013_265_920.0 * Tn * B / 256, launches=nl, us_per_time_step=sec / Tn * 1e6,
         **with_ref({}, sec, base))


def lstm(args, cls="LSTM"):
    """The LSTM / plain RNN of the recurrent family at the C5 shape (in512 / h512, batch-first input, loss on the last hidden state)."""
    np.random.seed(0)
    B, Tn = (64, 128) if args.small else (256, 1024)
    rnn, head = getattr(nn, cls)(512, 512, 1, batch_first=True, dtype=f32).to(DEV), nn.Linear(512, 1, dtype=f32).to(DEV)
    opt = Adam(list(rnn.parameters()) + list(head.parameters()), lr=0.01)
    X, Y = T(np.random.randn(B, Tn, 512).astype(f32)), T(np.random.randn(B, 1).astype(f32))

    def step():
        _, h = rnn(X, None)
        h = h[0] if cls == "LSTM" else h
        loss = F.mse_loss(head(h[:, 0, :]), Y)
        opt.zero_grad(); loss.backward(); opt.step()

    sec, nl = timed(step, max(2, args.steps // 3), warmup=2)
    Tc = 32 if args.small else 64

    def build(ns):
        np.random.seed(0)
        rg = getattr(ns["nn"], cls)(512, 512, 1, batch_first=True, dtype=f32)
        rh = ns["nn"].Linear(512, 1, dtype=f32)
        ropt = ns["refload"]._REF_PKG["mods"]["pydynet.optim"].Adam(list(rg.parameters()) + list(rh.parameters()), lr=0.01)
        rX, rY = ns["pdn"].Tensor(X.numpy()[:, :Tc], dtype=f32), ns["pdn"].Tensor(Y.numpy(), dtype=f32)

        def rstep():
            _, h = rg(rX, None)
            h = h[0] if cls == "LSTM" else h
            loss = ns["F"].mse_loss(rh(h[:, 0, :]), rY)
            ropt.zero_grad(); loss.backward(); ropt.step()
        return rstep
    base = cpu_ref(build, 1, f"reference nn.{cls} at T = {Tc} of {Tn} time steps, batch {B}; time scaled LINEARLY in T, which favours the CPU",
                   scale=Tn / Tc, warm=0)
    # per time step: x Wx and h Wh, G*H columns each (G = 4 gates, 1 for the plain RNN), forward + two backward products apiece
    cols = 2048 if cls == "LSTM" else 512
    emit(f"{cls} in512 h512 T{Tn} B{B} train step (persistent whole-sequence kernels)", sec, flops=3 * 2.0 * 2 * 512 * cols * Tn * B, launches=nl,
         us_per_time_step=sec / Tn * 1e6, **with_ref({}, sec, base))


def rnn(args):
    lstm(args, "RNN")


def c3(args):
    """BASELINE config 3 side views: prefill of a 256-token prompt (tokens/s, batch 1) and greedy generation to total
    length 256 at batch 1 / 32 / 128 — the reference's own llm/llama/model.py where it is staged, through the unchanged Module surface
    (the inference plan serves it); the unmodified reference on the host cores beside each line (batch 32 / 128: a bounded sample —
    the 4-token prefill + one decode step at contexts 5 / 130 / 255, as in bench.py's cpu_baseline)."""
    import bench
    Llama, how = bench.model_class()
    rng = np.random.default_rng(0)
    pdn.autograd.set_grad_enabled(False)
    try:
        # ---- prefill(256)
        net, _ = bench.build_model(Llama, 1, DEV)
        ids = rng.integers(1, bench.CFG["V"], (1, 256))
        tid = T(ids)
        sec, nl = timed(lambda: net(tid, 0), max(3, args.steps // 2), warmup=2)
        base = None
        if args.cpu:
            ref, kind = bench._reference_llama(1)
            t0 = time.perf_counter()
            ref(ids, 0)
            dt = time.perf_counter() - t0
            base = {"ms_per_step": dt * 1e3, "kind": kind, "cores": bench._host_threads(), "sample": "one forward over the same 256-token prompt"}
        emit(f"C3 Llama prefill, 256-token prompt, batch 1 ({how})", sec, flops=3.529e9, launches=nl, tokens_per_s=256 / sec, **with_ref({}, sec, base))
        del net
        # ---- greedy generation to total length 256
        for B in ((1, 8) if args.small else (1, 32, 128)):
            net, _ = bench.build_model(Llama, B, DEV)
            prompt = T(np.tile(np.array([[1, 100, 200, 300]]), (B, 1)))
            sec, nl = timed(lambda: bench.generate_resident(net, prompt), 3, warmup=2)
            base = None
            if args.cpu:
                cb = bench.cpu_baseline(B, steps=1, warmup=0, b1_steps=1) if B > 1 else None
                if B == 1:
                    ref, kind = bench._reference_llama(1)
                    t0 = time.perf_counter()
                    for _ in ref.generate(np.array([[1, 100, 200, 300]]), 256):
                        pass
                    dt = time.perf_counter() - t0
                    base = {"ms_per_step": dt * 1e3, "kind": kind, "cores": bench._host_threads(), "sample": "the full workload: batch 1, total length 256"}
                elif cb and cb.get("value"):
                    base = {"ms_per_step": B * 256 / cb["value"] * 1e3, "kind": cb.get("kind"), "cores": cb.get("cores"), "sample": cb.get("sample")}
            emit(f"C3 Llama greedy generation to total length 256, batch {B} ({how})", sec, launches=nl, tokens_per_s=B * 256 / sec,
                 us_per_token_step=sec / 252 * 1e6, **with_ref({}, sec, base))
            del net
    finally:
        pdn.autograd.set_grad_enabled(True)


def decode(args):
    """Decode-step micro-benchmarks at BASELINE config-3 shapes (batch 1024 sequences): the KV-cache attention of one layer at
    a context of 130 keys (HBM-bound: K and V rows once) and the four per-layer GEMMs + lm_head on cached weight planes."""
    from pydynet_b200.nn import _fused
    from pydynet_b200.backend.array import ndarray
    rng = np.random.default_rng(2)
    B, H, D, S = (64 if args.small else 1024), 6, 48, 1024
    for Bv, Lk in ([(B, 130)] if args.small else [(B, 130), (B, 256), (128, 130)]):
        q = T(rng.standard_normal((Bv, 1, H, D)).astype(f32))
        # several cache pairs so that consecutive launches do not find their K/V rows in the 126 MB L2
        n_caches = 1 if args.small else max(2, int(400e6 / (2 * Bv * Lk * H * D * 4)) + 1)
        caches = [(T(rng.standard_normal((Bv, S, H, D)).astype(f32)), T(rng.standard_normal((Bv, S, H, D)).astype(f32))) for _ in range(min(n_caches, 6))]
        it = [0]
        out = ndarray.empty((Bv, 1, H, D), f32)
        qs = _fused._bhl_strides(q.data)
        views = [(ck.data[:, :Lk], cv.data[:, :Lk]) for ck, cv in caches]
        cs = _fused._bhl_strides(views[0][0])

        def att():  # straight through the C ABI: the Python operator wrapper costs more than this kernel
            kv, vv = views[it[0] % len(views)]
            it[0] += 1
            lib.call("pdn_attention_fwd", q.data.ptr, kv.ptr, vv.ptr, None, out.ptr, None, Bv, H, 1, Lk, D, qs, cs, cs, None, 1.0 / D**.5, None, 0)

        os.environ["PDN_ATTN"] = "ffma"
        sec, nl = timed(att, 50)
        os.environ.pop("PDN_ATTN", None)
        emit(f"micro decode attention B{Bv} H{H} hd{D} ctx{Lk} (one layer, KV cache [B,{S},H,D] fp32, {len(caches)} rotating caches)", sec,
             nbytes=2.0 * Bv * Lk * H * D * 4 + 2.0 * Bv * H * D * 4, launches=nl)
    for (K, N) in [(288, 864), (288, 288), (288, 1536), (768, 288), (288, 32000)]:
        w = nn.Parameter(T(rng.standard_normal((K, N)).astype(f32) * 0.05))
        x = T(rng.standard_normal((B, K)).astype(f32))
        pl = _fused.rmsnorm_planes(x, T(np.ones(K, f32)), 1e-5)
        out = ndarray.empty((B, N), f32)
        handle = _fused._packed(w).handle

        def mm():
            lib.call("pdn_gemm_prepacked_planes", pl.ptr, pl.M, pl.Kp, handle, out.ptr, N, None, 0)

        sec, nl = timed(mm, 100)
        emit(f"micro decode GEMM [{B},{K}]x[{K},{N}] on cached planes (back-to-back launches)", sec, flops=2.0 * B * K * N, launches=nl)


def micro_att(args):
    micro(args, conv=False)


def micro(args, conv=True):
    rng = np.random.default_rng(1)
    for (N, Ci, O, HW) in [] if not conv else ([(32, 20, 50, 14)] if args.small else [(256, 20, 50, 14), (128, 64, 128, 56)]):
        conv = nn.Conv2d(Ci, O, 3, 1, 1, dtype=f32).to(DEV)
        x = T(rng.standard_normal((N, Ci, HW, HW)).astype(f32), True)

        def step():
            x.zero_grad(); conv.weight.zero_grad(); conv.bias.zero_grad()
            conv(x).sum().backward()

        sec, nl = timed(step, args.steps)
        Nc = min(N, 16)

        def build(ns, Ci=Ci, O=O, HW=HW, Nc=Nc, xh=x.numpy()):
            rconv = ns["nn"].Conv2d(Ci, O, 3, 1, 1, dtype=f32)
            rx = ns["pdn"].Tensor(xh[:Nc], dtype=f32, requires_grad=True)

            def rstep():
                rx.zero_grad(); rconv.weight.zero_grad(); rconv.bias.zero_grad()
                rconv(rx).sum().backward()
            return rstep
        base = cpu_ref(build, 1, f"reference nn.Conv2d (im2col + add.at) at batch {Nc} of {N}; time scaled linearly in the batch", scale=N / Nc)
        emit(f"micro Conv2d {Ci}->{O} {HW}x{HW} k3p1 b{N} fwd+bwd", sec, flops=3 * 2.0 * N * HW * HW * Ci * 9 * O, launches=nl, **with_ref({}, sec, base))
    B, H, S, D = (8, 8, 128, 64) if args.small else (128, 8, 512, 64)
    q, k, v = (T(rng.standard_normal((B, S, H, D)).astype(f32), True) for _ in range(3))

    from pydynet_b200.nn import _fused

    def att():
        for t in (q, k, v):
            t.zero_grad()
            t.data.buf.version += 1  # fresh q, k, v every step as far as the operand-plane cache can tell
        _fused.attention(q, k, v, None, 1.0 / D**.5).sum().backward()

    sec, nl = timed(att, max(2, args.steps // 2), warmup=2)
    Bc = min(B, 8)

    def build(ns, qh=q.numpy(), kh=k.numpy(), vh=v.numpy()):
        rq, rk, rv = (ns["pdn"].Tensor(a[:Bc], dtype=f32, requires_grad=True) for a in (qh, kh, vh))

        def rstep():
            for t in (rq, rk, rv):
                t.zero_grad()
            sc = rq.transpose(0, 2, 1, 3) @ rk.transpose(0, 2, 3, 1) / D**.5  # the operator chain of transformer.py:93-104
            out = ns["F"].softmax(sc, axis=-1) @ rv.transpose(0, 2, 1, 3)
            out.transpose(0, 2, 1, 3).reshape(Bc, S, -1).sum().backward()
        return rstep
    base = cpu_ref(build, 1, f"the reference's attention operator chain (q kT / sqrt(d) -> softmax -> @ v, transformer.py:93-104) at batch {Bc} of {B}; "
                             "time scaled linearly in the batch", scale=B / Bc)
    emit(f"micro attention core B{B} H{H} S{S} hd{D} fwd+bwd (fused tcgen05 flash kernels, BF16x3)", sec, flops=3 * 4.0 * B * H * S * S * D, launches=nl,
         **with_ref({}, sec, base))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c2,c3,c4,c5,lstm,rnn,micro")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-cpu", dest="cpu", action="store_false", help="skip the reference-on-host-cores run beside every config")
    args = ap.parse_args()
    ARGS = args
    for name in args.only.split(","):
        try:
            {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5, "lstm": lstm, "rnn": rnn, "micro": micro, "micro_att": micro_att, "gemm": gemm, "rows": rows, "decode": decode}[name](args)
        except Exception as e:  # keep going: one config must not hide the others
            import traceback
            traceback.print_exc()
            print(json.dumps({"config": name, "error": repr(e)[:300]}), flush=True)
        pdn.autograd.set_grad_enabled(True)
