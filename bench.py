"""bench.py — BASELINE metric: Llama3-6L (vocab 32000, dim 288, 6 heads, ffn 768) greedy-generation tokens/s.

A "step" is one pass of the hot path over one batch: B prompts (4 tokens each) are prefetched and decoded greedily with
the KV cache until the total length is 256 — the reference's own benchmark loop (llm/llama/infer.py:51-64, metric =
total length / elapsed, prompt tokens included), batched through the model's ``max_batch_size``.  Synthetic N(0, 0.05)
weights of the named architecture (no checkpoints are reachable offline).

  python bench.py [--gpus N --steps K --warmup W]   our arm: the reference's OWN model file (llm/llama/model.py, unmodified, staged under
                                                    baseline/_ref) exec'd with ``pydynet`` aliased to pydynet_b200, on cuda
                                                    (one process per GPU under torchrun)
  python bench.py --impl reference ...              CPU arm: the same file on the unmodified reference package (NumPy) on host cores

Prints ONE JSON line (rank 0): the headline is the batched configuration (``--batch`` sequences per GPU); ``b1`` holds the same
numbers for the reference's own llm/llama/infer.py configuration (max_batch_size = 1), ``token_check`` the comparison of the
generated ids with the CPU oracle (after the timed region), ``dp_train`` the data-parallel training step with the NCCL gradient
all-reduce (SURVEY.md §8 e).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(V=32000, D=288, H=6, FF=768, S=1024, L=6)
METRIC = "llama3_6L_greedy_generation_tokens_per_s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_attention_rows launch from `ncu --set full`, keyed by (batch, keys):
# profiles/r1g_ncu_extract.txt (605.26 MB read + 17.01 MB written at batch 1024, 256 keys; algorithmic 606.3 MB)
ATT_NCU_TRAFFIC = {(1024, 256): 622.27e6}
# the same for ONE k_decode_mega launch at batch 1, context 48 (profiles/r2_ncu_extract.txt: 61.32 MB read + 0.35 MB written; the fp32
# weights are 60.9 MB)
MEGA_NCU_TRAFFIC = 61.67e6
PROMPT_LEN, TOTAL_LEN = 4, int(os.environ.get("PDN_BENCH_TOTAL_LEN", 256))  # the env override exists for short ncu captures only


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- our arm
def model_class():
    """The model definition the product arm runs: the reference's own file, unmodified, where the staged copy travelled with
    the snapshot (baseline/_ref, see baseline/stage_reference.py); the repo's stand-in with the same structure otherwise."""
    try:
        from baseline import refload
        if refload.available():
            return refload.dropin_model("llm/llama/model.py")["Llama"], "reference file llm/llama/model.py, unmodified (pydynet aliased to pydynet_b200)"
    except Exception as e:  # noqa: BLE001
        print(f"note: reference model file not usable ({e}); stand-in definition used", file=sys.stderr)
    from workloads.llama import Llama
    return Llama, "stand-in workloads/llama.py (reference tree not staged)"


def build_model(Llama, B, device):
    from workloads.llama import synthetic_llama_params
    params = synthetic_llama_params(CFG["V"], CFG["D"], CFG["H"], CFG["FF"], CFG["L"], seed=0, std=0.05)
    net = Llama(CFG["V"], CFG["D"], CFG["H"], CFG["FF"], CFG["S"], B, CFG["L"], np.float32).to(device)
    for name, p in net._parameters.items():
        if name in params:
            with p.device:
                p.data[...] = params[name]
    net.eval()
    return net, params


def generate_resident(net, prompt_dev):
    """Device-resident pass: prompt already in HBM, token ids stay on the device (one D2H at the end, outside)."""
    return [t for t in net.generate(prompt_dev, TOTAL_LEN)]


def pinned_copy(lib, arr):
    """A NumPy view of page-locked host memory (pdn_malloc_host) holding a copy of `arr`: the e2e arm's inputs start there."""
    import ctypes as C
    p = C.c_void_p()
    lib.call("pdn_malloc_host", C.byref(p), max(int(arr.nbytes), 1))
    out = np.frombuffer((C.c_byte * arr.nbytes).from_address(p.value), dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def generate_e2e(net, prompt_host):
    """End-to-end pass the way reference llm/llama/infer.py:44-58 drives the model: the NumPy prompt goes straight into
    ``generate`` (host -> device inside), every generated id is read back to the host as it is produced."""
    out = []
    for t in net.generate(prompt_host, TOTAL_LEN):
        out.append(t.numpy())
    return np.concatenate(out, axis=1)


class KernelTimer:
    """CUDA-event bracket around every launch of one entry point inside the timed region (events are recorded on the
    library's compute stream, the stream the kernel is launched on).

    in_graph=False: brackets eager launches. in_graph=True: brackets the launches made while a decode step is being RECORDED
    into a CUDA graph — the two cudaEventRecord calls become event-record nodes of that graph, so every replay re-stamps them
    and, once the pass has finished, the pair holds the device time of the kernel in the LAST replayed decode step (context =
    total length). No host hook runs between kernels of a replay."""

    def __init__(self, lib, entry, predicate, in_graph=False):
        self.lib, self.entry, self.pred, self.pairs, self.on, self.in_graph = lib, entry, predicate, [], False, in_graph
        self.pool = []
        import pydynet_b200.cuda as cuda
        self.cuda = cuda

    def install(self):
        import ctypes as C
        L = self.lib
        orig = L.call
        timer = self

        def call(name, *args):
            if timer.on and name == timer.entry and timer.pred(args) and timer.cuda.is_capturing() == timer.in_graph:
                if timer.pool:
                    e0, e1 = timer.pool.pop()
                else:
                    e0, e1 = C.c_void_p(), C.c_void_p()
                    orig("pdn_event_create", C.byref(e0))
                    orig("pdn_event_create", C.byref(e1))
                orig("pdn_event_record", e0)
                orig(name, *args)
                orig("pdn_event_record", e1)
                timer.pairs.append((e0, e1))
            else:
                orig(name, *args)

        L.call = call  # every binding site resolves lib.call at call time

    def collect(self):
        import ctypes as C
        ms = C.c_float()
        tot, n = 0.0, 0
        for e0, e1 in self.pairs:
            self.lib.call("pdn_event_elapsed_ms", e0, e1, C.byref(ms))
            tot += ms.value
            n += 1
        self.pool.extend(self.pairs)
        self.pairs = []
        return tot, n


def _event_pair(lib):
    import ctypes as C
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    lib.call("pdn_event_create", C.byref(ev0))
    lib.call("pdn_event_create", C.byref(ev1))
    return ev0, ev1


def _elapsed_s(lib, ev0, ev1):
    import ctypes as C
    ms = C.c_float()
    lib.load().pdn_event_elapsed_ms(ev0, ev1, C.byref(ms))
    return max(ms.value / 1e3, 1e-9)


def bench_b1(Llama, how, device, steps, warmup):
    """The reference's own configuration (llm/llama/infer.py:19-37: max_batch_size 1): one prompt decoded greedily to total
    length 256 through the unchanged ``generate`` — rows < 32, so every token is ONE launch of the persistent decode kernel
    (csrc/decode_mega.cu). Returns the numbers and the generated ids (checked against the oracle by the caller)."""
    import pydynet_b200 as pdn
    from pydynet_b200.backend import lib
    net, _ = build_model(Llama, 1, device)
    prompt_host = np.array([[1, 100, 200, 300]])
    prompt_dev = pdn.Tensor(prompt_host, device=device)
    for _ in range(max(warmup, 3)):
        generate_resident(net, prompt_dev)
    pdn.cuda.synchronize()
    ev0, ev1 = _event_pair(lib)
    lib.reset_launch_count()
    lib.call("pdn_event_record", ev0)
    t0 = time.perf_counter()
    for _ in range(steps):
        toks = generate_resident(net, prompt_dev)
    lib.call("pdn_event_record", ev1)
    pdn.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_s = _elapsed_s(lib, ev0, ev1)
    launches = lib.launch_count()
    for _ in range(2):
        generate_e2e(net, prompt_host)
    pdn.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = generate_e2e(net, prompt_host)
    pdn.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ids = np.concatenate([t.numpy() for t in toks], axis=1)
    assert (ids == out).all(), "device-resident and end-to-end passes disagree"
    n_tok = TOTAL_LEN * steps
    # one decode step streams every weight once (fp32) + the cached K and V rows of the context: SURVEY.md §8(d) C3
    w_bytes = 4.0 * (CFG["L"] * (4 * CFG["D"] ** 2 + 3 * CFG["D"] * CFG["FF"] + 2 * CFG["D"]) + CFG["D"] * CFG["V"] + CFG["V"] + CFG["D"])
    kv_bytes_mean = 2.0 * 4 * CFG["D"] * CFG["L"] * (TOTAL_LEN / 2)
    hbm, _, which = _peaks()
    us_tok = dev_s / (steps * (TOTAL_LEN - PROMPT_LEN)) * 1e6
    return {"workload": "llm/llama/infer.py's own configuration: max_batch_size 1, prompt 4 tokens, greedy to total length 256",
            "model_class": how, "value": n_tok / dev_s, "unit": "tokens/s", "ms_per_step": dev_s / steps * 1e3,
            "wall_ms_per_step": wall / steps * 1e3,
            "e2e": {"value": n_tok / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": int(prompt_host.nbytes),
                    "d2h_bytes_per_step": int((TOTAL_LEN - PROMPT_LEN) * 8),
                    "note": "infer.py loop: NumPy prompt in, every id read back with .numpy() as it is produced (one stream sync per token)"},
            "gpu_launches": int(launches), "us_per_token_device": us_tok,
            "roofline": {"kernel": "k_decode_mega (whole decode step, one cooperative launch per token)", "bound": "hbm",
                         "achieved": (w_bytes + kv_bytes_mean) / (us_tok * 1e-6) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": (w_bytes + kv_bytes_mean) / (us_tok * 1e-6) / 1e9 / hbm, "traffic": MEGA_NCU_TRAFFIC, "peak_source": which,
                         "note": "algorithmic bytes per token = every fp32 weight once (60.9 MB) + K and V rows of the mean context; at one "
                                 "row the step is a chain of dependent grid-wide phases (latency-bound), not bandwidth-bound"}}, ids, prompt_host


def run_ours(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only: barrier + max of timings
    import ctypes as C
    import pydynet_b200 as pdn
    from pydynet_b200.backend import lib
    device = f"cuda:{local}"
    B = args.batch
    Llama, how = model_class()
    net, params = build_model(Llama, B, device)
    rng = np.random.default_rng(100 + rank)
    prompt_host = rng.integers(1, CFG["V"], (B, PROMPT_LEN))
    prompt_dev = pdn.Tensor(prompt_host, device=device)

    def barrier():
        pdn.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # Dominant kernel of a decode step at this batch (profiles/r1d_launches_b1024.csv): the KV-cache attention, one launch per
    # layer, HBM-bound (every cached K and V row of the batch is read once per layer and step). Decode steps 2.. are CUDA-graph
    # replays, so its CUDA-event bracket is recorded INTO the graph (KernelTimer, in_graph=True) and read after each pass: the
    # samples are layer 0's launch in the last decode step of every timed pass (context = TOTAL_LEN keys).
    seen = [0]

    from pydynet_b200.nn._plans import decode_branches
    NB = decode_branches(B)  # concurrent batch slices of the recorded decode step (layers are issued round-robin over them)
    BS = B // NB

    def first_layer_only(a):  # layer 0's launches of each recorded decode step (one per batch slice), not all 6 * NB
        seen[0] += 1
        return (seen[0] - 1) % (CFG["L"] * NB) < NB

    att_timer = KernelTimer(lib, "pdn_attention_fwd_dev", first_layer_only if os.environ.get("PDN_BENCH_ATT_ALL") is None else (lambda a: True), in_graph=True)
    att_timer.install()
    # Second view: the longest launch of the GEMM family, lm_head [B,288]x[288,32000] with the argmax epilogue, bracketed the same
    # way inside the recorded decode step.
    timer = KernelTimer(lib, "pdn_gemm_prepacked_planes_argmax", lambda a: int(a[1]) == BS, in_graph=True)
    timer.install()
    with pdn.no_grad():
        for _ in range(max(args.warmup, 3)):
            generate_resident(net, prompt_dev)
        barrier()
        # the decode graphs were recorded during warm-up: record them once more with the timers' event-record nodes inside (one pair
        # per timer and graph; there are two graphs, one per ping-pong slot). Every replay re-stamps the pairs, so after the timed
        # region they hold the device times of the kernels in the LAST TWO decode steps of the last timed pass.
        plan = net.__dict__.get("_pdn_plan")
        if plan:
            plan._drop_recorded()
        timer.on = True
        att_timer.on = os.environ.get("PDN_BENCH_NO_ATT_TIMER") is None
        generate_resident(net, prompt_dev)
        generate_resident(net, prompt_dev)
        timer.on = att_timer.on = False  # nothing is recorded any more; the pairs live in the graphs
        barrier()
        ev0, ev1 = _event_pair(lib)
        lib.reset_launch_count()
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            lib.call("pdn_event_record", ev0)
            for _ in range(args.steps):
                toks = generate_resident(net, prompt_dev)
            lib.call("pdn_event_record", ev1)
            barrier()
            wall = time.perf_counter() - t0
        dev_s = _elapsed_s(lib, ev0, ev1)
        launches = lib.launch_count()
        k_ms, k_n = timer.collect()
        a_ms, a_n = att_timer.collect()
        # end-to-end arm: prompt in pinned host memory -> device, every id read back to the host (reference infer.py loop)
        prompt_pinned = pinned_copy(lib, prompt_host)
        for _ in range(2):
            generate_e2e(net, prompt_pinned)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = generate_e2e(net, prompt_pinned)
        barrier()
        e2e_s = time.perf_counter() - t0
        ids = np.concatenate([t.numpy() for t in toks], axis=1)
        # split-K partial tiles are combined with floating-point atomics (order varies run to run at the 1e-7 level), so a sequence
        # may leave its twin at an fp32 near-tie of the top two logits; reported, and judged by the margin rule in token_check
        same_seq = int((ids == out).all(axis=1).sum())
        served = bool(plan and not plan.dead and plan.verified)
        b1 = b1_ids = b1_prompt = None
        if rank == 0 and args.b1:
            try:
                b1, b1_ids, b1_prompt = bench_b1(Llama, how, device, args.steps, args.warmup)
            except Exception as e:  # noqa: BLE001  (side view: the headline line is still printed)
                b1, b1_ids, b1_prompt = {"error": repr(e)[:300]}, None, None
    pdn.autograd.set_grad_enabled(True)
    dp = dp_train_bench(args, rank, world, local, dist) if args.dp_train else None
    if dist is not None:
        import torch
        t = torch.tensor([dev_s, e2e_s, wall], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, wall = (float(v) for v in t)
        lt = torch.tensor([launches], dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    tokens = world * B * TOTAL_LEN * args.steps
    hbm, tf, which = _peaks()
    # KV-cache attention of one layer in the last decode step: algorithmic bytes per launch = every cached K and V row of the
    # batch once (Lk = TOTAL_LEN keys x H*D fp32) + the query rows in + the output operand planes out (bf16 hi/lo = 4 B/element)
    HD = CFG["D"]
    att_bytes = 2.0 * BS * TOTAL_LEN * HD * 4 + BS * HD * 4 + BS * HD * 4
    a_avg_s = (a_ms / a_n) / 1e3 if a_n else float("nan")
    # lm_head GEMM fused with the greedy argmax: A [B,288] + W [288,32000] + bias (fp32-sized operands, 4 B/element as bf16 hi+lo
    # planes) + B int64 ids out; the [B,32000] logits never touch HBM
    alg_bytes = 4.0 * (BS * CFG["D"] + CFG["D"] * CFG["V"] + CFG["V"]) + 8.0 * BS
    alg_flops = 2.0 * BS * CFG["D"] * CFG["V"]
    k_avg_s = (k_ms / k_n) / 1e3 if k_n else float("nan")
    res = {
        "metric": METRIC, "value": tokens / dev_s, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world),
        "model_class": how, "served_by_inference_plan": served,
        "resident_vs_e2e": {"identical_sequences": same_seq, "of": B},
        "e2e": {"value": tokens / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": int(prompt_host.nbytes),
                "d2h_bytes_per_step": int(B * (TOTAL_LEN - PROMPT_LEN) * 8)},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall / args.steps * 1e3,
        "roofline": {"kernel": "k_attention_rows<16,4,4,8> (KV-cache decode attention, one launch per layer: q [B,1,6,48] over cache[:, :Lk] of "
                               "[B,1024,6,48] fp32, output as GEMM operand planes) - the top kernel of the decode step at this batch "
                               "(profiles/r1g_launches_b1024.csv)",
                     "bound": "hbm", "achieved": att_bytes / max(a_avg_s, 1e-12) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": att_bytes / max(a_avg_s, 1e-12) / 1e9 / hbm, "traffic": ATT_NCU_TRAFFIC.get((BS, TOTAL_LEN)), "peak_source": which,
                     "launch_us": a_avg_s * 1e6, "launches_timed": a_n, "batch_slices": NB, "sequences_per_launch": BS,
                     "note": "achieved = algorithmic bytes (K and V rows of the batch once at Lk = total length, + q in + planes out) / CUDA-event "
                             "time of the launch, bracketed by event-record nodes inside the replayed CUDA graphs (layer 0 of the last two decode "
                             "steps of the timed region); traffic = ncu dram bytes of one launch at the same shape (profiles/r1g_ncu_extract.txt), "
                             "null if that shape was not captured",
                     "gemm_view": {"kernel": "k_gemm_tc<256> with argmax epilogue (lm_head [B,288]x[288,32000] on cached bf16 hi/lo weight planes, "
                                             "tcgen05 BF16x3) - the longest launch of the GEMM family", "bound": "tensor",
                                   "achieved": alg_flops / max(k_avg_s, 1e-12) / 1e12, "peak": tf, "unit": "TFLOP/s",
                                   "frac": alg_flops / max(k_avg_s, 1e-12) / 1e12 / tf,
                                   "frac_of_bf16x3_ceiling": 3 * alg_flops / max(k_avg_s, 1e-12) / 1e12 / tf,
                                   "note": "algorithmic fp32 FLOPs (2*B*288*32000) / CUDA-event time; every product costs 3 BF16 MMAs (fp32 parity), "
                                           "so the MMA-issue fraction is 3x frac",
                                   "hbm_gbs_alg": alg_bytes / max(k_avg_s, 1e-12) / 1e9, "launch_us": k_avg_s * 1e6, "launches_timed": k_n}},
        "clocks": clk.summary(),
    }
    if rank == 0:
        if b1 is not None:
            res["b1"] = b1
        # rank-0-only side sections (no collectives inside): a failure there is reported in its field, the line is still printed
        if args.token_check:
            try:
                res["token_check"] = verify_tokens(params, prompt_host, ids, b1_prompt, b1_ids)
            except Exception as e:  # noqa: BLE001
                res["token_check"] = {"ok": False, "error": repr(e)[:300]}
        if dp is not None:
            res["dp_train"] = dp
        if args.cpu_baseline and world == 1:
            try:
                res["cpu_baseline"] = cpu_baseline(B, steps=1, warmup=1, b1_steps=1 if args.b1 else 0)
            except Exception as e:  # noqa: BLE001
                res["cpu_baseline"] = {"error": repr(e)[:300]}
        print(json.dumps(res))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


C4 = dict(E=512, H=8, S=512, V=8192, FFX=3, B=128)


def dp_train_bench(args, rank, world, local, dist):
    """BASELINE configs[3]: Transformer encoder (d_model 512, 8 heads, seq 512, 1 layer, ffn 1536) train step — forward, loss,
    backward, fused Adam — data-parallel over the ``world`` GPUs: batch sharded, ONE exchange step per iteration = NCCL all-reduce
    of the flat fp32 gradient buffer (6.8 M floats) in 4 ranges queued on a side stream WHILE backward is still running
    (pydynet_b200/distributed.py), cross-rank batch statistics in the batch-coupled norms. Weak scaling (batch 128 per GPU) is
    the reported line; the strong-scaling time (global batch 128) and a W-rank-vs-1-GPU parity check ride along."""
    import pydynet_b200 as pdn
    from pydynet_b200 import distributed as pd
    from pydynet_b200.backend import lib
    from pydynet_b200.optim import Adam
    how = "stand-in workloads/encoder.py"
    Transformer = None
    try:
        from baseline import refload
        if refload.available():
            Transformer = refload.dropin_model("examples/pydynet/transformer.py", lines=(52, 192), extra=refload.dropin_extra())["Transformer"]
            how = "reference file examples/pydynet/transformer.py:53-192, unmodified (pydynet aliased to pydynet_b200)"
    except Exception as e:  # noqa: BLE001
        print(f"note: reference transformer.py not usable ({e}); stand-in used", file=sys.stderr)
    if Transformer is None:
        from workloads.encoder import Transformer
    device = f"cuda:{local}"
    out = {"workload": "configs[3]: Transformer encoder d_model 512, 8 heads, seq 512, 1 layer, ffn 1536, vocab 8192 - train step "
                       "(forward, logistic loss, backward, Adam), data-parallel with NCCL gradient all-reduce",
           "model_class": how, "dtype": "f32 (tcgen05 BF16x3 GEMMs / attention)", "data": "synthetic"}
    pdn.autograd.set_grad_enabled(True)
    if world > 1:
        if os.environ.get("PDN_NCCL_QUIET") is None:  # the communicator's own init lines (nranks, transports) - on STDERR: stdout
            os.environ["NCCL_DEBUG"] = "INFO"         # carries the one JSON line only
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        pd.init_process_group("nccl", rank, world)
        pd.sync_batch_stats(True)
        import ctypes as C
        c_rank, c_world = C.c_int(-1), C.c_int(-1)
        lib.call("pdn_nccl_world", C.byref(c_rank), C.byref(c_world))
        out["nccl_comm"] = {"ncclCommCount": int(c_world.value), "ncclCommUserRank": int(c_rank.value)}
    E, S, V = C4["E"], C4["S"], C4["V"]
    fl_per_sample = 3 * (2.0 * S * E * E * 4 + 4.0 * C4["H"] * S * S * (E // C4["H"]) + 2.0 * S * E * E * C4["FFX"] * 2)

    def run(per_gpu, steps):
        np.random.seed(0)
        net = Transformer(E, 1, C4["H"], C4["FFX"], 0.05, V, S)
        net.word_embedding.reset_parameters()
        net.to(device)
        opt = Adam(net.parameters(), lr=5e-4)
        ddp = pd.DataParallel(net, opt, buckets=4, overlap=True)
        rng = np.random.default_rng(10 + rank)
        X = pdn.Tensor(rng.integers(1, V, (per_gpu, S)), device=device)
        y = pdn.Tensor(rng.choice([-1, 1], per_gpu).astype(np.float32), device=device)
        net.train()
        losses = []

        def step():
            loss = pdn.log(1 + pdn.exp(-y * pdn.squeeze(net(X, None)))).mean()  # reference transformer.py:244-245, unit weights
            opt.zero_grad()
            loss.backward()
            ddp.step()
            return loss

        for _ in range(3):
            losses.append(step())
        pdn.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev0, ev1 = _event_pair(lib)
        lib.reset_launch_count()
        lib.call("pdn_event_record", ev0)
        for _ in range(steps):
            losses.append(step())
        lib.call("pdn_event_record", ev1)
        pdn.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sec = _elapsed_s(lib, ev0, ev1) / steps
        nl = lib.launch_count() // steps
        if dist is not None:
            import torch
            t = torch.tensor([sec], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t[0])
        attn_plan = net.layers[0].attention.__dict__.get("_pdn_plan")
        return {"ms_per_step": sec * 1e3, "launches_per_step_per_gpu": int(nl), "loss_first": float(losses[0].item()),
                "loss_last": float(losses[-1].item()), "bucket_floats": int(ddp._flat.total) if ddp._flat is not None else None,
                "buckets": len(ddp._ranges), "queued_during_backward": sum(1 for (_, sw) in ddp.launch_log if sw == 0),
                "fused_attention_plan": bool(attn_plan and attn_plan.verified)}, sec

    steps = max(3, min(args.steps, 20))
    weak, sec = run(C4["B"], steps)
    _, tf, _ = _peaks()
    weak.update({"per_gpu_batch": C4["B"], "global_batch": C4["B"] * world, "scaling": "weak",
                 "tokens_per_s": world * C4["B"] * S / sec, "samples_per_s": world * C4["B"] / sec,
                 "tflops_alg": world * C4["B"] * fl_per_sample / sec / 1e12,
                 "frac_of_bf16x3_ceiling_per_gpu": 3 * C4["B"] * fl_per_sample / sec / 1e12 / tf})
    out["weak"] = weak
    if world > 1 and C4["B"] % world == 0:
        strong, sec_s = run(C4["B"] // world, steps)
        strong.update({"per_gpu_batch": C4["B"] // world, "global_batch": C4["B"], "scaling": "strong", "tokens_per_s": C4["B"] * S / sec_s})
        out["strong"] = strong
    out["nccl"] = {"ranks": out.get("nccl_comm", {}).get("ncclCommCount", 1), "collective": f"ncclAllReduce(sum, fp32) x {weak['buckets']} ranges of the flat gradient buffer per step on a side stream "
                                                 "(own communicator), 1/world folded into the fused Adam kernel" if world > 1 else "none (1 GPU)",
                   "bytes_per_step": int(4 * (weak["bucket_floats"] or 0)) if world > 1 else 0}
    if world > 1:
        try:
            from tools.dp_check import parity
            out["parity_vs_single_gpu"] = parity(device, rank, world)
        except Exception as e:  # noqa: BLE001
            out["parity_vs_single_gpu"] = {"ok": False, "error": repr(e)[:300]}
        pd.sync_batch_stats(False)
        pd.destroy_process_group()
    return out


def workload_config(B, world):
    return {"workload": "configs[2]: llm/llama 6-layer Llama3 inference, total length 256 (4 prompt + 252 greedy decode steps), "
                        f"batch {B} sequences per GPU, vocab 32000 dim 288 heads 6 ffn 768, KV cache",
            "batch_per_gpu": B, "seq_len": TOTAL_LEN, "parallelism": f"replicas x{world} (no data-path collective)",
            "l2_policy": "inputs larger than L2: per step the KV cache (B*14.2 MB) + weights (97.7 MB) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------- CPU arm + checker
def _host_threads():
    """The reference arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which OpenBLAS would obey."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        return max([i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()] or [n])
    except Exception:
        return n


def _reference_llama(B):
    """(model, kind): the UNMODIFIED reference package + its own llm/llama/model.py from baseline/_ref (kind 'reference'), or
    the NumPy restatement oracle/pdn_oracle.py (kind 'port') where the reference tree is not staged."""
    from oracle.pdn_oracle import synthetic_llama_params, LlamaOracle
    params = synthetic_llama_params(CFG["V"], CFG["D"], CFG["H"], CFG["FF"], CFG["L"], seed=0, std=0.05)
    try:
        from baseline import refload
        if refload.available():
            Llama = refload.ref_model("llm/llama/model.py")["Llama"]
            net = Llama(CFG["V"], CFG["D"], CFG["H"], CFG["FF"], CFG["S"], B, CFG["L"], np.float32)
            for name, p in net._parameters.items():
                if name in params:
                    p.data[...] = params[name]
            net.eval()  # reference module.py:45-59: also switches autograd off
            return net, "reference"
    except Exception as e:  # noqa: BLE001
        print(f"note: staged reference not usable ({e}); oracle port timed instead", file=sys.stderr)

    class Port:
        def __init__(self):
            self.m = LlamaOracle(params, CFG["H"], CFG["S"], B, CFG["L"])
            self.layers = []

        def __call__(self, ids, pos):
            return self.m.step(np.asarray(ids), pos)

        def generate(self, ids, total):
            return iter(self.m.generate(np.asarray(ids), total).T[:, :, None])

    return Port(), "port"


DECODE_SAMPLE_POS = (5, 130, 255)


def _batched_sample(net, kind, B, prompt):
    """One bounded sample of the batched workload on the host cores: the 4-token prefill + ONE decode step at each of the
    contexts 5 / 130 / 255 (cost is linear in the context, so their mean is the mean decode step of the pass), on the full
    batch. Returns (seconds spent, estimated seconds of the whole 256-token pass)."""
    t0 = time.perf_counter()
    logits = net(prompt, 0)
    t_pre = time.perf_counter() - t0
    nxt = (logits if isinstance(logits, np.ndarray) else logits.data)[:, -1, :].argmax(-1)[:, None]
    t_dec = []
    for pos in DECODE_SAMPLE_POS:
        t0 = time.perf_counter()
        net(nxt, pos)
        t_dec.append(time.perf_counter() - t0)
    spent = t_pre + sum(t_dec)
    return spent, t_pre + (TOTAL_LEN - PROMPT_LEN) * float(np.mean(t_dec))


def _fill_caches(net, kind):
    """Decode steps at contexts 130 / 255 attend over cache rows a bounded sample never computed: fill them with N(0, 1)
    (values do not change the arithmetic cost)."""
    rng = np.random.default_rng(7)
    if kind == "reference":
        arrays = [getattr(layer.attention, nm).data for layer in net.layers for nm in ("cache_k", "cache_v")]
    else:
        arrays = list(net.m.ck) + list(net.m.cv)
    block = None
    for arr in arrays:  # [B, S, H, hd]: one random [1, 256, H, hd] block broadcast over the batch (14.5 GB of caches at batch 1024)
        if block is None or block.shape[1:] != (TOTAL_LEN, ) + arr.shape[2:]:
            block = rng.standard_normal((1, TOTAL_LEN) + arr.shape[2:]).astype(arr.dtype)
        arr[:, :TOTAL_LEN] = block


def cpu_baseline(B, steps=1, warmup=1, b1_steps=1):
    """The reference's NumPy path on the host cores, on a BOUNDED sample of the product arm's workload (see _batched_sample);
    ``b1``: the reference's own configuration (max_batch_size 1) run in full."""
    threads = _host_threads()
    net, kind = _reference_llama(B)
    prompt = np.random.default_rng(100).integers(1, CFG["V"], (B, PROMPT_LEN))
    try:
        _fill_caches(net, kind)
    except Exception as e:  # noqa: BLE001
        print(f"note: KV caches left at zero ({e})", file=sys.stderr)
    for _ in range(warmup):
        _batched_sample(net, kind, B, prompt)
    spent = est = 0.0
    for _ in range(steps):
        s, e = _batched_sample(net, kind, B, prompt)
        spent += s
        est += e
    out = {"value": B * TOTAL_LEN * steps / est, "unit": "tokens/s", "cores": int(threads), "kind": kind,
           "sample": f"batch {B}: per step the 4-token prefill + one decode step at each context {DECODE_SAMPLE_POS} on the FULL batch, timed "
                     f"with perf_counter; tokens/s = batch*256 / (t_prefill + 252 * mean(t_decode)); {steps} step(s) after {warmup} warm-up, "
                     f"{spent:.1f} s of NumPy/OpenBLAS work, os.cpu_count()={os.cpu_count()}",
           "sample_s": spent}
    if b1_steps:
        net1, kind1 = _reference_llama(1)
        p1 = np.array([[1, 100, 200, 300]])
        t0 = time.perf_counter()
        for _ in range(b1_steps):
            n = PROMPT_LEN
            for _t in net1.generate(p1, TOTAL_LEN):
                n += 1
        dt = time.perf_counter() - t0
        out["b1"] = {"value": n * b1_steps / dt, "unit": "tokens/s", "kind": kind1, "cores": int(threads),
                     "sample": f"the full workload: batch 1, total length 256, {b1_steps} pass(es), {dt:.1f} s"}
    return out


def verify_tokens(params, prompt, ids, b1_prompt, b1_ids, n_seq=8, n_tok=64):
    """CHECKER (outside every timed region): the ids the timed passes produced against oracle/pdn_oracle.py under the margin
    rule of SURVEY.md §8(c) — exact up to the first step where the oracle's own top-1/top-2 margin is below 1e-4."""
    from oracle import pdn_oracle as O
    out = {"rule": "exact up to the first step whose oracle top-1/top-2 logit margin is < 1e-4 (relative to max |logit|)"}
    try:
        n_seq = min(n_seq, prompt.shape[0])
        ref, mar = O.LlamaOracle(params, CFG["H"], CFG["S"], n_seq, CFG["L"]).generate_with_margins(prompt[:n_seq], PROMPT_LEN + n_tok)
        exact, near = O.check_greedy_tokens(ids[:n_seq, :n_tok], ref, mar)
        out["batched"] = {"sequences": n_seq, "tokens_each": n_tok, "exact": exact, "diverged_at_near_tie": near, "ok": True}
        if b1_ids is not None:
            ref, mar = O.LlamaOracle(params, CFG["H"], CFG["S"], 1, CFG["L"]).generate_with_margins(b1_prompt, TOTAL_LEN)
            exact, near = O.check_greedy_tokens(b1_ids, ref, mar)
            out["b1"] = {"sequences": 1, "tokens_each": int(b1_ids.shape[1]), "exact": exact, "diverged_at_near_tie": near, "ok": True}
    except AssertionError as e:
        out["ok"] = False
        out["error"] = str(e)[:300]
        return out
    out["ok"] = True
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    B = args.batch
    t0 = time.perf_counter()
    base = cpu_baseline(B, steps=args.steps, warmup=args.warmup, b1_steps=min(args.steps, 3) if args.b1 else 0)
    world = int(os.environ.get("WORLD_SIZE", 1))
    res = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["sample_s"] / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(B, world),
           "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    if "b1" in base:
        res["b1"] = base["b1"]
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PDN_BENCH_BATCH", 1024)))
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-b1", dest="b1", action="store_false")
    ap.add_argument("--no-token-check", dest="token_check", action="store_false")
    ap.add_argument("--no-dp-train", dest="dp_train", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
