"""bench.py — BASELINE metric: Llama3-6L (vocab 32000, dim 288, 6 heads, ffn 768) greedy-generation tokens/s.

A "step" is one pass of the hot path over one batch: B prompts (4 tokens each) are prefetched and decoded greedily with
the KV cache until the total length is 256 — the reference's own benchmark loop (llm/llama/infer.py:51-64, metric =
total length / elapsed, prompt tokens included), batched through the model's ``max_batch_size``.  Synthetic N(0, 0.05)
weights of the named architecture (no checkpoints are reachable offline).

  python bench.py [--gpus N --steps K --warmup W]         our arm: pydynet_b200 on cuda (one process per GPU under torchrun)
  python bench.py --impl reference ...                    CPU arm: the reference algorithm (oracle port, NumPy) on host cores

Prints ONE JSON line (rank 0). See DESIGN.md §Measurement for every field.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_attention_rows launch from `ncu --set full`, keyed by (batch, keys):
# profiles/r1g_ncu_extract.txt (605.26 MB read + 17.01 MB written at batch 1024, 256 keys; algorithmic 606.3 MB)
ATT_NCU_TRAFFIC = {(1024, 256): 622.27e6}
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
def build_model(B, device):
    import pydynet_b200 as pdn
    from workloads.llama import Llama, synthetic_llama_params
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


def generate_e2e(net, prompt_host, device, pinned):
    """End-to-end pass through the public API the way reference llm/llama/infer.py:44-58 drives it: host prompt -> device,
    every generated id read back to the host as it is produced."""
    import pydynet_b200 as pdn
    ids = pdn.Tensor(prompt_host, device=device)
    out = []
    for t in net.generate(ids, TOTAL_LEN):
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
    net, params = build_model(B, device)
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
    # samples are the 6 layers' launches of the last decode step of every timed pass (context = TOTAL_LEN keys).
    seen = [0]

    def first_layer_only(a):  # layer 0's launch of each recorded decode step: two event nodes per graph, not twelve
        seen[0] += 1
        return seen[0] % CFG["L"] == 1

    att_timer = KernelTimer(lib, "pdn_attention_fwd_dev", first_layer_only if os.environ.get("PDN_BENCH_ATT_ALL") is None else (lambda a: True), in_graph=True)
    att_timer.install()
    # Second view: the longest launch of the GEMM family, lm_head [B,288]x[288,32000] with the argmax epilogue, bracketed where it
    # is launched eagerly inside the timed region (the first decode step of every pass: same kernel, same shapes, same stream).
    timer = KernelTimer(lib, "pdn_gemm_prepacked_planes_argmax", lambda a: int(a[1]) == B)
    timer.install()
    with pdn.no_grad():
        for _ in range(max(args.warmup, 3)):
            generate_resident(net, prompt_dev)
        barrier()
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        lib.call("pdn_event_create", C.byref(ev0))
        lib.call("pdn_event_create", C.byref(ev1))
        lib.reset_launch_count()
        timer.on = True
        att_timer.on = os.environ.get("PDN_BENCH_NO_ATT_TIMER") is None
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            lib.call("pdn_event_record", ev0)
            for _ in range(args.steps):
                toks = generate_resident(net, prompt_dev)
            lib.call("pdn_event_record", ev1)
            barrier()
            wall = time.perf_counter() - t0
        timer.on = att_timer.on = False
        ms = C.c_float()
        lib.load().pdn_event_elapsed_ms(ev0, ev1, C.byref(ms))
        launches = lib.launch_count()
        k_ms, k_n = timer.collect()
        a_ms, a_n = att_timer.collect()
        dev_s = max(ms.value / 1e3, 1e-9)
        # end-to-end arm: prompt in pinned host memory -> device, every id read back to the host (reference infer.py loop)
        prompt_host = pinned_copy(lib, prompt_host)
        for _ in range(2):
            generate_e2e(net, prompt_host, device, None)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = generate_e2e(net, prompt_host, device, None)
        barrier()
        e2e_s = time.perf_counter() - t0
    pdn.autograd.set_grad_enabled(True)
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
    att_bytes = 2.0 * B * TOTAL_LEN * HD * 4 + B * HD * 4 + B * HD * 4
    a_avg_s = (a_ms / a_n) / 1e3 if a_n else float("nan")
    # lm_head GEMM fused with the greedy argmax: A [B,288] + W [288,32000] + bias (fp32-sized operands, 4 B/element as bf16 hi+lo
    # planes) + B int64 ids out; the [B,32000] logits never touch HBM
    alg_bytes = 4.0 * (B * CFG["D"] + CFG["D"] * CFG["V"] + CFG["V"]) + 8.0 * B
    alg_flops = 2.0 * B * CFG["D"] * CFG["V"]
    k_avg_s = (k_ms / k_n) / 1e3 if k_n else float("nan")
    res = {
        "metric": "llama3_6L_greedy_generation_tokens_per_s", "value": tokens / dev_s, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: llm/llama 6-layer Llama3 inference, total length 256 (4 prompt + 252 greedy decode steps), "
                               f"batch {B} sequences per GPU, vocab 32000 dim 288 heads 6 ffn 768, KV cache",
                   "batch_per_gpu": B, "seq_len": TOTAL_LEN, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2_policy": "inputs larger than L2: per step the KV cache (B*14.2 MB) + weights (97.7 MB) exceed the 126 MB L2"},
        "e2e": {"value": tokens / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": int(prompt_host.nbytes),
                "d2h_bytes_per_step": int(B * (TOTAL_LEN - PROMPT_LEN) * 8)},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall / args.steps * 1e3,
        "roofline": {"kernel": "k_attention_rows<16,4,4,8> (KV-cache decode attention, one launch per layer: q [B,1,6,48] over cache[:, :Lk] of "
                               "[B,1024,6,48] fp32, output as GEMM operand planes) - the top kernel of the decode step at this batch "
                               "(profiles/r1g_launches_b1024.csv)",
                     "bound": "hbm", "achieved": att_bytes / max(a_avg_s, 1e-12) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": att_bytes / max(a_avg_s, 1e-12) / 1e9 / hbm, "traffic": ATT_NCU_TRAFFIC.get((B, TOTAL_LEN)), "peak_source": which,
                     "launch_us": a_avg_s * 1e6, "launches_timed": a_n,
                     "note": "achieved = algorithmic bytes (K and V rows of the batch once at Lk = total length, + q in + planes out) / CUDA-event "
                             "time of the launch, bracketed by event-record nodes inside the replayed CUDA graph (layer 0 of the last decode step "
                             "of every timed pass); traffic = ncu dram bytes of one launch at the same shape (profiles/r1g_ncu_extract.txt), "
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
        if args.cpu_baseline and world == 1:
            res["cpu_baseline"] = cpu_baseline(params, args.cpu_batch, args.cpu_total_len)
        print(json.dumps(res))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_baseline(params, B, total_len, steps=1):
    """The reference algorithm (oracle port: NumPy restatement of llm/llama/model.py, parity-pinned by
    tests/test_oracle.py) on the host cores, on a bounded sample of the same workload."""
    from oracle.pdn_oracle import LlamaOracle
    try:
        from threadpoolctl import threadpool_info
        threads = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count()
    rng = np.random.default_rng(100)
    prompt = rng.integers(1, CFG["V"], (B, PROMPT_LEN))
    t0 = time.perf_counter()
    for _ in range(steps):
        m = LlamaOracle(params, CFG["H"], CFG["S"], B, CFG["L"])
        m.generate(prompt, total_len)
    dt = time.perf_counter() - t0
    return {"value": B * total_len * steps / dt, "unit": "tokens/s", "cores": int(threads), "kind": "port",
            "sample": f"batch {B}, total length {total_len} (4 prompt + {total_len - PROMPT_LEN} decode steps) x {steps} pass(es), "
                      f"{dt:.1f} s of NumPy/OpenBLAS work, os.cpu_count()={os.cpu_count()}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle.pdn_oracle import synthetic_llama_params
    params = synthetic_llama_params(CFG["V"], CFG["D"], CFG["H"], CFG["FF"], CFG["L"], seed=0, std=0.05)
    B, total = args.cpu_batch, args.cpu_total_len
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(params, B, min(total, 8))
    t0 = time.perf_counter()
    base = cpu_baseline(params, B, total, steps=args.steps)
    dt = time.perf_counter() - t0
    res = {"impl": "reference", "metric": "llama3_6L_greedy_generation_tokens_per_s", "value": base["value"], "unit": "tokens/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[2]: llm/llama 6-layer Llama3 inference on host cores (oracle port of the reference NumPy path), "
                                  f"bounded sample: batch {B}, total length {total}"},
           "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PDN_BENCH_BATCH", 1024)))
    ap.add_argument("--cpu-batch", type=int, default=int(os.environ.get("PDN_BENCH_BATCH", 1024)))
    ap.add_argument("--cpu-total-len", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
