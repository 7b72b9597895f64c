"""bench.py's reference arm (CPU) prints ONE JSON line with the contract's keys (run here on a tiny sample), and the product arm's
source never imports oracle/ outside the cpu_baseline leg."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "llama3_6L_greedy_generation_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    # the unmodified reference package where it is staged (baseline/_ref or /root/reference), the NumPy port otherwise
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["b1"]["value"] > 0 and "batch 1" in d["b1"]["sample"]  # the reference's own configuration, run in full
    assert d["config"]["seq_len"] == 256 and d["config"]["batch_per_gpu"] == 2  # same config dict as the product arm
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_product_arm_does_not_use_the_oracle():
    src = open(os.path.join(ROOT, "bench.py")).read()
    # everything from the product-arm marker up to the "CPU arm + checker" marker: no import of oracle/, no use of its names
    ours = src[src.index("# ------------------------------------------------------------------------------------------------- our arm"):
               src.index("# ------------------------------------------------------------------------------------------------- CPU arm + checker")]
    assert "def run_ours(" in ours and "def bench_b1(" in ours
    assert not re.search(r"(from|import)\s+oracle\b|pdn_oracle|LlamaOracle", ours), "the product arm of bench.py must not import or call oracle/"
    for root, _, files in os.walk(os.path.join(ROOT, "pydynet_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle/"
    for f in os.listdir(os.path.join(ROOT, "workloads")):
        if f.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, "workloads", f)).read(), re.M), f
