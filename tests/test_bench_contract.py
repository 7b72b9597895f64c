"""bench.py's reference arm (CPU) prints ONE JSON line with the contract's keys (run here on a tiny sample), and the product arm's
source never imports oracle/ outside the cpu_baseline leg."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2",
                        "--cpu-total-len", "6"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "llama3_6L_greedy_generation_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_product_arm_does_not_use_the_oracle():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ours = src[src.index("def run_ours("):src.index("def cpu_baseline(")]
    assert "oracle" not in re.sub(r"#.*", "", ours), "the product arm of bench.py must not import or call oracle/"
    for root, _, files in os.walk(os.path.join(ROOT, "pydynet_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle/"
    for f in os.listdir(os.path.join(ROOT, "workloads")):
        if f.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, "workloads", f)).read(), re.M), f
