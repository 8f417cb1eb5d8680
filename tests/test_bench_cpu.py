"""bench.py pieces that run without a GPU: the reference arm (oracle port on the host cores), the
algorithmic-byte formula of SURVEY.md section 8(d), and the rule that the product never touches oracle/."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edge*feature/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_formula_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    n, nnz, F = 2_000_000, 52_000_000, 64
    per_step_stored = 4 * (n + 1) + 8 * nnz + 12 * n * F
    assert abs(per_step_stored - 1.960e9) < 5e6                                  # 1.960 GB
    assert bench.algorithmic_bytes_per_pass(n, nnz, F, value_free=False) == 20 * per_step_stored
    vf = bench.algorithmic_bytes_per_pass(n, nnz, F, value_free=True)
    assert vf == 20 * (per_step_stored - 4 * nnz) + 2 * 4 * nnz                  # values only in the 2 first steps


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ppnp_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+(oracle|ppnp_oracle)\b", txt, re.M) or "oracle/" in txt and fn.endswith(".py") and "ppnp_oracle" in txt:
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
