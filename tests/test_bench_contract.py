"""CPU: the parts of bench.py's JSON contract that need no GPU — the reference arm line (driver keys, `impl`,
`cpu_baseline`, zero-copy `e2e`), ranks other than 0 staying silent, and the algorithmic-byte model of DESIGN.md §4."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_line_has_the_contract_keys():
    lines = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-pairs", "4", "--rings", "16",
                       "--cols", "512"])
    d = json.loads(lines[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    assert run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                     env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    ne, npl, it = np.array([10, 12, 14]), np.array([100, 110, 120]), np.array([2, 3])
    ab = bench.algorithmic_bytes(1000, ne, npl, it)
    F = ne + npl
    assert ab["extract"] == 16 * 1000 * 3 + 4 * F.sum()
    assert ab["nn_build"] == 32 * F[:-1].sum()
    assert ab["knn"] == (it * F[1:]).sum() * (16 + 16 * 5 + 8)
    assert ab["lm"] == (it * F[1:]).sum() * 48 and ab["misc"] == 0


import pytest


@pytest.mark.gpu
def test_our_arm_line_has_the_contract_keys():
    lines = run_bench(["--steps", "2", "--warmup", "1", "--scans", "12", "--rings", "16", "--cols", "512",
                       "--cpu-seconds", "0.5", "--no-configs"])
    d = json.loads(lines[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 2 and d["scaling"] == "weak" and d["value"] > 0
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 12 * 16 * 512 * 16 and e["d2h_bytes_per_step"] > 0
    assert e["value_each_call_waited"] > 0 and "asynchronous" in e["mode"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] in d["kernel_ms_per_step"] and r["launches"] > 0
    cb = d["cpu_baseline"]
    assert cb["cores"] == 1 and cb["kind"] == "port" and cb["value"] > 0 and "pairs" in cb["sample"]
    assert abs(sum(d["kernel_share"].values()) - 1.0) < 1e-9
    assert "workload" in d["config"] and "l2" in d["config"]
    assert d["metric"] == "extract+register scans/sec at 16x512"
    assert cb["reference_readme"]["ms_per_scan"] == 16.5
    # the run checks its own results against the CPU oracle on the cpu_baseline sample (non-zero exit on violation)
    pc = d["parity_check"]
    assert pc["ok"] and pc["pairs"] >= 4 and pc["max_rad"] < 1e-6 and pc["max_m"] < 1e-5
    assert pc["terminations_equal"] and pc["outer_iterations_equal"] and pc["feature_counts_equal"]
    assert pc["indices_equal"]["equal"]
    assert r["traffic"] is None  # no ncu capture exists for this shape: never a number measured on another one
