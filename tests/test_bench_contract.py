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
