"""GPU: the code paths an environment switch selects give the same results as the defaults.

The switches are read once per process, so each case runs a short sequence in a child process and compares the
poses / terminations / feature counts with this process's result bit for bit (a tree's shape, where its boxes were
kept during the build, or whether the LM solve used its moment sums for the inlier planes may change speed or — for
the moment sums — the last bits, never an index or a count)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
from loam_b200 import _capi, synth
R, P, n = 16, 512, 24  # (>= 16 pairs: the batched shared-memory walk runs by default)
scans = np.stack([synth.make_scan(R, P, k) for k in range(n)]).astype(np.float32)
ctx = _capi.Context(0)
lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
poses, term, iters, ne, npl = ctx.odometry_host(scans, lp, fe, rp)
print(json.dumps({"poses": np.asarray(poses).tolist(), "term": np.asarray(term).tolist(), "iters": np.asarray(iters).tolist(),
                  "ne": np.asarray(ne).tolist(), "np": np.asarray(npl).tolist()}))
""" % ROOT


def run_child(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def baseline():
    return run_child({})


@pytest.mark.parametrize("env", [{"LOAMGPU_BUILD_GLOBAL_BOXES": "1"}, {"LOAMGPU_BUILD_LEGACY": "1"},
                                 {"LOAMGPU_KNN_SMEM_MIN_PAIRS": "100000"}])
def test_build_and_walk_variants_are_bit_identical(baseline, env):
    got = run_child(env)
    assert got == baseline


@pytest.mark.parametrize("env", [{"LOAMGPU_LM_MOMENTS": "0"}, {"LOAMGPU_QUERY_ORDER": "original"}])
def test_summation_order_variants_agree_to_rounding(baseline, env):
    """Streaming every record instead of using the moment sums, or walking the queries in source-index order instead of
    Morton order, changes the order in which the residual sums are formed: same associations, poses equal to rounding."""
    got = run_child(env)
    assert got["term"] == baseline["term"] and got["iters"] == baseline["iters"]
    assert got["ne"] == baseline["ne"] and got["np"] == baseline["np"]
    np.testing.assert_allclose(np.array(got["poses"]), np.array(baseline["poses"]), rtol=0, atol=1e-9)
