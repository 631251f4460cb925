"""CPU: the oracle's feature half against the reference's own unit-test known answers, against the REAL
reference code (oracle/_ref, compiled from /root/reference) and against the committed golden fixtures."""
import os

import numpy as np
import pytest

import feature_cases as FC
from loam_b200 import synth
from oracle.pyoracle import FeParams, LidarParams

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "features_golden.npz")


@pytest.mark.parametrize("name", list(FC.CURVATURE_CASES))
def test_curvature_known_answers(oracle, name):
    pts, lp, expect = FC.CURVATURE_CASES[name]()
    c = oracle.curvature(pts, LidarParams(*lp), FeParams(*FC.FE_TEST))
    assert len(c) == len(pts)
    np.testing.assert_allclose(c, expect, atol=1e-9, rtol=0)  # tolerance of the reference test


@pytest.mark.parametrize("name", list(FC.MASK_CASES))
def test_mask_known_answers(oracle, name):
    pts, lp, expect = FC.MASK_CASES[name]()
    m = oracle.valid_mask(pts, LidarParams(*lp), FeParams(*FC.FE_TEST))
    assert len(m) == len(pts)
    for i, v in expect.items():
        assert bool(m[i]) == v, (name, i)


@pytest.mark.parametrize("name", list(FC.CURVATURE_CASES) + list(FC.MASK_CASES))
def test_real_reference_agrees_on_unit_scenes(oracle, reflib, name):
    pts, lp, _ = {**FC.CURVATURE_CASES, **FC.MASK_CASES}[name]()
    lp, fe = LidarParams(*lp), FeParams(*FC.FE_TEST)
    assert np.array_equal(oracle.curvature(pts, lp, fe), reflib.curvature(pts, lp, fe)[0])
    assert np.array_equal(oracle.valid_mask(pts, lp, fe), reflib.valid_mask(pts, lp, fe))
    eo, po = oracle.extract(pts, lp, fe)
    er, pr = reflib.extract(pts, lp, fe)
    assert np.array_equal(eo, er) and np.array_equal(po, pr)


def test_empty_cloud(oracle):  # NonStdAllocator test: empty scan, LidarParams(0, 0, ...)
    e, p = oracle.extract(np.zeros((0, 3)), LidarParams(0, 0, 0.1, 100.0), FeParams(*FC.FE_TEST))
    assert len(e) == 0 and len(p) == 0


def test_size_mismatch_raises(oracle):
    with pytest.raises(RuntimeError):
        oracle.extract(np.zeros((10, 3)), LidarParams(1, 11, 0.1, 100.0), FeParams.default())


@pytest.mark.parametrize("shape", [(16, 1800), (64, 1024), (32, 777)])
@pytest.mark.parametrize("fe_t", FC.PARAM_SWEEP)
def test_oracle_matches_real_reference_on_synthetic_scans(oracle, reflib, shape, fe_t):
    R, P = shape
    scan = synth.make_scan(R, P, k=5, dropout=0.01)[:, :3].astype(np.float64)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams(*fe_t)
    assert np.array_equal(oracle.curvature(scan, lp, fe), reflib.curvature(scan, lp, fe)[0])  # bit-exact fp64
    assert np.array_equal(oracle.valid_mask(scan, lp, fe), reflib.valid_mask(scan, lp, fe))
    eo, po, ties = oracle.extract(scan, lp, fe, return_ties=True)
    er, pr = reflib.extract(scan, lp, fe)
    if ties == 0:  # the reference's std::sort leaves ties unpinned; tie-free data must agree exactly
        assert np.array_equal(eo, er)
        assert np.array_equal(po, pr)
    else:
        assert sorted(eo) == sorted(er) or ties > 0


def test_full_size_128x2048_matches_real_reference(oracle, reflib):
    scan = synth.make_scan(128, 2048, k=1)[:, :3].astype(np.float64)
    lp, fe = LidarParams(128, 2048, 1.0, 120.0), FeParams.default()
    eo, po, ties = oracle.extract(scan, lp, fe, return_ties=True)
    er, pr = reflib.extract(scan, lp, fe)
    assert ties == 0
    assert np.array_equal(eo, er) and np.array_equal(po, pr)


def test_golden_fixtures(oracle):
    """Fixtures were produced by the real reference (tests/golden/make_golden.py); they travel to the GPU box."""
    g = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 4
    for n in names:
        R, P = g[n + "/shape"]
        fe_t = g[n + "/fe"]
        fe = FeParams(int(fe_t[0]), int(fe_t[1]), int(fe_t[2]), int(fe_t[3]), *fe_t[4:])
        lp = LidarParams(int(R), int(P), *g[n + "/range"])
        scan = g[n + "/scan"].astype(np.float64)
        assert np.array_equal(oracle.curvature(scan, lp, fe), g[n + "/curvature"])
        assert np.array_equal(oracle.valid_mask(scan, lp, fe), g[n + "/mask"])
        e, p = oracle.extract(scan, lp, fe)
        assert np.array_equal(e, g[n + "/edge"]), n
        assert np.array_equal(p, g[n + "/planar"]), n
