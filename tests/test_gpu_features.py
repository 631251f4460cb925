"""GPU parity (through the C-ABI): feature extraction — bit-exact indices, curvature and masks against the oracle,
the reference's unit-test known answers and the golden fixtures produced by the real reference code."""
import os

import numpy as np
import pytest

import feature_cases as FC
import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "features_golden.npz")


def gpu_extract(ctx, pts, lp, fe):
    return ctx.extract(pts, H.to_capi(lp), H.to_capi(fe))


@pytest.mark.parametrize("name", list(FC.CURVATURE_CASES))
def test_curvature_known_answers(ctx, oracle, name):
    pts, lp, expect = FC.CURVATURE_CASES[name]()
    lp, fe = LidarParams(*lp), FeParams(*FC.FE_TEST)
    c = ctx.curvature(pts, H.to_capi(lp), H.to_capi(fe))
    np.testing.assert_allclose(c, expect, atol=1e-9, rtol=0)
    assert np.array_equal(c, oracle.curvature(pts, lp, fe))


@pytest.mark.parametrize("name", list(FC.MASK_CASES))
def test_mask_known_answers(ctx, oracle, name):
    pts, lp, expect = FC.MASK_CASES[name]()
    lp, fe = LidarParams(*lp), FeParams(*FC.FE_TEST)
    m = ctx.valid_mask(pts, H.to_capi(lp), H.to_capi(fe))
    for i, v in expect.items():
        assert bool(m[i]) == v, (name, i)
    assert np.array_equal(m, oracle.valid_mask(pts, lp, fe))
    e, p = gpu_extract(ctx, pts, lp, fe)
    eo, po = oracle.extract(pts, lp, fe)
    assert np.array_equal(e, eo) and np.array_equal(p, po)


def test_empty_cloud_and_size_mismatch(ctx):
    lp0 = LidarParams(0, 0, 0.1, 100.0)
    e, p = gpu_extract(ctx, np.zeros((0, 3)), lp0, FeParams(*FC.FE_TEST))  # NonStdAllocator scene
    assert len(e) == 0 and len(p) == 0
    with pytest.raises(_capi.LoamGpuError) as ei:
        gpu_extract(ctx, np.zeros((10, 3)), LidarParams(1, 11, 0.1, 100.0), FeParams.default())
    assert ei.value.code == _capi.ERR_SIZE_MISMATCH
    # message text of the reference's std::runtime_error (common.h:106-111)
    assert str(ei.value) == "LOAM: provided lidar scan size ( 10)  does not match provided lidar parameters (1 x 11)"
    with pytest.raises(_capi.LoamGpuError) as ei:
        fe = FeParams.default()
        fe.number_sectors = 0
        gpu_extract(ctx, np.zeros((11, 3)), LidarParams(1, 11, 0.1, 100.0), fe)
    assert ei.value.code == _capi.ERR_INVALID


@pytest.mark.parametrize("shape", [(16, 1800), (64, 1024), (128, 2048), (32, 777), (3, 50)])
@pytest.mark.parametrize("fe_t", FC.PARAM_SWEEP)
def test_indices_bit_exact_vs_oracle(ctx, oracle, shape, fe_t):
    R, P = shape
    scan = synth.make_scan(R, P, k=5, dropout=0.01)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams(*fe_t)
    xyz = scan[:, :3].astype(np.float64)
    eo, po, ties = oracle.extract(xyz, lp, fe, return_ties=True)
    e, p = gpu_extract(ctx, scan, lp, fe)  # float4-shaped input, TMA-staged
    assert np.array_equal(e, eo), "edge indices differ"
    assert np.array_equal(p, po), "planar indices differ"
    # same tie-break is used on both sides, so ties do not excuse a difference; report them anyway
    assert ties >= 0
    assert np.array_equal(ctx.curvature(scan, H.to_capi(lp), H.to_capi(fe)), oracle.curvature(xyz, lp, fe))
    assert np.array_equal(ctx.valid_mask(scan, H.to_capi(lp), H.to_capi(fe)), oracle.valid_mask(xyz, lp, fe))


@pytest.mark.parametrize("layout,P", [("f32x4", 9000), ("f32x3", 9600), ("f64x3", 6900)])
def test_largest_rings_that_fit_a_cta(ctx, oracle, layout, P):
    """One ring is staged in one CTA's shared memory (include/loamgpu.h, limits): rings close to the limit of each
    record type still give the reference's indices, with a whole SM's shared memory taken by one ring."""
    R = 2
    scan = synth.make_scan(R, P, k=3, dropout=0.01)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    xyz = scan[:, :3].astype(np.float64)
    eo, po = oracle.extract(xyz, lp, fe)
    cloud = {"f32x4": scan, "f32x3": np.ascontiguousarray(scan[:, :3]), "f64x3": xyz}[layout]
    e, p = gpu_extract(ctx, cloud, lp, fe)
    assert np.array_equal(e, eo) and np.array_equal(p, po)


def test_ring_too_long_is_refused_loudly(ctx):
    R, P = 1, 20000
    with pytest.raises(_capi.LoamGpuError) as ei:
        gpu_extract(ctx, np.zeros((R * P, 3), dtype=np.float32), LidarParams(R, P, 1.0, 120.0), FeParams.default())
    assert ei.value.code == _capi.ERR_UNSUPPORTED
    assert "shared memory" in str(ei.value)


@pytest.mark.parametrize("layout", ["f64x3", "f32x3", "f64x4_strided", "f32_stride20"])
def test_input_layouts_agree(ctx, oracle, layout):
    R, P = 16, 901  # odd P: f64 ring bytes not 16-aligned -> strided-load path
    scan = synth.make_scan(R, P, k=9)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    xyz = scan[:, :3].astype(np.float64)
    eo, po = oracle.extract(xyz, lp, fe)
    if layout == "f64x3":
        arr = xyz
    elif layout == "f32x3":
        arr = np.ascontiguousarray(scan[:, :3])
    elif layout == "f64x4_strided":
        arr = np.zeros((R * P, 4))
        arr[:, :3] = xyz
    else:
        arr = np.zeros((R * P, 5), dtype=np.float32)
        arr[:, :3] = scan[:, :3]
    e, p = gpu_extract(ctx, arr, lp, fe)
    assert np.array_equal(e, eo) and np.array_equal(p, po)


def test_f64_cloud_not_float32_representable(ctx, oracle):
    """The Python path of the reference hands float64 clouds: unfused fp64 arithmetic must still match bit-for-bit."""
    R, P = 32, 1024
    rng = np.random.RandomState(7)
    xyz = synth.make_scan(R, P, k=2)[:, :3].astype(np.float64) + rng.normal(0, 1e-4, (R * P, 3))
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    assert np.array_equal(ctx.curvature(xyz, H.to_capi(lp), H.to_capi(fe)), oracle.curvature(xyz, lp, fe))
    e, p = gpu_extract(ctx, xyz, lp, fe)
    eo, po = oracle.extract(xyz, lp, fe)
    assert np.array_equal(e, eo) and np.array_equal(p, po)


def test_golden_fixtures_from_real_reference(ctx):
    g = np.load(GOLDEN)
    for n in sorted({k.split("/")[0] for k in g.files}):
        R, P = g[n + "/shape"]
        fe_t = g[n + "/fe"]
        fe = FeParams(int(fe_t[0]), int(fe_t[1]), int(fe_t[2]), int(fe_t[3]), *fe_t[4:])
        lp = LidarParams(int(R), int(P), *g[n + "/range"])
        scan = np.ascontiguousarray(g[n + "/scan"])  # float32 (n,3)
        assert np.array_equal(ctx.curvature(scan, H.to_capi(lp), H.to_capi(fe)), g[n + "/curvature"]), n
        assert np.array_equal(ctx.valid_mask(scan, H.to_capi(lp), H.to_capi(fe)), g[n + "/mask"]), n
        e, p = gpu_extract(ctx, scan, lp, fe)
        assert np.array_equal(e, g[n + "/edge"]), n
        assert np.array_equal(p, g[n + "/planar"]), n


def test_real_reference_library_agrees_when_present(ctx, reflib):
    scan = synth.make_scan(64, 1024, k=33)
    lp, fe = LidarParams(64, 1024, 1.0, 120.0), FeParams.default()
    er, pr = reflib.extract(scan[:, :3].astype(np.float64), lp, fe)
    e, p = gpu_extract(ctx, scan, lp, fe)
    assert np.array_equal(e, er) and np.array_equal(p, pr)


def test_properties_at_full_size(ctx):
    """Size-independent properties at 128x2048: picks are unique, valid at selection time, respect thresholds,
    respect the per-sector cap (max+1) and the +-(N-1) suppression distance inside a sector."""
    R, P = 128, 2048
    scan = synth.make_scan(R, P, k=4)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    e, p = gpu_extract(ctx, scan, lp, fe)
    c = ctx.curvature(scan, H.to_capi(lp), H.to_capi(fe))
    m = ctx.valid_mask(scan, H.to_capi(lp), H.to_capi(fe))
    allp = np.concatenate([e, p])
    assert len(np.unique(allp)) == len(allp)
    assert m[allp].all()
    assert (c[e] > fe.edge_feat_threshold).all() and (c[p] < fe.planar_feat_threshold).all()
    pps = P // 6
    for idx, cap in ((e, 11), (p, 51)):
        ring, col = idx // P, idx % P
        sector = np.minimum(col // pps, 5)
        key = ring * 6 + sector
        assert np.bincount(key).max() <= cap
        assert (np.diff(key) >= 0).all()  # line-major, sector-major output order
    order = np.sort(allp)
    same_ring = (order[1:] // P) == (order[:-1] // P)
    assert (np.diff(order)[same_ring] >= 3).all()  # N = 3: picks are at least N apart within a ring


def test_python_module_mirror(ctx, oracle):
    import loam_b200 as loam
    R, P = 16, 512
    scan = synth.make_scan(R, P, k=1)
    lp = loam.LidarParams(R, P, 1.0, 120.0)
    feats = loam.extractFeatures(scan, lp)
    eo, po = oracle.extract(scan[:, :3].astype(np.float64), LidarParams(R, P, 1.0, 120.0), FeParams.default())
    assert np.array_equal(feats.edge_points, scan[eo]) and np.array_equal(feats.planar_points, scan[po])
    # sequence-of-points input, as the reference bindings marshal it
    feats2 = loam.extractFeatures([row for row in scan[:, :3].astype(np.float64)], lp)
    assert np.array_equal(feats2.planar_points, scan[po][:, :3].astype(np.float64))
    with pytest.raises(RuntimeError):
        loam.extractFeatures(scan[:-1], lp)
    cv = loam.computeCurvature(scan, lp)
    assert cv["index"][5] == 5 and cv["curvature"][0] == -1
    assert loam.computeValidPoints(scan, lp).dtype == bool
