"""GPU parity (through the C-ABI): loamgpu_extract_dewarped — the de-warp fused into the extraction kernel's ring
staging (extension).  The moved points must equal orc_dewarp bit for bit and the indices must equal
the reference's extraction run on those moved points."""
import numpy as np
import pytest

import helpers as H
import loam_b200 as L
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, RegParams

pytestmark = pytest.mark.gpu

MOTIONS = [
    np.r_[H.axis_angle(0.02, (0, 0, 1)), 0.09, -0.03, 0.0],
    np.r_[H.axis_angle(0.05, (0.3, -0.2, 1.0)), 0.2, 0.1, -0.05],
    np.r_[-H.axis_angle(0.03, (0, 1, 0.5)), 0.1, 0.0, -0.02],
]


def layouts(scan):
    xyz = scan[:, :3].astype(np.float64)
    f64x4 = np.zeros((len(scan), 4))
    f64x4[:, :3] = xyz
    f32s = np.zeros((len(scan), 5), dtype=np.float32)
    f32s[:, :3] = scan[:, :3]
    return {"f32x4": scan, "f32x3": np.ascontiguousarray(scan[:, :3]), "f64x3": xyz, "f64x4": f64x4, "f32_stride20": f32s}


@pytest.mark.parametrize("shape", [(16, 1800), (64, 1024), (128, 2048), (5, 333)])
@pytest.mark.parametrize("mi", range(len(MOTIONS)))
def test_points_and_indices_bit_exact(ctx, oracle, shape, mi):
    R, P = shape
    m = MOTIONS[mi]
    scan = synth.make_scan(R, P, k=11 + mi, dropout=0.01)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    moved_o = oracle.dewarp(scan[:, :3].astype(np.float64), P, m)
    eo, po = oracle.extract(moved_o, lp, fe)
    e, p, moved = ctx.extract_dewarped(scan, H.to_capi(lp), H.to_capi(fe), m)
    assert np.array_equal(moved, moved_o), "de-warped points differ from the oracle"
    assert np.array_equal(e, eo) and np.array_equal(p, po)


def test_real_reference_on_the_moved_points(ctx, oracle, reflib):
    R, P = 64, 1024
    m = MOTIONS[1]
    scan = synth.make_warped_scan(R, P, 21, np.r_[H.axis_angle(0.04, (0, 0, 1)), 0.3, 0.1, 0.0], sigma=0.01)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    er, pr = reflib.extract(oracle.dewarp(scan[:, :3].astype(np.float64), P, m), lp, fe)
    e, p, _ = ctx.extract_dewarped(scan, H.to_capi(lp), H.to_capi(fe), m)
    assert np.array_equal(e, er) and np.array_equal(p, pr)


def test_input_layouts_agree(ctx, oracle):
    R, P = 16, 901
    m = MOTIONS[0]
    scan = synth.make_scan(R, P, k=9)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    moved_o = oracle.dewarp(scan[:, :3].astype(np.float64), P, m)
    eo, po = oracle.extract(moved_o, lp, fe)
    for name, arr in layouts(scan).items():
        e, p, moved = ctx.extract_dewarped(arr, H.to_capi(lp), H.to_capi(fe), m)
        assert np.array_equal(moved, moved_o), name
        assert np.array_equal(e, eo) and np.array_equal(p, po), name


def test_identity_motion_equals_plain_extract(ctx):
    R, P = 64, 1024
    scan = synth.make_scan(R, P, k=2)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    e0, p0 = ctx.extract(scan, H.to_capi(lp), H.to_capi(fe))
    e, p, moved = ctx.extract_dewarped(scan, H.to_capi(lp), H.to_capi(fe), [0, 0, 0, 1, 0, 0, 0])
    assert np.array_equal(e, e0) and np.array_equal(p, p0)
    assert np.array_equal(moved, scan[:, :3].astype(np.float64))
    e, p, moved = ctx.extract_dewarped(scan, H.to_capi(lp), H.to_capi(fe), [0, 0, 0, 1, 0, 0, 0], want_points=False)
    assert moved is None and np.array_equal(e, e0) and np.array_equal(p, p0)
    # a plain extraction after a de-warped one is unaffected by it
    e, p = ctx.extract(scan, H.to_capi(lp), H.to_capi(fe))
    assert np.array_equal(e, e0) and np.array_equal(p, p0)


def test_errors(ctx):
    lp, fe = LidarParams(4, 64, 1.0, 120.0), FeParams.default()
    scan = synth.make_scan(4, 64, k=0)
    with pytest.raises(_capi.LoamGpuError) as ei:
        ctx.extract_dewarped(scan, H.to_capi(lp), H.to_capi(fe), [0, 0, 0, np.nan, 0, 0, 0])
    assert ei.value.code == _capi.ERR_INVALID
    with pytest.raises(_capi.LoamGpuError) as ei:
        ctx.extract_dewarped(scan[:-1], H.to_capi(lp), H.to_capi(fe), [0, 0, 0, 1, 0, 0, 0])
    assert ei.value.code == _capi.ERR_SIZE_MISMATCH
    e, p, moved = ctx.extract_dewarped(np.zeros((0, 3)), H.to_capi(LidarParams(0, 0, 1.0, 120.0)), H.to_capi(fe),
                                       [0, 0, 0, 1, 0, 0, 0])
    assert len(e) == 0 and len(p) == 0 and moved.shape == (0, 3)


def test_dewarped_sweep_registers_onto_a_static_view(ctx):
    """A sweep recorded while the sensor moves 0.5 m / 0.05 rad, registered onto a static view taken at the pose of
    the sweep start.  Raw, the smear biases the estimate by about half the motion; de-warped, it is the identity."""
    R, P = 64, 1024
    lpp = L.LidarParams(R, P, 1.0, 120.0)
    m = np.r_[H.axis_angle(0.05, (0, 0, 1)), 0.5, 0.1, 0.0]
    ident = np.r_[0, 0, 0, 1.0, 0, 0, 0]
    A = (0.5, -0.3, 0.2)
    static = synth.make_warped_scan(R, P, 0, ident, sigma=0.01, seed=1, start_pose=A)
    moving = synth.make_warped_scan(R, P, 0, m, sigma=0.01, seed=2, start_pose=A)
    tgt = L.extractFeatures(static, lpp)
    src = L.extractFeaturesDewarped(moving, lpp, L.Pose3d(L.Quaterniond.from_coeffs(m[:4]), m[4:7]))
    assert src.edge_points.dtype == np.float64 and src.edge_points.shape[1] == 3
    est = L.registerFeatures(src, tgt, L.Pose3d.Identity())._to7()
    raw = L.registerFeatures(L.extractFeatures(moving, lpp), tgt, L.Pose3d.Identity())._to7()
    assert H.angular_distance(est[:4], ident[:4]) < 2e-3 and np.linalg.norm(est[4:]) < 1e-2, est
    assert H.angular_distance(raw[:4], ident[:4]) > 1e-2 and np.linalg.norm(raw[4:]) > 0.1, raw


def test_sequence_with_per_sweep_motions_matches_oracle(ctx, oracle):
    """loamgpu_odometry_host_dewarped: every scan de-warped with its own motion inside the extraction kernel, the moved
    points carried to registration by the pack kernel.  Checked against the CPU chain orc_dewarp -> extract -> register."""
    R, P, n = 32, 1024, 5
    lp, fe, rp = LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()
    motions = np.stack([synth.relative_pose(k, k + 1) for k in range(n)])
    motions[2, :4] = -motions[2, :4]  # one on the far hemisphere
    scans = np.stack([synth.make_warped_scan(R, P, k, synth.relative_pose(k, k + 1), sigma=0.01) for k in range(n)])
    poses, term, its, ne, npl = ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp), sweep_motions=motions)
    feats = []
    for k in range(n):
        moved = oracle.dewarp(scans[k][:, :3].astype(np.float64), P, motions[k])
        e, p = oracle.extract(moved, lp, fe)
        assert ne[k] == len(e) and npl[k] == len(p), k
        ge, gp, gm = ctx.extract_dewarped(scans[k], H.to_capi(lp), H.to_capi(fe), motions[k])
        assert np.array_equal(ge, e) and np.array_equal(gp, p) and np.array_equal(gm, moved)
        feats.append((moved[e], moved[p]))
    for k in range(n - 1):
        ref = oracle.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp)
        assert H.angular_distance(ref[:4], poses[k][:4]) < 1e-6, k
        assert np.abs(ref[4:] - poses[k][4:]).max() < 1e-5, k
    # the single-pair path fed with the moved points the single-scan call returns agrees with the sequence call
    one = ctx.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], [0, 0, 0, 1, 0, 0, 0], H.to_capi(rp))
    assert H.angular_distance(one[:4], poses[0][:4]) < 1e-9 and np.abs(one[4:] - poses[0][4:]).max() < 1e-9


def test_sequence_identity_motions_bit_identical_to_plain(ctx):
    R, P, n = 16, 1800, 4
    lp, fe, rp = LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    a = ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp))
    ident = np.tile(np.r_[0, 0, 0, 1.0, 0, 0, 0], (n, 1))
    b = ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp), sweep_motions=ident)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    with pytest.raises(_capi.LoamGpuError) as ei:
        bad = ident.copy()
        bad[1, 5] = np.inf
        ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp), sweep_motions=bad)
    assert ei.value.code == _capi.ERR_INVALID


def test_golden_fixture_from_real_reference(ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dewarp_golden.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        R, P = (int(v) for v in g[n + "/shape"])
        lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
        e, p, moved = ctx.extract_dewarped(np.ascontiguousarray(g[n + "/scan"]), H.to_capi(lp), H.to_capi(fe), g[n + "/motion"])
        assert np.array_equal(moved, g[n + "/moved"]), n
        assert np.array_equal(e, g[n + "/edge"]) and np.array_equal(p, g[n + "/planar"]), n
