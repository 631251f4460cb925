"""GPU parity of the explicit batch entry points (loamgpu_extract_batch, loamgpu_register_pairs): every element of a
batch must equal the corresponding single call (loamgpu_extract: bit-exact indices; loamgpu_register: same termination,
iteration count and — the 6x6 sums being reduced by one CTA instead of a cluster of eight — poses to rounding)."""
import numpy as np
import pytest

import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, RegParams

pytestmark = pytest.mark.gpu
IDENT = np.array([0, 0, 0, 1, 0, 0, 0.0])


@pytest.mark.parametrize("shape,n", [((64, 1024), 5), ((16, 1800), 7), ((32, 777), 3)])
def test_extract_batch_equals_single_calls_and_oracle(ctx, oracle, shape, n):
    R, P = shape
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    scans = np.stack([synth.make_scan(R, P, k=20 + k, dropout=0.01 if k % 2 else 0.0) for k in range(n)])
    got = ctx.extract_batch(scans, H.to_capi(lp), H.to_capi(fe))
    assert len(got) == n
    for k in range(n):
        e1, p1 = ctx.extract(scans[k], H.to_capi(lp), H.to_capi(fe))
        assert np.array_equal(got[k][0], e1) and np.array_equal(got[k][1], p1)
    eo, po = oracle.extract(scans[0][:, :3].astype(np.float64), lp, fe)
    assert np.array_equal(got[0][0], eo) and np.array_equal(got[0][1], po)


def test_extract_batch_chunks_layouts_and_errors(ctx):
    R, P = 16, 512
    lp, fe = H.to_capi(LidarParams(R, P, 1.0, 120.0)), H.to_capi(FeParams.default())
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(7)])
    ref = ctx.extract_batch(scans, lp, fe)
    ctx.set_chunk_pairs(3)  # 7 scans in chunks of 3, 3, 1
    try:
        chunked = ctx.extract_batch(scans, lp, fe)
        f64 = ctx.extract_batch(scans[:, :, :3].astype(np.float64), lp, fe)  # f64x3 layout, same values
    finally:
        ctx.set_chunk_pairs(0)  # back to automatic
    for a, b, c in zip(ref, chunked, f64):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    assert ctx.extract_batch(np.zeros((0, R * P, 4), dtype=np.float32), lp, fe) == []
    with pytest.raises(_capi.LoamGpuError) as ei:  # wrong geometry: the reference's std::runtime_error text
        ctx.extract_batch(scans[:, :-1], lp, fe)
    assert ei.value.code == _capi.ERR_SIZE_MISMATCH and "does not match provided lidar parameters" in str(ei.value)


def test_register_pairs_equals_single_calls(ctx, oracle):
    ed, pl = H.simple_scene()
    rp = _capi.default_reg_params()
    pairs, inits = [], []
    for name, sTt, init, max_it, _, _ in H.REG_SCENARIOS[:4]:
        pairs.append((H.transform(ed, sTt), H.transform(pl, sTt), ed, pl))
        inits.append(IDENT)
    # ragged batch: a planar-only pair, a pair with too few associations, a lidar-shaped pair with other sizes
    pairs.append((np.zeros((0, 3)), pl[:3600], np.zeros((0, 3)), pl[:3600]))
    inits.append(IDENT)
    far = np.r_[H.axis_angle(0.3, [0, 1, 0]), [1.0, 2.0, 3.0]]
    pairs.append((ed + 100.0, pl + 100.0, ed, pl))
    inits.append(far)
    R, P = 16, 1800
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    s0 = synth.make_scan(R, P, k=10)[:, :3].astype(np.float64)
    s1 = synth.make_scan(R, P, k=11)[:, :3].astype(np.float64)
    (e0, p0), (e1, p1) = oracle.extract(s0, lp, fe), oracle.extract(s1, lp, fe)
    pairs.append((s1[e1], s1[p1], s0[e0], s0[p0]))
    inits.append(IDENT)
    poses, term, its = ctx.register_pairs(pairs, np.array(inits), rp)
    for k, (pr, init) in enumerate(zip(pairs, inits)):
        single, det = ctx.register(*pr, init, rp, want_detail=True)
        assert term[k] == det["termination"] and its[k] == det["n_iters"], k
        assert H.angular_distance(single[:4], poses[k][:4]) < 1e-9 and np.abs(single[4:] - poses[k][4:]).max() < 1e-9, k
    assert term[5] == 2 and np.array_equal(poses[5], far)  # INSUFFICIENT_ASSOCIATIONS: estimate untouched
    # chunked (2 pairs at a time) and default-identity initial poses
    ctx.set_chunk_pairs(2)
    try:
        poses2, term2, its2 = ctx.register_pairs(pairs, np.array(inits), rp)
        poses3, _, _ = ctx.register_pairs(pairs[:4], None, rp)
    finally:
        ctx.set_chunk_pairs(0)  # back to automatic
    assert np.array_equal(poses, poses2) and np.array_equal(term, term2) and np.array_equal(its, its2)
    assert np.array_equal(poses3, poses[:4])
    # and against the CPU oracle
    po = oracle.register(*pairs[6], IDENT, RegParams.default())
    assert H.angular_distance(po[:4], poses[6][:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - poses[6][4:]).max() < H.POSE_TOL_M


def test_register_pairs_empty_and_invalid(ctx):
    rp = _capi.default_reg_params()
    poses, term, its = ctx.register_pairs([], None, rp)
    assert poses.shape == (0, 7)
    ed, pl = H.simple_scene()
    bad = _capi.default_reg_params()
    bad.num_plane_neighbors = 0
    with pytest.raises(_capi.LoamGpuError):
        ctx.register_pairs([(ed, pl, ed, pl)], None, bad)
