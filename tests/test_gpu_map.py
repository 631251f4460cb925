"""GPU parity (through the C-ABI) of the large-target path: multi-CTA NN build (bvh_big.cu), the device-resident local
map (loamgpu_map_*) and scan-to-map registration (BASELINE.json config 5).  The reference has no separate code for
this — a map is just a large `target` of registerFeatures (registration.h:128-131, README.md:63) — so the checks are:
same k-NN lists as brute force / the oracle's k-d tree, and the same registration results as the small-target path
and the CPU oracle."""
import os

import numpy as np
import pytest

import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, RegParams
from test_gpu_registration import IDENT, check_knn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_ctx():
    """A context that sends EVERY target through the multi-CTA build (threshold 1 point)."""
    os.environ["LOAMGPU_BIG_TARGET_MIN"] = "1"
    try:
        c = _capi.Context(0)
    finally:
        del os.environ["LOAMGPU_BIG_TARGET_MIN"]
    yield c
    c.close()


@pytest.fixture(scope="module")
def scene():
    return H.simple_scene()


def assert_same_registration(a, b, exact=True):
    (pa, da), (pb, db) = a, b
    if exact:
        assert np.array_equal(pa, pb)
    else:
        assert H.angular_distance(pa[:4], pb[:4]) < H.POSE_TOL_RAD and np.abs(pa[4:] - pb[4:]).max() < H.POSE_TOL_M
    n = da["n_iters"] if isinstance(da, dict) else da.n_iters
    get = lambda d, k: d[k] if isinstance(d, dict) else getattr(d, k)
    assert n == get(db, "n_iters") and get(da, "termination") == get(db, "termination")
    for i in range(n):
        assert np.array_equal(get(da, "edge_assoc")[i], get(db, "edge_assoc")[i]), ("edge assoc", i)
        assert np.array_equal(get(da, "plane_assoc")[i], get(db, "plane_assoc")[i]), ("plane assoc", i)


@pytest.mark.parametrize("k,max_dist", [(5, 2.0), (5, 0.15), (1, 2.0), (8, 0.5), (5, -1.0), (32, 3.0)])
def test_big_build_knn_exact_vs_brute_force(big_ctx, oracle, scene, k, max_dist):
    ed, pl = scene
    rng = np.random.RandomState(100 + k)
    q = np.concatenate([pl[rng.choice(len(pl), 300)] + rng.normal(0, 0.07, (300, 3)), rng.uniform(-6, 8, (60, 3)),
                        np.array([[1e3, -1e3, 50.0]])])
    check_knn(big_ctx, oracle, pl, q, k, max_dist)
    check_knn(big_ctx, oracle, ed, q[::4], k, max_dist)


def test_big_build_degenerate_targets(big_ctx, oracle):
    q = np.array([[0.1, 0.2, 0.3], [5, 5, 5.0]])
    check_knn(big_ctx, oracle, np.array([[0.0, 0, 0]]), q, 5, 1.0)                 # one point: no internal node
    check_knn(big_ctx, oracle, np.array([[0.0, 0, 0], [1, 1, 1.0]]), q, 5, -1.0)   # a single internal node
    check_knn(big_ctx, oracle, np.tile([[1.0, 2, 3]], (5000, 1)), q, 5, -1.0)     # identical codes: index tie-break
    line = np.c_[np.zeros(3000), np.zeros(3000), np.linspace(0, 1, 3000)]         # zero-extent bbox on two axes
    check_knn(big_ctx, oracle, line, q, 5, 2.0)
    idx, cnt = big_ctx.knn(np.zeros((0, 3)), q, 5, 1.0)                           # empty target
    assert (cnt == 0).all()


def test_big_build_300k_points_vs_oracle_kdtree(big_ctx, ctx, oracle):
    """Several radix-sort tiles per pass (147 blocks), duplicated points, clustered + uniform density."""
    rng = np.random.RandomState(7)
    pts = np.concatenate([rng.uniform(-15, 15, (200_000, 3)) * [1, 0.66, 0.2],
                          rng.normal(0, 0.05, (80_000, 3)) + rng.uniform(-10, 10, (80_000, 1)) * [1, 0.3, 0.0]])
    pts = np.concatenate([pts, pts[:20_000]])  # exact duplicates -> ties resolve by ascending index
    q = np.concatenate([pts[rng.choice(len(pts), 3000)] + rng.normal(0, 0.02, (3000, 3)), rng.uniform(-20, 20, (500, 3))])
    ref, rcnt = oracle.knn_tree_batch(pts, q, 5, 1.0)
    for c in (big_ctx, ctx):  # 300k >= the default threshold too: the plain context takes the same path
        idx, cnt = c.knn(pts, q, 5, 1.0)
        assert np.array_equal(cnt, rcnt)
        # the oracle's k-d tree resolves exact ties (the duplicated points) by traversal order, the CUDA path by
        # ascending index: rows that differ must hold the same distances
        same = (idx == ref).all(axis=1)
        for i in np.nonzero(~same)[0]:
            d2 = lambda ii: ((pts[ii] - q[i]) ** 2).sum(axis=1)
            assert np.array_equal(d2(idx[i, :cnt[i]]), d2(ref[i, :rcnt[i]])), i
        assert same.mean() > 0.75


@pytest.mark.parametrize("case", H.REG_SCENARIOS[:2] + H.REG_SCENARIOS[4:], ids=lambda c: c[0])
def test_big_target_path_equals_small_target_path(ctx, big_ctx, scene, case):
    """Tree shape never changes k-NN results, everything downstream is the same code: bit-identical outputs."""
    _, sTt, init, max_it, _, _ = case
    ed, pl = scene
    rp = _capi.default_reg_params()
    rp.max_iterations = max_it
    init7 = IDENT if init is None else init
    se, sp = H.transform(ed, sTt), H.transform(pl, sTt)
    small = ctx.register(se, sp, ed, pl, init7, rp, want_detail=True)
    big = big_ctx.register(se, sp, ed, pl, init7, rp, want_detail=True)
    assert_same_registration(big, small)
    assert np.array_equal(big[1]["lm_cost"], small[1]["lm_cost"])


def test_map_register_equals_plain_register_and_oracle(ctx, oracle, scene):
    ed, pl = scene
    sTt = H.REG_SCENARIOS[2][1]
    se, sp = H.transform(ed, sTt), H.transform(pl, sTt)
    rp = _capi.default_reg_params()
    m = ctx.map_create(ed, pl)
    assert m.size() == (len(ed), len(pl))
    got = ctx.register_to_map(m, se, sp, IDENT, rp, want_detail=True)
    assert_same_registration(got, ctx.register(se, sp, ed, pl, IDENT, rp, want_detail=True))
    assert_same_registration(got, oracle.register(se, sp, ed, pl, IDENT, RegParams.default(), want_detail=True), exact=False)
    again = ctx.register_to_map(m, se, sp, IDENT, rp, want_detail=True)  # the map is not consumed
    assert_same_registration(again, got)
    assert np.array_equal(ctx.register_to_map(m, se, sp, IDENT, rp), got[0])
    m.close()


def test_map_edge_cases(ctx, scene):
    ed, pl = scene
    rp = _capi.default_reg_params()
    empty = np.zeros((0, 3))
    m = ctx.map_create(empty, pl[:3600])  # planar-only map (test_registration.cpp:177-199 shape)
    pose, det = ctx.register_to_map(m, empty, pl[:3600], IDENT, rp, want_detail=True)
    assert det["termination"] == 0 and det["n_iters"] == 1
    assert H.angular_distance(pose[:4], IDENT[:4]) < 1e-4 and np.all(np.abs(pose[4:]) < 1e-3)
    m.close()
    m = ctx.map_create(empty, empty)  # empty map: no associations
    init = np.r_[H.axis_angle(0.3, [0, 1, 0]), [1.0, 2.0, 3.0]]
    pose, det = ctx.register_to_map(m, ed, pl, init, rp, want_detail=True)
    assert det["termination"] == 2 and det["n_iters"] == 0 and np.array_equal(pose, init)
    m.close()


def test_map_update_appends_transforms_and_evicts(ctx, scene):
    ed, pl = scene
    rp = _capi.default_reg_params()
    sTt = H.REG_SCENARIOS[0][1]
    se, sp = H.transform(ed, sTt), H.transform(pl, sTt)
    shift = np.array([0, 0, 0, 1, 0.5, -0.25, 0.125])  # exactly representable translation: device == numpy bit for bit
    m = ctx.map_create(ed[:100], pl[:4000])
    add_e, add_p = ed[100:] - shift[4:], pl[4000:] - shift[4:]
    m.update(add_e, add_p, pose=shift)  # identity rotation: the device computes exactly p + t
    ed2, pl2 = np.concatenate([ed[:100], add_e + shift[4:]]), np.concatenate([pl[:4000], add_p + shift[4:]])
    assert m.size() == (len(ed), len(pl))
    full = ctx.register(se, sp, ed2, pl2, IDENT, rp, want_detail=True)
    assert_same_registration(ctx.register_to_map(m, se, sp, IDENT, rp, want_detail=True), full)
    ed, pl = ed2, pl2
    # sliding window: keep the newest points only; indices are relative to the new front
    m.update(np.zeros((0, 3)), np.zeros((0, 3)), max_edge=120, max_planar=6000)
    assert m.size() == (120, 6000)
    win = ctx.register(se, sp, ed[-120:], pl[-6000:], IDENT, rp, want_detail=True)
    assert_same_registration(ctx.register_to_map(m, se, sp, IDENT, rp, want_detail=True), win)
    m.update(ed[:7], pl[:11], max_edge=120, max_planar=6000)  # append + evict in one call, no pose
    exp_e, exp_p = np.concatenate([ed[-120:], ed[:7]])[-120:], np.concatenate([pl[-6000:], pl[:11]])[-6000:]
    assert_same_registration(ctx.register_to_map(m, se, sp, IDENT, rp, want_detail=True),
                             ctx.register(se, sp, exp_e, exp_p, IDENT, rp, want_detail=True))
    m.close()


def test_scan_to_local_map_matches_oracle(ctx, oracle):
    """Config 5 in small: target = features of 12 consecutive scans moved into the frame of scan 0 by the
    ground-truth poses (rounded to float32), source = the next scan, initial estimate = the previous
    scan's ground-truth pose.  ~200k planar target points -> multi-CTA build on the default context."""
    R, P, n_map = 64, 1024, 12
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    te, tp = [], []
    for k in range(n_map):
        s = synth.make_scan(R, P, k=k)[:, :3].astype(np.float64)
        e, p = oracle.extract(s, lp, fe)
        T = synth.relative_pose(0, k)
        te.append(H.transform(s[e], T).astype(np.float32).astype(np.float64))
        tp.append(H.transform(s[p], T).astype(np.float32).astype(np.float64))
    te, tp = np.concatenate(te), np.concatenate(tp)
    assert len(tp) > 150_000
    s = synth.make_scan(R, P, k=n_map)[:, :3].astype(np.float64)
    e, p = oracle.extract(s, lp, fe)
    init = synth.relative_pose(0, n_map - 1)
    rp = RegParams.default()
    ref = oracle.register(s[e], s[p], te, tp, init, rp, want_detail=True)
    got = ctx.register(s[e], s[p], te, tp, init, H.to_capi(rp), want_detail=True)
    assert_same_registration(got, ref, exact=False)
    m = ctx.map_create(te, tp)
    assert_same_registration(ctx.register_to_map(m, s[e], s[p], init, H.to_capi(rp), want_detail=True), got)
    m.close()
    # and the answer is the motion: within a few mm / 1e-3 rad of the ground truth
    gt = synth.relative_pose(0, n_map)
    assert H.angular_distance(got[0][:4], gt[:4]) < 2e-3 and np.abs(got[0][4:] - gt[4:]).max() < 2e-2


def test_python_module_local_map(ctx, scene):
    import loam_b200 as loam
    ed, pl = scene
    sTt = H.REG_SCENARIOS[0][1]
    src = loam.LoamFeatures(H.transform(ed, sTt), H.transform(pl, sTt))
    lmap = loam.LocalMap(loam.LoamFeatures(ed[:50], pl[:5000]))
    lmap.insert(loam.LoamFeatures(ed[50:], pl[5000:]), loam.Pose3d.Identity())
    assert lmap.size() == (len(ed), len(pl))
    d_map, d_plain = loam.RegistrationDetail(), loam.RegistrationDetail()
    via_map = loam.registerFeatures(src, lmap, loam.Pose3d(), loam.RegistrationParams(), d_map)
    plain = loam.registerFeatures(src, loam.LoamFeatures(ed, pl), loam.Pose3d(), loam.RegistrationParams(), d_plain)
    assert np.array_equal(via_map._to7(), plain._to7())
    assert d_map.termination_type == d_plain.termination_type == loam.CONVERGED
    assert d_map.iteration_info[0].plane_associations == d_plain.iteration_info[0].plane_associations
