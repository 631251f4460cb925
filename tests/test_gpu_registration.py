"""GPU parity (through the C-ABI): kNN index lists, associations and poses of registerFeatures."""
import numpy as np
import pytest

import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, RegParams

pytestmark = pytest.mark.gpu
IDENT = np.array([0, 0, 0, 1, 0, 0, 0.0])


@pytest.fixture(scope="module")
def scene():
    return H.simple_scene()


def check_knn(ctx, oracle, targets, queries, k, max_dist):
    idx, cnt = ctx.knn(targets, queries, k, max_dist)
    for i, q in enumerate(queries):
        bi, _ = oracle.knn_brute(targets, q, k, max_dist)
        assert cnt[i] == len(bi), (i, cnt[i], len(bi))
        assert np.array_equal(idx[i, :cnt[i]], bi), i
        assert (idx[i, cnt[i]:] == 0xFFFFFFFF).all()


@pytest.mark.parametrize("k,max_dist", [(5, 2.0), (5, 1.0), (5, 0.15), (1, 2.0), (8, 0.5), (5, -1.0), (5, 0.0), (12, 1.0), (32, 3.0)])
def test_knn_full_lists_exact_vs_brute_force(ctx, oracle, scene, k, max_dist):
    ed, pl = scene
    rng = np.random.RandomState(k)
    q = np.concatenate([pl[rng.choice(len(pl), 300)] + rng.normal(0, 0.07, (300, 3)),
                        rng.uniform(-6, 8, (60, 3)),          # far from the surfaces / outside the bbox
                        np.array([[1e3, -1e3, 50.0]])])         # far outside
    check_knn(ctx, oracle, pl, q, k, max_dist)
    check_knn(ctx, oracle, ed, q[::4], k, max_dist)


def test_knn_on_lidar_feature_sets(ctx, oracle):
    R, P = 64, 1024
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    s0 = synth.make_scan(R, P, k=0)[:, :3].astype(np.float64)
    s1 = synth.make_scan(R, P, k=1)[:, :3].astype(np.float64)
    e0, p0 = oracle.extract(s0, lp, fe)
    e1, p1 = oracle.extract(s1, lp, fe)
    check_knn(ctx, oracle, s0[p0], s1[p1][::7], 5, 2.0)
    check_knn(ctx, oracle, s0[e0], s1[e1][::3], 5, 1.0)


def test_knn_degenerate_targets(ctx, oracle):
    q = np.array([[0.1, 0.2, 0.3], [5, 5, 5.0]])
    idx, cnt = ctx.knn(np.zeros((0, 3)), q, 5, 1.0)          # empty target (test_registration.cpp:177-199 has empty edges)
    assert (cnt == 0).all()
    check_knn(ctx, oracle, np.array([[0.0, 0, 0]]), q, 5, 1.0)             # single point
    check_knn(ctx, oracle, np.tile([[1.0, 2, 3]], (40, 1)), q, 5, -1.0)    # all identical: ties -> ascending index
    line = np.c_[np.zeros(50), np.zeros(50), np.linspace(0, 1, 50)]       # zero-extent bbox on two axes
    check_knn(ctx, oracle, line, q, 5, 2.0)


@pytest.mark.parametrize("case", H.REG_SCENARIOS, ids=[c[0] for c in H.REG_SCENARIOS])
def test_reference_registration_scenarios(ctx, oracle, scene, case):
    name, sTt, init, max_it, rtol, ttol = case
    ed, pl = scene
    rp = RegParams.default()
    rp.max_iterations = max_it
    init7 = IDENT if init is None else init
    se, sp = H.transform(ed, sTt), H.transform(pl, sTt)
    pose, det = ctx.register(se, sp, ed, pl, init7, H.to_capi(rp), want_detail=True)
    # (1) the reference's own assertion: ground truth within its tolerance
    ang, t = H.pose_error(sTt, pose)
    assert ang < rtol and np.all(np.abs(t) < ttol)
    # (2) parity with the oracle: poses within 1e-6 rad / 1e-5 m, correspondences bit-exact
    po, do = oracle.register(se, sp, ed, pl, init7, rp, want_detail=True)
    assert H.angular_distance(po[:4], pose[:4]) < H.POSE_TOL_RAD
    assert np.abs(po[4:] - pose[4:]).max() < H.POSE_TOL_M
    assert det["n_iters"] == do.n_iters and det["termination"] == do.termination
    assert np.array_equal(det["lm_iters"], do.lm_iters)
    for i in range(do.n_iters):
        assert np.array_equal(det["edge_assoc"][i], do.edge_assoc[i]), ("edge assoc", i)
        assert np.array_equal(det["plane_assoc"][i], do.plane_assoc[i]), ("plane assoc", i)
        assert H.angular_distance(det["iter_update"][i][:4], do.iter_update[i][:4]) < H.POSE_TOL_RAD
        assert np.abs(det["iter_est"][i] - do.iter_est[i]).max() < H.POSE_TOL_M
    np.testing.assert_allclose(det["lm_cost"], do.lm_cost, rtol=1e-9, atol=1e-18)


def test_planar_only_self_registration(ctx, scene):  # NonStandardAllocator: empty edge sets
    _, pl = scene
    pl1 = pl[:3600]
    e = np.zeros((0, 3))
    pose, det = ctx.register(e, pl1, e, pl1, IDENT, _capi.default_reg_params(), want_detail=True)
    assert H.angular_distance(pose[:4], IDENT[:4]) < 1e-4 and np.all(np.abs(pose[4:]) < 1e-3)
    assert det["termination"] == 0 and det["n_iters"] == 1


def test_insufficient_associations_and_zero_iterations(ctx, scene):
    ed, pl = scene
    init = np.r_[H.axis_angle(0.3, [0, 1, 0]), [1.0, 2.0, 3.0]]
    pose, det = ctx.register(ed + 100.0, pl + 100.0, ed, pl, init, _capi.default_reg_params(), want_detail=True)
    assert det["termination"] == 2 and det["n_iters"] == 0
    assert np.array_equal(pose, init)  # estimate unchanged (registration-inl.h:45-48)
    rp = _capi.default_reg_params()
    rp.max_iterations = 0
    pose, det = ctx.register(ed, pl, ed, pl, init, rp, want_detail=True)
    assert det["termination"] == 1 and det["n_iters"] == 0 and np.array_equal(pose, init)


@pytest.mark.parametrize("kw", [dict(num_plane_neighbors=8, min_plane_fit_points=5, max_plane_neighbor_dist=1.0),
                                dict(num_edge_neighbors=3, min_line_fit_points=2, max_edge_neighbor_dist=0.0),
                                dict(num_plane_neighbors=12, num_edge_neighbors=10, max_iterations=3),
                                dict(min_associations=20000)])
def test_parameter_variants_match_oracle(ctx, oracle, scene, kw):
    ed, pl = scene
    sTt = H.REG_SCENARIOS[1][1]
    rp = RegParams.default()
    for k, v in kw.items():
        setattr(rp, k, v)
    se, sp = H.transform(ed, sTt), H.transform(pl, sTt)
    pose, det = ctx.register(se, sp, ed, pl, IDENT, H.to_capi(rp), want_detail=True)
    po, do = oracle.register(se, sp, ed, pl, IDENT, rp, want_detail=True)
    assert det["termination"] == do.termination and det["n_iters"] == do.n_iters
    assert H.angular_distance(po[:4], pose[:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - pose[4:]).max() < H.POSE_TOL_M
    for i in range(do.n_iters):
        assert np.array_equal(det["edge_assoc"][i], do.edge_assoc[i])
        assert np.array_equal(det["plane_assoc"][i], do.plane_assoc[i])


@pytest.mark.parametrize("shape", [(16, 1800), (64, 1024)])
def test_scan_to_scan_on_lidar_shapes(ctx, oracle, shape):
    R, P = shape
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    s0 = synth.make_scan(R, P, k=10)[:, :3].astype(np.float64)
    s1 = synth.make_scan(R, P, k=11)[:, :3].astype(np.float64)
    e0, p0 = oracle.extract(s0, lp, fe)
    e1, p1 = oracle.extract(s1, lp, fe)
    pose, det = ctx.register(s1[e1], s1[p1], s0[e0], s0[p0], IDENT, _capi.default_reg_params(), want_detail=True)
    po, do = oracle.register(s1[e1], s1[p1], s0[e0], s0[p0], IDENT, RegParams.default(), want_detail=True)
    assert det["n_iters"] == do.n_iters and det["termination"] == do.termination
    assert H.angular_distance(po[:4], pose[:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - pose[4:]).max() < H.POSE_TOL_M
    for i in range(do.n_iters):
        assert np.array_equal(det["edge_assoc"][i], do.edge_assoc[i])
        assert np.array_equal(det["plane_assoc"][i], do.plane_assoc[i])


def test_python_module_mirror(ctx, oracle, scene):
    import loam_b200 as loam
    ed, pl = scene
    sTt = H.REG_SCENARIOS[0][1]
    src = loam.LoamFeatures(H.transform(ed, sTt), H.transform(pl, sTt))
    tgt = loam.LoamFeatures(ed, pl)
    detail = loam.RegistrationDetail()
    pose = loam.registerFeatures(src, tgt, loam.Pose3d(), loam.RegistrationParams(), detail)
    ang, t = H.pose_error(sTt, pose._to7())
    assert ang < 1e-4 and np.all(np.abs(t) < 1e-4)
    _, do = oracle.register(H.transform(ed, sTt), H.transform(pl, sTt), ed, pl, want_detail=True)
    assert detail.termination_type == loam.CONVERGED and len(detail.iteration_info) == do.n_iters >= 2
    info = detail.iteration_info[0]
    assert len(info.plane_associations) > 8000 and isinstance(info.edge_associations[0], tuple)
    # left composition of each recorded update reproduces the next recorded estimate (registration-inl.h:65)
    nxt = info.estimate_update.compose(info.target_T_source_init)
    np.testing.assert_allclose(nxt._to7(), detail.iteration_info[1].target_T_source_init._to7(), atol=1e-15)


def test_two_contexts_on_two_host_threads_agree(scene):
    """The reference is stateless and re-entrant (callers may invoke it from many threads); here that
    means one loamgpu context per thread.  Two threads, each with its own context and stream, register different
    problems concurrently; every result equals the single-threaded one bit for bit."""
    import threading
    ed, pl = scene
    rp = _capi.default_reg_params()
    problems = [(H.transform(ed, c[1]), H.transform(pl, c[1]), ed, pl) for c in H.REG_SCENARIOS[:4]]
    solo = _capi.Context(0)
    expect = [solo.register(*p, IDENT, rp) for p in problems]
    solo.close()
    results, errors = {}, []

    def worker(tid):
        try:
            c = _capi.Context(0)
            for rep in range(3):
                for i, p in enumerate(problems):
                    j = (i + tid) % len(problems)
                    results[(tid, rep, j)] = c.register(*problems[j], IDENT, rp)
                    e, pidx = c.extract(synth.make_scan(16, 512, k=tid), _capi.CLidarParams(16, 512, 1.0, 120.0),
                                        _capi.default_fe_params())
                    results[(tid, rep, "n_feat")] = (len(e), len(pidx))
            c.close()
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (tid, rep, j), v in results.items():
        if j == "n_feat":
            assert v == results[(tid, 0, "n_feat")]
        else:
            assert np.array_equal(v, expect[j]), (tid, rep, j)
