"""Fuzz (-m gpu): k-NN lists against brute force and registerFeatures against the oracle on random scenes."""
import numpy as np
import pytest

import helpers as H
from oracle.pyoracle import RegParams

pytestmark = pytest.mark.gpu
IDENT = np.array([0, 0, 0, 1, 0, 0, 0.0])


def random_cloud(rng, n):
    kind = rng.randint(0, 5)
    if kind == 0:
        pts = rng.uniform(-5, 5, (n, 3))
    elif kind == 1:  # clusters of very different density
        c = rng.uniform(-5, 5, (6, 3))
        pts = c[rng.randint(0, 6, n)] + rng.normal(0, 1, (n, 3)) * rng.choice([0.01, 0.1, 1.0], (n, 1))
    elif kind == 2:  # coplanar + collinear parts
        pts = np.c_[rng.uniform(-5, 5, (n, 2)), np.zeros(n)]
        pts[: n // 3] = np.c_[np.linspace(-3, 3, n // 3), np.zeros((n // 3, 2))]
    elif kind == 3:  # quantised: many exactly equal distances (ties -> ascending index)
        pts = np.round(rng.uniform(-3, 3, (n, 3)) * 2) / 2
    else:  # duplicates
        base = rng.uniform(-5, 5, (max(n // 4, 1), 3))
        pts = base[rng.randint(0, len(base), n)]
    return pts


def test_knn_fuzz_vs_brute_force(ctx, oracle):
    for seed in range(60):
        rng = np.random.RandomState(1000 + seed)
        n = int(rng.choice([1, 2, 7, 8, 9, 17, 100, 1000, 5000]))
        pts = random_cloud(rng, n)
        q = np.concatenate([pts[rng.randint(0, n, 40)] + rng.normal(0, 0.05, (40, 3)), rng.uniform(-8, 8, (10, 3)),
                            pts[rng.randint(0, n, 5)]])  # the last ones coincide with target points
        k = int(rng.choice([1, 3, 5, 8, 12]))
        md = float(rng.choice([-1.0, 0.05, 0.3, 1.0, 3.0]))
        idx, cnt = ctx.knn(pts, q, k, md)
        for i in range(len(q)):
            bi, _ = oracle.knn_brute(pts, q[i], k, md)
            assert cnt[i] == len(bi) and np.array_equal(idx[i, :cnt[i]], bi), (seed, i)


def random_scene(rng):
    """Planes and vertical / horizontal poles sampled with noise: planar and edge feature sets of a small room."""
    planes, edges = [], []
    for _ in range(int(rng.randint(2, 5))):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        o = n * rng.uniform(2, 6)
        u = np.cross(n, [0.3, 0.5, 0.8])
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        s = rng.uniform(-2, 2, (int(rng.randint(300, 1500)), 2))
        planes.append(o + s[:, :1] * u + s[:, 1:] * v + rng.normal(0, 0.004, (len(s), 3)))
    for _ in range(int(rng.randint(1, 4))):
        a, d = rng.uniform(-4, 4, 3), rng.normal(size=3)
        d /= np.linalg.norm(d)
        t = np.linspace(-1.5, 1.5, int(rng.randint(40, 120)))
        edges.append(a + t[:, None] * d + rng.normal(0, 0.004, (len(t), 3)))
    return np.concatenate(edges), np.concatenate(planes)


def test_register_fuzz_vs_oracle(ctx, oracle):
    checked = 0
    for seed in range(24):
        rng = np.random.RandomState(5000 + seed)
        ed, pl = random_scene(rng)
        sTt = np.r_[H.axis_angle(rng.uniform(0, 0.06), rng.normal(size=3)), rng.uniform(-0.08, 0.08, 3)]
        se, sp = H.transform(ed, sTt) + rng.normal(0, 0.003, ed.shape), H.transform(pl, sTt) + rng.normal(0, 0.003, pl.shape)
        rp = RegParams.default()
        rp.num_plane_neighbors = int(rng.choice([5, 5, 6, 8]))
        rp.num_edge_neighbors = int(rng.choice([5, 3, 6]))
        rp.min_plane_fit_points = min(int(rng.choice([4, 5])), rp.num_plane_neighbors)
        rp.max_plane_neighbor_dist = float(rng.choice([2.0, 0.5, 1.0]))
        rp.max_iterations = int(rng.choice([10, 10, 3, 1]))
        rp.min_associations = int(rng.choice([100, 100, 10]))
        init = IDENT if seed % 3 else np.r_[H.axis_angle(0.01, [0, 0, 1]), [0.01, -0.02, 0.0]]
        pose, det = ctx.register(se, sp, ed, pl, init, H.to_capi(rp), want_detail=True)
        po, do = oracle.register(se, sp, ed, pl, init, rp, want_detail=True)
        assert det["termination"] == do.termination and det["n_iters"] == do.n_iters, seed
        assert np.array_equal(det["lm_iters"], do.lm_iters), seed
        for i in range(do.n_iters):
            assert np.array_equal(det["edge_assoc"][i], do.edge_assoc[i]), (seed, i)
            assert np.array_equal(det["plane_assoc"][i], do.plane_assoc[i]), (seed, i)
        assert H.angular_distance(po[:4], pose[:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - pose[4:]).max() < H.POSE_TOL_M, seed
        checked += do.n_iters
    assert checked > 24


def test_batched_walk_fuzz_vs_single_calls(ctx):
    """>= 16 pairs in one call take the batched shared-memory k-NN walk (compact records, integer box tests); single
    registrations take the general walk over the float node boxes.  Both must find the same neighbours — on noisy rooms
    and on quantised / duplicate-riddled sets where equal distances are everywhere and only the index tie-break decides."""
    rp = H.to_capi(RegParams.default())
    pairs, inits = [], []
    for seed in range(20):
        rng = np.random.RandomState(7000 + seed)
        ed, pl = random_scene(rng)
        if seed % 5 == 3:  # quantised coordinates: ties
            ed, pl = np.round(ed * 8) / 8, np.round(pl * 8) / 8
        if seed % 5 == 4:  # duplicated points
            pl = np.concatenate([pl, pl[rng.randint(0, len(pl), len(pl) // 3)]])
        sTt = np.r_[H.axis_angle(rng.uniform(0, 0.05), rng.normal(size=3)), rng.uniform(-0.06, 0.06, 3)]
        pairs.append((H.transform(ed, sTt) + rng.normal(0, 0.002, ed.shape),
                      H.transform(pl, sTt) + rng.normal(0, 0.002, pl.shape), ed, pl))
        inits.append(IDENT)
    poses, term, its = ctx.register_pairs(pairs, np.array(inits), rp)
    for k, pr in enumerate(pairs):
        single, det = ctx.register(*pr, IDENT, rp, want_detail=True)
        assert term[k] == det["termination"] and its[k] == det["n_iters"], k
        assert H.angular_distance(single[:4], poses[k][:4]) < 1e-9 and np.abs(single[4:] - poses[k][4:]).max() < 1e-9, k
