#!/usr/bin/env python
"""One-off soak on a GPU box: the fuzz generators of tests/ run far beyond the seeds the test suite uses.
usage: python tools/soak_gpu.py [n_feature_cases] [n_knn_cases] [n_register_cases]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_cases as FZ  # noqa: E402
import helpers as H  # noqa: E402
import test_fuzz_registration as FR  # noqa: E402
from loam_b200 import _capi  # noqa: E402
from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import FeParams, LidarParams, RegParams  # noqa: E402


def main():
    nf, nk, nr = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 3000), (2, 400), (3, 120)))
    ctx = _capi.Context(0)
    orc = pyoracle.Oracle()
    t0 = time.time()
    for seed in range(240, 240 + nf):
        pts, lp, fe = FZ.make_case(seed)
        lp, fe = LidarParams(*lp), FeParams(*fe)
        eo, po = orc.extract(pts, lp, fe)
        e, p = ctx.extract(pts, H.to_capi(lp), H.to_capi(fe))
        assert np.array_equal(e, eo) and np.array_equal(p, po), ("features", seed)
        f4 = np.zeros((len(pts), 4), dtype=np.float32)
        f4[:, :3] = pts
        if np.array_equal(f4[:, :3].astype(np.float64), pts):
            e4, p4 = ctx.extract(f4, H.to_capi(lp), H.to_capi(fe))
            assert np.array_equal(e4, eo) and np.array_equal(p4, po), ("features float4", seed)
    print(f"features: {nf} cases ok, {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    for seed in range(60, 60 + nk):
        rng = np.random.RandomState(1000 + seed)
        n = int(rng.choice([1, 2, 7, 8, 9, 17, 100, 1000, 5000]))
        pts = FR.random_cloud(rng, n)
        q = np.concatenate([pts[rng.randint(0, n, 40)] + rng.normal(0, 0.05, (40, 3)), rng.uniform(-8, 8, (10, 3)),
                            pts[rng.randint(0, n, 5)]])
        k = int(rng.choice([1, 3, 5, 8, 12]))
        md = float(rng.choice([-1.0, 0.05, 0.3, 1.0, 3.0]))
        idx, cnt = ctx.knn(pts, q, k, md)
        for i in range(len(q)):
            bi, _ = orc.knn_brute(pts, q[i], k, md)
            assert cnt[i] == len(bi) and np.array_equal(idx[i, :cnt[i]], bi), ("knn", seed, i)
    print(f"knn: {nk} cases ok, {time.time() - t0:.1f} s", flush=True)
    # batched (shared-memory) walk against single calls (general walk) and the oracle
    t0 = time.time()
    rp = RegParams.default()
    crp = H.to_capi(rp)
    ident = np.array([0, 0, 0, 1, 0, 0, 0.0])
    for base in range(0, nr, 24):
        pairs = []
        for seed in range(base, min(base + 24, nr)):
            rng = np.random.RandomState(9000 + seed)
            ed, pl = FR.random_scene(rng)
            if seed % 5 == 3:
                ed, pl = np.round(ed * 8) / 8, np.round(pl * 8) / 8
            if seed % 5 == 4:
                pl = np.concatenate([pl, pl[rng.randint(0, len(pl), len(pl) // 3)]])
            sTt = np.r_[H.axis_angle(rng.uniform(0, 0.05), rng.normal(size=3)), rng.uniform(-0.06, 0.06, 3)]
            pairs.append((H.transform(ed, sTt) + rng.normal(0, 0.002, ed.shape),
                          H.transform(pl, sTt) + rng.normal(0, 0.002, pl.shape), ed, pl))
        poses, term, its = ctx.register_pairs(pairs, np.tile(ident, (len(pairs), 1)), crp)
        for k, pr in enumerate(pairs):
            single, det = ctx.register(*pr, ident, crp, want_detail=True)
            assert term[k] == det["termination"] and its[k] == det["n_iters"], ("batch vs single", base + k)
            assert H.angular_distance(single[:4], poses[k][:4]) < 1e-9 and np.abs(single[4:] - poses[k][4:]).max() < 1e-9, (base + k)
            if k % 4 == 0:
                po, do = orc.register(*pr, ident, rp, want_detail=True)
                assert det["termination"] == do.termination and det["n_iters"] == do.n_iters, ("oracle", base + k)
                for i in range(do.n_iters):
                    assert np.array_equal(det["edge_assoc"][i], do.edge_assoc[i]), ("oracle edge assoc", base + k, i)
                    assert np.array_equal(det["plane_assoc"][i], do.plane_assoc[i]), ("oracle plane assoc", base + k, i)
                assert H.angular_distance(po[:4], single[:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - single[4:]).max() < H.POSE_TOL_M
    print(f"register: {nr} pairs ok (batched walk = single calls; every 4th against the oracle), {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
