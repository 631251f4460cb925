"""Residual-block builders shared by the Jacobian / solver parity tests (CPU and GPU).

A block is what associateEdges / associatePlanes hand to Ceres (registration.cpp:52-57, 93-98): the source point
ALREADY transformed by the current estimate, plus the fitted line (a, b = centre +- 0.1 dir, geometry.cpp:57) or
plane (unit normal, d).  Arrays: is_plane [M] int32, P [M,3], A [M,3] (a | normal), B [M,3] (b | (d, 0, 0)).
"""
import numpy as np

import helpers as H


def random_blocks(seed, n_edge, n_plane, noise=0.05, outliers=0.1, motion=None):
    """Blocks around a box-room-like cloud: residuals of ~`noise` metres, a fraction `outliers` beyond Huber's 1 m."""
    rng = np.random.RandomState(seed)
    M = n_edge + n_plane
    is_plane = np.r_[np.zeros(n_edge, np.int32), np.ones(n_plane, np.int32)]
    P = rng.uniform(-15, 15, (M, 3))
    A, B = np.zeros((M, 3)), np.zeros((M, 3))
    off = rng.normal(0, noise, M)
    far = rng.uniform(size=M) < outliers
    off[far] = rng.uniform(1.2, 3.0, far.sum()) * rng.choice([-1, 1], far.sum())
    for i in range(M):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if is_plane[i]:
            A[i] = d                           # normal
            B[i, 0] = d @ P[i] - off[i]        # d: the point sits `off` from the plane
        else:
            o = np.cross(d, rng.normal(size=3))
            o /= np.linalg.norm(o)
            c = P[i] + off[i] * o + rng.uniform(-0.3, 0.3) * d   # line centre: |off| from the point
            A[i], B[i] = c + 0.1 * d, c - 0.1 * d
    if motion is not None:  # the cloud seen from a moved sensor: the solve has something to recover
        P = H.transform(P, motion)
    return is_plane, P, A, B


def scene_blocks(oracle, source_T_target, every=6):
    """Blocks of the first outer iteration of a reference registration scenario (tests/test_registration.cpp:8-87),
    built with the oracle's own kNN and fits from every `every`-th source point, identity initial estimate."""
    ed, pl = H.simple_scene()
    se, sp = H.transform(ed, source_T_target), H.transform(pl, source_T_target)
    kinds, P, A, B = [], [], [], []
    for src, tgt, k, md, need, plane in ((se, ed, 5, 1.0, 3, 0), (sp[::every], pl, 5, 2.0, 4, 1)):
        idx, cnt = oracle.knn_tree_batch(tgt, src, k, md)
        for i, q in enumerate(src):
            if cnt[i] < need:
                continue
            nb = tgt[idx[i, :cnt[i]]]
            if plane:
                n, d, avg = oracle.fit_plane(nb)
                if avg > 0.1:
                    continue
                A.append(n)
                B.append([d, 0, 0])
            else:
                a, b, _ = oracle.fit_line(nb)
                A.append(a)
                B.append(b)
            kinds.append(plane)
            P.append(q)
    return np.array(kinds, np.int32), np.array(P), np.array(A), np.array(B)
