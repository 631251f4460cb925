"""Debug helper (GPU box): loamgpu_knn vs brute force on real lidar feature sets."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, Oracle, RegParams
from scipy.spatial import cKDTree

k0 = int(sys.argv[1]) if len(sys.argv) > 1 else 40
R, P = 64, 1024
lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
orc = Oracle()
ctx = _capi.Context(0)
s0 = synth.make_scan(R, P, k=k0)[:, :3].astype(np.float64)
s1 = synth.make_scan(R, P, k=k0 + 1)[:, :3].astype(np.float64)
e0, p0 = orc.extract(s0, lp, fe)
e1, p1 = orc.extract(s1, lp, fe)
for name, T, Q, r in (("planar", s0[p0], s1[p1], 2.0), ("edge", s0[e0], s1[e1], 1.0)):
    idx, cnt = ctx.knn(T, Q, 5, r)
    tree = cKDTree(T)
    d, ii = tree.query(Q, k=5, distance_upper_bound=r)
    ref_cnt = (d < r).sum(1)
    bad = np.nonzero(cnt != ref_cnt)[0]
    print(name, "n", len(T), "queries", len(Q), "count mismatches", len(bad))
    for b in bad[:5]:
        print("  q", b, Q[b], "gpu cnt", cnt[b], idx[b], "ref cnt", ref_cnt[b], ii[b], d[b])
    same = sum(np.array_equal(np.sort(idx[i, :cnt[i]]), np.sort(ii[i, :ref_cnt[i]])) for i in range(len(Q)))
    print("  identical neighbour sets:", same, "/", len(Q))
