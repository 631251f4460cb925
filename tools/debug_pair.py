"""Debug helper (GPU box): compare single-pair register / odometry against the oracle for one scan pair."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, Oracle, RegParams

k0 = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
R, P = 64, 1024
scans = np.stack([synth.make_scan(R, P, k=k0 + k) for k in range(n)])
lp, fe, rp = LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()
orc = Oracle()
ctx = _capi.Context(0)
feats = []
for s in scans:
    xyz = s[:, :3].astype(np.float64)
    e, p = orc.extract(xyz, lp, fe)
    feats.append((xyz[e], xyz[p]))
poses, term, its, ne, npl = ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp))
for k in range(n - 1):
    po, do = orc.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp, want_detail=True)
    pg, dg = ctx.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], H.IDENT if hasattr(H, "IDENT") else np.array([0, 0, 0, 1, 0, 0, 0.0]), H.to_capi(rp), want_detail=True)
    print("pair", k, "oracle iters", do.n_iters, "gpu single iters", dg["n_iters"], "odo iters", its[k])
    print("  ang(oracle, single) =", H.angular_distance(po[:4], pg[:4]), " ang(oracle, odo) =", H.angular_distance(po[:4], poses[k][:4]))
    for it in range(min(do.n_iters, dg["n_iters"])):
        ea, eb = do.edge_assoc[it], dg["edge_assoc"][it]
        pa, pb = do.plane_assoc[it], dg["plane_assoc"][it]
        print("  iter", it, "edge assoc equal", np.array_equal(ea, eb), len(ea), len(eb), "plane equal", np.array_equal(pa, pb), len(pa), len(pb))

# which associations are missing at iteration 0 of pair 0?
k = 0
po, do = orc.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], None, rp, want_detail=True)
pg, dg = ctx.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], np.array([0, 0, 0, 1, 0, 0, 0.0]), H.to_capi(rp), want_detail=True)
a = {int(r[0]): int(r[1]) for r in do.plane_assoc[0]}
b = {int(r[0]): int(r[1]) for r in dg["plane_assoc"][0]}
missing = sorted(set(a) - set(b))
print("missing", len(missing), missing[:20])
extra = sorted(set(b) - set(a))
print("extra", len(extra))
diff = [i for i in a if i in b and a[i] != b[i]]
print("different nearest", len(diff))
idx, cnt = ctx.knn(feats[0][1], feats[1][1][missing[:10]], 5, 2.0)
print(cnt, idx)
mi = np.array(missing)
print("missing idx stats: min", mi.min(), "max", mi.max(), "n_src", len(feats[1][1]), "n_tgt", len(feats[0][1]))
print(np.diff(mi)[:40])
