"""One tiny pass through every entry-point family (and so every kernel) of libloamgpu.so on cuda:0, with cheap
consistency checks: single-scan extraction (bulk-copy and strided staging, de-warped staging), curvature / mask,
sequence odometry (plain and with per-sweep motions), single registration (clustered LM), explicit batches, the k-NN
entry, the device-resident map (multi-CTA NN build).  Sizes are small (16x512 scans).

Written to run under compute-sanitizer (memcheck / racecheck / synccheck); on this pool the tool is closed to jobs
("runs under it have left GPUs needing a reset"), so tools/gpu_check.sh runs it plain as a smoke test."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loam_b200 import _capi, synth  # noqa: E402


def main():
    R, P, n = 16, 512, 4
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    lp = _capi.CLidarParams(R, P, 1.0, 120.0)
    fe, rp = _capi.CFeParams(), _capi.CRegParams()
    lib = _capi.load_library()
    lib.loamgpu_default_fe_params(_capi.C.byref(fe))
    lib.loamgpu_default_reg_params(_capi.C.byref(rp))
    ctx = _capi.Context(0)
    before = ctx.launch_count

    e, p = ctx.extract(scans[0], lp, fe)                                   # TMA bulk staging
    e2, p2 = ctx.extract(scans[0][:, :3].astype(np.float64), lp, fe)       # f64x3 bulk staging
    odd = np.zeros((R * P, 5), dtype=np.float32)
    odd[:, :3] = scans[0][:, :3]
    e3, p3 = ctx.extract(odd, lp, fe)                                      # strided staging
    assert np.array_equal(e, e2) and np.array_equal(e, e3) and np.array_equal(p, p2) and np.array_equal(p, p3)
    ed, pd, moved = ctx.extract_dewarped(scans[0], lp, fe, [0, 0, 0.01, 0.99995, 0.1, 0.0, 0.0])
    assert moved.shape == (R * P, 3) and len(ed) and len(pd)
    ctx.curvature(scans[0], lp, fe)
    ctx.valid_mask(scans[0], lp, fe)

    poses, term, its, ne, npl = ctx.odometry_host(scans, lp, fe, rp)
    assert len(poses) == n - 1
    ident = np.tile(np.r_[0, 0, 0, 1.0, 0, 0, 0], (n, 1))
    poses_i = ctx.odometry_host(scans, lp, fe, rp, sweep_motions=ident)[0]
    assert np.array_equal(poses, poses_i), "identity sweep motions must not change anything"
    mo = np.stack([synth.relative_pose(k, k + 1) for k in range(n)])
    assert len(ctx.odometry_host(scans, lp, fe, rp, sweep_motions=mo)[0]) == n - 1

    feats = []
    for k in range(n):
        ek, pk = ctx.extract(scans[k], lp, fe)
        xyz = scans[k][:, :3].astype(np.float64)
        feats.append((xyz[ek], xyz[pk]))
    one = ctx.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], [0, 0, 0, 1, 0, 0, 0], rp, want_detail=True)
    assert np.allclose(one[0], poses[0], atol=1e-6)

    ctx.extract_batch(scans, lp, fe)
    pairs = [(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1]) for k in range(n - 1)]
    ctx.register_pairs(pairs, None, rp)
    ctx.knn(feats[0][1], feats[1][1][:200], 5, 1.0)

    m = ctx.map_create(feats[0][0], feats[0][1])
    m.update(feats[1][0], feats[1][1], pose=[0, 0, 0, 1, 0, 0, 0])
    ctx.register_to_map(m, feats[2][0], feats[2][1], [0, 0, 0, 1, 0, 0, 0], rp)
    m.close()
    print(f"entry smoke ok: {ctx.launch_count - before} launches")
    ctx.close()


if __name__ == "__main__":
    main()
