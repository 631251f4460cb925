"""Generate tests/golden/features_golden.npz from the REAL reference feature code.

Run in the dev container (where /root/reference exists):  python tests/golden/make_golden.py
Each case stores the float32 scan itself (so nothing depends on libm reproducibility across hosts),
the parameters, and what the reference's own extractFeatures / computeCurvature / computeValidPoints
return for it (through oracle/_ref/libloam_ref.so, see oracle/ref_shim.cpp).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loam_b200 import synth  # noqa: E402
from oracle.pyoracle import FeParams, LidarParams, RefLib  # noqa: E402

CASES = [
    # name, rings, cols, scan k, dropout, (N, S, maxE, maxP, edge_thr, planar_thr, occ, par), (min_range, max_range)
    ("r8x256_default", 8, 256, 0, 0.0, (3, 6, 10, 50, 100.0, 1.0, 0.5, 1.0), (1.0, 120.0)),
    ("r16x600_default", 16, 600, 7, 0.0, (3, 6, 10, 50, 100.0, 1.0, 0.5, 1.0), (1.0, 120.0)),
    ("r4x1030_ragged_n5_s4", 4, 1030, 11, 0.0, (5, 4, 4, 20, 50.0, 0.5, 0.3, 0.5), (1.0, 120.0)),
    ("r8x512_dropout_s7", 8, 512, 21, 0.02, (3, 7, 2, 9, 10.0, 2.0, 0.5, 1.0), (2.0, 14.0)),
]


def main():
    ref = RefLib()
    out = {}
    for name, R, P, k, drop, fe_t, rng_t in CASES:
        scan = synth.make_scan(R, P, k=k, dropout=drop)
        lp = LidarParams(R, P, *rng_t)
        fe = FeParams(*fe_t)
        xyz = scan[:, :3].astype(np.float64)
        e, p = ref.extract(xyz, lp, fe)
        c, _ = ref.curvature(xyz, lp, fe)
        m = ref.valid_mask(xyz, lp, fe)
        out[name + "/scan"] = scan[:, :3].copy()
        out[name + "/shape"] = np.array([R, P], dtype=np.int64)
        out[name + "/fe"] = np.array(fe_t, dtype=np.float64)
        out[name + "/range"] = np.array(rng_t, dtype=np.float64)
        out[name + "/edge"] = e
        out[name + "/planar"] = p
        out[name + "/curvature"] = c
        out[name + "/mask"] = m
        print(name, len(e), len(p))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "features_golden.npz"), **out)


if __name__ == "__main__":
    main()
