"""Generate tests/golden/dewarp_golden.npz: de-warp fixtures (extension).

The reference has no de-warp code, so there is nothing of the reference to capture for the de-warp itself; the fixture
pins (a) the definition — the moved points orc_dewarp produces for stored float32 scans and motions, so an accidental
change of the operation order shows up as a diff in bits — and (b) what the REAL reference feature code
(oracle/_ref/libloam_ref.so) extracts from those moved points, which is what loamgpu_extract_dewarped must return.

Run in the dev container (where /root/reference exists):  python tests/golden/make_dewarp_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from loam_b200 import synth  # noqa: E402
from oracle.pyoracle import FeParams, LidarParams, Oracle, RefLib  # noqa: E402

CASES = [
    # name, rings, cols, start pose (x, y, yaw), start_T_end
    ("r8x256_yaw", 8, 256, (0.5, -0.3, 0.2), np.r_[H.axis_angle(0.05, (0, 0, 1)), 0.4, 0.1, 0.0]),
    ("r16x600_far_hemisphere", 16, 600, (-1.0, 0.4, -0.1), np.r_[-H.axis_angle(0.03, (0, 0, 1)), -0.2, 0.15, 0.0]),
    ("r4x1030_general", 4, 1030, (0.0, 0.0, 0.0), np.r_[H.axis_angle(0.04, (0.3, -0.2, 1.0)), 0.2, 0.1, -0.05]),
]


def main():
    orc, ref = Oracle(), RefLib()
    fe = FeParams.default()
    out = {}
    for name, R, P, start, m in CASES:
        planar = m[0] == 0.0 and m[1] == 0.0 and m[6] == 0.0
        scan = (synth.make_warped_scan(R, P, 0, m, sigma=0.01, seed=5, start_pose=start) if planar
                else synth.make_scan(R, P, k=3))  # (the generator only smears planar motions)
        lp = LidarParams(R, P, 1.0, 120.0)
        moved = orc.dewarp(scan[:, :3].astype(np.float64), P, m)
        e, p = ref.extract(moved, lp, fe)
        out[name + "/scan"] = scan[:, :3].copy()
        out[name + "/shape"] = np.array([R, P], dtype=np.int64)
        out[name + "/motion"] = np.asarray(m, dtype=np.float64)
        out[name + "/moved"] = moved
        out[name + "/edge"] = e.astype(np.uint32)
        out[name + "/planar"] = p.astype(np.uint32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dewarp_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("/edge")})


if __name__ == "__main__":
    main()
