"""Writes registration_inputs.txt: the cases make_registration_golden.cpp feeds to the REAL reference.

Cases: the reference's own six registration scenes (tests/test_registration.cpp:69-199) and two scan-to-scan pairs of
the synthetic sequence (features extracted by the reference's own feature code through oracle/_ref when it is built,
else by the oracle port — the same indices either way, tests/test_oracle_features.py).  Every number is written with
17 significant digits, so the C++ side reads back bit-identical doubles.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import helpers as H  # noqa: E402
from loam_b200 import synth  # noqa: E402
from oracle.pyoracle import FeParams, LidarParams, Oracle, RegParams  # noqa: E402

PARAM_FIELDS = [f for f, _ in RegParams._fields_]


def cases():
    ed, pl = H.simple_scene()
    ident = np.array([0, 0, 0, 1, 0, 0, 0.0])
    for name, sTt, init, max_it, *_ in H.REG_SCENARIOS:
        rp = RegParams.default()
        rp.max_iterations = max_it
        yield name, rp, ident if init is None else init, H.transform(ed, sTt), H.transform(pl, sTt), ed, pl
    e = np.zeros((0, 3))
    yield "planar_only_self", RegParams.default(), ident, e, pl[:3600], e, pl[:3600]
    orc = Oracle()
    R, P = 32, 512
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    for k in (0, 7):
        scans = [synth.make_scan(R, P, k=k + j)[:, :3].astype(np.float64) for j in (0, 1)]
        f = [orc.extract(s, lp, fe) for s in scans]
        yield (f"synthetic_pair_{k}", RegParams.default(), ident, scans[1][f[1][0]], scans[1][f[1][1]],
               scans[0][f[0][0]], scans[0][f[0][1]])


def main(path=os.path.join(HERE, "registration_inputs.txt")):
    all_cases = list(cases())
    with open(path, "w") as f:
        f.write(f"cases {len(all_cases)}\n")
        for name, rp, init, se, sp, te, tp in all_cases:
            f.write(f"case {name}\n")
            f.write("params " + " ".join(repr(getattr(rp, k)) for k in PARAM_FIELDS) + "\n")
            f.write("init " + " ".join(f"{v:.17g}" for v in init) + "\n")
            for tag, cloud in (("source_edge", se), ("source_planar", sp), ("target_edge", te), ("target_planar", tp)):
                f.write(f"{tag} {len(cloud)}\n")
                for p in cloud:
                    f.write(f"{p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n")
    print(f"wrote {len(all_cases)} cases to {path}")


if __name__ == "__main__":
    main(*sys.argv[1:])
