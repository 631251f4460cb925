#!/usr/bin/env python
"""Latency of ONE extractFeatures / registerFeatures call through the C-ABI with host buffers (BASELINE.json config 1:
a single 64x1024 scan pair, the reference's own usage pattern, README.md:46-59).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import _capi, synth  # noqa: E402


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, float(np.median(ts)) * 1e3, float(np.min(ts)) * 1e3


def main():
    out = {"config": "single scan pair through the host-buffer C-ABI calls (one call = one reference call)"}
    ctx = _capi.Context(0)
    for R, P in ((64, 1024), (16, 1800), (128, 2048)):
        lp = _capi.CLidarParams(R, P, 1.0, 120.0)
        fe, rp = _capi.default_fe_params(), _capi.default_reg_params()
        s0, s1 = synth.make_scan(R, P, k=0), synth.make_scan(R, P, k=1)
        (e0, p0), ext_ms, ext_best = timed(lambda: ctx.extract(s0, lp, fe))
        e1, p1 = ctx.extract(s1, lp, fe)
        x0, x1 = s0[:, :3].astype(np.float64), s1[:, :3].astype(np.float64)
        args = (x1[e1], x1[p1], x0[e0], x0[p0], np.array([0, 0, 0, 1, 0, 0, 0.0]), rp)
        before = ctx.launch_count
        (pose, det), reg_ms, reg_best = timed(lambda: ctx.register(*args, want_detail=True), 30)
        launches = (ctx.launch_count - before) // 35
        _, reg2_ms, reg2_best = timed(lambda: ctx.register(*args), 30)
        out[f"{R}x{P}"] = {"extract_ms": {"median": ext_ms, "best": ext_best},
                           "register_ms": {"median": reg2_ms, "best": reg2_best, "with_detail_median": reg_ms,
                                           "outer_iterations": int(det["n_iters"]), "kernel_launches": int(launches)},
                           "features": [int(len(e0)), int(len(p0))]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
