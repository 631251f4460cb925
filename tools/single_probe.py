#!/usr/bin/env python
"""Where a single loamgpu_extract / loamgpu_register call spends its time: wall clock per call next to the sum of its
kernel times (CUDA events per launch).  One GPU, 64x1024."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import _capi, synth  # noqa: E402


def main():
    R, P = 64, 1024
    scans = [np.ascontiguousarray(synth.make_scan(R, P, k), dtype=np.float32) for k in range(3)]
    ctx = _capi.Context(0)
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    out = {}
    feats = []
    for prof in (False, True):
        ctx.set_profiling(prof)
        for s in scans[:2]:
            ctx.extract(s, lp, fe)
        ctx.kernel_times()
        t0 = time.perf_counter()
        n = 200
        for i in range(n):
            e, p = ctx.extract(scans[i % 2], lp, fe)
        dt = (time.perf_counter() - t0) / n * 1e3
        kt = ctx.kernel_times()
        out["extract_prof%d" % prof] = {"wall_ms": dt, "kernel_ms": {k: v[0] / n for k, v in kt.items() if v[1]} if prof else None}
    for s in scans[:2]:
        e, p = ctx.extract(s, lp, fe)
        xyz = s[:, :3].astype(np.float64)
        feats.append((np.ascontiguousarray(xyz[e]), np.ascontiguousarray(xyz[p])))
    init = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    for prof in (False, True):
        ctx.set_profiling(prof)
        for _ in range(3):
            ctx.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], init, rp)
        ctx.kernel_times()
        t0 = time.perf_counter()
        n = 200
        for i in range(n):
            r = ctx.register(feats[1][0], feats[1][1], feats[0][0], feats[0][1], init, rp)
        dt = (time.perf_counter() - t0) / n * 1e3
        kt = ctx.kernel_times()
        out["register_prof%d" % prof] = {"wall_ms": dt, "kernel_ms": {k: v[0] / n for k, v in kt.items() if v[1]} if prof else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
