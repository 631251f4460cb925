#!/usr/bin/env python
"""Cost of the fused de-warp: the 64x1024 sequence step of bench.py, device-resident, once plain
(loamgpu_odometry_device) and once with per-sweep motions (loamgpu_odometry_device_dewarped), with the per-kernel-class
CUDA-event times of both.  The scans are the regular synthetic ones and the motions the trajectory's own per-step
poses: the arithmetic does not depend on whether a scan is really smeared.  Prints one JSON line.

A separate de-warp pass before the path would read the float4 scan and write it back (or write doubles): 2 x 16 B
(or 16 + 24 B) per point on top of the extraction's own 16 B read; fused, the extra traffic is zero and the cost is the
per-point interpolation + rotation in the staging loop, which replaces the TMA bulk copy by per-thread loads."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import _capi, synth  # noqa: E402


def main():
    import torch

    R, P = 64, 1024
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps, warmup = 5, 3
    dev = torch.device("cuda", 0)
    lp = _capi.CLidarParams(R, P, 1.0, 120.0)
    fe, rp = _capi.default_fe_params(), _capi.default_reg_params()
    ctx = _capi.Context(0)
    d_scans = synth.make_scans_torch(R, P, 0, n, dev)
    motions = np.stack([synth.relative_pose(k, k + 1) for k in range(n)])
    d_mo = torch.from_numpy(motions).to(dev)
    d_pose = torch.zeros((n - 1, 7), dtype=torch.float64, device=dev)
    d_term = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    d_iter = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    d_ne = torch.zeros(n, dtype=torch.int32, device=dev)
    d_np = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    out_ptrs = (d_pose.data_ptr(), d_term.data_ptr(), d_iter.data_ptr(), d_ne.data_ptr(), d_np.data_ptr())

    def plain():
        ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, *out_ptrs)

    def dewarped():
        ctx.odometry_device_dewarped_ptr(d_scans.data_ptr(), n, d_mo.data_ptr(), lp, fe, rp, *out_ptrs)

    out = {"config": f"synthetic 64x1024 sequence, {n} scans/step, device-resident, {steps} steps after {warmup} warm-up",
           "l2": f"inputs larger than L2 ({n * R * P * 16 / 2**20:.0f} MiB of scans per step)"}
    for name, fn in (("plain", plain), ("dewarped", dewarped)):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ctx.kernel_times()
        ctx.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kt = ctx.kernel_times()
        ctx.set_profiling(False)
        out[name] = {"scans_per_s": n / (ms / 1e3), "ms_per_step": ms,
                     "kernel_ms_per_step": {k: v[0] / steps for k, v in kt.items()},
                     "converged": int((d_term.cpu().numpy() == 0).sum()),
                     "mean_outer_iterations": float(d_iter.cpu().numpy().mean())}
    ext_p, ext_d = out["plain"]["kernel_ms_per_step"]["extract"], out["dewarped"]["kernel_ms_per_step"]["extract"]
    out["extract_GBps"] = {"plain": n * R * P * 16 / 1e9 / (ext_p / 1e3), "dewarped": n * R * P * 16 / 1e9 / (ext_d / 1e3)}
    out["separate_pass_extra_bytes_per_scan"] = R * P * 32
    print(json.dumps(out))
    ctx.set_stream(None)
    ctx.close()


if __name__ == "__main__":
    main()
