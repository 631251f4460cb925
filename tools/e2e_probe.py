#!/usr/bin/env python
"""Where the host-buffer pipeline loses time: K asynchronous calls (wall clock) next to the sum of their kernel times
(CUDA events per launch), for several chunk sizes and record sizes.  One GPU."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import _capi, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    R, P, n = 64, 1024, 1024
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    d4 = synth.make_scans_torch(R, P, 0, n, dev)
    out = []
    for pb in (16, 12):
        d = d4 if pb == 16 else d4[:, :, :3].contiguous()
        h = torch.empty(d.shape, dtype=torch.float32, pin_memory=True)
        h.copy_(d)
        hs = [torch.zeros((n - 1, 7), dtype=torch.float64, pin_memory=True), torch.zeros(n - 1, dtype=torch.int32, pin_memory=True),
              torch.zeros(n - 1, dtype=torch.int32, pin_memory=True), torch.zeros(n, dtype=torch.int32, pin_memory=True),
              torch.zeros(n, dtype=torch.int32, pin_memory=True)]
        ds = [t.to(dev) for t in hs]
        for chunk in (0, 256, 512, 1024):
            for prof in (False, True):
                ctx = _capi.Context(0)
                stream = torch.cuda.Stream(device=dev)
                ctx.set_stream(stream.cuda_stream)
                if chunk:
                    ctx.set_chunk_pairs(chunk)
                for _ in range(3):
                    ctx.odometry_host_async_ptr(h.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in hs), stride=pb)
                    ctx.odometry_device_ptr(d.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in ds), stride=pb)
                ctx.synchronize()
                torch.cuda.synchronize()
                ctx.kernel_times()
                ctx.set_profiling(prof)
                K = 12
                t0 = time.perf_counter()
                for _ in range(K):
                    ctx.odometry_host_async_ptr(h.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in hs), stride=pb)
                t_enq = time.perf_counter() - t0
                ctx.synchronize()
                torch.cuda.synchronize()
                wall = time.perf_counter() - t0
                kt = ctx.kernel_times()
                ctx.set_profiling(False)
                # the same chunking with device-resident scans
                t0 = time.perf_counter()
                for _ in range(K):
                    ctx.odometry_device_ptr(d.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in ds), stride=pb)
                ctx.synchronize()
                torch.cuda.synchronize()
                wall_dev = time.perf_counter() - t0
                out.append({"point_bytes": pb, "chunk": chunk, "profiling": prof, "e2e_scans_per_s": K * n / wall,
                            "enqueue_ms_per_call": 1e3 * t_enq / K, "wall_ms_per_call": 1e3 * wall / K,
                            "kernel_ms_per_call": sum(v[0] for v in kt.values()) / K if prof else None,
                            "device_resident_ms_per_call": 1e3 * wall_dev / K})
                print(json.dumps(out[-1]), flush=True)
                ctx.set_stream(None)
                ctx.close()


if __name__ == "__main__":
    main()
