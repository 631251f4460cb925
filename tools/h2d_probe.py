#!/usr/bin/env python
"""Host-to-device ceiling of the box, next to what the host-buffer sequence calls reach.

    python tools/h2d_probe.py                                   # 1 GPU
    torchrun --nproc-per-node N tools/h2d_probe.py              # N ranks copying at once (aggregate fabric ceiling)

Per rank, 1 GiB of page-locked memory: (a) H2D copies alone, (b) the same copies while the device-resident sequence
call of this library runs back to back on another stream (what the asynchronous host pipeline does), (c) the
device-resident call reading its scans straight from the page-locked host buffer (zero-copy: the extraction kernel's
bulk copies pull the rings over PCIe themselves).  Rank 0 prints one JSON line; rates are max-over-ranks times.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import _capi, synth  # noqa: E402


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def tmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R, P, n = 64, 1024, 1024
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    d_scans = synth.make_scans_torch(R, P, rank * n, n, dev)
    h_scans = torch.empty(d_scans.shape, dtype=torch.float32, pin_memory=True)
    h_scans.copy_(d_scans)
    d_dst = torch.empty_like(d_scans)
    outs = [torch.zeros((n - 1, 7), dtype=torch.float64, device=dev), torch.zeros(n - 1, dtype=torch.int32, device=dev),
            torch.zeros(n - 1, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev),
            torch.zeros(n, dtype=torch.int32, device=dev)]
    ptrs = [t.data_ptr() for t in outs]
    ctx = _capi.Context(local)
    work, copy = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ctx.set_stream(work.cuda_stream)
    gib = h_scans.numel() * 4 / 2**30
    reps = 6

    def copies():
        with torch.cuda.stream(copy):
            for _ in range(reps):
                d_dst.copy_(h_scans, non_blocking=True)

    for _ in range(2):
        copies()
        ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, *ptrs)
    barrier()
    t0 = time.perf_counter()
    copies()
    torch.cuda.synchronize()
    alone = tmax(time.perf_counter() - t0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps + 4):  # keep the SMs busy for longer than the copies take
        ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, *ptrs)
    e0.record(copy)
    copies()
    e1.record(copy)
    torch.cuda.synchronize()
    loaded = tmax(e0.elapsed_time(e1) / 1e3)
    barrier()
    # kernels alone, then with the scans read in place from page-locked host memory
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(work)
    for _ in range(3):
        ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, *ptrs)
    k1.record(work)
    torch.cuda.synchronize()
    resident = tmax(k0.elapsed_time(k1) / 3e3)
    ref_pose = outs[0].clone()
    zero_copy, zc_equal = None, None
    try:
        ctx.odometry_device_ptr(h_scans.data_ptr(), n, lp, fe, rp, *ptrs)
        torch.cuda.synchronize()
        k0.record(work)
        for _ in range(3):
            ctx.odometry_device_ptr(h_scans.data_ptr(), n, lp, fe, rp, *ptrs)
        k1.record(work)
        torch.cuda.synchronize()
        zero_copy = tmax(k0.elapsed_time(k1) / 3e3)
        zc_equal = bool(torch.equal(ref_pose, outs[0]))
    except Exception as e:  # noqa: BLE001
        zero_copy = str(e)
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "bytes_per_copy_GiB": gib,
            "h2d_alone_GBps_per_gpu": gib * reps * 2**30 / 1e9 / alone, "h2d_alone_GBps_total": world * gib * reps * 2**30 / 1e9 / alone,
            "h2d_under_kernels_GBps_per_gpu": gib * reps * 2**30 / 1e9 / loaded,
            "h2d_under_kernels_GBps_total": world * gib * reps * 2**30 / 1e9 / loaded,
            "scans_per_s_ceiling_float4_under_kernels": world * gib * reps * 2**30 / loaded / (R * P * 16),
            "device_resident_scans_per_s": world * n / resident,
            "zero_copy_scans_per_s": (world * n / zero_copy) if isinstance(zero_copy, float) else zero_copy,
            "zero_copy_results_equal": zc_equal}), flush=True)
    ctx.set_stream(None)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
