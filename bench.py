#!/usr/bin/env python
"""bench.py — extract + register throughput (scans/sec) on synthetic organised 64x1024 scans.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU implementation, host cores

A "step" is one pass of the hot path over one batch: `--scans` consecutive scans of the synthetic sequence
per GPU (every scan extracted once; pair k registers scan k+1 onto scan k from an identity initial
estimate — the README loop of the reference).  Ranks own disjoint segments of the sequence; there is no
data-path collective (torch.distributed is used for the barrier and the max-over-ranks timing only).

  value : whole-job scans/sec with the scans already resident in HBM (loamgpu_odometry_device),
          device time from CUDA events on the launching stream, max over ranks.
  e2e   : the same through the host-buffer C-ABI call: pinned host scans in, poses/terminations/iteration counts/
          feature counts out to pinned host memory, every copy inside the timed region.  `e2e.value` = K asynchronous
          calls (loamgpu_odometry_host_async: how a recording is streamed through in pieces; the copies of a call run
          under the kernels of the previous one) and one wait after the last; `e2e.value_each_call_waited` = the
          synchronous call (loamgpu_odometry_host) K times.
  roofline     : the dominant kernel class (by CUDA-event time inside the timed region): algorithmic bytes
                  / its summed launch time, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline : the reference feature code (oracle/_ref, real reference sources) + the restated registration
                 (oracle/) on ONE host core over a bounded sample of the same sequence (rank 0, N=1 only).

oracle/ is executed here only for cpu_baseline and for `--impl reference` (the task's two allowed places).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "extract+register scans/sec at 64x1024"
UNIT = "scans/s"
K_NEIGH = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rings", type=int, default=64)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--scans", type=int, default=1024, help="scans per step per GPU")
    ap.add_argument("--chunk-pairs", type=int, default=0, help="pairs per internal chunk (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--ref-pairs", type=int, default=0, help="--impl reference: pairs per step (0 = 2 x host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return f"synthetic {a.rings}x{a.cols} sequence, scan-to-scan extract+register, {a.scans} scans/step/GPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + clock event reasons of one GPU every 20 ms while the timed region runs."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs closest to its GPU (NVML's ideal CPU affinity) BEFORE the pinned
    staging buffers are allocated, so they are first-touched on that NUMA node: with 8 ranks streaming 1 MiB per
    scan each, buffers on the wrong socket make the host side the bottleneck of `e2e`.  Returns the CPU count."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------ CPU legs
def _cpu_setup(a):
    from oracle.pyoracle import FeParams, LidarParams, Oracle, RefLib, RegParams
    lp, fe, rp = LidarParams(a.rings, a.cols, 1.0, 120.0), FeParams.default(), RegParams.default()
    orc = Oracle()
    ref = RefLib() if RefLib.available() else None
    return lp, fe, rp, orc, ref


def _cpu_extract(scan, lp, fe, orc, ref):
    """One scan through the reference's feature code (real sources when oracle/_ref exists)."""
    if ref is not None:
        _, _, e, p = ref.extract_timed_f32x4(scan, lp, fe, reps=1)
    else:
        e, p = orc.extract(scan[:, :3].astype(np.float64), lp, fe)
    xyz = scan[:, :3].astype(np.float64)
    return xyz[e], xyz[p]


def cpu_pairs_time(scans, lp, fe, rp, orc, ref):
    """Single-thread time of extract(every scan once) + register(each consecutive pair)."""
    t0 = time.perf_counter()
    feats = [_cpu_extract(s, lp, fe, orc, ref) for s in scans]
    for k in range(len(scans) - 1):
        orc.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp)
    return time.perf_counter() - t0


def cpu_baseline_leg(a, host_scans):
    """Bounded single-core sample of the same workload: the first n+1 scans of rank 0's segment."""
    lp, fe, rp, orc, ref = _cpu_setup(a)
    probe = 3
    t_probe = cpu_pairs_time(host_scans[:probe + 1], lp, fe, rp, orc, ref)  # also warms caches
    per_pair = t_probe / probe
    n = int(max(4, min(len(host_scans) - 1, a.cpu_seconds / max(per_pair, 1e-6))))
    t = cpu_pairs_time(host_scans[:n + 1], lp, fe, rp, orc, ref)
    # every scan of a long sequence is extracted once and registered once: scans/s = pairs / time
    return {"value": n / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n} pairs ({n + 1} scans) of rank 0's segment, {t:.1f} s on one core; extract = "
                      f"{'real reference features code (oracle/_ref)' if ref is not None else 'oracle port'}, "
                      f"register = restated CPU port (Ceres/nanoflann absent)",
            "ms_per_scan": 1e3 * t / n}


def reference_arm(a):
    """--impl reference: the reference's CPU implementation on all host cores (independent pair blocks per thread)."""
    from concurrent.futures import ThreadPoolExecutor

    from loam_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lp, fe, rp, orc, ref = _cpu_setup(a)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    pairs = a.ref_pairs or 2 * cores
    per = max(1, pairs // cores)
    pairs = per * cores
    # each thread owns a contiguous block of `per` pairs (+1 halo scan) of the sequence: same sharding as the GPU arm
    blocks = [np.stack([synth.make_scan(a.rings, a.cols, k=t * per + j) for j in range(per + 1)]) for t in range(cores)]

    def work(b):
        return cpu_pairs_time(b, lp, fe, rp, orc, ref)

    def step():
        with ThreadPoolExecutor(cores) as ex:  # ctypes releases the GIL inside the C/C++ calls
            list(ex.map(work, blocks))

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = pairs * a.steps / dt
    kind = "port"  # registration (≈90 % of the CPU time) is the restated port; only the feature half is real reference code
    sample = (f"{pairs} pairs/step ({per} per thread + halo scan) of the same synthetic sequence; extract = "
              f"{'real reference features code (oracle/_ref)' if ref is not None else 'oracle port'}, register = "
              f"restated CPU port (Ceres 2.2.0 / nanoflann 1.5.5 are not in the image)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "rings": a.rings, "cols": a.cols, "host_threads": cores,
                       "pairs_per_step": pairs},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def algorithmic_bytes(n_points, ne, npl, iters, k=K_NEIGH):
    """DESIGN.md §4 / SURVEY.md §8(d): bytes each kernel class must move, from the run's own counts."""
    ne, npl, iters = ne.astype(np.int64), npl.astype(np.int64), iters.astype(np.int64)
    F = ne + npl                         # features per scan
    n_scans = len(ne)
    S = F[1:]                            # source features of pair k = scan k+1
    T = F[:-1]
    return {
        "extract": 16 * n_points * n_scans + 4 * int(F.sum()),
        "pack": (4 + 16) * int(F.sum()),
        "nn_build": 32 * int(T.sum()),
        "knn": int((iters * S).sum()) * (16 + 16 * k + 8),
        "fit": int((iters * S).sum()) * (16 * k + 48),
        "lm": int((iters * S).sum()) * 48,
        "misc": 0,
    }


def ours(a):
    import torch
    import torch.distributed as dist

    from loam_b200 import _capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    numa_cpus = bind_to_gpu_numa_node(physical_gpu_index(local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from loam_b200.sharding import shard_sequence
    R, P = a.rings, a.cols
    # weak scaling: a virtual sequence of world * --scans scans; rank r owns a contiguous block of its pairs and
    # therefore extracts its own scans plus one halo scan (the first scan of the next block)
    shard = shard_sequence(a.scans * world, world, rank)
    n = shard.n_scans
    n_points = R * P
    lp = _capi.CLidarParams(R, P, 1.0, 120.0)
    fe, rp = _capi.default_fe_params(), _capi.default_reg_params()
    ctx = _capi.Context(local)
    if a.chunk_pairs:
        ctx.set_chunk_pairs(a.chunk_pairs)

    # this rank's scans of the synthetic sequence (generated on the device, then mirrored to pinned host)
    d_scans = synth.make_scans_torch(R, P, shard.scan_lo, n, dev)
    h_scans = torch.empty(d_scans.shape, dtype=torch.float32, pin_memory=True)
    h_scans.copy_(d_scans)
    d_pose = torch.zeros((n - 1, 7), dtype=torch.float64, device=dev)
    d_term = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    d_iter = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    d_ne = torch.zeros(n, dtype=torch.int32, device=dev)
    d_np = torch.zeros(n, dtype=torch.int32, device=dev)
    h_pose = torch.zeros((n - 1, 7), dtype=torch.float64, pin_memory=True)
    h_term = torch.zeros(n - 1, dtype=torch.int32, pin_memory=True)
    h_iter = torch.zeros(n - 1, dtype=torch.int32, pin_memory=True)
    h_ne = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    h_np = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()

    # a dedicated non-default stream: the library treats a NULL stream handle as "use the context's own stream",
    # and the CUDA events below must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    def step_device():
        ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, d_pose.data_ptr(), d_term.data_ptr(),
                                d_iter.data_ptr(), d_ne.data_ptr(), d_np.data_ptr())

    def step_host():
        ctx.odometry_host_ptr(h_scans.data_ptr(), n, lp, fe, rp, h_pose.data_ptr(), h_term.data_ptr(),
                              h_iter.data_ptr(), h_ne.data_ptr(), h_np.data_ptr())

    # ---------------- device-resident timing (value)
    for _ in range(a.warmup):
        step_device()
    barrier()
    ctx.kernel_times()  # reset accumulators
    ctx.set_profiling(True)
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step_device()
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    ktimes = ctx.kernel_times()
    ctx.set_profiling(False)

    # ---------------- end-to-end timing through the host-buffer C-ABI call (e2e)
    # (a) the way a recording is streamed through: K asynchronous calls (pinned host scans in, results out to pinned
    #     host memory, all copies inside the timed region), one wait at the end — the copies of a call overlap the
    #     kernels of the previous one;  (b) every call waited for before the next one starts.
    def step_host_async():
        ctx.odometry_host_async_ptr(h_scans.data_ptr(), n, lp, fe, rp, h_pose.data_ptr(), h_term.data_ptr(),
                                    h_iter.data_ptr(), h_ne.data_ptr(), h_np.data_ptr())

    for _ in range(max(1, min(a.warmup, 2))):  # both forms (they use different chunk sizes: buffers grow once)
        step_host()
        step_host_async()
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host_async()
    ctx.synchronize()
    torch.cuda.synchronize()
    t_host = time.perf_counter() - t0
    barrier()
    e2e_ms = max_over_ranks(1e3 * t_host)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    torch.cuda.synchronize()
    t_sync = time.perf_counter() - t0
    barrier()
    e2e_sync_ms = max_over_ranks(1e3 * t_sync)
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # results of the last step (host copies from the e2e call)
    ne, npl = h_ne.numpy().astype(np.int64), h_np.numpy().astype(np.int64)
    iters, term = h_iter.numpy().astype(np.int64), h_term.numpy()
    assert np.array_equal(h_pose.numpy(), d_pose.cpu().numpy()), "host and device entry points disagree"

    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes(n_points, ne, npl, iters)
        dom = max(ktimes, key=lambda k: ktimes[k][0])
        dom_ms, dom_n = ktimes[dom]
        achieved = (ab[dom] * a.steps / 1e9) / (dom_ms / 1e3) if dom_ms > 0 else 0.0
        traffic, limiter = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):  # per-launch DRAM bytes and the limiter named by the committed `ncu --set full` capture
            try:
                tj = json.load(open(tp))
                limiter = tj.get("_limiter", {}).get(dom)
                units = n if tj.get("unit", {}).get(dom) == "scan" else n - 1
                per_unit = tj.get("per_unit", {}).get(dom)
                traffic = per_unit * units * a.steps / max(dom_n, 1) if per_unit else None
            except Exception:
                traffic = None
        total_scans = a.scans * world * a.steps  # halo scans (extracted by two ranks) are counted once
        kernel_ms_total = sum(v[0] for v in ktimes.values())
        whole = sum(ab.values()) * a.steps
        line = {
            "metric": METRIC, "value": total_scans / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "rings": R, "cols": P, "scans_per_step_per_gpu": a.scans,
                       "params": "default FeatureExtractionParams / RegistrationParams, identity init",
                       "l2": f"inputs larger than L2 ({n * n_points * 16 / 2**20:.0f} MiB of scans per step per GPU)",
                       "sharding": "contiguous sequence segments per rank, no data-path collective",
                       "host_affinity": (f"each rank bound to its GPU's NUMA-local CPUs ({numa_cpus})" if numa_cpus
                                         else "unbound")},
            "e2e": {"value": total_scans / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(world * n * n_points * 16),
                    "d2h_bytes_per_step": int(world * ((n - 1) * (56 + 4 + 4) + n * 8)),
                    "ms_per_step": e2e_ms / a.steps,
                    "mode": "K asynchronous host-buffer calls (loamgpu_odometry_host_async), one wait after the last",
                    "value_each_call_waited": total_scans / (e2e_sync_ms / 1e3)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ab[dom] * a.steps / max(dom_n, 1),
                         "avg_launch_ms": dom_ms / max(dom_n, 1), "launches": dom_n,
                         "whole_path_GBps": (whole / 1e9) / (dev_ms / 1e3),
                         "whole_path_frac": (whole / 1e9) / (dev_ms / 1e3) / peak,
                         "limiter": limiter},
            "kernel_ms_per_step": {k: v[0] / a.steps for k, v in ktimes.items()},
            "kernel_share": {k: (v[0] / kernel_ms_total if kernel_ms_total else 0.0) for k, v in ktimes.items()},
            "results": {"mean_edge": float(ne.mean()), "mean_planar": float(npl.mean()),
                        "mean_outer_iterations": float(iters.mean()),
                        "terminations": {str(k): int((term == k).sum()) for k in np.unique(term)}},
        }
        if sampler:
            line["clocks"] = sampler.summary()
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(a, h_scans.numpy())
        print(json.dumps(line), flush=True)
    ctx.set_stream(None)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)
    return ours(a)


if __name__ == "__main__":
    sys.exit(main())
