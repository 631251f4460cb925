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

UNIT = "scans/s"
README_MS_PER_SCAN = 16.5  # the reference's only published figure: ~3.5 ms extract + ~13 ms register (README.md:31)


def metric_name(a):
    """BASELINE.json's metric, quoted on the shape actually run (the default is its 64x1024 headline)."""
    return f"extract+register scans/sec at {a.rings}x{a.cols}"
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
    ap.add_argument("--no-configs", action="store_true", help="skip the block of the other named BASELINE.json configs")
    ap.add_argument("--point-bytes", type=int, default=16, choices=[12, 16],
                    help="bytes per float record of the headline run: 16 = {x,y,z,.} (SURVEY §8d spec, default), 12 = packed xyz")
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
        if scan.shape[1] == 3:  # packed xyz records: the reference shim reads 16-byte FieldAccessor points
            scan = np.concatenate([scan, np.zeros((len(scan), 1), np.float32)], axis=1)
        _, _, e, p = ref.extract_timed_f32x4(scan, lp, fe, reps=1)
    else:
        e, p = orc.extract(scan[:, :3].astype(np.float64), lp, fe)
    xyz = scan[:, :3].astype(np.float64)
    return xyz[e], xyz[p]


def cpu_pairs_time(scans, lp, fe, rp, orc, ref, results=None):
    """Single-thread time of extract(every scan once) + register(each consecutive pair).  `results` (optional dict)
    receives what the CPU computed: feature counts per scan and pose / termination / outer iterations per pair."""
    t0 = time.perf_counter()
    feats = [_cpu_extract(s, lp, fe, orc, ref) for s in scans]
    out = []
    for k in range(len(scans) - 1):
        if results is None:
            orc.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp)
        else:
            out.append(orc.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp,
                                    want_detail=True))
    dt = time.perf_counter() - t0
    if results is not None:
        results["n_edge"] = np.array([len(f[0]) for f in feats])
        results["n_planar"] = np.array([len(f[1]) for f in feats])
        results["poses"] = np.array([o[0] for o in out])
        results["termination"] = np.array([o[1].termination for o in out])
        results["iterations"] = np.array([o[1].n_iters for o in out])
    return dt


def cpu_baseline_leg(a, host_scans):
    """Bounded single-core sample of the same workload: the first n+1 scans of rank 0's segment.  Returns the
    cpu_baseline object and the CPU results of those pairs (bench.py compares the GPU's with them: parity_check)."""
    lp, fe, rp, orc, ref = _cpu_setup(a)
    probe = 3
    t_probe = cpu_pairs_time(host_scans[:probe + 1], lp, fe, rp, orc, ref)  # also warms caches
    per_pair = t_probe / probe
    n = int(max(4, min(len(host_scans) - 1, a.cpu_seconds / max(per_pair, 1e-6))))
    res = {}
    t = cpu_pairs_time(host_scans[:n + 1], lp, fe, rp, orc, ref, results=res)
    # every scan of a long sequence is extracted once and registered once: scans/s = pairs / time
    base = {"value": n / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n} pairs ({n + 1} scans) of rank 0's segment, {t:.1f} s on one core; extract = "
                      f"{'real reference features code (oracle/_ref)' if ref is not None else 'oracle port'}, "
                      f"register = restated CPU port (Ceres/nanoflann absent)",
            "ms_per_scan": 1e3 * t / n,
            # the only published number for this path, always printed beside the restated CPU figure
            "reference_readme": {"ms_per_scan": README_MS_PER_SCAN, "scans_per_s": 1e3 / README_MS_PER_SCAN,
                                 "source": "README.md:31 of the reference: ~3.5 ms extractFeatures + ~13 ms "
                                           "registerFeatures, Ouster-64, author's laptop, 1 thread"}}
    return base, res


def parity_check(cpu, poses, term, iters, ne, npl, index_check):
    """GPU results of this very run against the CPU oracle on the cpu_baseline sample (same pairs, same inputs)."""
    n = len(cpu["poses"])
    dq = np.abs(np.sum(cpu["poses"][:, :4] * poses[:n, :4], axis=1))
    rad = 2.0 * np.arccos(np.minimum(1.0, dq))
    out = {"pairs": int(n), "max_rad": float(rad.max()), "max_m": float(np.abs(cpu["poses"][:, 4:] - poses[:n, 4:]).max()),
           "terminations_equal": bool(np.array_equal(cpu["termination"], term[:n])),
           "outer_iterations_equal": bool(np.array_equal(cpu["iterations"], iters[:n])),
           "feature_counts_equal": bool(np.array_equal(cpu["n_edge"], ne[:n + 1]) and
                                        np.array_equal(cpu["n_planar"], npl[:n + 1])),
           "indices_equal": index_check, "tolerance": {"rad": 1e-6, "m": 1e-5},
           "against": "CPU oracle: real reference feature code + restated registration (parity with the real Ceres "
                      "solve itself is unpinned, DESIGN.md §7)"}
    out["ok"] = bool(out["max_rad"] < 1e-6 and out["max_m"] < 1e-5 and out["terminations_equal"] and
                     out["outer_iterations_equal"] and out["feature_counts_equal"] and index_check["equal"])
    return out


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
    line = {"impl": "reference", "metric": metric_name(a), "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
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


class Seq:
    """One synthetic sequence segment on one GPU: device-resident scans, their pinned host mirror, result buffers."""

    def __init__(self, torch, synth, dev, R, P, scan_lo, n, point_bytes=16):
        self.R, self.P, self.n, self.n_points, self.point_bytes = R, P, n, R * P, point_bytes
        self.d_scans = synth.make_scans_torch(R, P, scan_lo, n, dev)
        if point_bytes == 12:  # packed xyz records: the fourth float of a sensor record is never read
            self.d_scans = self.d_scans[:, :, :3].contiguous()
        self.h_scans = torch.empty(self.d_scans.shape, dtype=torch.float32, pin_memory=True)
        self.h_scans.copy_(self.d_scans)
        mk = lambda shape, dt, **kw: torch.zeros(shape, dtype=dt, **kw)  # noqa: E731
        self.d = [mk((n - 1, 7), torch.float64, device=dev), mk(n - 1, torch.int32, device=dev),
                  mk(n - 1, torch.int32, device=dev), mk(n, torch.int32, device=dev), mk(n, torch.int32, device=dev)]
        self.h = [mk((n - 1, 7), torch.float64, pin_memory=True), mk(n - 1, torch.int32, pin_memory=True),
                  mk(n - 1, torch.int32, pin_memory=True), mk(n, torch.int32, pin_memory=True),
                  mk(n, torch.int32, pin_memory=True)]
        torch.cuda.synchronize()


def time_sequence(torch, ctx, stream, seq, lp, fe, rp, steps, warmup, barrier, max_over_ranks, profile=True):
    """`steps` timed passes over `seq`: device-resident (CUDA events on the launching stream) and through the
    host-buffer calls (wall clock around K asynchronous calls + one wait, and around K synchronous calls)."""
    n = seq.n
    dp, hp = [t.data_ptr() for t in seq.d], [t.data_ptr() for t in seq.h]

    def step_device():
        ctx.odometry_device_ptr(seq.d_scans.data_ptr(), n, lp, fe, rp, *dp, stride=seq.point_bytes)

    def step_host():
        ctx.odometry_host_ptr(seq.h_scans.data_ptr(), n, lp, fe, rp, *hp, stride=seq.point_bytes)

    def step_host_async():
        ctx.odometry_host_async_ptr(seq.h_scans.data_ptr(), n, lp, fe, rp, *hp, stride=seq.point_bytes)

    for _ in range(warmup):
        step_device()
    barrier()
    ctx.kernel_times()  # reset accumulators
    ctx.set_profiling(profile)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    barrier()
    out = {"dev_ms": max_over_ranks(e0.elapsed_time(e1)), "launches": ctx.launch_count - launches0,
           "ktimes": ctx.kernel_times()}
    ctx.set_profiling(False)
    # (a) the way a recording is streamed through: K asynchronous calls (pinned host scans in, results out to pinned
    #     host memory, all copies inside the timed region), one wait at the end — the copies of a call overlap the
    #     kernels of the previous one;  (b) every call waited for before the next one starts.
    for _ in range(max(1, min(warmup, 2))):  # both forms (they use different chunk sizes: buffers grow once)
        step_host()
        step_host_async()
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host_async()
    ctx.synchronize()
    torch.cuda.synchronize()
    out["e2e_ms"] = max_over_ranks(1e3 * (time.perf_counter() - t0))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    torch.cuda.synchronize()
    out["e2e_sync_ms"] = max_over_ranks(1e3 * (time.perf_counter() - t0))
    barrier()
    assert np.array_equal(seq.h[0].numpy(), seq.d[0].cpu().numpy()), "host and device entry points disagree"
    if seq.point_bytes == 16 and profile:
        # the same asynchronous calls on packed {x,y,z} records (loamgpu_odometry_host_async_strided, 12 bytes per
        # point): 25 % fewer host-to-device bytes; results identical
        ref_pose = seq.h[0].numpy().copy()
        h3 = torch.empty(seq.h_scans.shape[:2] + (3,), dtype=torch.float32, pin_memory=True)
        h3.copy_(seq.h_scans[:, :, :3])

        def step_packed():
            ctx.odometry_host_async_ptr(h3.data_ptr(), n, lp, fe, rp, *hp, stride=12)

        for _ in range(2):
            step_packed()
        ctx.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_packed()
        ctx.synchronize()
        torch.cuda.synchronize()
        out["e2e_packed_ms"] = max_over_ranks(1e3 * (time.perf_counter() - t0))
        barrier()
        assert np.array_equal(seq.h[0].numpy(), ref_pose), "packed-xyz and float4 records disagree"
    return out


def traffic_per_launch(dom, dom_n, shape, n, steps):
    """Per-launch DRAM bytes of the dominant kernel class from the committed `ncu --set full` capture — only when that
    capture was taken on the shape being run (profiles/traffic.json names its shape and the commit it was taken at)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, None, None
    try:
        tj = json.load(open(tp))
        if list(tj.get("shape", [64, 1024])) != list(shape):
            return None, None, f"no capture for {shape[0]}x{shape[1]} (profiles/traffic.json: {tj.get('shape')})"
        units = n if tj.get("unit", {}).get(dom) == "scan" else n - 1
        per_unit = tj.get("per_unit", {}).get(dom)
        traffic = per_unit * units * steps / max(dom_n, 1) if per_unit else None
        return traffic, tj.get("_limiter", {}).get(dom), f"ncu --set full capture at commit {tj.get('commit', '?')}"
    except Exception:
        return None, None, None


def other_configs(torch, synth, _capi, ctx, stream, dev, a, barrier, max_over_ranks):
    """The other named configs of BASELINE.json on this GPU (3 timed steps each after 3 warm-ups): 16x1800, 128x2048
    and scan-to-local-map.  Parity for these shapes is what the -m gpu tests cover; here they are measured."""
    out = {}
    fe, rp = _capi.default_fe_params(), _capi.default_reg_params()
    for name, R, P, n in (("16x1800", 16, 1800, 1024), ("128x2048", 128, 2048, 256)):
        seq = Seq(torch, synth, dev, R, P, 0, n)
        lp = _capi.CLidarParams(R, P, 1.0, 120.0)
        t = time_sequence(torch, ctx, stream, seq, lp, fe, rp, 3, 3, barrier, max_over_ranks, profile=False)
        it = seq.h[2].numpy()
        out[name] = {"metric": f"extract+register scans/sec at {R}x{P}", "scans_per_step": n, "steps": 3, "warmup": 3,
                     "value": 3 * n / (t["dev_ms"] / 1e3), "e2e": 3 * n / (t["e2e_ms"] / 1e3),
                     "e2e_each_call_waited": 3 * n / (t["e2e_sync_ms"] / 1e3), "unit": UNIT,
                     "ms_per_step": t["dev_ms"] / 3, "mean_edge": float(seq.h[3].numpy().mean()),
                     "mean_planar": float(seq.h[4].numpy().mean()), "mean_outer_iterations": float(it.mean()),
                     "h2d_bytes_per_step": int(n * R * P * 16)}
        del seq
        torch.cuda.empty_cache()
    # scan-to-local-map (config 5): target = features of 20 accumulated 128x2048 scans (max 400 planar / 60 edge per
    # sector) in the frame of scan 0, resident on the device; source = one 64x1024 scan; one C-ABI call per registration
    R, P, n_map = 128, 2048, 20
    lp = _capi.CLidarParams(R, P, 1.0, 120.0)
    fe_map = _capi.default_fe_params()
    fe_map.max_planar_feats_per_sector, fe_map.max_edge_feats_per_sector = 400, 60

    def moved(points, pose):
        q, tr = np.asarray(pose[:4]), np.asarray(pose[4:7])
        uv = 2.0 * np.cross(q[:3], points)
        return (points + q[3] * uv + np.cross(q[:3], uv) + tr).astype(np.float32).astype(np.float64)

    te, tp = [], []
    for k in range(n_map):
        sc = synth.make_scan(R, P, k=k)
        e, p = ctx.extract(sc, lp, fe_map)
        xyz = sc[:, :3].astype(np.float64)
        te.append(moved(xyz[e], synth.relative_pose(0, k)))
        tp.append(moved(xyz[p], synth.relative_pose(0, k)))
    te, tp = np.concatenate(te), np.concatenate(tp)
    sc = synth.make_scan(64, 1024, k=n_map)
    e, p = ctx.extract(sc, _capi.CLidarParams(64, 1024, 1.0, 120.0), fe)
    se, sp = sc[:, :3].astype(np.float64)[e], sc[:, :3].astype(np.float64)[p]
    init, gt = synth.relative_pose(0, n_map - 1), synth.relative_pose(0, n_map)
    t0 = time.perf_counter()
    m = ctx.map_create(te, tp)
    create_ms = 1e3 * (time.perf_counter() - t0)
    for _ in range(3):
        pose = ctx.register_to_map(m, se, sp, init, rp)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        pose = ctx.register_to_map(m, se, sp, init, rp)
        ts.append(1e3 * (time.perf_counter() - t0))
    m.close()
    dq = abs(float(np.dot(pose[:4], gt[:4])))
    out["scan_to_map"] = {"metric": "scan-to-local-map registrations/sec (one host-buffer call each)",
                          "target_points": int(len(te) + len(tp)), "source_points": int(len(se) + len(sp)),
                          "value": 1e3 / float(np.median(ts)), "unit": "registrations/s",
                          "register_to_map_ms": float(np.median(ts)), "map_create_ms": create_ms,
                          "error_vs_ground_truth": {"rad": 2.0 * float(np.arccos(min(1.0, dq))),
                                                    "m": float(np.abs(pose[4:] - gt[4:]).max())}}
    return out


def single_call_block(host_scans, R, P):
    """ONE loam::extractFeatures + ONE loam::registerFeatures call per scan through the C++ API (include/loam/*.h),
    the reference's own usage pattern and what its README figure (3.5 + 13 ms) is quoted on.  Timed by the C++ program
    tests/cpp/bench_single_call (host vectors in, features / pose out, every copy inside the timed call)."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "tests", "cpp", "bench_single_call")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-C", os.path.dirname(exe), "build"], capture_output=True, text=True)
        if r.returncode != 0:
            return {"unavailable": "tests/cpp/bench_single_call does not build: " + r.stderr[-200:]}
    scans = host_scans[:2]
    if scans.shape[2] == 3:
        scans = np.concatenate([scans, np.zeros(scans.shape[:2] + (1,), np.float32)], axis=2)
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        np.ascontiguousarray(scans, dtype=np.float32).tofile(f.name)
        r = subprocess.run([exe, f.name, str(R), str(P), "40"], capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        return {"unavailable": "bench_single_call failed: " + (r.stderr or r.stdout)[-200:]}
    d = json.loads(r.stdout.strip().splitlines()[-1])
    out = {"extract_ms": d["extract_ms"], "register_ms": d["register_ms"],
           "register_with_detail_ms": d["register_with_detail_ms"], "outer_iterations": d["outer_iterations"],
           "scans_per_s": 1e3 / (d["extract_ms"] + d["register_ms"]), "shape": [R, P],
           "api": "loam::extractFeatures(std::vector<PointF>) + loam::registerFeatures(LoamFeatures<PointF>) "
                  "(include/loam/*.h over the C-ABI), host buffers, median of 40 calls"}
    if (R, P) == (64, 1024):  # the README figure is for a 64-ring Ouster scan
        out["vs_readme"] = README_MS_PER_SCAN / (d["extract_ms"] + d["register_ms"])
    return out


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
    seq = Seq(torch, synth, dev, R, P, shard.scan_lo, n, a.point_bytes)

    # a dedicated non-default stream: the library treats a NULL stream handle as "use the context's own stream",
    # and the CUDA events below must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    t = time_sequence(torch, ctx, stream, seq, lp, fe, rp, a.steps, a.warmup, barrier, max_over_ranks)
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    dev_ms, e2e_ms, e2e_sync_ms, ktimes, launches = t["dev_ms"], t["e2e_ms"], t["e2e_sync_ms"], t["ktimes"], t["launches"]

    # results of the last step (host copies from the e2e call)
    h_pose, h_term, h_iter, h_ne, h_np = (x.numpy() for x in seq.h)
    ne, npl = h_ne.astype(np.int64), h_np.astype(np.int64)
    iters, term = h_iter.astype(np.int64), h_term

    rc = 0
    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes(n_points, ne, npl, iters)
        dom = max(ktimes, key=lambda k: ktimes[k][0])
        dom_ms, dom_n = ktimes[dom]
        achieved = (ab[dom] * a.steps / 1e9) / (dom_ms / 1e3) if dom_ms > 0 else 0.0
        traffic, limiter, traffic_src = traffic_per_launch(dom, dom_n, (R, P), n, a.steps)
        total_scans = a.scans * world * a.steps  # halo scans (extracted by two ranks) are counted once
        kernel_ms_total = sum(v[0] for v in ktimes.values())
        whole = sum(ab.values()) * a.steps
        line = {
            "metric": metric_name(a), "value": total_scans / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "rings": R, "cols": P, "scans_per_step_per_gpu": a.scans,
                       "params": "default FeatureExtractionParams / RegistrationParams, identity init",
                       "point_bytes": a.point_bytes,
                       "l2": f"inputs larger than L2 ({n * n_points * a.point_bytes / 2**20:.0f} MiB of scans per step per GPU)",
                       "sharding": "contiguous sequence segments per rank, no data-path collective",
                       "host_affinity": (f"each rank bound to its GPU's NUMA-local CPUs ({numa_cpus})" if numa_cpus
                                         else "unbound")},
            "e2e": {"value": total_scans / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(world * n * n_points * a.point_bytes),
                    "d2h_bytes_per_step": int(world * ((n - 1) * (56 + 4 + 4) + n * 8)),
                    "ms_per_step": e2e_ms / a.steps,
                    "mode": "K asynchronous host-buffer calls (loamgpu_odometry_host_async), one wait after the last",
                    "value_each_call_waited": total_scans / (e2e_sync_ms / 1e3),
                    **({"packed_xyz": {"value": total_scans / (t["e2e_packed_ms"] / 1e3), "unit": UNIT,
                                       "h2d_bytes_per_step": int(world * n * n_points * 12),
                                       "note": "the same calls on 12-byte {x,y,z} records: identical results"}}
                       if "e2e_packed_ms" in t else {})},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
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
            host = seq.h_scans.numpy()
            line["cpu_baseline"], cpu = cpu_baseline_leg(a, host)
            # feature indices of the first scans through the single-scan entry point against the reference's feature code
            lp_o, fe_o, _, orc, ref = _cpu_setup(a)
            eq = True
            for k in range(2):
                xyz = host[k][:, :3].astype(np.float64)
                ce, cp = (ref.extract(xyz, lp_o, fe_o) if ref is not None else orc.extract(xyz, lp_o, fe_o))
                ge, gp = ctx.extract(host[k], lp, fe)
                eq = eq and np.array_equal(ce, ge) and np.array_equal(cp, gp)
            line["parity_check"] = parity_check(cpu, h_pose, term, iters, ne, npl, {
                "scans": 2, "equal": bool(eq),
                "against": "real reference feature code (oracle/_ref)" if ref is not None else "oracle port"})
            if not line["parity_check"]["ok"]:
                rc = 3
        if world == 1 and not a.no_configs:
            line["single_call"] = single_call_block(seq.h_scans.numpy(), R, P)
            del seq
            torch.cuda.empty_cache()
            line["configs"] = other_configs(torch, synth, _capi, ctx, stream, dev, a, barrier, max_over_ranks)
        print(json.dumps(line), flush=True)
        if rc:
            sys.stderr.write("bench.py: PARITY VIOLATION " + json.dumps(line["parity_check"]) + "\n")
    ctx.set_stream(None)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)
    return ours(a)


if __name__ == "__main__":
    sys.exit(main())
