"""ctypes binding of libloamgpu.so (the C-ABI declared in include/loamgpu.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc, and if no CUDA
device is present `Context()` raises (loamgpu_create fails with LOAMGPU_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

u64, f64, u32, i32 = C.c_uint64, C.c_double, C.c_uint32, C.c_int32
PD = C.POINTER(C.c_double)
PU32 = C.POINTER(C.c_uint32)

OK, ERR_SIZE_MISMATCH, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = range(5)
F32, F64 = 0, 1

# every symbol include/loamgpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "loamgpu_create", "loamgpu_destroy", "loamgpu_last_error", "loamgpu_set_stream", "loamgpu_default_fe_params",
    "loamgpu_default_reg_params", "loamgpu_launch_count", "loamgpu_extract", "loamgpu_curvature",
    "loamgpu_valid_mask", "loamgpu_register", "loamgpu_knn", "loamgpu_odometry_host", "loamgpu_odometry_device",
    "loamgpu_set_chunk_pairs", "loamgpu_set_profiling", "loamgpu_kernel_times", "loamgpu_map_create",
    "loamgpu_map_destroy", "loamgpu_map_size", "loamgpu_map_update", "loamgpu_register_to_map",
    "loamgpu_extract_batch", "loamgpu_register_pairs", "loamgpu_odometry_host_async", "loamgpu_synchronize",
    "loamgpu_extract_dewarped", "loamgpu_odometry_host_dewarped", "loamgpu_odometry_device_dewarped",
    "loamgpu_odometry_host_strided", "loamgpu_odometry_host_async_strided", "loamgpu_odometry_device_strided",
    "loamgpu_multi_create", "loamgpu_multi_destroy", "loamgpu_multi_last_error", "loamgpu_multi_device_count",
    "loamgpu_multi_odometry_host", "loamgpu_debug_problem_eval",
]
KERNEL_CLASSES = ["extract", "pack", "nn_build", "knn", "lm", "misc", "fit"]


class CLidarParams(C.Structure):
    _fields_ = [("scan_lines", u64), ("points_per_line", u64), ("min_range", f64), ("max_range", f64)]


class CFeParams(C.Structure):
    _fields_ = [("neighbor_points", u64), ("number_sectors", u64), ("max_edge_feats_per_sector", u64),
                ("max_planar_feats_per_sector", u64), ("edge_feat_threshold", f64), ("planar_feat_threshold", f64),
                ("occlusion_thresh", f64), ("parallel_thresh", f64)]


class CRegParams(C.Structure):
    _fields_ = [("num_edge_neighbors", u64), ("max_edge_neighbor_dist", f64), ("min_line_fit_points", u64),
                ("min_line_condition_number", f64), ("num_plane_neighbors", u64), ("max_plane_neighbor_dist", f64),
                ("min_plane_fit_points", u64), ("max_avg_point_plane_dist", f64), ("max_iterations", u64),
                ("rotation_convergence_thresh", f64), ("position_convergence_thresh", f64), ("min_associations", u64)]


class CDetail(C.Structure):
    _fields_ = [("max_iters_cap", u32), ("n_src_edge", u32), ("n_src_planar", u32), ("n_iters", u32),
                ("termination", i32), ("iter_est", PD), ("iter_update", PD), ("n_edge_assoc", PU32),
                ("n_plane_assoc", PU32), ("edge_assoc", PU32), ("plane_assoc", PU32), ("lm_iters", PU32),
                ("lm_cost", PD)]


class LoamGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m loam_b200.build` (needs nvcc)")
    lib = C.CDLL(path)
    lib.loamgpu_last_error.restype = C.c_char_p
    lib.loamgpu_last_error.argtypes = [C.c_void_p]
    lib.loamgpu_launch_count.restype = u64
    lib.loamgpu_launch_count.argtypes = [C.c_void_p]
    lib.loamgpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.loamgpu_destroy.argtypes = [C.c_void_p]
    lib.loamgpu_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.loamgpu_set_chunk_pairs.argtypes = [C.c_void_p, u32]
    lib.loamgpu_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.loamgpu_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    vp = C.c_void_p
    lib.loamgpu_extract.argtypes = [vp, vp, C.c_int, C.c_size_t, u64, vp, vp, vp, u64, vp, vp, u64, vp]
    lib.loamgpu_extract_dewarped.argtypes = [vp, vp, C.c_int, C.c_size_t, u64, vp, vp, vp, vp, u64, vp, vp, u64, vp, vp]
    lib.loamgpu_curvature.argtypes = [vp, vp, C.c_int, C.c_size_t, u64, vp, vp, vp]
    lib.loamgpu_valid_mask.argtypes = [vp, vp, C.c_int, C.c_size_t, u64, vp, vp, vp]
    lib.loamgpu_register.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, u64, vp, vp, vp, vp]
    lib.loamgpu_knn.argtypes = [vp, vp, u64, vp, u64, u32, f64, vp, vp]
    lib.loamgpu_extract_batch.argtypes = [vp, vp, C.c_int, C.c_size_t, u64, u64, vp, vp, vp, u64, vp, vp, u64, vp]
    lib.loamgpu_register_pairs.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_map_create.argtypes = [vp, vp, u64, vp, u64, C.POINTER(C.c_void_p)]
    lib.loamgpu_map_destroy.argtypes = [vp, vp]
    lib.loamgpu_map_destroy.restype = None
    lib.loamgpu_map_size.argtypes = [vp, vp, vp]
    lib.loamgpu_map_update.argtypes = [vp, vp, vp, u64, vp, u64, vp, u64, u64]
    lib.loamgpu_register_to_map.argtypes = [vp, vp, vp, u64, vp, u64, vp, vp, vp, vp]
    lib.loamgpu_odometry_host.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_odometry_host_async.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_synchronize.argtypes = [vp]
    lib.loamgpu_odometry_device.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    for f in (lib.loamgpu_odometry_host_strided, lib.loamgpu_odometry_host_async_strided,
              lib.loamgpu_odometry_device_strided):
        f.argtypes = [vp, vp, C.c_size_t, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_odometry_host_dewarped.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_odometry_device_dewarped.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.loamgpu_debug_problem_eval.argtypes = [vp, u64, vp, vp, vp, u64, vp, vp, vp, vp, C.c_int, vp]
    lib.loamgpu_multi_create.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p)]
    lib.loamgpu_multi_destroy.argtypes = [vp]
    lib.loamgpu_multi_destroy.restype = None
    lib.loamgpu_multi_last_error.argtypes = [vp]
    lib.loamgpu_multi_last_error.restype = C.c_char_p
    lib.loamgpu_multi_device_count.argtypes = [vp]
    lib.loamgpu_multi_odometry_host.argtypes = [vp, vp, C.c_size_t, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cloud(points):
    """Return (array, dtype code, stride) for an (N,>=3) float32/float64 array without changing values."""
    a = np.asarray(points)
    if a.size == 0:
        return np.zeros((0, 3), dtype=np.float64), F64, 24
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("point cloud must be an (N, >=3) array")
    if a.dtype == np.float32:
        a = np.ascontiguousarray(a)
        return a, F32, a.strides[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, F64, a.strides[0]


def _xyz64(points):
    a = np.asarray(points, dtype=np.float64)
    if a.size == 0:
        return np.zeros((0, 3), dtype=np.float64)
    return np.ascontiguousarray(a.reshape(len(a), -1)[:, :3])


def _check_scan_size(s, lp):
    """validateLidarScan (common.h:104-113) for a [n_scans, points, C] array: the C-ABI sequence calls take the scan size
    from the lidar parameters, so a mismatch must be caught before any pointer arithmetic."""
    want = int(lp.scan_lines) * int(lp.points_per_line)
    if s.shape[1] != want:
        raise LoamGpuError(ERR_SIZE_MISMATCH,
                           f"LOAM: provided lidar scan size ( {s.shape[1]})  does not match provided lidar parameters "
                           f"({int(lp.scan_lines)} x {int(lp.points_per_line)})")


class Context:
    """One loamgpu context = one device + one stream + its device buffers."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.loamgpu_create(device, C.byref(h))
        if rc != OK:
            raise LoamGpuError(rc, "loamgpu_create failed: " + self.lib.loamgpu_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.loamgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise LoamGpuError(rc, self.lib.loamgpu_last_error(self.h).decode())

    @property
    def launch_count(self) -> int:
        return int(self.lib.loamgpu_launch_count(self.h))

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.loamgpu_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def set_profiling(self, on: bool):
        self._check(self.lib.loamgpu_set_profiling(self.h, 1 if on else 0))

    def kernel_times(self):
        """{class: (milliseconds, launches)} since the previous call (synchronises the stream)."""
        ms = np.zeros(len(KERNEL_CLASSES))
        n = np.zeros(len(KERNEL_CLASSES), dtype=np.uint64)
        self._check(self.lib.loamgpu_kernel_times(self.h, _ptr(ms), _ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def set_chunk_pairs(self, n: int):
        self._check(self.lib.loamgpu_set_chunk_pairs(self.h, n))

    # ---------------------------------------------------------------- features
    def extract(self, points, lp: CLidarParams, fe: CFeParams):
        a, dt, stride = _cloud(points)
        n = len(a)
        e = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        ne, npl = u64(0), u64(0)
        self._check(self.lib.loamgpu_extract(self.h, _ptr(a), dt, stride, n, C.addressof(lp), C.addressof(fe),
                                             _ptr(e), len(e), C.addressof(ne), _ptr(p), len(p), C.addressof(npl)))
        return e[:ne.value].copy(), p[:npl.value].copy()

    def extract_dewarped(self, points, lp: CLidarParams, fe: CFeParams, start_T_end, want_points: bool = True):
        """loamgpu_extract_dewarped: (edge_idx, planar_idx, de-warped n x 3 float64 points or None)."""
        a, dt, stride = _cloud(points)
        n = len(a)
        m = np.ascontiguousarray(start_T_end, dtype=np.float64).reshape(7)
        e = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        moved = np.empty((n, 3), dtype=np.float64) if want_points else None
        ne, npl = u64(0), u64(0)
        self._check(self.lib.loamgpu_extract_dewarped(
            self.h, _ptr(a), dt, stride, n, C.addressof(lp), C.addressof(fe), _ptr(m), _ptr(e), len(e),
            C.addressof(ne), _ptr(p), len(p), C.addressof(npl), _ptr(moved) if want_points and n else None))
        return e[:ne.value].copy(), p[:npl.value].copy(), moved

    def curvature(self, points, lp, fe):
        a, dt, stride = _cloud(points)
        out = np.empty(len(a), dtype=np.float64)
        self._check(self.lib.loamgpu_curvature(self.h, _ptr(a), dt, stride, len(a), C.addressof(lp), C.addressof(fe),
                                               _ptr(out)))
        return out

    def valid_mask(self, points, lp, fe):
        a, dt, stride = _cloud(points)
        out = np.empty(len(a), dtype=np.uint8)
        self._check(self.lib.loamgpu_valid_mask(self.h, _ptr(a), dt, stride, len(a), C.addressof(lp),
                                                C.addressof(fe), _ptr(out)))
        return out.astype(bool)

    # ---------------------------------------------------------------- registration
    @staticmethod
    def _detail_buffers(rp, n_se, n_sp):
        cap = max(int(rp.max_iterations), 1)
        bufs = dict(
            iter_est=np.zeros((cap, 7)), iter_update=np.zeros((cap, 7)),
            n_edge_assoc=np.zeros(cap, dtype=np.uint32), n_plane_assoc=np.zeros(cap, dtype=np.uint32),
            edge_assoc=np.zeros((cap, max(n_se, 1), 2), dtype=np.uint32),
            plane_assoc=np.zeros((cap, max(n_sp, 1), 2), dtype=np.uint32),
            lm_iters=np.zeros(cap, dtype=np.uint32), lm_cost=np.zeros((cap, 2)))
        det = CDetail(cap, max(n_se, 1), max(n_sp, 1), 0, 1,
                      bufs["iter_est"].ctypes.data_as(PD), bufs["iter_update"].ctypes.data_as(PD),
                      bufs["n_edge_assoc"].ctypes.data_as(PU32), bufs["n_plane_assoc"].ctypes.data_as(PU32),
                      bufs["edge_assoc"].ctypes.data_as(PU32), bufs["plane_assoc"].ctypes.data_as(PU32),
                      bufs["lm_iters"].ctypes.data_as(PU32), bufs["lm_cost"].ctypes.data_as(PD))
        return det, bufs

    @staticmethod
    def _detail_info(det, bufs):
        n = det.n_iters
        return dict(n_iters=n, termination=det.termination, iter_est=bufs["iter_est"][:n].copy(),
                    iter_update=bufs["iter_update"][:n].copy(), lm_iters=bufs["lm_iters"][:n].copy(),
                    lm_cost=bufs["lm_cost"][:n].copy(),
                    edge_assoc=[bufs["edge_assoc"][i, :bufs["n_edge_assoc"][i]].copy() for i in range(n)],
                    plane_assoc=[bufs["plane_assoc"][i, :bufs["n_plane_assoc"][i]].copy() for i in range(n)])

    def register(self, src_edge, src_planar, tgt_edge, tgt_planar, init_pose, rp: CRegParams, want_detail=False):
        se, sp, te, tp = (_xyz64(x) for x in (src_edge, src_planar, tgt_edge, tgt_planar))
        init = np.ascontiguousarray(init_pose, dtype=np.float64)
        out = np.empty(7, dtype=np.float64)
        det, bufs = self._detail_buffers(rp, len(se), len(sp)) if want_detail else (None, None)
        self._check(self.lib.loamgpu_register(self.h, _ptr(se), len(se), _ptr(sp), len(sp), _ptr(te), len(te),
                                              _ptr(tp), len(tp), _ptr(init), C.addressof(rp), _ptr(out),
                                              C.addressof(det) if det is not None else None))
        return (out, self._detail_info(det, bufs)) if want_detail else out

    # ---------------------------------------------------------------- explicit batches
    def extract_batch(self, scans, lp: CLidarParams, fe: CFeParams):
        """scans: [n_scans, R*P, >=3] float32/float64.  Returns a list of (edge_idx, planar_idx) per scan."""
        a = np.asarray(scans)
        if a.ndim != 3:
            raise ValueError("scans must be [n_scans, points, >=3]")
        n_scans, n_per = a.shape[0], a.shape[1]
        a = np.ascontiguousarray(a if a.dtype == np.float32 else a.astype(np.float64))
        dt = F32 if a.dtype == np.float32 else F64
        stride = a.strides[1] if n_per else (16 if dt == F32 else 24)
        cap_e = max(1, int(lp.scan_lines * fe.number_sectors * (fe.max_edge_feats_per_sector + 1)))
        cap_p = max(1, int(lp.scan_lines * fe.number_sectors * (fe.max_planar_feats_per_sector + 1)))
        cap_e, cap_p = min(cap_e, max(n_per, 1)), min(cap_p, max(n_per, 1))
        e = np.zeros((n_scans, cap_e), dtype=np.uint32)
        p = np.zeros((n_scans, cap_p), dtype=np.uint32)
        ne = np.zeros(n_scans, dtype=np.uint32)
        npl = np.zeros(n_scans, dtype=np.uint32)
        self._check(self.lib.loamgpu_extract_batch(self.h, _ptr(a), dt, stride, n_scans, n_per, C.addressof(lp), C.addressof(fe),
                                                   _ptr(e), cap_e, _ptr(ne), _ptr(p), cap_p, _ptr(npl)))
        return [(e[i, :ne[i]].copy(), p[i, :npl[i]].copy()) for i in range(n_scans)]

    def register_pairs(self, pairs, init_poses, rp: CRegParams):
        """pairs: sequence of (src_edge, src_planar, tgt_edge, tgt_planar); init_poses: [n,7] or None.
        Returns poses[n,7], termination[n], iterations[n]."""
        n = len(pairs)
        cols = [[_xyz64(pr[k]) for pr in pairs] for k in range(4)]
        cat = [np.ascontiguousarray(np.concatenate(c)) if n else np.zeros((0, 3)) for c in cols]
        cnt = [np.array([len(x) for x in c], dtype=np.uint64) for c in cols]
        init = None if init_poses is None else np.ascontiguousarray(init_poses, dtype=np.float64).reshape(n, 7)
        poses = np.zeros((n, 7))
        term = np.zeros(n, dtype=np.int32)
        its = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.loamgpu_register_pairs(self.h, n, _ptr(cat[0]), _ptr(cnt[0]), _ptr(cat[1]), _ptr(cnt[1]),
                                                    _ptr(cat[2]), _ptr(cnt[2]), _ptr(cat[3]), _ptr(cnt[3]), _ptr(init),
                                                    C.addressof(rp), _ptr(poses), _ptr(term), _ptr(its)))
        return poses, term, its

    # ---------------------------------------------------------------- device-resident local map
    def map_create(self, edge, planar) -> "DeviceMap":
        e, p = _xyz64(edge), _xyz64(planar)
        h = C.c_void_p()
        self._check(self.lib.loamgpu_map_create(self.h, _ptr(e), len(e), _ptr(p), len(p), C.byref(h)))
        return DeviceMap(self, h)

    def register_to_map(self, dmap: "DeviceMap", src_edge, src_planar, init_pose, rp: CRegParams, want_detail=False):
        se, sp = _xyz64(src_edge), _xyz64(src_planar)
        init = np.ascontiguousarray(init_pose, dtype=np.float64)
        out = np.empty(7, dtype=np.float64)
        det, bufs = self._detail_buffers(rp, len(se), len(sp)) if want_detail else (None, None)
        self._check(self.lib.loamgpu_register_to_map(self.h, dmap.h, _ptr(se), len(se), _ptr(sp), len(sp), _ptr(init),
                                                     C.addressof(rp), _ptr(out),
                                                     C.addressof(det) if det is not None else None))
        return (out, self._detail_info(det, bufs)) if want_detail else out

    def knn(self, targets, queries, k: int, max_dist: float):
        t, q = _xyz64(targets), _xyz64(queries)
        idx = np.full((len(q), k), 0xFFFFFFFF, dtype=np.uint32)
        cnt = np.zeros(len(q), dtype=np.uint32)
        self._check(self.lib.loamgpu_knn(self.h, _ptr(t), len(t), _ptr(q), len(q), k, max_dist, _ptr(idx), _ptr(cnt)))
        return idx, cnt

    # ---------------------------------------------------------------- sequence odometry
    def odometry_host(self, scans: np.ndarray, lp, fe, rp, sweep_motions=None):
        """scans: float32 [n_scans, R*P, 4] host array.  Returns poses[n-1,7], termination, iterations, n_edge, n_planar.
        sweep_motions: optional [n_scans, 7] start_T_end per sweep — the scans are de-warped inside the extraction
        kernel (loamgpu_odometry_host_dewarped)."""
        s = np.ascontiguousarray(scans, dtype=np.float32)
        n = s.shape[0]
        if s.ndim != 3 or s.shape[2] not in (3, 4):
            raise ValueError("scans must be [n_scans, points, 3 or 4] float32 records")
        stride = 4 * s.shape[2]  # packed xyz (12 bytes) or float4 (16 bytes)
        _check_scan_size(s, lp)
        poses = np.zeros((max(n - 1, 0), 7))
        term = np.zeros(max(n - 1, 0), dtype=np.int32)
        its = np.zeros(max(n - 1, 0), dtype=np.uint32)
        ne = np.zeros(n, dtype=np.uint32)
        npl = np.zeros(n, dtype=np.uint32)
        if sweep_motions is None:
            self._check(self.lib.loamgpu_odometry_host_strided(self.h, _ptr(s), stride, n, C.addressof(lp),
                                                               C.addressof(fe), C.addressof(rp), _ptr(poses), _ptr(term),
                                                               _ptr(its), _ptr(ne), _ptr(npl)))
        else:
            m = np.ascontiguousarray(sweep_motions, dtype=np.float64)
            if m.shape != (n, 7):
                raise ValueError("sweep_motions must be [n_scans, 7]")
            if stride != 16:
                raise ValueError("the de-warping sequence call takes float4 records")
            self._check(self.lib.loamgpu_odometry_host_dewarped(self.h, _ptr(s), n, _ptr(m), C.addressof(lp),
                                                                C.addressof(fe), C.addressof(rp), _ptr(poses), _ptr(term),
                                                                _ptr(its), _ptr(ne), _ptr(npl)))
        return poses, term, its, ne, npl

    def odometry_host_ptr(self, scans_ptr: int, n_scans: int, lp, fe, rp, poses_ptr, term_ptr, iters_ptr, ne_ptr,
                          np_ptr, stride: int = 16):
        """`stride` = bytes per float record: 16 (x y z .) or 12 (packed x y z)."""
        self._check(self.lib.loamgpu_odometry_host_strided(self.h, scans_ptr, stride, n_scans, C.addressof(lp),
                                                           C.addressof(fe), C.addressof(rp), poses_ptr, term_ptr,
                                                           iters_ptr, ne_ptr, np_ptr))

    def odometry_host_async_ptr(self, scans_ptr: int, n_scans: int, lp, fe, rp, poses_ptr, term_ptr, iters_ptr, ne_ptr,
                                np_ptr, stride: int = 16):
        """Enqueue only (page-locked host buffers, valid until synchronize()); consecutive calls pipeline."""
        self._check(self.lib.loamgpu_odometry_host_async_strided(self.h, scans_ptr, stride, n_scans, C.addressof(lp),
                                                                 C.addressof(fe), C.addressof(rp), poses_ptr, term_ptr,
                                                                 iters_ptr, ne_ptr, np_ptr))

    def synchronize(self):
        self._check(self.lib.loamgpu_synchronize(self.h))

    def debug_problem_eval(self, is_plane, P, A, B, x, mode=1):
        """TEST HOOK: one evaluation of the LM kernel on explicit residual blocks (tests/residual_cases.py layout) at
        the iterate x.  Returns (H [6,6] = J^T J, g [6] = J^T r, cost, used_moments, planes_streamed)."""
        k = np.asarray(is_plane) != 0
        P, A, B = (np.asarray(v, dtype=np.float64).reshape(len(k), 3) for v in (P, A, B))
        ep, ea, eb = (np.ascontiguousarray(v[~k]) for v in (P, A, B))
        pp, pn = np.ascontiguousarray(P[k]), np.ascontiguousarray(A[k])
        pd = np.ascontiguousarray(B[k, 0])
        xx = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(30)
        self._check(self.lib.loamgpu_debug_problem_eval(self.h, len(ep), _ptr(ep), _ptr(ea), _ptr(eb), len(pp), _ptr(pp),
                                                        _ptr(pn), _ptr(pd), _ptr(xx), int(mode), _ptr(out)))
        H = np.zeros((6, 6))
        t = 0
        for i in range(6):
            for j in range(i, 6):
                H[i, j] = H[j, i] = out[t]
                t += 1
        return H, out[21:27].copy(), float(out[27]), bool(out[28]), int(out[29])

    def odometry_device_ptr(self, scans_ptr: int, n_scans: int, lp, fe, rp, poses_ptr, term_ptr, iters_ptr, ne_ptr,
                            np_ptr, stride: int = 16):
        """All pointers are device addresses (e.g. torch tensor .data_ptr()); asynchronous on the context stream."""
        self._check(self.lib.loamgpu_odometry_device_strided(self.h, scans_ptr, stride, n_scans, C.addressof(lp),
                                                             C.addressof(fe), C.addressof(rp), poses_ptr, term_ptr,
                                                             iters_ptr, ne_ptr, np_ptr))

    def odometry_device_dewarped_ptr(self, scans_ptr: int, n_scans: int, motions_ptr: int, lp, fe, rp, poses_ptr,
                                     term_ptr, iters_ptr, ne_ptr, np_ptr):
        """odometry_device_ptr on sweeps de-warped in the extraction kernel; motions_ptr = device [n_scans, 7] doubles."""
        self._check(self.lib.loamgpu_odometry_device_dewarped(self.h, scans_ptr, n_scans, motions_ptr, C.addressof(lp),
                                                              C.addressof(fe), C.addressof(rp), poses_ptr, term_ptr,
                                                              iters_ptr, ne_ptr, np_ptr))


class MultiContext:
    """One sequence over several GPUs of one box (loamgpu_multi_*): contiguous pair blocks, one host thread + context per
    device, no collective.  `devices` may name a device more than once."""

    def __init__(self, devices):
        self.lib = load_library()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = self.lib.loamgpu_multi_create(arr, len(self.devices), C.byref(h))
        if rc != OK:
            raise LoamGpuError(rc, self.lib.loamgpu_last_error(None).decode())
        self.h = h

    def odometry_host(self, scans: np.ndarray, lp, fe, rp):
        s = np.ascontiguousarray(scans, dtype=np.float32)
        if s.ndim != 3 or s.shape[2] not in (3, 4):
            raise ValueError("scans must be [n_scans, points, 3 or 4] float32 records")
        n = s.shape[0]
        _check_scan_size(s, lp)
        poses = np.zeros((max(n - 1, 0), 7))
        term = np.zeros(max(n - 1, 0), dtype=np.int32)
        its = np.zeros(max(n - 1, 0), dtype=np.uint32)
        ne, npl = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        self.odometry_host_ptr(s.ctypes.data, n, lp, fe, rp, poses.ctypes.data, term.ctypes.data, its.ctypes.data,
                               ne.ctypes.data, npl.ctypes.data, stride=4 * s.shape[2])
        return poses, term, its, ne, npl

    def odometry_host_ptr(self, scans_ptr, n_scans, lp, fe, rp, poses_ptr, term_ptr, iters_ptr, ne_ptr, np_ptr, stride=16):
        rc = self.lib.loamgpu_multi_odometry_host(self.h, scans_ptr, stride, n_scans, C.addressof(lp), C.addressof(fe),
                                                  C.addressof(rp), poses_ptr, term_ptr, iters_ptr, ne_ptr, np_ptr)
        if rc != OK:
            raise LoamGpuError(rc, self.lib.loamgpu_multi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.loamgpu_multi_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMap:
    """Handle of a loamgpu_map: registration target kept on the device (points + NN structures)."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def size(self):
        ne, npl = u64(0), u64(0)
        self.ctx._check(self.ctx.lib.loamgpu_map_size(self.h, C.byref(ne), C.byref(npl)))
        return int(ne.value), int(npl.value)

    def update(self, edge, planar, pose=None, max_edge: int = 0, max_planar: int = 0):
        """Append features (moved into the map frame by `pose` if given), keep the newest max_* points, rebuild."""
        e, p = _xyz64(edge), _xyz64(planar)
        ps = None if pose is None else np.ascontiguousarray(pose, dtype=np.float64)
        self.ctx._check(self.ctx.lib.loamgpu_map_update(self.ctx.h, self.h, _ptr(e), len(e), _ptr(p), len(p), _ptr(ps),
                                                        max_edge, max_planar))

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.loamgpu_map_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_fe_params() -> CFeParams:
    p = CFeParams()
    load_library().loamgpu_default_fe_params(C.byref(p))
    return p


def default_reg_params() -> CRegParams:
    p = CRegParams()
    load_library().loamgpu_default_reg_params(C.byref(p))
    return p
