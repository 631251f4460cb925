"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracle.

`Oracle`  wraps oracle/liboracle.so      (C restatement, loam_oracle.c)
`RefLib`  wraps oracle/_ref/libloam_ref.so (the REAL reference feature code; see ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (loam_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u64, f64, u32, i32 = C.c_uint64, C.c_double, C.c_uint32, C.c_int32
PD = C.POINTER(C.c_double)
PU32 = C.POINTER(C.c_uint32)


class LidarParams(C.Structure):
    _fields_ = [("scan_lines", u64), ("points_per_line", u64), ("min_range", f64), ("max_range", f64)]


class FeParams(C.Structure):
    _fields_ = [("neighbor_points", u64), ("number_sectors", u64), ("max_edge_feats_per_sector", u64),
                ("max_planar_feats_per_sector", u64), ("edge_feat_threshold", f64), ("planar_feat_threshold", f64),
                ("occlusion_thresh", f64), ("parallel_thresh", f64)]

    @staticmethod
    def default():
        return FeParams(3, 6, 10, 50, 100.0, 1.0, 0.5, 1.0)


class RegParams(C.Structure):
    _fields_ = [("num_edge_neighbors", u64), ("max_edge_neighbor_dist", f64), ("min_line_fit_points", u64),
                ("min_line_condition_number", f64), ("num_plane_neighbors", u64), ("max_plane_neighbor_dist", f64),
                ("min_plane_fit_points", u64), ("max_avg_point_plane_dist", f64), ("max_iterations", u64),
                ("rotation_convergence_thresh", f64), ("position_convergence_thresh", f64), ("min_associations", u64)]

    @staticmethod
    def default():
        return RegParams(5, 1.0, 3, 10.0, 5, 2.0, 4, 0.1, 10, 1e-3, 1e-2, 100)


class _Detail(C.Structure):
    _fields_ = [("max_iters_cap", u32), ("n_src_edge", u32), ("n_src_planar", u32), ("n_iters", u32),
                ("termination", i32), ("iter_est", PD), ("iter_update", PD), ("n_edge_assoc", PU32),
                ("n_plane_assoc", PU32), ("edge_assoc", PU32), ("plane_assoc", PU32), ("lm_iters", PU32),
                ("lm_cost", PD)]


@dataclass
class Detail:
    """Python view of RegistrationDetail (registration.h:79-109 of the reference)."""
    n_iters: int = 0
    termination: int = 1
    iter_est: np.ndarray = None
    iter_update: np.ndarray = None
    edge_assoc: list = field(default_factory=list)   # per iteration: [n,2] uint32
    plane_assoc: list = field(default_factory=list)
    lm_iters: np.ndarray = None
    lm_cost: np.ndarray = None


def _as_xyz(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.size == 0:
        return np.zeros((0, 3))
    a = a.reshape(len(a), -1)[:, :3]
    return np.ascontiguousarray(a)


def ensure_built():
    """Build liboracle.so (and _ref when the reference tree is present)."""
    if not os.path.exists(os.path.join(HERE, "liboracle.so")) or (
            os.path.isdir("/root/reference/loam/include")
            and not os.path.exists(os.path.join(HERE, "_ref", "libloam_ref.so"))):
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self):
        ensure_built()
        self.lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L = self.lib
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_build.argtypes = [C.c_void_p, u64]
        L.orc_kdtree_free.argtypes = [C.c_void_p]
        L.orc_kdtree_knn.restype = u32
        L.orc_kdtree_knn.argtypes = [C.c_void_p, C.c_void_p, u32, f64, C.c_void_p, C.c_void_p]
        L.orc_knn_brute.restype = u32
        L.orc_knn_brute.argtypes = [C.c_void_p, u64, C.c_void_p, u32, f64, C.c_void_p, C.c_void_p]
        L.orc_fit_line.restype = f64
        L.orc_fit_plane.restype = f64
        L.orc_point_to_line.restype = f64
        L.orc_point_to_plane.restype = f64
        L.orc_point_to_plane.argtypes = [C.c_void_p, C.c_void_p, f64]
        L.orc_quat_angular_distance.restype = f64
        L.orc_residual_eval.restype = f64
        L.orc_residual_eval.argtypes = [i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_problem_eval.restype = f64
        L.orc_problem_eval.argtypes = [C.c_void_p] * 4 + [u64] + [C.c_void_p] * 4
        L.orc_manifold_plus.argtypes = [C.c_void_p] * 3
        L.orc_lm_solve.restype = u32
        L.orc_lm_solve.argtypes = [C.c_void_p] * 4 + [u64, C.c_void_p, C.c_int, C.c_void_p]

    # ---- features ----
    def curvature(self, xyz, lp: LidarParams, fe: FeParams):
        xyz = _as_xyz(xyz)
        out = np.empty(len(xyz), dtype=np.float64)
        rc = self.lib.orc_curvature(xyz.ctypes, u64(len(xyz)), C.byref(lp), C.byref(fe), out.ctypes)
        if rc:
            raise RuntimeError("scan size mismatch")
        return out

    def valid_mask(self, xyz, lp, fe):
        xyz = _as_xyz(xyz)
        out = np.empty(len(xyz), dtype=np.uint8)
        rc = self.lib.orc_valid_mask(xyz.ctypes, u64(len(xyz)), C.byref(lp), C.byref(fe), out.ctypes)
        if rc:
            raise RuntimeError("scan size mismatch")
        return out.astype(bool)

    def extract(self, xyz, lp, fe, return_ties=False):
        xyz = _as_xyz(xyz)
        n = len(xyz)
        e = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        ne, npl, ties = u64(0), u64(0), u64(0)
        rc = self.lib.orc_extract(xyz.ctypes, u64(n), C.byref(lp), C.byref(fe), e.ctypes, C.byref(ne), p.ctypes,
                                  C.byref(npl), C.byref(ties))
        if rc:
            raise RuntimeError("orc_extract failed rc=%d" % rc)
        if return_ties:
            return e[:ne.value].copy(), p[:npl.value].copy(), ties.value
        return e[:ne.value].copy(), p[:npl.value].copy()

    # ---- kNN / geometry ----
    def knn_brute(self, pts, q, k, max_dist):
        pts = _as_xyz(pts)
        q = np.ascontiguousarray(q, dtype=np.float64)
        idx = np.empty(k, dtype=np.uint32)
        d2 = np.empty(k, dtype=np.float64)
        m = self.lib.orc_knn_brute(pts.ctypes, len(pts), q.ctypes, k, max_dist, idx.ctypes, d2.ctypes)
        return idx[:m].copy(), d2[:m].copy()

    def knn_tree_batch(self, pts, queries, k, max_dist):
        pts = _as_xyz(pts)
        queries = _as_xyz(queries)
        t = self.lib.orc_kdtree_build(pts.ctypes, len(pts))
        out = np.full((len(queries), k), 0xFFFFFFFF, dtype=np.uint32)
        cnt = np.zeros(len(queries), dtype=np.uint32)
        idx = np.empty(k, dtype=np.uint32)
        for i in range(len(queries)):
            m = self.lib.orc_kdtree_knn(t, queries[i].ctypes, k, max_dist, idx.ctypes, None)
            out[i, :m] = idx[:m]
            cnt[i] = m
        self.lib.orc_kdtree_free(t)
        return out, cnt

    def fit_line(self, pts):
        pts = _as_xyz(pts)
        a, b = np.empty(3), np.empty(3)
        cond = self.lib.orc_fit_line(pts.ctypes, u32(len(pts)), a.ctypes, b.ctypes)
        return a, b, cond

    def fit_plane(self, pts):
        pts = _as_xyz(pts)
        n = np.empty(3)
        d = f64(0)
        avg = self.lib.orc_fit_plane(pts.ctypes, u32(len(pts)), n.ctypes, C.byref(d))
        return n, d.value, avg

    def point_to_line(self, p, a, b):
        p, a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (p, a, b))
        return self.lib.orc_point_to_line(p.ctypes, a.ctypes, b.ctypes)

    def point_to_plane(self, p, n, d):
        p, n = (np.ascontiguousarray(v, dtype=np.float64) for v in (p, n))
        return self.lib.orc_point_to_plane(p.ctypes, n.ctypes, d)

    def pose_compose(self, p1, p2):
        p1, p2 = (np.ascontiguousarray(v, dtype=np.float64) for v in (p1, p2))
        o = np.empty(7)
        self.lib.orc_pose_compose(p1.ctypes, p2.ctypes, o.ctypes)
        return o

    def pose_inverse(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        o = np.empty(7)
        self.lib.orc_pose_inverse(p.ctypes, o.ctypes)
        return o

    def pose_act(self, pose, pt):
        pose, pt = (np.ascontiguousarray(v, dtype=np.float64) for v in (pose, pt))
        o = np.empty(3)
        self.lib.orc_pose_act(pose.ctypes, pt.ctypes, o.ctypes)
        return o

    def dewarp(self, xyz, points_per_line, start_T_end):
        xyz = _as_xyz(xyz)
        m = np.ascontiguousarray(start_T_end, dtype=np.float64)
        out = np.empty_like(xyz)
        self.lib.orc_dewarp(xyz.ctypes, u64(len(xyz)), u64(points_per_line), m.ctypes, out.ctypes)
        return out

    def angular_distance(self, q1, q2):
        q1, q2 = (np.ascontiguousarray(v, dtype=np.float64) for v in (q1, q2))
        return self.lib.orc_quat_angular_distance(q1.ctypes, q2.ctypes)

    # ---- test hooks: residuals / Jacobians / the restated ceres::Solve on explicit residual blocks ----
    @staticmethod
    def _blocks(is_plane, P, A, B):
        k = np.ascontiguousarray(is_plane, dtype=np.int32)
        P, A, B = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(len(k), 3)) for v in (P, A, B))
        return k, P, A, B

    def residual_eval(self, is_plane, p, a, b, x):
        p, a, b, x = (np.ascontiguousarray(v, dtype=np.float64) for v in (p, a, b, x))
        J7 = np.empty(7)
        r = self.lib.orc_residual_eval(int(is_plane), p.ctypes.data, a.ctypes.data, b.ctypes.data, x.ctypes.data,
                                       J7.ctypes.data)
        return r, J7

    def problem_eval(self, is_plane, P, A, B, x):
        k, P, A, B = self._blocks(is_plane, P, A, B)
        x = np.ascontiguousarray(x, dtype=np.float64)
        M = len(k)
        r, J, g = np.empty(M), np.empty((M, 6)), np.empty(6)
        cost = self.lib.orc_problem_eval(k.ctypes.data, P.ctypes.data, A.ctypes.data, B.ctypes.data, M, x.ctypes.data,
                                         r.ctypes.data, J.ctypes.data, g.ctypes.data)
        return cost, r, J, g

    def manifold_plus(self, x, delta):
        x, delta = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, delta))
        out = np.empty(7)
        self.lib.orc_manifold_plus(x.ctypes.data, delta.ctypes.data, out.ctypes.data)
        return out

    def lm_solve(self, is_plane, P, A, B, x0=None, armed_flag=True):
        k, P, A, B = self._blocks(is_plane, P, A, B)
        x = np.array([0, 0, 0, 1, 0, 0, 0.0] if x0 is None else x0, dtype=np.float64)
        cost2 = np.zeros(2)
        it = self.lib.orc_lm_solve(k.ctypes.data, P.ctypes.data, A.ctypes.data, B.ctypes.data, len(k), x.ctypes.data,
                                   1 if armed_flag else 0, cost2.ctypes.data)
        return x, int(it), cost2

    # ---- registration ----
    def register(self, src_edge, src_planar, tgt_edge, tgt_planar, init_pose=None, rp: RegParams | None = None,
                 want_detail=False, use_kdtree=True, armed_flag=True):
        se, sp, te, tp = (_as_xyz(a) for a in (src_edge, src_planar, tgt_edge, tgt_planar))
        rp = rp or RegParams.default()
        init = np.ascontiguousarray(init_pose if init_pose is not None else [0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
        out = np.empty(7)
        det_ptr = None
        if want_detail:
            cap = max(int(rp.max_iterations), 1)
            bufs = dict(
                iter_est=np.zeros((cap, 7)), iter_update=np.zeros((cap, 7)),
                n_edge_assoc=np.zeros(cap, dtype=np.uint32), n_plane_assoc=np.zeros(cap, dtype=np.uint32),
                edge_assoc=np.zeros((cap, max(len(se), 1), 2), dtype=np.uint32),
                plane_assoc=np.zeros((cap, max(len(sp), 1), 2), dtype=np.uint32),
                lm_iters=np.zeros(cap, dtype=np.uint32), lm_cost=np.zeros((cap, 2)))
            d = _Detail(cap, max(len(se), 1), max(len(sp), 1), 0, 1,
                        bufs["iter_est"].ctypes.data_as(PD), bufs["iter_update"].ctypes.data_as(PD),
                        bufs["n_edge_assoc"].ctypes.data_as(PU32), bufs["n_plane_assoc"].ctypes.data_as(PU32),
                        bufs["edge_assoc"].ctypes.data_as(PU32), bufs["plane_assoc"].ctypes.data_as(PU32),
                        bufs["lm_iters"].ctypes.data_as(PU32), bufs["lm_cost"].ctypes.data_as(PD))
            det_ptr = C.byref(d)
        rc = self.lib.orc_register(se.ctypes, u64(len(se)), sp.ctypes, u64(len(sp)), te.ctypes, u64(len(te)),
                                   tp.ctypes, u64(len(tp)), init.ctypes, C.byref(rp), out.ctypes, det_ptr,
                                   C.c_int(1 if use_kdtree else 0), C.c_int(1 if armed_flag else 0))
        if rc:
            raise RuntimeError("orc_register failed rc=%d" % rc)
        if not want_detail:
            return out
        n = d.n_iters
        det = Detail(n_iters=n, termination=d.termination, iter_est=bufs["iter_est"][:n].copy(),
                     iter_update=bufs["iter_update"][:n].copy(), lm_iters=bufs["lm_iters"][:n].copy(),
                     lm_cost=bufs["lm_cost"][:n].copy())
        for i in range(n):
            det.edge_assoc.append(bufs["edge_assoc"][i, :bufs["n_edge_assoc"][i]].copy())
            det.plane_assoc.append(bufs["plane_assoc"][i, :bufs["n_plane_assoc"][i]].copy())
        return out, det


class RefLib:
    """The real reference feature-extraction code (compiled from /root/reference, see oracle/Makefile)."""

    def __init__(self):
        ensure_built()
        path = os.path.join(HERE, "_ref", "libloam_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_last_error.restype = C.c_char_p

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libloam_ref.so")) or os.path.isdir(
            "/root/reference/loam/include")

    @staticmethod
    def _pack(lp: LidarParams, fe: FeParams):
        lu = (u64 * 2)(lp.scan_lines, lp.points_per_line)
        ld = (f64 * 2)(lp.min_range, lp.max_range)
        fu = (u64 * 4)(fe.neighbor_points, fe.number_sectors, fe.max_edge_feats_per_sector,
                       fe.max_planar_feats_per_sector)
        fd = (f64 * 4)(fe.edge_feat_threshold, fe.planar_feat_threshold, fe.occlusion_thresh, fe.parallel_thresh)
        return lu, ld, fu, fd

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def extract(self, xyz, lp, fe):
        xyz = _as_xyz(xyz)
        n = len(xyz)
        e = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        ne, npl = u64(0), u64(0)
        self._check(self.lib.ref_extract_f64(xyz.ctypes, u64(n), *self._pack(lp, fe), e.ctypes, C.byref(ne),
                                             p.ctypes, C.byref(npl)))
        return e[:ne.value].copy(), p[:npl.value].copy()

    def curvature(self, xyz, lp, fe):
        xyz = _as_xyz(xyz)
        c = np.empty(len(xyz))
        idx = np.empty(len(xyz), dtype=np.uint64)
        self._check(self.lib.ref_curvature_f64(xyz.ctypes, u64(len(xyz)), *self._pack(lp, fe), c.ctypes, idx.ctypes))
        return c, idx

    def valid_mask(self, xyz, lp, fe):
        xyz = _as_xyz(xyz)
        m = np.empty(len(xyz), dtype=np.uint8)
        self._check(self.lib.ref_valid_f64(xyz.ctypes, u64(len(xyz)), *self._pack(lp, fe), m.ctypes))
        return m.astype(bool)

    def extract_timed_f32x4(self, scan_f32x4, lp, fe, reps=5):
        s = np.ascontiguousarray(scan_f32x4, dtype=np.float32)
        n = len(s)
        e = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        ne, npl, best, mean = u64(0), u64(0), f64(0), f64(0)
        self._check(self.lib.ref_extract_f32x4_timed(s.ctypes, u64(n), *self._pack(lp, fe), C.c_int(reps),
                                                     C.byref(best), C.byref(mean), e.ctypes, C.byref(ne), p.ctypes,
                                                     C.byref(npl)))
        return best.value, mean.value, e[:ne.value].copy(), p[:npl.value].copy()
