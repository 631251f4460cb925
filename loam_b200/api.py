"""Host-side mirror of the reference's Python module `loam_python` (python/loam_bindings.cpp:11-145).

Same names, argument order and error behaviour as the reference bindings:
  extractFeatures(input_scan, lidar_params, params)            loam_bindings.cpp:85-86
  computeCurvature / computeValidPoints                        loam_bindings.cpp:88-92
  registerFeatures(source, target, target_T_source_init, params, detail)   loam_bindings.cpp:141-144
All arithmetic of the hot path happens in the CUDA kernels behind the C-ABI (include/loamgpu.h);
this file only packs numpy buffers.  Pose3d/Quaterniond are the boundary value types
(geometry.h:27-50) and are plain host objects, as in the reference.

Fast path: `input_scan` may be one contiguous (N,3)/(N,4) float32 or float64 ndarray, which
is handed to the C-ABI without a per-point conversion; a sequence of per-point arrays (what the reference
bindings marshal) is also accepted.
"""
from __future__ import annotations

import enum
import math
import threading

import numpy as np

from . import _capi


# ----------------------------------------------------------------------------------------------- params
class LidarParams:
    """common.h:29-41 (members are read-only, like the reference's const members)."""

    def __init__(self, scan_lines: int, points_per_line: int, min_range: float, max_range: float):
        self._c = _capi.CLidarParams(int(scan_lines), int(points_per_line), float(min_range), float(max_range))

    scan_lines = property(lambda s: s._c.scan_lines)
    points_per_line = property(lambda s: s._c.points_per_line)
    min_range = property(lambda s: s._c.min_range)
    max_range = property(lambda s: s._c.max_range)


def _struct_params(cls_name, cstruct, defaults, doc):
    fields = [f for f, _ in cstruct._fields_]

    class P:
        __doc__ = doc

        def __init__(self, **kw):
            for f in fields:
                setattr(self, f, defaults[f])
            for k, v in kw.items():
                if k not in fields:
                    raise TypeError(f"{cls_name} has no field {k}")
                setattr(self, k, v)

        def _to_c(self):
            c = cstruct()
            for f, t in cstruct._fields_:
                v = getattr(self, f)
                setattr(c, f, int(v) if t is _capi.u64 else float(v))
            return c

        def __repr__(self):
            return cls_name + "(" + ", ".join(f"{f}={getattr(self, f)!r}" for f in fields) + ")"

    P.__name__ = P.__qualname__ = cls_name
    return P


FeatureExtractionParams = _struct_params(
    "FeatureExtractionParams", _capi.CFeParams,
    dict(neighbor_points=3, number_sectors=6, max_edge_feats_per_sector=10, max_planar_feats_per_sector=50,
         edge_feat_threshold=100.0, planar_feat_threshold=1.0, occlusion_thresh=0.5, parallel_thresh=1.0),
    "features.h:37-66")

RegistrationParams = _struct_params(
    "RegistrationParams", _capi.CRegParams,
    dict(num_edge_neighbors=5, max_edge_neighbor_dist=1.0, min_line_fit_points=3, min_line_condition_number=10.0,
         num_plane_neighbors=5, max_plane_neighbor_dist=2.0, min_plane_fit_points=4, max_avg_point_plane_dist=0.1,
         max_iterations=10, rotation_convergence_thresh=1e-3, position_convergence_thresh=1e-2,
         min_associations=100),
    "registration.h:40-75")


# ----------------------------------------------------------------------------------------------- geometry value types
class Quaterniond:
    """Eigen::Quaterniond as bound by the reference: ctor (w, x, y, z), accessors w() x() y() z()."""

    def __init__(self, w=1.0, x=0.0, y=0.0, z=0.0):
        self._q = np.array([x, y, z, w], dtype=np.float64)  # Eigen coeffs() order

    def w(self): return float(self._q[3])
    def x(self): return float(self._q[0])
    def y(self): return float(self._q[1])
    def z(self): return float(self._q[2])

    def coeffs(self):
        return self._q.copy()

    @staticmethod
    def from_coeffs(xyzw):
        q = Quaterniond()
        q._q = np.array(xyzw, dtype=np.float64)
        return q

    def __mul__(self, o):
        if isinstance(o, Quaterniond):
            a, b = self._q, o._q
            return Quaterniond(a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2],
                               a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                               a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                               a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0])
        v = np.asarray(o, dtype=np.float64)
        u = self._q[:3]
        uv = 2.0 * np.cross(u, v)
        return v + self._q[3] * uv + np.cross(u, uv)

    def conjugate(self):
        return Quaterniond(self._q[3], -self._q[0], -self._q[1], -self._q[2])

    def inverse(self):
        n2 = float(self._q @ self._q)
        return Quaterniond(self._q[3] / n2, -self._q[0] / n2, -self._q[1] / n2, -self._q[2] / n2)

    def angularDistance(self, other):
        d = self * other.conjugate()
        return 2.0 * math.atan2(float(np.linalg.norm(d._q[:3])), abs(float(d._q[3])))

    def __repr__(self):
        return f"Quaterniond(w={self.w()}, x={self.x()}, y={self.y()}, z={self.z()})"


class Pose3d:
    """geometry.h:27-50 / geometry.cpp:10-29."""

    def __init__(self, rotation: Quaterniond | None = None, translation=None):
        self.rotation = rotation if rotation is not None else Quaterniond()
        self.translation = np.zeros(3) if translation is None else np.array(translation, dtype=np.float64)

    @staticmethod
    def Identity():
        return Pose3d()

    def inverse(self):
        inv = self.rotation.inverse()
        return Pose3d(inv, inv * (-self.translation))

    def compose(self, other):
        return Pose3d(self.rotation * other.rotation, self.translation + self.rotation * other.translation)

    def act(self, point):
        return self.rotation * np.asarray(point, dtype=np.float64) + self.translation

    def matrix(self):
        x, y, z, w = self.rotation._q
        m = np.eye(4)
        m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        m[:3, 3] = self.translation
        return m

    def _to7(self):
        return np.concatenate([self.rotation._q, self.translation])

    @staticmethod
    def _from7(v):
        return Pose3d(Quaterniond.from_coeffs(v[:4]), np.array(v[4:7]))

    def __repr__(self):
        return f"Pose3d({self.rotation!r}, {self.translation!r})"


# ----------------------------------------------------------------------------------------------- features
class LoamFeatures:
    """features.h:70-76: edge_points / planar_points (here (n, d) arrays whose rows are copies of input points)."""

    def __init__(self, edge_points=None, planar_points=None):
        self.edge_points = np.zeros((0, 3)) if edge_points is None else edge_points
        self.planar_points = np.zeros((0, 3)) if planar_points is None else planar_points


class RegistrationTerminationType(enum.IntEnum):
    CONVERGED = 0
    MAX_ITER = 1
    INSUFFICIENT_ASSOCIATIONS = 2


class RegistrationIterationInfo:
    """registration.h:86-104"""

    def __init__(self, target_T_source_init, edge_associations, plane_associations, estimate_update):
        self.target_T_source_init = target_T_source_init
        self.edge_associations = edge_associations
        self.plane_associations = plane_associations
        self.estimate_update = estimate_update


class RegistrationDetail:
    """registration.h:79-109: appended to by registerFeatures (registration-inl.h:59-61,76)."""

    def __init__(self):
        self.iteration_info = []
        self.termination_type = RegistrationTerminationType.MAX_ITER


# ----------------------------------------------------------------------------------------------- context
_tls = threading.local()


class _ThreadContexts:
    """This thread's contexts, one per device.  Lives in threading.local(): when the thread ends the object is collected
    and its contexts (each owns device workspaces) are closed, like the C++ side's thread_local ThreadContext."""

    def __init__(self):
        self.by_device = {}

    def __del__(self):
        for c in self.by_device.values():
            try:
                c.close()
            except Exception:
                pass


def get_context(device: int = 0) -> _capi.Context:
    """Per-(thread, device) context: the reference is stateless/re-entrant; so are these calls."""
    tc = getattr(_tls, "contexts", None)
    if tc is None:
        tc = _tls.contexts = _ThreadContexts()
    c = tc.by_device.get(device)
    if c is None:
        c = tc.by_device[device] = _capi.Context(device)
    return c


def release_context(device=None):
    """Close this thread's context(s) now (device=None: all of them) instead of at thread exit."""
    tc = getattr(_tls, "contexts", None)
    if tc is None:
        return
    for d in list(tc.by_device) if device is None else [device]:
        c = tc.by_device.pop(d, None)
        if c is not None:
            c.close()


def _as_cloud(input_scan):
    if isinstance(input_scan, np.ndarray) and input_scan.ndim == 2:
        return input_scan
    pts = [np.asarray(p, dtype=np.float64).reshape(-1) for p in input_scan]
    if not pts:
        return np.zeros((0, 3))
    return np.stack(pts)


def _map_error(e: _capi.LoamGpuError):
    # reference: std::runtime_error on size mismatch (common.h:104-113) -> pybind RuntimeError
    if e.code == _capi.ERR_SIZE_MISMATCH:
        return RuntimeError(str(e))
    if e.code == _capi.ERR_INVALID:
        return ValueError(str(e))
    return e


# ----------------------------------------------------------------------------------------------- entry points
def extractFeatureIndices(input_scan, lidar_params: LidarParams, params=None, device: int = 0):
    """Indices (edge, planar) into input_scan in the reference's output order."""
    params = params or FeatureExtractionParams()
    cloud = _as_cloud(input_scan)
    try:
        return get_context(device).extract(cloud, lidar_params._c, params._to_c())
    except _capi.LoamGpuError as e:
        raise _map_error(e) from None


def extractFeatures(input_scan, lidar_params: LidarParams, params=None, device: int = 0) -> LoamFeatures:
    """loam::extractFeatures (features.h:108-111)."""
    cloud = _as_cloud(input_scan)
    e, p = extractFeatureIndices(cloud, lidar_params, params, device)
    return LoamFeatures(cloud[e].copy(), cloud[p].copy())


def extractFeaturesDewarped(input_scan, lidar_params: LidarParams, start_T_end: Pose3d, params=None,
                            device: int = 0) -> LoamFeatures:
    """Extension (mirrors loam::extractFeaturesDewarped, include/loam/features.h): the de-warp the reference leaves to
    its caller (README.md:63) fused into extractFeatures.  Column c of each ring is moved into the frame of the sweep
    start by interp(Identity, start_T_end, c / points_per_line); the returned feature points are the moved points
    (float64), ready for registerFeatures."""
    params = params or FeatureExtractionParams()
    cloud = _as_cloud(input_scan)
    try:
        e, p, moved = get_context(device).extract_dewarped(cloud, lidar_params._c, params._to_c(), start_T_end._to7())
    except _capi.LoamGpuError as err:
        raise _map_error(err) from None
    return LoamFeatures(moved[e], moved[p])


def computeCurvature(input_scan, lidar_params: LidarParams, params=None, device: int = 0):
    """loam::computeCurvature (features.h:119-122): structured array with fields index, curvature."""
    params = params or FeatureExtractionParams()
    cloud = _as_cloud(input_scan)
    try:
        c = get_context(device).curvature(cloud, lidar_params._c, params._to_c())
    except _capi.LoamGpuError as e:
        raise _map_error(e) from None
    out = np.empty(len(c), dtype=[("index", np.uint64), ("curvature", np.float64)])
    out["index"] = np.arange(len(c), dtype=np.uint64)
    out["curvature"] = c
    return out


def computeValidPoints(input_scan, lidar_params: LidarParams, params=None, device: int = 0):
    """loam::computeValidPoints (features.h:166-169)."""
    params = params or FeatureExtractionParams()
    cloud = _as_cloud(input_scan)
    try:
        return get_context(device).valid_mask(cloud, lidar_params._c, params._to_c())
    except _capi.LoamGpuError as e:
        raise _map_error(e) from None


def _fill_detail(detail, info):
    for i in range(info["n_iters"]):
        detail.iteration_info.append(RegistrationIterationInfo(
            Pose3d._from7(info["iter_est"][i]), [tuple(map(int, r)) for r in info["edge_assoc"][i]],
            [tuple(map(int, r)) for r in info["plane_assoc"][i]], Pose3d._from7(info["iter_update"][i])))
    detail.termination_type = RegistrationTerminationType(info["termination"])


class LocalMap:
    """Device-resident registration target (extension, mirrors include/loam/local_map.h): the reference leaves
    "maintain a local map of points" to its caller (README.md:63), who passes the accumulated map as `target` and pays
    a KD-tree build over all of it per call.  A LocalMap keeps the points and their NN structures on the GPU."""

    def __init__(self, features: LoamFeatures | None = None, device: int = 0):
        features = features or LoamFeatures()
        self._device = device
        try:
            self._map = get_context(device).map_create(_as_cloud(features.edge_points), _as_cloud(features.planar_points))
        except _capi.LoamGpuError as e:
            raise _map_error(e) from None

    def insert(self, features: LoamFeatures, map_T_features: Pose3d | None = None, max_edge: int = 0, max_planar: int = 0):
        """Append features (moved into the map frame by map_T_features), keep the newest max_* points, rebuild."""
        try:
            self._map.update(_as_cloud(features.edge_points), _as_cloud(features.planar_points),
                             None if map_T_features is None else map_T_features._to7(), max_edge, max_planar)
        except _capi.LoamGpuError as e:
            raise _map_error(e) from None

    def size(self):
        return self._map.size()


def registerFeatures(source: LoamFeatures, target, target_T_source_init: Pose3d, params=None,
                     detail: RegistrationDetail | None = None, device: int = 0) -> Pose3d:
    """loam::registerFeatures (registration.h:128-131); `target` may also be a LocalMap."""
    params = params or RegistrationParams()
    src = (_as_cloud(source.edge_points), _as_cloud(source.planar_points))
    try:
        if isinstance(target, LocalMap):
            ctx = get_context(target._device)
            call = lambda **kw: ctx.register_to_map(target._map, *src, target_T_source_init._to7(), params._to_c(), **kw)
        else:
            ctx = get_context(device)
            call = lambda **kw: ctx.register(*src, _as_cloud(target.edge_points), _as_cloud(target.planar_points),
                                             target_T_source_init._to7(), params._to_c(), **kw)
        if detail is None:
            return Pose3d._from7(call())
        pose, info = call(want_detail=True)
    except _capi.LoamGpuError as e:
        raise _map_error(e) from None
    _fill_detail(detail, info)
    return Pose3d._from7(pose)


_multi_lock = threading.Lock()
_multis: dict = {}


def get_multi_context(devices) -> _capi.MultiContext:
    """One cached set of per-device contexts per (thread, device list)."""
    key = (threading.get_ident(), tuple(int(d) for d in devices))
    with _multi_lock:
        m = _multis.get(key)
        if m is None:
            m = _capi.MultiContext(key[1])
            _multis[key] = m
        return m


def odometry(scans, lidar_params: LidarParams, fe_params=None, reg_params=None, device: int = 0, sweep_motions=None,
             devices=None):
    """Batched extract + scan-to-scan registration over a float32 [n_scans, R*P, 4] (or packed [n_scans, R*P, 3])
    sequence (the README loop of the reference, run for the whole sequence in one call).
    Returns (poses[n-1,7] as qx qy qz qw tx ty tz, termination, iterations, n_edge, n_planar).
    sweep_motions ([n_scans, 7] start_T_end per sweep, optional): de-warp every scan inside the extraction kernel.
    devices (list of CUDA device indices, optional): shard the pairs of the sequence over several GPUs of this box —
    contiguous pair blocks, one host thread and context per device, no collective; same results as one device."""
    fe_params = fe_params or FeatureExtractionParams()
    reg_params = reg_params or RegistrationParams()
    try:
        if devices is not None and len(devices) > 0:
            if sweep_motions is not None:
                raise ValueError("sweep_motions is not supported together with devices=")
            return get_multi_context(devices).odometry_host(scans, lidar_params._c, fe_params._to_c(), reg_params._to_c())
        return get_context(device).odometry_host(scans, lidar_params._c, fe_params._to_c(), reg_params._to_c(),
                                                 sweep_motions)
    except _capi.LoamGpuError as e:
        raise _map_error(e) from None
