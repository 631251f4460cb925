"""Shared scenario builders for the parity tests (scenes follow the reference's own tests)."""
import numpy as np

from oracle.pyoracle import FeParams, LidarParams, RegParams  # noqa: F401  (ctypes structs, same layout as the C-ABI)


def frange(a, b, s):
    """C-style `for (double v = a; v < b; v += s)` with float accumulation (tests/test_registration.cpp:27-53)."""
    out = []
    v = float(a)
    while v < b:
        out.append(v)
        v += s
    return out


def simple_scene():
    """constructSimpleScene(), tests/test_registration.cpp:8-56 of the reference: 162 edge + 8941 planar points."""
    pl, ed = [], []
    for y in frange(3, 6, 0.05):
        for z in frange(-1, 2, 0.05):
            pl.append((-3.0, y, z))
    for x in frange(-1, 2, 0.05):
        for z in frange(-1, 2, 0.05):
            pl.append((x, 5.0, z))
    for x in frange(1, 3, 0.05):
        for y in frange(1, 3, 0.05):
            pl.append((x, y, -1.0))
    for z in frange(-1, 3, 0.05):
        ed.append((-1.0, 4.0, z))
    for z in frange(-1, 3, 0.05):
        ed.append((3.0, 2.0, z))
    return np.array(ed), np.array(pl)


def axis_angle(angle, axis):
    axis = np.array(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    return np.r_[np.sin(angle / 2) * axis, np.cos(angle / 2)]


def quat_rotate(q, v):
    """Eigen's q * v for q = (x, y, z, w), rows of v."""
    u = q[:3]
    uv = 2.0 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


def transform(points, pose):
    return quat_rotate(np.asarray(pose[:4]), np.asarray(points)) + np.asarray(pose[4:7])


def quat_mul(a, b):
    return np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                     a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                     a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                     a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])


def angular_distance(q1, q2):
    c = np.array([-q2[0], -q2[1], -q2[2], q2[3]])
    d = quat_mul(q1, c)
    return 2.0 * np.arctan2(np.linalg.norm(d[:3]), abs(d[3]))


def pose_error(source_T_target, target_T_source):
    """err = source_T_target ∘ target_T_source, as in tests/test_registration.cpp:80-86."""
    q = quat_mul(source_T_target[:4], target_T_source[:4])
    t = quat_rotate(source_T_target[:4], target_T_source[4:7]) + source_T_target[4:7]
    return angular_distance(q, np.array([0, 0, 0, 1.0])), t


Q_SIMPLE = [0.014692022378442412, 0.030140550562090015, 0.009544316157523478, 0.9993921140970299]

# (name, source_T_target, init, max_iterations, rot tol, trans tol) — tests/test_registration.cpp:69-175
REG_SCENARIOS = [
    ("simple", np.r_[Q_SIMPLE, [0.01, 0.03, -0.01]], None, 10, 1e-4, 1e-4),
    ("large_translation", np.r_[Q_SIMPLE, [-0.1, 0.1, 0.0]], None, 10, 1e-4, 1e-3),
    ("larger_translation", np.r_[Q_SIMPLE, [-0.3, 0.2, 0.1]], None, 10, 1e-4, 1e-3),
    ("large_rotation", np.r_[axis_angle(0.2, [1, 3, 1]), [-0.01, 0.02, 0.1]], None, 10, 1e-4, 1e-3),
    ("composition_direction", np.r_[axis_angle(0.1, [0, 0, 1]), [0, 0, 0]],
     np.r_[axis_angle(-0.1, [0, 0, 1]), [0.1, 0, 0]], 1, 1e-4, 1e-3),
]

# pose parity tolerance between the CUDA path and the oracle (BASELINE.json north_star)
POSE_TOL_RAD = 1e-6
POSE_TOL_M = 1e-5


def to_capi(s):
    """Oracle ctypes struct -> the C-ABI struct of the same layout (LidarParams / FeParams / RegParams)."""
    import ctypes as C

    from loam_b200 import _capi
    cls = {"LidarParams": _capi.CLidarParams, "FeParams": _capi.CFeParams, "RegParams": _capi.CRegParams}[type(s).__name__]
    assert C.sizeof(cls) == C.sizeof(s)
    c = cls()
    C.memmove(C.addressof(c), C.addressof(s), C.sizeof(s))
    return c
