"""The reference's feature-extraction unit scenes (tests/test_feature_extraction.cpp) as data.

Each case: (name, points (n,3), LidarParams args, expected) — run through the oracle, the real reference
(oracle/_ref) and the CUDA path by the tests.
"""
import numpy as np

# FeatureExtractionParams{5, 6, 5, 5, 100, 0.1, 0.25, 0.02} used by every reference test
FE_TEST = (5, 6, 5, 5, 100.0, 0.1, 0.25, 0.02)


def curvature_plane():  # test_feature_extraction.cpp:27-53
    pts = np.array([(i, 1, 0.0) for i in range(-5, 6)], dtype=np.float64)
    expect = np.array([-1] * 5 + [0.0] + [-1] * 5, dtype=np.float64)
    return pts, (1, 11, 0.1, 10.0), expect


def curvature_corner():  # :55-84  known answer 900.0
    pts = np.array([(i, abs(i) + 1, 0.0) for i in range(-5, 6)], dtype=np.float64)
    expect = np.array([-1] * 5 + [900.0] + [-1] * 5, dtype=np.float64)
    return pts, (1, 11, 0.1, 50.0), expect


def invalid_edges():  # :96-122
    pts = np.array([(i * 0.1, 1, 0.0) for i in range(-5, 6)], dtype=np.float64)
    expect = {i: False for i in list(range(5)) + list(range(6, 11))}
    expect[5] = True
    return pts, (1, 11, 0.1, 50.0), expect


def invalid_ranges():  # :124-155
    pts = [(i, 1, 0.0) for i in range(-5, 0)] + [(-0.5, 20.0, 0.0), (0.0, 0.2, 0.0)] + [(i, 1, 0.0) for i in range(1, 6)]
    expect = {i: False for i in list(range(5)) + [10 - i for i in range(5)]}
    expect[5] = False
    expect[6] = False
    return np.array(pts, dtype=np.float64), (1, 12, 0.5, 6.0), expect


def occlusion_case1():  # :157-190
    pts = [(i * 0.1, 4.0, 0.0) for i in range(-15, 0)] + [(i * 0.1, 6.0, 0.0) for i in range(0, 15)]
    expect = {}
    for i in range(5):
        expect[i] = False
        expect[29 - i] = False
    for i in range(5, 15):
        expect[i] = True
    for i in range(15, 20):
        expect[i] = False
    for i in range(20, 25):
        expect[i] = True
    return np.array(pts, dtype=np.float64), (1, 30, 0.1, 100.0), expect


def occlusion_case2():  # :192-225
    pts = [(i * 0.1, 6.0, 0.0) for i in range(-15, 0)] + [(i * 0.1, 4.0, 0.0) for i in range(0, 15)]
    expect = {}
    for i in range(5):
        expect[i] = False
        expect[29 - i] = False
    for i in range(5, 10):
        expect[i] = True
    for i in range(10, 15):
        expect[i] = False
    for i in range(15, 25):
        expect[i] = True
    return np.array(pts, dtype=np.float64), (1, 30, 0.1, 100.0), expect


def _parallel(y_left, y_right):
    pts = [(i * 0.1, y_left, 0.0) for i in range(-15, 0)] + [(0, 0, 2.05)] + [(i * 0.1, y_right, 0.0) for i in range(1, 16)]
    expect = {}
    for i in range(5):
        expect[i] = False
        expect[30 - i] = False
    for i in range(5, 15):
        expect[i] = True
    for i in range(16, 26):
        expect[i] = True
    expect[15] = False
    return np.array(pts, dtype=np.float64), (1, 31, 0.1, 100.0), expect


def parallel_case1():  # :227-262
    return _parallel(2.0, 2.1)


def parallel_case2():  # :264-299
    return _parallel(2.1, 2.0)


CURVATURE_CASES = {"curvature_plane": curvature_plane, "curvature_corner": curvature_corner}
MASK_CASES = {"invalid_edges": invalid_edges, "invalid_ranges": invalid_ranges, "occlusion_case1": occlusion_case1,
              "occlusion_case2": occlusion_case2, "parallel_case1": parallel_case1, "parallel_case2": parallel_case2}

# extra parameter sets exercised against the oracle on synthetic scans: (N, S, maxE, maxP, eth, pth, occ, par)
PARAM_SWEEP = [
    (3, 6, 10, 50, 100.0, 1.0, 0.5, 1.0),   # defaults
    (5, 4, 4, 20, 50.0, 0.5, 0.3, 0.5),
    (2, 8, 0, 0, 100.0, 1.0, 0.5, 1.0),     # max = 0 still emits one per sector (reference off-by-one)
    (6, 1, 30, 200, 20.0, 5.0, 1.0, 2.0),   # a single sector = the whole ring
    (3, 7, 2, 9, 10.0, 2.0, 0.5, 1.0),      # ragged sector remainder
    (1, 6, 10, 50, 1.0, 0.01, 0.5, 1.0),
    (3, 6, 10, 50, 0.5, 1.0, 0.5, 1.0),     # thresholds overlap: a point can be an edge AND a planar candidate
    (3, 200, 1, 2, 100.0, 1.0, 0.5, 1.0),   # 400 walks per ring
    (3, 6, 3, 12, 100.0, 1.0, 0.5, 1.0),    # every walk hits its cap: exercises the truncate-and-reopen repair
]
