"""De-warp extension (the reference leaves it to its caller, README.md:63 — so there is no reference
behaviour to pin: orc_dewarp is the definition, checked here against an independent numpy statement of the same
model, its exact fixed points, and the physics it is meant to undo)."""
import numpy as np
import pytest

import helpers as H
from loam_b200 import synth
from oracle.pyoracle import FeParams, LidarParams

MOTIONS = {
    "yaw_forward": np.r_[H.axis_angle(0.02, (0, 0, 1)), 0.09, -0.03, 0.0],
    "general": np.r_[H.axis_angle(0.05, (0.3, -0.2, 1.0)), 0.2, 0.1, -0.05],
    "far_hemisphere": np.r_[-H.axis_angle(0.03, (0, 1, 0.5)), 0.1, 0.0, -0.02],  # -q: same rotation, w < 0
}


def numpy_dewarp(xyz, P, m):
    sp = synth.sweep_poses(P, m)
    c = np.arange(len(xyz)) % P
    u, w = sp[c, :3], sp[c, 3:4]
    uv = 2.0 * np.cross(u, xyz)
    return xyz + w * uv + np.cross(u, uv) + sp[c, 4:7]


def test_identity_motion_is_bitwise_noop(oracle):
    xyz = synth.make_scan(8, 257, k=3)[:, :3].astype(np.float64)
    out = oracle.dewarp(xyz, 257, [0, 0, 0, 1, 0, 0, 0])
    assert np.array_equal(out, xyz)
    out = oracle.dewarp(xyz, 257, [0, 0, 0, -1, 0, 0, 0])  # the same rotation on the other hemisphere
    assert np.array_equal(out, xyz)


@pytest.mark.parametrize("name", list(MOTIONS))
def test_matches_numpy_statement(oracle, name):
    R, P = 8, 300
    xyz = synth.make_scan(R, P, k=4)[:, :3].astype(np.float64)
    out = oracle.dewarp(xyz, P, MOTIONS[name])
    np.testing.assert_allclose(out, numpy_dewarp(xyz, P, MOTIONS[name]), rtol=0, atol=1e-12)
    assert np.array_equal(out[0::P], xyz[0::P]), "column 0 is measured at the sweep start: untouched"
    # the sweep approaches start_T_end as the column approaches P
    end = H.transform(xyz[P - 1::P], MOTIONS[name])
    assert np.abs(out[P - 1::P] - end).max() < np.abs(xyz).max() * 0.06 / P * 4 + 1e-3


def test_pure_translation_is_linear_in_the_column(oracle):
    P = 64
    xyz = np.random.RandomState(0).uniform(-20, 20, size=(2 * P, 3))
    t = np.array([0.5, -0.25, 0.125])
    out = oracle.dewarp(xyz, P, np.r_[0, 0, 0, 1, t])
    s = (np.arange(2 * P) % P) / P
    np.testing.assert_allclose(out, xyz + s[:, None] * t, rtol=0, atol=1e-14)


def test_undoes_the_sweep_of_a_moving_sensor(oracle):
    """Noise-free scan recorded while the sensor moves: raw points are smeared off the scene surfaces; de-warped
    with the true motion they fall back onto them (as seen from the pose at the sweep start)."""
    R, P, k = 16, 900, 7
    m = np.r_[H.axis_angle(0.06, (0, 0, 1)), 0.35, -0.12, 0.0]
    scan = synth.make_warped_scan(R, P, k, m)
    xyz = scan[:, :3].astype(np.float64)
    world_T_start = synth.pose_of_scan(k)
    raw = synth.distance_to_scene(H.transform(xyz, world_T_start))
    fixed = synth.distance_to_scene(H.transform(oracle.dewarp(xyz, P, m), world_T_start))
    assert fixed.max() < 1e-4, fixed.max()  # float32 storage of ~20 m ranges
    assert np.percentile(raw, 90) > 0.05  # the warp is large compared with that


def test_extraction_differs_and_is_defined_on_the_moved_points(oracle):
    R, P = 16, 900
    m = np.r_[H.axis_angle(0.06, (0, 0, 1)), 0.35, -0.12, 0.0]
    xyz = synth.make_warped_scan(R, P, 7, m, sigma=0.01)[:, :3].astype(np.float64)
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    e0, p0 = oracle.extract(xyz, lp, fe)
    e1, p1 = oracle.extract(oracle.dewarp(xyz, P, m), lp, fe)
    assert len(e1) > 0 and len(p1) > 0
    assert not (np.array_equal(e0, e1) and np.array_equal(p0, p1))


def test_golden_fixture(oracle):
    """tests/golden/dewarp_golden.npz (made by make_dewarp_golden.py): the definition is pinned bit for bit, and the
    restated extraction agrees with what the real reference code extracted from the moved points."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dewarp_golden.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 3
    for n in names:
        R, P = (int(v) for v in g[n + "/shape"])
        moved = oracle.dewarp(g[n + "/scan"].astype(np.float64), P, g[n + "/motion"])
        assert np.array_equal(moved, g[n + "/moved"]), n
        e, p = oracle.extract(moved, LidarParams(R, P, 1.0, 120.0), FeParams.default())
        assert np.array_equal(e, g[n + "/edge"]) and np.array_equal(p, g[n + "/planar"]), n
