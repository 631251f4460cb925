"""The pybind11 module `loam_python` (python/loam_b200_bindings.cpp): the Python surface of the reference's
python/loam_bindings.cpp over the C-ABI, with the contiguous-ndarray fast path."""
import os
import re

import numpy as np
import pytest

import helpers as H
from loam_b200 import synth

# every name the reference's python/loam_bindings.cpp registers (classes :24,41,51,69,80,104,119,129,136; functions
# :85-92,141-144) and the attributes it gives them
REFERENCE_SURFACE = {
    "LidarParams": ["scan_lines", "points_per_line", "min_range", "max_range"],
    "Pose3d": ["Identity", "inverse", "compose", "act", "rotation", "translation"],
    "Quaterniond": ["w", "x", "y", "z"],
    "FeatureExtractionParams": ["neighbor_points", "number_sectors", "max_edge_feats_per_sector",
                                "max_planar_feats_per_sector", "edge_feat_threshold", "planar_feat_threshold",
                                "occlusion_thresh", "parallel_thresh"],
    "LoamFeatures": ["edge_points", "planar_points"],
    "RegistrationParams": ["num_edge_neighbors", "max_edge_neighbor_dist", "min_line_fit_points",
                           "min_line_condition_number", "num_plane_neighbors", "max_plane_neighbor_dist",
                           "min_plane_fit_points", "max_avg_point_plane_dist", "max_iterations",
                           "rotation_convergence_thresh", "position_convergence_thresh", "min_associations"],
    "RegistrationIterationInfo": ["target_T_source_init", "edge_associations", "plane_associations", "estimate_update"],
    "RegistrationTerminationType": ["CONVERGED", "MAX_ITER", "INSUFFICIENT_ASSOCIATIONS"],
    "RegistrationDetail": ["iteration_info", "termination_type"],
}
REFERENCE_FUNCTIONS = ["extractFeatures", "computeCurvature", "computeValidPoints", "registerFeatures"]


@pytest.fixture(scope="module")
def mod():
    from loam_b200 import build
    build.build_python_module()
    from loam_b200 import loam_python
    return loam_python


def test_surface_matches_the_reference_bindings(mod):
    for cls, attrs in REFERENCE_SURFACE.items():
        assert hasattr(mod, cls), cls
        for a in attrs:
            assert hasattr(getattr(mod, cls), a), (cls, a)
    for f in REFERENCE_FUNCTIONS:
        assert callable(getattr(mod, f))
    assert mod.CONVERGED == mod.RegistrationTerminationType.CONVERGED  # export_values(), loam_bindings.cpp:133
    ref = "/root/reference/python/loam_bindings.cpp"
    if os.path.exists(ref):  # dev container only: nothing the reference binds is missing from the table above
        src = open(ref).read()
        for name in re.findall(r'm,\s*"(\w+)"\)', src) + re.findall(r'm\.def\("(\w+)"', src):
            assert hasattr(mod, name), name
        for name in re.findall(r'\.def_read(?:write|only)\("(\w+)"', src):
            assert any(name in attrs for attrs in REFERENCE_SURFACE.values()), name


def test_value_types_and_defaults(mod):
    p = mod.Pose3d(mod.Quaterniond(*np.r_[H.Q_SIMPLE[3], H.Q_SIMPLE[:3]]), [0.01, 0.03, -0.01])
    inv = p.inverse()
    back = inv.compose(p)
    assert abs(back.rotation.w() - 1) < 1e-12 and np.abs(back.translation).max() < 1e-12
    np.testing.assert_allclose(p.act([1.0, 2.0, 3.0]), H.transform(np.array([[1.0, 2.0, 3.0]]),
                                                                   np.r_[H.Q_SIMPLE, [0.01, 0.03, -0.01]])[0])
    fe, rp = mod.FeatureExtractionParams(), mod.RegistrationParams()
    assert (fe.neighbor_points, fe.number_sectors, fe.max_edge_feats_per_sector, fe.max_planar_feats_per_sector) == (3, 6, 10, 50)
    assert (rp.num_edge_neighbors, rp.max_plane_neighbor_dist, rp.max_iterations, rp.min_associations) == (5, 2.0, 10, 100)
    lp = mod.LidarParams(64, 1024, 1.0, 120.0)
    with pytest.raises(AttributeError):
        lp.scan_lines = 3  # const members, def_readonly (loam_bindings.cpp:27-30)
    pc = mod.PointCurvature(7, 2.5)  # registered here (the reference forgot it)
    assert pc.index == 7 and pc.curvature == 2.5
    with pytest.raises(RuntimeError, match=r"LOAM: provided lidar scan size \( 10\)"):
        mod.extractFeatures(np.zeros((10, 3)), mod.LidarParams(2, 4, 1.0, 10.0))  # the reference's message


@pytest.mark.gpu
def test_extract_fast_path_and_point_sequences(mod, oracle):
    from oracle.pyoracle import FeParams, LidarParams
    R, P = 16, 512
    scan = synth.make_scan(R, P, k=3)
    e, p = oracle.extract(scan[:, :3].astype(np.float64), LidarParams(R, P, 1.0, 120.0), FeParams.default())
    lp = mod.LidarParams(R, P, 1.0, 120.0)
    f = mod.extractFeatures(scan, lp)  # float32 {x,y,z,0} records, zero copy
    assert f.edge_points.dtype == np.float32 and np.array_equal(f.edge_points, scan[e]) and np.array_equal(f.planar_points, scan[p])
    xyz = scan[:, :3].astype(np.float64)
    f64 = mod.extractFeatures(xyz, lp, mod.FeatureExtractionParams())
    assert np.array_equal(f64.edge_points, xyz[e]) and np.array_equal(f64.planar_points, xyz[p])
    # what the reference accepts: any sequence of points (here a list of 1-D arrays, one per point)
    fl = mod.extractFeatures([row for row in xyz], lp)
    assert np.array_equal(fl.planar_points, xyz[p])
    ie, ip = mod.extractFeatureIndices(scan[:, :3], lp)  # strided float32 view (row pitch 16 bytes)
    assert np.array_equal(ie, e) and np.array_equal(ip, p)
    curv = mod.computeCurvature(xyz, lp)
    assert isinstance(curv[0], mod.PointCurvature) and curv[5].index == 5
    np.testing.assert_array_equal(mod.computeCurvatureArray(xyz, lp), oracle.curvature(xyz, LidarParams(R, P, 1.0, 120.0), FeParams.default()))
    valid = mod.computeValidPoints(xyz, lp)
    assert isinstance(valid, list) and np.array_equal(np.array(valid), mod.computeValidPointsArray(xyz, lp))


@pytest.mark.gpu
def test_register_features_mirrors_the_reference_call(mod, oracle):
    ed, pl = H.simple_scene()
    sTt = H.REG_SCENARIOS[0][1]
    src = mod.LoamFeatures(H.transform(ed, sTt), H.transform(pl, sTt))
    tgt = mod.LoamFeatures(ed, pl)
    detail = mod.RegistrationDetail()
    pose = mod.registerFeatures(src, tgt, mod.Pose3d.Identity(), mod.RegistrationParams(), detail)
    out = np.r_[pose.rotation.x(), pose.rotation.y(), pose.rotation.z(), pose.rotation.w(), pose.translation]
    ang, t = H.pose_error(sTt, out)
    assert ang < 1e-4 and np.all(np.abs(t) < 1e-4)  # the reference's own assertion (test_registration.cpp:84-87)
    po, do = oracle.register(H.transform(ed, sTt), H.transform(pl, sTt), ed, pl, want_detail=True)
    assert H.angular_distance(po[:4], out[:4]) < H.POSE_TOL_RAD and np.abs(po[4:] - out[4:]).max() < H.POSE_TOL_M
    assert detail.termination_type == mod.CONVERGED and len(detail.iteration_info) == do.n_iters
    for i, info in enumerate(detail.iteration_info):
        assert np.array_equal(np.array(info.plane_associations, dtype=np.uint32), do.plane_assoc[i])
        assert np.array_equal(np.array(info.edge_associations, dtype=np.uint32), do.edge_assoc[i])
    # default detail argument, and a second call appends (registration-inl.h:60)
    mod.registerFeatures(src, tgt, mod.Pose3d.Identity())
    mod.registerFeatures(src, tgt, mod.Pose3d.Identity(), mod.RegistrationParams(), detail)
    assert len(detail.iteration_info) == 2 * do.n_iters


@pytest.mark.gpu
def test_odometry_extension_matches_the_ctypes_path(mod, ctx):
    from loam_b200 import _capi
    R, P, n = 16, 512, 4
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    poses, term, iters, ne, npl = mod.odometry(scans, mod.LidarParams(R, P, 1.0, 120.0))
    ref = ctx.odometry_host(scans, _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params())
    for a, b in zip((poses, term, iters, ne, npl), ref):
        assert np.array_equal(a, b)
