"""CPU: host-side mirror of the reference's Python module — value types and parameter structs (no kernels)."""
import numpy as np
import pytest

import loam_b200 as loam


def test_names_of_the_reference_module_exist():
    # python/loam_bindings.cpp:24-144
    for n in ["LidarParams", "Pose3d", "Quaterniond", "FeatureExtractionParams", "LoamFeatures", "extractFeatures",
              "computeCurvature", "computeValidPoints", "RegistrationParams", "RegistrationIterationInfo",
              "RegistrationTerminationType", "RegistrationDetail", "registerFeatures", "CONVERGED", "MAX_ITER",
              "INSUFFICIENT_ASSOCIATIONS"]:
        assert hasattr(loam, n), n


def test_lidar_params_readonly():
    lp = loam.LidarParams(64, 1024, 1.0, 120.0)
    assert (lp.scan_lines, lp.points_per_line, lp.min_range, lp.max_range) == (64, 1024, 1.0, 120.0)
    with pytest.raises(AttributeError):
        lp.scan_lines = 3


def test_param_defaults_and_fields():
    fe = loam.FeatureExtractionParams()
    assert fe.neighbor_points == 3 and fe.number_sectors == 6 and fe.max_planar_feats_per_sector == 50
    fe.neighbor_points = 5
    assert fe._to_c().neighbor_points == 5
    rp = loam.RegistrationParams(max_iterations=1)
    assert rp.max_iterations == 1 and rp.num_plane_neighbors == 5 and rp.min_associations == 100
    with pytest.raises(TypeError):
        loam.RegistrationParams(bogus=1)


def test_pose3d_compose_inverse_matrix_gtsam_constants():
    # tests/test_geometry.cpp:31-79 of the reference
    q1 = loam.Quaterniond(0.7473257838894183, 0.38405116269438366, -0.17015746936361906, -0.5148352287741462)
    q2 = loam.Quaterniond(0.8378767472656409, -0.040374739652255895, -0.40934599608063865, 0.3588429911288663)
    p1, p2 = loam.Pose3d(q1, [-0.4, 3.0, -8.9]), loam.Pose3d(q2, [4, -5, 1])
    c = p1.compose(p2)
    np.testing.assert_allclose(c.translation, [-2.59584795, -1.87410099, -12.56352171], rtol=1e-8)
    np.testing.assert_allclose([c.rotation.w(), c.rotation.x(), c.rotation.y(), c.rotation.z()],
                               [0.7567645973045605, 0.019808900212688513, -0.5655135339985058, -0.32727571648894294],
                               rtol=1e-8)
    inv = p1.inverse()
    np.testing.assert_allclose(inv.translation, [1.60941772, 6.39896027, 6.69575105], rtol=1e-8)
    p3 = loam.Pose3d(loam.Quaterniond(0.9693342323515085, 0.018781217536151106, 0.15609411554196426,
                                      0.18887307630401792), [1.0, -5.0, 2.0])
    expected = np.array([[0.87992318, -0.360299, 0.30970927, 1.], [0.37202555, 0.92794845, 0.0225534, -5.],
                         [-0.29552021, 0.09537451, 0.95056379, 2.], [0, 0, 0, 1.]])
    np.testing.assert_allclose(p3.matrix(), expected, atol=1e-6)
    np.testing.assert_allclose(p1.compose(p1.inverse()).translation, 0, atol=1e-12)


def test_pose3d_copy_semantics():  # test_geometry.cpp:17-29
    pa = loam.Pose3d()
    pb = loam.Pose3d(loam.Quaterniond.from_coeffs(pa.rotation.coeffs()), pa.translation.copy())
    pa.translation[0] = 1
    assert pb.translation[0] == 0 and pa.translation[0] == 1
    np.testing.assert_allclose(loam.Pose3d.Identity().act([1, 2, 3]), [1, 2, 3])
