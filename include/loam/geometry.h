// loam/geometry.h — Pose3d, the boundary value type of registerFeatures (reference: loam/geometry.h:27-50,
// src/geometry.cpp:10-29) and the two distance templates of geometry-inl.h:21-33.  Header-only here: the pose
// algebra is a few flops on the host; the fits and residuals of the hot path run in the CUDA kernels.
#pragma once
#include <Eigen/Dense>

namespace loam {

struct Pose3d {
  Eigen::Quaterniond rotation;
  Eigen::Vector3d translation;

  Pose3d(Eigen::Quaterniond rot, Eigen::Vector3d trans) : rotation(rot), translation(trans) {}
  Pose3d() : rotation(Eigen::Quaterniond::Identity()), translation(Eigen::Vector3d::Zero()) {}
  static Pose3d Identity() { return Pose3d(); }

  Pose3d inverse() const {
    const Eigen::Quaterniond inv = rotation.inverse();
    return Pose3d(inv, inv * (-translation));
  }
  /// this ∘ other
  Pose3d compose(const Pose3d& other) const {
    return Pose3d(rotation * other.rotation, translation + rotation * other.translation);
  }
  Eigen::Vector3d act(const Eigen::Vector3d& p) const { return rotation * p + translation; }
  Eigen::Matrix4d matrix() const {
    Eigen::Matrix4d m = Eigen::Matrix4d::Identity();
    const Eigen::Vector3d ex = rotation * Eigen::Vector3d(1, 0, 0), ey = rotation * Eigen::Vector3d(0, 1, 0),
                          ez = rotation * Eigen::Vector3d(0, 0, 1);
    for (int i = 0; i < 3; i++) {
      m(i, 0) = ex(i);
      m(i, 1) = ey(i);
      m(i, 2) = ez(i);
      m(i, 3) = translation(i);
    }
    return m;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

namespace geometry_internal {

/// |(p - a) x (p - b)| / |a - b|
template <typename Vec3>
double pointToLineDistance(const Vec3& point, const Vec3& line_a, const Vec3& line_b) {
  const Vec3 da = point - line_a, db = point - line_b, ab = line_a - line_b;
  return da.cross(db).norm() / ab.norm();
}
/// |n . p - d|
template <typename Vec3>
double pointToPlaneDistance(const Vec3& point, const Vec3& normal, const double distance) {
  const double s = normal.dot(point) - distance;
  return s < 0 ? -s : s;
}

}  // namespace geometry_internal
}  // namespace loam
