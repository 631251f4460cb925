// loam/common.h — shared types of the B200-native LOAM hot path.
//
// Same public names and meaning as the reference's loam/common.h (LidarParams :29-41, the three accessor
// policies :55-78, pointRange :81-86, pointToEigen :89-93, validateLidarScan :104-113) so that user code and
// python/loam_bindings.cpp compile unchanged.  Nothing here touches the GPU.
#pragma once
#include <Eigen/Dense>
#include <cmath>
#include <cstddef>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace loam {

/// Intrinsics of the organised scan: scan_lines rings of points_per_line points, row-major; ranges outside
/// [min_range, max_range] are invalid for feature extraction.  Members are const, as in the reference.
struct LidarParams {
  const size_t scan_lines;
  const size_t points_per_line;
  const double min_range;
  const double max_range;
  LidarParams(size_t lines, size_t per_line, double rmin, double rmax)
      : scan_lines(lines), points_per_line(per_line), min_range(rmin), max_range(rmax) {}
};

// Accessor policies: static x/y/z taking the point BY VALUE and widening to double (reference semantics).
template <typename PointType>
struct FieldAccessor {  // pt.x / pt.y / pt.z  (PCL-style points)
  static double x(PointType pt) { return pt.x; }
  static double y(PointType pt) { return pt.y; }
  static double z(PointType pt) { return pt.z; }
};
template <typename PointType>
struct ParenAccessor {  // pt(0..2)  (Eigen vectors)
  static double x(PointType pt) { return pt(0); }
  static double y(PointType pt) { return pt(1); }
  static double z(PointType pt) { return pt(2); }
};
template <typename PointType>
struct AtAccessor {  // pt.at(0..2)  (std::vector, py::array_t)
  static double x(PointType pt) { return pt.at(0); }
  static double y(PointType pt) { return pt.at(1); }
  static double z(PointType pt) { return pt.at(2); }
};

template <template <typename> class Accessor = FieldAccessor, typename PointType>
double pointRange(const PointType& pt) {
  const double x = Accessor<PointType>::x(pt), y = Accessor<PointType>::y(pt), z = Accessor<PointType>::z(pt);
  return std::sqrt(x * x + y * y + z * z);
}

template <template <typename> class Accessor = FieldAccessor, typename PointType>
Eigen::Vector3d pointToEigen(const PointType& pt) {
  return Eigen::Vector3d(Accessor<PointType>::x(pt), Accessor<PointType>::y(pt), Accessor<PointType>::z(pt));
}

/// Throws std::runtime_error with the reference's message when the scan size does not match the intrinsics.
template <typename PointType, template <typename> class Alloc>
void validateLidarScan(const std::vector<PointType, Alloc<PointType>>& input_scan, const LidarParams& lidar_params) {
  if (input_scan.size() == lidar_params.scan_lines * lidar_params.points_per_line) return;
  std::stringstream msg;
  msg << "LOAM: provided lidar scan size ( " << input_scan.size() << ")  does not match provided lidar parameters ("
      << lidar_params.scan_lines << " x " << lidar_params.points_per_line << ")";
  throw std::runtime_error(msg.str());
}

}  // namespace loam
