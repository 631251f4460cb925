// loam/features.h — feature extraction entry points, same signatures as the reference (loam/features.h:108-111,
// 119-122, 166-169) but executed by the sm_100a kernels behind loamgpu_extract / loamgpu_curvature /
// loamgpu_valid_mask.  The template parameter kinds and order are kept because python/loam_bindings.cpp takes the
// address of explicit instantiations <loam::AtAccessor, py::array_t<double>, std::allocator>.
#pragma once
#include <memory>
#include <vector>

#include "loam/common.h"
#include "loam/detail/gpu.h"
#include "loam/geometry.h"

namespace loam {

/// Same fields and defaults as the reference (features.h:37-66).
struct FeatureExtractionParams {
  size_t neighbor_points{3};              ///< N: half-width of the curvature stencil / suppression radius + 1
  size_t number_sectors{6};               ///< sectors per ring; the remainder columns join the last sector
  size_t max_edge_feats_per_sector{10};   ///< the walk accepts max + 1 (reference behaviour, features-inl.h:155)
  size_t max_planar_feats_per_sector{50};
  double edge_feat_threshold{100.0};      ///< curvature above this => edge candidate
  double planar_feat_threshold{1.0};      ///< curvature below this => planar candidate
  double occlusion_thresh{0.5};
  double parallel_thresh{1.0};
};

template <typename PointType, template <typename> class Alloc = std::allocator>
struct LoamFeatures {
  std::vector<PointType, Alloc<PointType>> edge_points;
  std::vector<PointType, Alloc<PointType>> planar_points;
};

struct PointCurvature {
  size_t index;
  double curvature;
  PointCurvature(size_t i, double c) : index(i), curvature(c) {}
  PointCurvature() = default;
};
inline bool curvatureComparator(const PointCurvature& lhs, const PointCurvature& rhs) { return lhs.curvature < rhs.curvature; }

namespace gpu {
inline loamgpu_fe_params toC(const FeatureExtractionParams& p) {
  loamgpu_fe_params c;
  c.neighbor_points = p.neighbor_points;
  c.number_sectors = p.number_sectors;
  c.max_edge_feats_per_sector = p.max_edge_feats_per_sector;
  c.max_planar_feats_per_sector = p.max_planar_feats_per_sector;
  c.edge_feat_threshold = p.edge_feat_threshold;
  c.planar_feat_threshold = p.planar_feat_threshold;
  c.occlusion_thresh = p.occlusion_thresh;
  c.parallel_thresh = p.parallel_thresh;
  return c;
}

/// Indices (into input_scan) of the edge / planar features in the reference's output order.
template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
std::pair<std::vector<uint32_t>, std::vector<uint32_t>> extractFeatureIndices(
    const std::vector<PointType, Alloc<PointType>>& input_scan, const LidarParams& lidar_params,
    const FeatureExtractionParams& params = FeatureExtractionParams()) {
  validateLidarScan(input_scan, lidar_params);
  std::pair<std::vector<uint32_t>, std::vector<uint32_t>> out;
  if (input_scan.empty()) return out;
  loamgpu_ctx* ctx = ThreadContext::get();
  const CloudView view = makeCloudView<Accessor>(input_scan);
  const loamgpu_lidar_params lp = toC(lidar_params);
  const loamgpu_fe_params fp = toC(params);
  out.first.resize(input_scan.size());
  out.second.resize(input_scan.size());
  uint64_t ne = 0, np = 0;
  check(ctx, loamgpu_extract(ctx, view.data, view.dtype, view.stride, input_scan.size(), &lp, &fp, out.first.data(),
                             out.first.size(), &ne, out.second.data(), out.second.size(), &np));
  out.first.resize(ne);
  out.second.resize(np);
  return out;
}
}  // namespace gpu

template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
LoamFeatures<PointType, Alloc> extractFeatures(const std::vector<PointType, Alloc<PointType>>& input_scan,
                                               const LidarParams& lidar_params,
                                               const FeatureExtractionParams& params = FeatureExtractionParams()) {
  const auto idx = gpu::extractFeatureIndices<Accessor>(input_scan, lidar_params, params);
  LoamFeatures<PointType, Alloc> out;  // feature points are copies of the input elements, as in the reference
  out.edge_points.reserve(idx.first.size());
  out.planar_points.reserve(idx.second.size());
  for (uint32_t i : idx.first) out.edge_points.push_back(input_scan[i]);
  for (uint32_t i : idx.second) out.planar_points.push_back(input_scan[i]);
  return out;
}

/// EXTENSION (no counterpart in the reference, which leaves de-warping to its caller — README.md:63): extractFeatures
/// of a motion-compensated scan, the compensation fused into the kernel's ring staging (loamgpu_extract_dewarped).
/// Column c of every ring is moved into the frame of the sweep start by interp(Identity, start_T_end, c / P).  The
/// moved points are not PointType-representable in general, so the features come back widened (the type
/// featuresToEigen produces, features.h:188-198) and go to registerFeatures as they are.
template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
LoamFeatures<Eigen::Vector3d> extractFeaturesDewarped(const std::vector<PointType, Alloc<PointType>>& input_scan,
                                                      const LidarParams& lidar_params, const Pose3d& start_T_end,
                                                      const FeatureExtractionParams& params = FeatureExtractionParams()) {
  validateLidarScan(input_scan, lidar_params);
  LoamFeatures<Eigen::Vector3d> out;
  if (input_scan.empty()) return out;
  loamgpu_ctx* ctx = gpu::ThreadContext::get();
  const gpu::CloudView view = gpu::makeCloudView<Accessor>(input_scan);
  const loamgpu_lidar_params lp = gpu::toC(lidar_params);
  const loamgpu_fe_params fp = gpu::toC(params);
  const double motion[7] = {start_T_end.rotation.x(),    start_T_end.rotation.y(),    start_T_end.rotation.z(),
                            start_T_end.rotation.w(),    start_T_end.translation(0), start_T_end.translation(1),
                            start_T_end.translation(2)};
  std::vector<uint32_t> e(input_scan.size()), p(input_scan.size());
  std::vector<double> moved(3 * input_scan.size());
  uint64_t ne = 0, np = 0;
  gpu::check(ctx, loamgpu_extract_dewarped(ctx, view.data, view.dtype, view.stride, input_scan.size(), &lp, &fp, motion,
                                           e.data(), e.size(), &ne, p.data(), p.size(), &np, moved.data()));
  out.edge_points.reserve(ne);
  out.planar_points.reserve(np);
  for (uint64_t i = 0; i < ne; i++)
    out.edge_points.emplace_back(moved[3 * e[i]], moved[3 * e[i] + 1], moved[3 * e[i] + 2]);
  for (uint64_t i = 0; i < np; i++)
    out.planar_points.emplace_back(moved[3 * p[i]], moved[3 * p[i] + 1], moved[3 * p[i] + 2]);
  return out;
}

template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
std::vector<PointCurvature> computeCurvature(const std::vector<PointType, Alloc<PointType>>& input_scan,
                                             const LidarParams& lidar_params,
                                             const FeatureExtractionParams& params = FeatureExtractionParams()) {
  validateLidarScan(input_scan, lidar_params);
  std::vector<PointCurvature> out;
  if (input_scan.empty()) return out;
  loamgpu_ctx* ctx = gpu::ThreadContext::get();
  const gpu::CloudView view = gpu::makeCloudView<Accessor>(input_scan);
  const loamgpu_lidar_params lp = gpu::toC(lidar_params);
  const loamgpu_fe_params fp = gpu::toC(params);
  std::vector<double> c(input_scan.size());
  gpu::check(ctx, loamgpu_curvature(ctx, view.data, view.dtype, view.stride, input_scan.size(), &lp, &fp, c.data()));
  out.reserve(c.size());
  for (size_t i = 0; i < c.size(); i++) out.emplace_back(i, c[i]);
  return out;
}

template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
std::vector<bool> computeValidPoints(const std::vector<PointType, Alloc<PointType>>& input_scan,
                                     const LidarParams& lidar_params,
                                     const FeatureExtractionParams& params = FeatureExtractionParams()) {
  validateLidarScan(input_scan, lidar_params);
  std::vector<bool> out;
  if (input_scan.empty()) return out;
  loamgpu_ctx* ctx = gpu::ThreadContext::get();
  const gpu::CloudView view = gpu::makeCloudView<Accessor>(input_scan);
  const loamgpu_lidar_params lp = gpu::toC(lidar_params);
  const loamgpu_fe_params fp = gpu::toC(params);
  std::vector<uint8_t> m(input_scan.size());
  gpu::check(ctx, loamgpu_valid_mask(ctx, view.data, view.dtype, view.stride, input_scan.size(), &lp, &fp, m.data()));
  out.assign(m.begin(), m.end());
  return out;
}

namespace features_internal {
/// Widen a feature set to Eigen::Vector3d (reference: features.h:188-198).
template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
LoamFeatures<Eigen::Vector3d> featuresToEigen(const LoamFeatures<PointType, Alloc>& in) {
  LoamFeatures<Eigen::Vector3d> out;
  for (const PointType& pt : in.edge_points) out.edge_points.push_back(pointToEigen<Accessor>(pt));
  for (const PointType& pt : in.planar_points) out.planar_points.push_back(pointToEigen<Accessor>(pt));
  return out;
}
}  // namespace features_internal

}  // namespace loam
