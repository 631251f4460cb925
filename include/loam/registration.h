// loam/registration.h — registerFeatures with the reference's signature (loam/registration.h:128-131), executed by
// the sm_100a kernels behind loamgpu_register: GPU nearest-neighbour structure over the target sets, batched k-NN +
// line/plane fits, and the Levenberg-Marquardt solve (Ceres 2.2.0 semantics) resident on the device.
#pragma once
#include <memory>
#include <utility>
#include <vector>

#include "loam/common.h"
#include "loam/detail/gpu.h"
#include "loam/features.h"
#include "loam/geometry.h"

namespace loam {

/// Same fields and defaults as the reference (registration.h:40-75).
struct RegistrationParams {
  size_t num_edge_neighbors{5};
  double max_edge_neighbor_dist{1.0};
  size_t min_line_fit_points{3};
  double min_line_condition_number{10};
  size_t num_plane_neighbors{5};
  double max_plane_neighbor_dist{2.0};
  size_t min_plane_fit_points{4};
  double max_avg_point_plane_dist{0.1};
  size_t max_iterations{10};
  double rotation_convergence_thresh{1e-3};
  double position_convergence_thresh{1e-2};
  size_t min_associations{100};
};

/// Per-call diagnostics (registration.h:79-109): appended to, never cleared, exactly like the reference.
struct RegistrationDetail {
  enum TerminationType { CONVERGED, MAX_ITER, INSUFFICIENT_ASSOCIATIONS };
  struct IterationInfo {
    Pose3d target_T_source_init;
    std::vector<std::pair<size_t, size_t>> edge_associations;   ///< (source index, nearest target index)
    std::vector<std::pair<size_t, size_t>> plane_associations;
    Pose3d estimate_update;
    IterationInfo(const Pose3d init, const std::vector<std::pair<size_t, size_t>> edges,
                  const std::vector<std::pair<size_t, size_t>> planes, const Pose3d update)
        : target_T_source_init(init), edge_associations(edges), plane_associations(planes), estimate_update(update) {}
  };
  std::vector<IterationInfo> iteration_info;
  TerminationType termination_type;
};

namespace gpu {
inline loamgpu_reg_params toC(const RegistrationParams& p) {
  loamgpu_reg_params c;
  c.num_edge_neighbors = p.num_edge_neighbors;
  c.max_edge_neighbor_dist = p.max_edge_neighbor_dist;
  c.min_line_fit_points = p.min_line_fit_points;
  c.min_line_condition_number = p.min_line_condition_number;
  c.num_plane_neighbors = p.num_plane_neighbors;
  c.max_plane_neighbor_dist = p.max_plane_neighbor_dist;
  c.min_plane_fit_points = p.min_plane_fit_points;
  c.max_avg_point_plane_dist = p.max_avg_point_plane_dist;
  c.max_iterations = p.max_iterations;
  c.rotation_convergence_thresh = p.rotation_convergence_thresh;
  c.position_convergence_thresh = p.position_convergence_thresh;
  c.min_associations = p.min_associations;
  return c;
}
inline void poseTo7(const Pose3d& p, double* o) {
  o[0] = p.rotation.x();
  o[1] = p.rotation.y();
  o[2] = p.rotation.z();
  o[3] = p.rotation.w();
  o[4] = p.translation(0);
  o[5] = p.translation(1);
  o[6] = p.translation(2);
}
inline Pose3d poseFrom7(const double* v) {
  return Pose3d(Eigen::Quaterniond(v[3], v[0], v[1], v[2]), Eigen::Vector3d(v[4], v[5], v[6]));
}
template <template <typename> class Accessor, typename PointType, template <typename> class Alloc>
std::vector<double> widen(const std::vector<PointType, Alloc<PointType>>& pts) {
  std::vector<double> out(pts.size() * 3);
  for (size_t i = 0; i < pts.size(); i++) {
    out[3 * i + 0] = Accessor<PointType>::x(pts[i]);
    out[3 * i + 1] = Accessor<PointType>::y(pts[i]);
    out[3 * i + 2] = Accessor<PointType>::z(pts[i]);
  }
  return out;
}
}  // namespace gpu

namespace gpu {
/// Host buffers behind a loamgpu_detail and their conversion into RegistrationDetail (registration.h:79-109).
struct DetailBuffers {
  uint32_t cap, ne, np;
  std::vector<double> iter_est, iter_upd;
  std::vector<uint32_t> n_ea, n_pa, ea, pa;
  loamgpu_detail d;
  DetailBuffers(const RegistrationParams& params, size_t n_src_edge, size_t n_src_planar)
      : cap(static_cast<uint32_t>(params.max_iterations ? params.max_iterations : 1)),
        ne(static_cast<uint32_t>(n_src_edge ? n_src_edge : 1)),
        np(static_cast<uint32_t>(n_src_planar ? n_src_planar : 1)),
        iter_est(7 * cap), iter_upd(7 * cap), n_ea(cap), n_pa(cap), ea(static_cast<size_t>(cap) * ne * 2),
        pa(static_cast<size_t>(cap) * np * 2) {
    std::memset(&d, 0, sizeof d);
    d.max_iters_cap = cap;
    d.n_src_edge = ne;
    d.n_src_planar = np;
    d.iter_est = iter_est.data();
    d.iter_update = iter_upd.data();
    d.n_edge_assoc = n_ea.data();
    d.n_plane_assoc = n_pa.data();
    d.edge_assoc = ea.data();
    d.plane_assoc = pa.data();
  }
  /// appends one IterationInfo per recorded iteration and overwrites termination_type (registration-inl.h:60,76)
  void appendTo(RegistrationDetail& detail) const {
    for (uint32_t it = 0; it < d.n_iters && it < cap; it++) {
      std::vector<std::pair<size_t, size_t>> edges(n_ea[it]), planes(n_pa[it]);
      for (uint32_t k = 0; k < n_ea[it]; k++) {
        const uint32_t* r = ea.data() + (static_cast<size_t>(it) * ne + k) * 2;
        edges[k] = {r[0], r[1]};
      }
      for (uint32_t k = 0; k < n_pa[it]; k++) {
        const uint32_t* r = pa.data() + (static_cast<size_t>(it) * np + k) * 2;
        planes[k] = {r[0], r[1]};
      }
      detail.iteration_info.emplace_back(poseFrom7(iter_est.data() + 7 * it), edges, planes,
                                         poseFrom7(iter_upd.data() + 7 * it));
    }
    detail.termination_type = static_cast<RegistrationDetail::TerminationType>(d.termination);
  }
};
}  // namespace gpu

template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
Pose3d registerFeatures(const LoamFeatures<PointType, Alloc>& source, const LoamFeatures<PointType, Alloc>& target,
                        const Pose3d& target_T_source_init, const RegistrationParams& params = RegistrationParams(),
                        std::shared_ptr<RegistrationDetail> detail = nullptr) {
  loamgpu_ctx* ctx = gpu::ThreadContext::get();
  const std::vector<double> se = gpu::widen<Accessor>(source.edge_points), sp = gpu::widen<Accessor>(source.planar_points),
                            te = gpu::widen<Accessor>(target.edge_points), tp = gpu::widen<Accessor>(target.planar_points);
  const loamgpu_reg_params rp = gpu::toC(params);
  double init[7], out[7];
  gpu::poseTo7(target_T_source_init, init);
  if (!detail) {
    gpu::check(ctx, loamgpu_register(ctx, se.data(), se.size() / 3, sp.data(), sp.size() / 3, te.data(), te.size() / 3,
                                     tp.data(), tp.size() / 3, init, &rp, out, nullptr));
    return gpu::poseFrom7(out);
  }
  gpu::DetailBuffers buf(params, se.size() / 3, sp.size() / 3);
  gpu::check(ctx, loamgpu_register(ctx, se.data(), se.size() / 3, sp.data(), sp.size() / 3, te.data(), te.size() / 3,
                                   tp.data(), tp.size() / 3, init, &rp, out, &buf.d));
  buf.appendTo(*detail);
  return gpu::poseFrom7(out);
}

}  // namespace loam
