// loam/local_map.h — device-resident registration target (extension; not part of the reference API).
//
// The reference leaves "maintain a local map of points" to its caller (README.md:63): the accumulated map is passed
// as the `target` of registerFeatures (registration.h:128-131), which rebuilds two KD-trees over the whole map on
// every call (registration-inl.h:20-23).  A LocalMap keeps the map's feature points and their nearest-neighbour
// structures on the GPU between calls (C-ABI: loamgpu_map_*, include/loamgpu.h); registering onto it gives exactly
// the results of registerFeatures(source, <the map's points as LoamFeatures>, ...).
#pragma once
#include <memory>
#include <utility>

#include "loam/registration.h"

namespace loam {

class LocalMap {
 public:
  /// Map holding `features` (in the map frame).  Belongs to the calling thread's device (gpu::setDevice).
  template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
  explicit LocalMap(const LoamFeatures<PointType, Alloc>& features) {
    loamgpu_ctx* ctx = gpu::ThreadContext::get();
    const std::vector<double> e = gpu::widen<Accessor>(features.edge_points), p = gpu::widen<Accessor>(features.planar_points);
    gpu::check(ctx, loamgpu_map_create(ctx, e.data(), e.size() / 3, p.data(), p.size() / 3, &map_));
  }
  /// Empty map.
  LocalMap() {
    loamgpu_ctx* ctx = gpu::ThreadContext::get();
    gpu::check(ctx, loamgpu_map_create(ctx, nullptr, 0, nullptr, 0, &map_));
  }
  ~LocalMap() {
    if (map_) loamgpu_map_destroy(gpu::ThreadContext::get(), map_);
  }
  LocalMap(const LocalMap&) = delete;
  LocalMap& operator=(const LocalMap&) = delete;
  LocalMap(LocalMap&& o) noexcept : map_(o.map_) { o.map_ = nullptr; }
  LocalMap& operator=(LocalMap&& o) noexcept {
    std::swap(map_, o.map_);
    return *this;
  }

  /// Append `features` moved into the map frame by `map_T_features` (Pose3d::act on the device), then keep only the
  /// newest max_edge / max_planar points (0 = unbounded) and rebuild the NN structures.  Point indices reported in
  /// RegistrationDetail refer to the map's current order: insertion order, after eviction.
  template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
  void insert(const LoamFeatures<PointType, Alloc>& features, const Pose3d& map_T_features, size_t max_edge = 0,
              size_t max_planar = 0) {
    loamgpu_ctx* ctx = gpu::ThreadContext::get();
    const std::vector<double> e = gpu::widen<Accessor>(features.edge_points), p = gpu::widen<Accessor>(features.planar_points);
    double pose[7];
    gpu::poseTo7(map_T_features, pose);
    gpu::check(ctx, loamgpu_map_update(ctx, map_, e.data(), e.size() / 3, p.data(), p.size() / 3, pose, max_edge, max_planar));
  }

  size_t numEdgePoints() const { return sizes().first; }
  size_t numPlanarPoints() const { return sizes().second; }
  const loamgpu_map* handle() const { return map_; }

 private:
  std::pair<size_t, size_t> sizes() const {
    uint64_t e = 0, p = 0;
    loamgpu_map_size(map_, &e, &p);
    return {static_cast<size_t>(e), static_cast<size_t>(p)};
  }
  loamgpu_map* map_ = nullptr;
};

/// registerFeatures with a device-resident target: same semantics, parameters and RegistrationDetail as
/// registration.h:128-131 with the map's points as `target`.
template <template <typename> class Accessor = FieldAccessor, typename PointType, template <typename> class Alloc>
Pose3d registerFeatures(const LoamFeatures<PointType, Alloc>& source, const LocalMap& target,
                        const Pose3d& target_T_source_init, const RegistrationParams& params = RegistrationParams(),
                        std::shared_ptr<RegistrationDetail> detail = nullptr) {
  loamgpu_ctx* ctx = gpu::ThreadContext::get();
  const std::vector<double> se = gpu::widen<Accessor>(source.edge_points), sp = gpu::widen<Accessor>(source.planar_points);
  const loamgpu_reg_params rp = gpu::toC(params);
  double init[7], out[7];
  gpu::poseTo7(target_T_source_init, init);
  if (!detail) {
    gpu::check(ctx, loamgpu_register_to_map(ctx, target.handle(), se.data(), se.size() / 3, sp.data(), sp.size() / 3, init,
                                            &rp, out, nullptr));
    return gpu::poseFrom7(out);
  }
  gpu::DetailBuffers buf(params, se.size() / 3, sp.size() / 3);
  gpu::check(ctx, loamgpu_register_to_map(ctx, target.handle(), se.data(), se.size() / 3, sp.data(), sp.size() / 3, init,
                                          &rp, out, &buf.d));
  buf.appendTo(*detail);
  return gpu::poseFrom7(out);
}

}  // namespace loam
