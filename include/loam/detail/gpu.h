// loam/detail/gpu.h — thin host glue between the loam:: templates and the C-ABI (include/loamgpu.h).
// One loamgpu context per (thread, device): the reference's entry points are stateless and re-entrant, and so
// are these.  There is no CPU path: when libloamgpu cannot create a context the call throws.
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "loam/common.h"
#include "loamgpu.h"

namespace loam {
namespace gpu {

/// Device used by this thread's subsequent loam:: calls (default: $LOAMGPU_DEVICE or 0).
inline int& threadDevice() {
  static thread_local int dev = [] {
    const char* e = std::getenv("LOAMGPU_DEVICE");
    return e ? std::atoi(e) : 0;
  }();
  return dev;
}
inline void setDevice(int device) { threadDevice() = device; }

class ThreadContext {
 public:
  static loamgpu_ctx* get() {
    static thread_local ThreadContext tc;
    const int dev = threadDevice();
    if (tc.ctx_ && tc.dev_ != dev) tc.reset();
    if (!tc.ctx_) {
      loamgpu_ctx* c = nullptr;
      const int rc = loamgpu_create(dev, &c);
      if (rc != LOAMGPU_OK)
        throw std::runtime_error(std::string("LOAM (loamgpu): cannot create a CUDA context: ") + loamgpu_last_error(nullptr));
      tc.ctx_ = c;
      tc.dev_ = dev;
    }
    return tc.ctx_;
  }
  ~ThreadContext() { reset(); }

 private:
  void reset() {
    if (ctx_) loamgpu_destroy(ctx_);
    ctx_ = nullptr;
  }
  loamgpu_ctx* ctx_ = nullptr;
  int dev_ = 0;
};

/// Maps a loamgpu status to the exception the reference would have thrown (common.h:104-113: std::runtime_error).
inline void check(loamgpu_ctx* ctx, int rc) {
  if (rc == LOAMGPU_OK) return;
  const std::string msg = loamgpu_last_error(ctx);
  if (rc == LOAMGPU_ERR_SIZE_MISMATCH) throw std::runtime_error(msg);
  if (rc == LOAMGPU_ERR_INVALID) throw std::invalid_argument("LOAM (loamgpu): " + msg);
  throw std::runtime_error("LOAM (loamgpu): " + msg);
}

// ---- zero-copy detection: a point type whose x, y, z are float members at byte offsets 0 / 4 / 8 read through
// FieldAccessor (PCL's PointXYZ family) is handed to the C-ABI as LOAMGPU_F32 with stride sizeof(PointType).
template <typename P, typename = void>
struct HasFloatXYZ : std::false_type {};
template <typename P>
struct HasFloatXYZ<P, std::void_t<decltype(std::declval<P>().x), decltype(std::declval<P>().y), decltype(std::declval<P>().z)>>
    : std::integral_constant<bool, std::is_same<decltype(P::x), float>::value && std::is_same<decltype(P::y), float>::value &&
                                       std::is_same<decltype(P::z), float>::value && std::is_standard_layout<P>::value &&
                                       std::is_trivially_copyable<P>::value> {};

template <template <typename> class Accessor, typename PointType>
constexpr bool usesFieldAccessor() {
  return std::is_same<Accessor<PointType>, FieldAccessor<PointType>>::value;
}

/// A scan as the C-ABI wants it: either a borrowed view of the caller's float records or a packed copy of doubles.
struct CloudView {
  const void* data = nullptr;
  int dtype = LOAMGPU_F64;
  size_t stride = 24;
  std::vector<double> packed;
};

template <template <typename> class Accessor, typename PointType, template <typename> class Alloc>
CloudView makeCloudView(const std::vector<PointType, Alloc<PointType>>& pts) {
  CloudView v;
  if constexpr (HasFloatXYZ<PointType>::value && usesFieldAccessor<Accessor, PointType>()) {
    if (!pts.empty()) {
      const char* base = reinterpret_cast<const char*>(&pts[0]);
      if (reinterpret_cast<const char*>(&pts[0].x) == base && reinterpret_cast<const char*>(&pts[0].y) == base + 4 &&
          reinterpret_cast<const char*>(&pts[0].z) == base + 8 && sizeof(PointType) % 4 == 0) {
        v.data = base;
        v.dtype = LOAMGPU_F32;
        v.stride = sizeof(PointType);
        return v;
      }
    }
  }
  v.packed.resize(pts.size() * 3);
  for (size_t i = 0; i < pts.size(); i++) {  // the accessors widen to double exactly as the reference does
    v.packed[3 * i + 0] = Accessor<PointType>::x(pts[i]);
    v.packed[3 * i + 1] = Accessor<PointType>::y(pts[i]);
    v.packed[3 * i + 2] = Accessor<PointType>::z(pts[i]);
  }
  v.data = v.packed.data();
  return v;
}

inline loamgpu_lidar_params toC(const LidarParams& p) {
  loamgpu_lidar_params c;
  c.scan_lines = p.scan_lines;
  c.points_per_line = p.points_per_line;
  c.min_range = p.min_range;
  c.max_range = p.max_range;
  return c;
}

}  // namespace gpu
}  // namespace loam
