// Shared device/host helpers for the loamgpu kernels (sm_100a).
// All arithmetic that decides an index (curvature, range, squared distance) is written with
// explicit round-to-nearest intrinsics so it can never be contracted into an FMA: the reference
// is built for baseline x86-64 (no FMA), and feature / correspondence indices must be bit-exact.
// The whole library is additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "loamgpu.h"

namespace loamgpu {

// ------------------------------------------------------------------ fp64, never fused
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// common.h:81-86 : sqrt((x*x + y*y) + z*z)
__device__ __forceinline__ double point_range(double x, double y, double z) {
  return __dsqrt_rn(dadd(dadd(dmul(x, x), dmul(y, y)), dmul(z, z)));
}
// nanoflann L2_Simple_Adaptor::evalMetric : diff = query - point ; ((d0^2) + d1^2) + d2^2
__device__ __forceinline__ double sqdist(double qx, double qy, double qz, double px, double py, double pz) {
  const double d0 = dsub(qx, px), d1 = dsub(qy, py), d2 = dsub(qz, pz);
  return dadd(dadd(dmul(d0, d0), dmul(d1, d1)), dmul(d2, d2));
}

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm(const V3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// Pose = (qx qy qz qw tx ty tz).  Eigen QuaternionBase::_transformVector, no normalisation
// (geometry.cpp:21): uv = 2 (u x v) ; v + w uv + u x uv
__device__ __forceinline__ V3 quat_rotate(const double* q, const V3& v) {
  const V3 u{q[0], q[1], q[2]};
  V3 uv = cross(u, v);
  uv.x += uv.x;
  uv.y += uv.y;
  uv.z += uv.z;
  const V3 uuv = cross(u, uv);
  return V3{v.x + q[3] * uv.x + uuv.x, v.y + q[3] * uv.y + uuv.y, v.z + q[3] * uv.z + uuv.z};
}
__device__ __forceinline__ V3 pose_act(const double* pose, const V3& p) {
  const V3 r = quat_rotate(pose, p);
  return V3{r.x + pose[4], r.y + pose[5], r.z + pose[6]};
}
// De-warp of one point of an organised scan (extension).  mq = quaternion of start_T_end on the
// hemisphere of Identity (dewarp_hemisphere), mt = its translation; column / P is the fraction of the sweep at which
// the point was measured.  Every operation is an IEEE + - * / sqrt in a fixed order: the CPU definition replays it bit
// for bit, and the extraction and pack kernels both call this one function.
__device__ __forceinline__ void dewarp_hemisphere(const double* motion, double* mq) {
  const bool flip = motion[3] < 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) mq[i] = flip ? -motion[i] : motion[i];
}
__device__ __forceinline__ V3 dewarp_point(const double* mq, const double* mt, uint32_t column, uint32_t P, const V3& p) {
  const double s = (double)column / (double)P;
  double pose[7];
  pose[0] = dmul(s, mq[0]);
  pose[1] = dmul(s, mq[1]);
  pose[2] = dmul(s, mq[2]);
  pose[3] = dadd(dsub(1.0, s), dmul(s, mq[3]));
  const double nn = __dsqrt_rn(
      dadd(dadd(dadd(dmul(pose[0], pose[0]), dmul(pose[1], pose[1])), dmul(pose[2], pose[2])), dmul(pose[3], pose[3])));
  pose[0] = __ddiv_rn(pose[0], nn);
  pose[1] = __ddiv_rn(pose[1], nn);
  pose[2] = __ddiv_rn(pose[2], nn);
  pose[3] = __ddiv_rn(pose[3], nn);
  pose[4] = dmul(s, mt[0]);
  pose[5] = dmul(s, mt[1]);
  pose[6] = dmul(s, mt[2]);
  return pose_act(pose, p);
}

// Eigen quaternion product a*b with (x,y,z,w) storage
__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x;
  o[1] = y;
  o[2] = z;
  o[3] = w;
}

// ------------------------------------------------------------------ TMA bulk copy + mbarrier (sm_90+/sm_100a PTX)
// One elected thread arms the barrier with the byte count and issues cp.async.bulk (SASS: UBLKCP); the data lands
// in shared memory through the async proxy and every waiting thread sees it once the barrier phase completes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------ NN structure: LBVH (binary radix tree)
// Points are sorted by 30-bit Morton code; a binary radix tree over the sorted (code, position) keys has n - 1
// internal nodes, root = node 0.  Node i covers a key range [first, last] (derived top-down during traversal) and
// splits it into [first, split] / [split + 1, last]; its children are nodes `split` and `split + 1` — adjacent
// records — unless the child range is a single point (flag bits).  Boxes are float32 rounded outward, so box tests
// run on the fp32 pipe and still give a true lower bound on the fp64 point distances.
#ifndef BVH_LEAF
#define BVH_LEAF 8
#endif
constexpr int kBvhLeaf = BVH_LEAF;  // subtrees of at most this many points are scanned, not descended
constexpr uint32_t kLeftLeaf = 0x80000000u, kRightLeaf = 0x40000000u, kSplitMask = 0x3FFFFFFFu;
struct BvhHdr {
  uint32_t n;  // points
  uint32_t pad[3];
};
struct __align__(32) BvhNode {
  float lo[3];
  uint32_t split;  // split position | kLeftLeaf | kRightLeaf
  float hi[3];
  uint32_t pad;    // the other end of the node's key range (node i covers [min(i, pad), max(i, pad)])
};
// Compact traversal records: what the batched k-NN kernel keeps in SHARED MEMORY while it walks one target set.  One
// record per internal node that covers more than kBvhLeaf points (~ n / 4.6 of them), holding the boxes of its two
// children on a 15-bit grid over the set's bounding box (rounded outward: a lower bound stays a lower bound) and what
// each child is: another record, or a run of <= kBvhLeaf points of the Morton-sorted copy to scan.  32 bytes per
// record, so a 64x1024 scan's two sets (2.4 k + 14.3 k points) take ~120 KB.  Written by the build kernel into the
// (then dead) sort scratch of the set.
constexpr uint32_t kRefLeaf = 0x80000000u;   // child ref: kRefLeaf | (count - 1) << 24 | first point  /  record index
constexpr uint32_t kNoRecs = 0xFFFFFFFFu;    // BvhQuant::n_rec of a set without compact records (multi-CTA build, overflow)
struct __align__(16) BvhRec {
  uint32_t lo[3];   // [axis]: lower cell of child L | lower cell of child R << 16
  uint32_t nhi[3];  // [axis]: minus the upper cell of child L (as int16) | minus the upper cell of child R << 16
  uint32_t ref[2];  // child L, child R
};
// Cells are 15-bit (0 .. kRecCellMax) so that every difference of two cell numbers fits an int16: the walk tests both
// children of a record with the 16x2 SIMD integer instructions of sm_100a (VIADDMNMX.S16x2[.RELU]: two per axis).
constexpr uint32_t kRecCellMax = 32767u;
constexpr double kRecCells = 32766.0;  // cells across the longest axis of the set's box (plus the margins)
// packs the boxes of two children (per axis: lower cell | upper cell << 16) into the record's six box words
__host__ __device__ inline void rec_pack_boxes(const uint32_t* boxL, const uint32_t* boxR, uint32_t* w6) {
  for (int d = 0; d < 3; d++) {
    w6[d] = (boxL[d] & 0xFFFFu) | (boxR[d] << 16);
    w6[3 + d] = ((0u - (boxL[d] >> 16)) & 0xFFFFu) | ((0u - (boxR[d] >> 16)) << 16);
  }
}
struct BvhQuant {   // grid of one set: cell index of coordinate x on axis d = (x - org[d]) * inv_cell
  double org[3];
  double inv_cell;
  double inv_cell2;  // inv_cell^2 (1 + 1e-12): squared metres -> squared cells, never rounded below the exact value
  uint32_t n_rec;    // records of this set (0: at most kBvhLeaf points, scanned directly)
  uint32_t pad;
};

// ------------------------------------------------------------------ per-pair ICF state
struct PairState {
  double est[7];       // target_T_source_est
  int32_t status;      // -1 active ; 0 CONVERGED ; 1 MAX_ITER ; 2 INSUFFICIENT_ASSOCIATIONS
  uint32_t iters;      // outer iterations recorded (iteration_info.size())
  uint32_t n_edge_assoc, n_plane_assoc;  // of the current outer iteration
};

// flattened parameter block handed to the registration kernels
struct RegP {
  int ke, kp;                // neighbours
  double re, rp;             // max neighbour distance (<=0: unbounded)
  int min_line, min_plane;   // fit guards
  double min_cond, max_avg;  // fit quality guards
  int max_iterations;
  double rot_thr, pos_thr;
  uint64_t min_assoc;
};

}  // namespace loamgpu
