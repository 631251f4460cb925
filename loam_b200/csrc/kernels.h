// Host-visible launch interface of the loamgpu kernels (internal to libloamgpu.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace loamgpu {

#ifndef EXTRACT_THREADS
#define EXTRACT_THREADS 256
#endif
constexpr int kExtractThreads = EXTRACT_THREADS;
constexpr int kAssocThreads = 128;
#ifndef LM_THREADS
#define LM_THREADS 256
#endif
#ifndef LM_MINBLOCKS
#define LM_MINBLOCKS 2
#endif
constexpr int kLmThreads = LM_THREADS;
constexpr int kLmMinBlocks = LM_MINBLOCKS;
#ifndef BUILD_THREADS
#define BUILD_THREADS 1024
#endif
#ifndef BUILD_MINBLOCKS
#define BUILD_MINBLOCKS 1
#endif
constexpr int kBuildThreads = BUILD_THREADS;  // one CTA per feature set (single-CTA NN build)
constexpr int kKnnSmall = 5;     // neighbour counts up to this use the 5-slot register top-k
constexpr int kKnnRegMax = 8;    // ... up to this the 8-slot one
constexpr int kKnnMax = 32;      // hard limit on num_*_neighbors
constexpr uint32_t kLmNcCap = 2048;  // planes per pair the LM kernel's moment sums may leave uncovered (else it streams all)

struct ExtractArgs {
  const unsigned char* pts;     // first scan of this launch
  uint64_t scan_stride_bytes;   // bytes between scans
  uint32_t stride;              // bytes between points
  int dtype;
  int use_bulk;                 // ring can be staged with one cp.async.bulk
  uint32_t R, P, N, S;
  uint32_t maxE, maxP;          // max_*_feats_per_sector clipped to P
  double edge_thr, planar_thr, occ, par, min_range, max_range;
  uint32_t capE_ring, capP_ring;
  uint32_t* ring_edge;          // [scan][R][capE_ring] scan-level point index
  uint32_t* ring_planar;        // [scan][R][capP_ring]
  uint32_t* ring_counts;        // [scan][R][2]
  double* curv_out;             // optional [scan][R*P]  (computeCurvature)
  uint8_t* mask_out;            // optional [scan][R*P]  (computeValidPoints)
  // de-warp fused into the ring staging (extension): every point is moved into the frame of the sweep start by
  // interp(Identity, motion, column / P) before anything else looks at it
  int dewarp;
  double motion[7];             // start_T_end: qx qy qz qw tx ty tz (one motion for every scan of the launch)
  const double* motions;        // or: device [scan][7], one motion per scan of the launch (overrides `motion`)
  double* dewarp_out;           // optional [scan][R*P][3] de-warped points
};

struct PackArgs {
  const unsigned char* pts;
  uint64_t scan_stride_bytes;
  uint32_t stride;
  int dtype;
  uint32_t R;
  uint32_t capE_ring, capP_ring, capE_scan, capP_scan;
  const uint32_t* ring_edge;
  const uint32_t* ring_planar;
  const uint32_t* ring_counts;
  uint64_t scan0;               // global index of the first scan of this launch
  uint32_t n_slots;             // feature slots form a ring buffer: slot = scan % n_slots
  uint32_t* edge_idx;           // [slot][capE_scan]
  uint32_t* planar_idx;         // [slot][capP_scan]
  double4* edge_pts;            // [slot][capE_scan]  (x,y,z,0)
  double4* planar_pts;          // [slot][capP_scan]
  uint32_t* feat_counts;        // [slot][2]
  uint32_t* n_edge_out;         // optional [global scan] feature counts for the caller
  uint32_t* n_planar_out;
  // de-warped extraction: the gathered feature points are the moved points (same arithmetic as the extract kernel)
  int dewarp;
  uint32_t P;
  double motion[7];
  const double* motions;        // device [scan][7] or null (then `motion`)
};

// stage_bytes: staging record size in shared memory (0 = the default of the dtype: float4 / 3 doubles)
size_t extract_smem_bytes(int dtype, uint32_t P, uint32_t S, uint32_t stage_bytes = 0);
cudaError_t launch_extract(const ExtractArgs& a, uint32_t n_scans, cudaStream_t st);
cudaError_t launch_pack(const PackArgs& a, uint32_t n_scans, cudaStream_t st);

// One NN structure ("set") = header + node records + Morton-sorted points (+ build scratch).
struct BvhSetArrays {
  BvhHdr* hdr;        // [n_sets]
  BvhNode* nodes;     // [n_sets][pt_cap]     internal nodes 0 .. n-2, root = 0
  double4* sorted;    // [n_sets][pt_cap]     (x,y,z, bits(original index)), Morton order
  uint2* keys;        // [n_sets][2][pt_cap]  radix-sort ping-pong scratch: (morton, original index); after the build
                      //                      the first pt_cap / 2 BvhRec slots of a set's scratch hold its compact records
  int* aux;           // [n_sets][pt_cap]     build scratch: per-node readiness, then compact record numbers
  uint32_t pt_cap;
  BvhQuant* quant;    // [n_sets]             grid + record count of the compact records (null: none, e.g. map targets)
};
// compact records of set `set` (they overlay its sort scratch)
__host__ __device__ inline const BvhRec* bvh_recs(const BvhSetArrays& g, uint32_t set) {
  return reinterpret_cast<const BvhRec*>(g.keys + (size_t)set * 2 * g.pt_cap);
}
__host__ __device__ inline uint32_t bvh_rec_cap(const BvhSetArrays& g) { return g.pt_cap / 2; }

// Build NN structures for `n_sets` point sets.  Set s reads points
// pts[(slot0 + s) % n_slots][0 .. counts[((slot0+s) % n_slots)*2 + kind]).
struct BvhBuildArgs {
  const double4* pts;      // [slot][pt_stride]
  const uint32_t* counts;  // [slot][2]
  uint32_t pt_stride;
  int kind;                // 0 edge, 1 planar
  uint64_t slot0;
  uint32_t n_slots;
  BvhSetArrays g;
  int smem_tree;           // set by launch_bvh_build: sorted codes + readiness flags fit in shared memory
};
cudaError_t launch_bvh_build(const BvhBuildArgs& a, uint32_t n_sets, cudaStream_t st);
// the edge and the planar sets of the same slots in one launch when both fit the shared-memory build (else two launches)
cudaError_t launch_bvh_build2(const BvhBuildArgs& edge, const BvhBuildArgs& planar, uint32_t n_sets, cudaStream_t st,
                              uint64_t* launches);

// Multi-CTA build of ONE large set (bvh_big.cu): same layout as launch_bvh_build produces for set 0 of `g`
// (g.hdr, g.nodes, g.sorted; g.keys is the sort scratch).  `scratch` holds bvh_big_scratch_bytes(n) bytes.
size_t bvh_big_scratch_bytes(uint32_t n);
cudaError_t launch_bvh_build_big(const double4* pts, uint32_t n, const BvhSetArrays& g, void* scratch, cudaStream_t st,
                                 uint64_t* launches);

struct AssocArgs {
  // source / target feature slots of pair p: src = (pair0 + p + 1) % n_slots, tgt = (pair0 + p) % n_slots
  const double4* edge_pts;
  const double4* planar_pts;
  const uint32_t* feat_counts;
  uint32_t capE_scan, capP_scan;
  uint64_t pair0;
  uint32_t n_slots;
  int src_offset;          // 1 for sequence odometry; explicit-pair calls use slots 1 (src) / 0 (tgt)
  BvhSetArrays ge, gp;     // NN structures: set `pair` = target of the pair, set `pair + src_offset` = its source
  PairState* state;        // [pair]
  double4* rec_p;          // [pair][capE+capP]  transformed point, w = 0 invalid / 1 edge / 2 plane (after the fit kernel)
  double4* rec_a;          // [pair][capE+capP]  edge: line point a ; plane: normal, w = d
  double4* rec_b;          // [pair][capE]       edge: line point b
  // Per-source-feature arrays (rec_*, nn_*) are in QUERY order: edge feature at position m of the source set's Morton
  // order -> slot m, planar -> slot capE + m.  The k-NN kernel stores the feature's original index in rec_p.w; the fit
  // kernel replaces it by the residual kind.  (Threads walk the queries in that order, so every write coalesces.)
  uint32_t* nn_idx;        // [pair][capE+capP][nn_stride] neighbour indices of the current outer iteration
  uint32_t* nn_cnt;        // [pair][capE+capP]            neighbours inside the radius
  uint32_t nn_stride;      // max(num_edge_neighbors, num_plane_neighbors)
  int morton_queries;      // 1: walk the source set in its Morton order (default), 0: original order
  uint32_t n_pairs;        // pairs of this launch
  const uint32_t* active;  // [0] = number of pairs still iterating, [1..] their indices; null = all, in order
  uint32_t* leftover;      // [0] = count, [1..] pairs the shared-memory k-NN kernel left to the general one
  int32_t* nearest;        // optional [outer_iter][pair][capE+capP] nearest target index or -1 (detail)
  // external target (device-resident local map): when ext_target != 0 every pair registers onto set 0 of te / tp,
  // whose points in original order are te_pts / tp_pts; ge / gp then hold only the source sets (set = pair)
  int ext_target;
  BvhSetArrays te, tp;
  const double4* te_pts;
  const double4* tp_pts;
  RegP rp;
};
cudaError_t launch_assoc_knn(const AssocArgs& a, uint32_t n_pairs, int outer_iter, cudaStream_t st);
cudaError_t launch_assoc_fit(const AssocArgs& a, uint32_t n_pairs, int outer_iter, cudaStream_t st);

struct LmArgs {
  PairState* state;
  const double4* rec_p;
  const double4* rec_a;
  const double4* rec_b;
  const uint32_t* feat_counts;
  uint32_t capE_scan, capP_scan;
  uint64_t pair0;
  uint32_t n_slots;
  int src_offset;
  int outer_iter;
  uint32_t n_pairs;
  const uint32_t* active;  // see AssocArgs
  uint32_t cluster;        // CTAs (thread-block cluster size, 1..8) sharing one pair
  // moment path (register.cu: plane_moments_pass): per pair the planes NOT covered by the moment sums; null = stream all
  double4* nc_p;           // [pair][nc_cap]
  double4* nc_a;           // [pair][nc_cap]
  uint32_t nc_cap;
  RegP rp;
  // optional detail (single-pair API): per outer iteration rows
  double* d_iter_est;      // [cap][7]
  double* d_iter_update;   // [cap][7]
  uint32_t* d_assoc_n;     // [cap][2]
  uint32_t* d_lm_iters;    // [cap]
  double* d_lm_cost;       // [cap][2]
};
cudaError_t launch_lm(const LmArgs& a, uint32_t n_pairs, cudaStream_t st);
// test hook: the 28 sums (H[21] g[6] cost) of ONE evaluation of pair 0's records at the iterate x (+ 2 status words)
cudaError_t launch_lm_debug_eval(const LmArgs& a, const double* x_dev, int mode, double* out_dev, cudaStream_t st);
cudaError_t launch_compact_active(const PairState* st, uint32_t n_pairs, uint32_t* active, cudaStream_t s);

struct KnnArgs {
  const double* queries;  // [n][3]
  uint64_t n_queries;
  BvhSetArrays g;         // set 0
  int k;
  double max_dist;
  uint32_t* idx_out;      // [n][k]
  uint32_t* count_out;    // [n]
};
cudaError_t launch_knn(const KnnArgs& a, cudaStream_t st);

// Widen up to four packed n x 3 double clouds (device) into double4 feature slots (featuresToEigen, features.h:188-198)
struct WidenArgs {
  const double* src[4];
  double4* dst[4];
  uint32_t n[4];
};
cudaError_t launch_widen(const WidenArgs& a, cudaStream_t s);
// p <- pose * p for n double4 points in place (map insertion: scan frame -> map frame, Pose3d::act, geometry.cpp:21)
cudaError_t launch_transform_points(double4* pts, uint32_t n, const double* pose_dev, cudaStream_t s);
// init poses: null = identity for every pair; init_stride 0 = the same pose for every pair, 7 = one pose per pair
cudaError_t launch_init_pairs(PairState* st, uint32_t n_pairs, const double* init_pose_or_null, uint32_t init_stride,
                              cudaStream_t s);
// Copy per-pair results to flat output arrays (device pointers; any may be null) and finalise MAX_ITER.
cudaError_t launch_finish_pairs(const PairState* st, uint32_t n_pairs, double* poses, int32_t* term, uint32_t* iters,
                                cudaStream_t s);

}  // namespace loamgpu
