// Registration kernels (replaces loam/registration-inl.h, src/registration.cpp, src/kdtree.cpp,
// src/geometry.cpp:42-73 of the reference, plus the Ceres 2.2.0 trust-region solve it delegates to).
//
//   K3  bvh_build_kernel    implicit LBVH over one target feature set (bvh.cuh; replaces the nanoflann
//                           KD-tree build, registration-inl.h:20-23)
//   K4  knn_bvh()           exact k-NN + strict radius filter (bvh.cuh; kdtree.cpp:10-28)
//   K5  fit_line/fit_plane  per-query PCA line / column-pivoted-QR plane (geometry.cpp:42-73)
//   K4  assoc_knn_kernel    transform + k-NN for every source feature of every active pair, queries in the
//                           source set's Morton order (registration.cpp:34-37,75-78)
//   K5  assoc_fit_kernel    line / plane fit + guards -> residual records (registration.cpp:39-57,80-98)
//   K6/K7 lm_kernel         one CTA per pair: residuals + analytic SE(3) Jacobians + Huber corrector,
//                           6x6 J^T J / J^T r / cost reduced warp-shuffle -> shared memory, and the
//                           Levenberg-Marquardt controller (Ceres TrustRegionMinimizer semantics) entirely
//                           on the device; then the ICF update / convergence test (registration-inl.h:59-73)
//
// Tie-break (documented, deterministic): equal squared distances resolve by ascending target index
// (nanoflann resolves them by tree-traversal order, i.e. unpinned in the reference).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "bvh.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace loamgpu {

namespace {

// ============================================================================ fits (K5)

// Cyclic Jacobi on a symmetric 3x3; returns the eigenvector of the largest eigenvalue
// (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>, geometry.cpp:49-51).
__device__ __forceinline__ V3 principal_axis(double A[3][3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double tr = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-20 * tr) break;
#pragma unroll
    for (int p = 0; p < 2; p++) {
#pragma unroll
      for (int q = p + 1; q < 3; q++) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        A[p][p] = A[p][p] - t * apq;
        A[q][q] = A[q][q] + t * apq;
        A[p][q] = 0.0;
        A[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = c * arp - s * arq;
        A[p][r] = A[r][p];
        A[r][q] = s * arp + c * arq;
        A[q][r] = A[r][q];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  // column of the largest eigenvalue, selected without run-time indexing (keeps A and V in registers)
  double best = A[0][0];
  V3 axis{V[0][0], V[1][0], V[2][0]};
  if (A[1][1] > best) {
    best = A[1][1];
    axis = V3{V[0][1], V[1][1], V[2][1]};
  }
  if (A[2][2] > best) axis = V3{V[0][2], V[1][2], V[2][2]};
  return axis;
}

// geometry.cpp:42-59.  The condition number is never produced (the reference computes and discards it,
// leaving DBL_MAX), so the min_line_condition_number guard can only fire for a threshold above DBL_MAX.
template <int KMAX>
__device__ __noinline__ void fit_line(const double (&P)[KMAX][3], int K, V3& la, V3& lb) {
  double c[3] = {0, 0, 0};
  for (int k = 0; k < K; k++) {
    c[0] += P[k][0];
    c[1] += P[k][1];
    c[2] += P[k][2];
  }
  c[0] /= (double)K;
  c[1] /= (double)K;
  c[2] /= (double)K;
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < K; k++) {
    const double v[3] = {P[k][0] - c[0], P[k][1] - c[1], P[k][2] - c[2]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += v[i] * v[j];
  }
  const V3 dir = principal_axis(S);
  la = V3{c[0] + 0.1 * dir.x, c[1] + 0.1 * dir.y, c[2] + 0.1 * dir.z};
  lb = V3{c[0] - 0.1 * dir.x, c[1] - 0.1 * dir.y, c[2] - 0.1 * dir.z};
}

// geometry.cpp:62-73 : column-pivoted Householder QR least squares of  points * abc = 1  (Eigen
// ColPivHouseholderQR semantics incl. its rank threshold), then normal = abc/|abc|, d = 1/|abc| and the
// SIGNED mean distance.
// __noinline__ on purpose: inlined into assoc_fit_kernel, nvcc 12.9 / sm_100a produced a wrong signed mean distance
// for near-collinear neighbour sets (normal and d were right, the final loop over P was not); as a real call the
// results match the CPU oracle bit for bit (tests/test_gpu_odometry.py::test_sequence_matches_oracle pins this).
template <int KMAX>
__device__ __noinline__ double fit_plane(const double (&P)[KMAX][3], int K, V3& nrm, double& dist) {
  double A[KMAX][3], c[KMAX];
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < K; k++) {
    A[k][0] = P[k][0];
    A[k][1] = P[k][1];
    A[k][2] = P[k][2];
    c[k] = 1.0;
  }
  const int size = K < 3 ? K : 3;
  double maxnorm = 0.0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    double s = 0;
    for (int k = 0; k < K; k++) s += A[k][j] * A[k][j];
    s = sqrt(s);
    if (s > maxnorm) maxnorm = s;
  }
  const double th = maxnorm * 2.220446049250313e-16;
  const double threshold_helper = th * th / (double)K;  // Eigen: abs2(maxnorm * eps) / rows
  int nonzero = size;
  for (int k = 0; k < size; k++) {
    int big = k;
    double bigsq = -1.0;
    for (int j = k; j < 3; j++) {
      double s = 0;
      for (int r = k; r < K; r++) s += A[r][j] * A[r][j];
      if (s > bigsq) {
        bigsq = s;
        big = j;
      }
    }
    if (nonzero == size && bigsq < threshold_helper * (double)(K - k)) nonzero = k;
    if (big != k) {
      for (int r = 0; r < K; r++) {
        const double t = A[r][k];
        A[r][k] = A[r][big];
        A[r][big] = t;
      }
      const int t = perm[k];
      perm[k] = perm[big];
      perm[big] = t;
    }
    double tail = 0;
    for (int r = k + 1; r < K; r++) tail += A[r][k] * A[r][k];
    const double c0 = A[k][k];
    double tau, beta;
    if (tail <= 2.2250738585072014e-308) {
      tau = 0;
      beta = c0;
      for (int r = k + 1; r < K; r++) A[r][k] = 0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      for (int r = k + 1; r < K; r++) A[r][k] = A[r][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    for (int j = k + 1; j < 3; j++) {
      double w = A[k][j];
      for (int r = k + 1; r < K; r++) w += A[r][k] * A[r][j];
      w *= tau;
      A[k][j] -= w;
      for (int r = k + 1; r < K; r++) A[r][j] -= w * A[r][k];
    }
    {
      double w = c[k];
      for (int r = k + 1; r < K; r++) w += A[r][k] * c[r];
      w *= tau;
      c[k] -= w;
      for (int r = k + 1; r < K; r++) c[r] -= w * A[r][k];
    }
  }
  double y[3] = {0, 0, 0};
  for (int i = nonzero - 1; i >= 0; i--) {
    double s = c[i];
    for (int j = i + 1; j < nonzero; j++) s -= A[i][j] * y[j];
    y[i] = s / A[i][i];
  }
  double abc[3] = {0, 0, 0};
  for (int i = 0; i < nonzero; i++) abc[perm[i]] = y[i];
  const double nn = sqrt(abc[0] * abc[0] + abc[1] * abc[1] + abc[2] * abc[2]);
  nrm = V3{abc[0] / nn, abc[1] / nn, abc[2] / nn};
  dist = 1.0 / nn;
  double sum = 0;
  for (int k = 0; k < K; k++) sum += (P[k][0] * nrm.x + P[k][1] * nrm.y + P[k][2] * nrm.z) - dist;
  return sum / (double)K;
}

// ---------------------------------------------------------------------------- register-resident fits (K <= 8)
// Same arithmetic, in the same order, as fit_line / fit_plane above, but every loop is unrolled over the compile-time
// capacity KMAX with row / rank predicates, so the neighbour matrix, the QR workspace and the permutation live in
// registers (the generic versions index them dynamically, which puts ~280 B per thread in local memory).
template <int KMAX>
__device__ __forceinline__ void fit_line_reg(const double (&P)[KMAX][3], int K, V3& la, V3& lb) {
  double c[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    if (k < K) {
      c[0] += P[k][0];
      c[1] += P[k][1];
      c[2] += P[k][2];
    }
  }
  c[0] /= (double)K;
  c[1] /= (double)K;
  c[2] /= (double)K;
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    if (k < K) {
      const double v[3] = {P[k][0] - c[0], P[k][1] - c[1], P[k][2] - c[2]};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] += v[i] * v[j];
    }
  }
  const V3 dir = principal_axis(S);
  la = V3{c[0] + 0.1 * dir.x, c[1] + 0.1 * dir.y, c[2] + 0.1 * dir.z};
  lb = V3{c[0] - 0.1 * dir.x, c[1] - 0.1 * dir.y, c[2] - 0.1 * dir.z};
}

template <int KMAX>
__device__ __forceinline__ double fit_plane_reg(const double (&P)[KMAX][3], int K, V3& nrm, double& dist) {
  double A[KMAX][3], c[KMAX];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    const bool in = k < K;
    A[k][0] = in ? P[k][0] : 0.0;
    A[k][1] = in ? P[k][1] : 0.0;
    A[k][2] = in ? P[k][2] : 0.0;
    c[k] = in ? 1.0 : 0.0;
  }
  const int size = K < 3 ? K : 3;
  double maxnorm = 0.0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    double sq = 0;
#pragma unroll
    for (int k = 0; k < KMAX; k++)
      if (k < K) sq += A[k][j] * A[k][j];
    sq = sqrt(sq);
    if (sq > maxnorm) maxnorm = sq;
  }
  const double th = maxnorm * 2.220446049250313e-16;
  const double threshold_helper = th * th / (double)K;  // Eigen: abs2(maxnorm * eps) / rows
  int nonzero = size;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k < size) {
      int big = k;
      double bigsq = -1.0;
#pragma unroll
      for (int j = k; j < 3; j++) {
        double sq = 0;
#pragma unroll
        for (int r = k; r < KMAX; r++)
          if (r < K) sq += A[r][j] * A[r][j];
        if (sq > bigsq) {
          bigsq = sq;
          big = j;
        }
      }
      if (nonzero == size && bigsq < threshold_helper * (double)(K - k)) nonzero = k;
#pragma unroll
      for (int j = k + 1; j < 3; j++) {  // swap columns k and big (big is a run-time value: one static swap per case)
        if (big == j) {
#pragma unroll
          for (int r = 0; r < KMAX; r++) {
            const double t = A[r][k];
            A[r][k] = A[r][j];
            A[r][j] = t;
          }
          const int t = perm[k];
          perm[k] = perm[j];
          perm[j] = t;
        }
      }
      double tail = 0;
#pragma unroll
      for (int r = k + 1; r < KMAX; r++)
        if (r < K) tail += A[r][k] * A[r][k];
      const double c0 = A[k][k];
      double tau, beta;
      if (tail <= 2.2250738585072014e-308) {
        tau = 0;
        beta = c0;
#pragma unroll
        for (int r = k + 1; r < KMAX; r++) A[r][k] = 0;
      } else {
        beta = sqrt(c0 * c0 + tail);
        if (c0 >= 0) beta = -beta;
#pragma unroll
        for (int r = k + 1; r < KMAX; r++)
          if (r < K) A[r][k] = A[r][k] / (c0 - beta);
        tau = (beta - c0) / beta;
      }
      A[k][k] = beta;
#pragma unroll
      for (int j = k + 1; j < 3; j++) {
        double w = A[k][j];
#pragma unroll
        for (int r = k + 1; r < KMAX; r++)
          if (r < K) w += A[r][k] * A[r][j];
        w *= tau;
        A[k][j] -= w;
#pragma unroll
        for (int r = k + 1; r < KMAX; r++)
          if (r < K) A[r][j] -= w * A[r][k];
      }
      {
        double w = c[k];
#pragma unroll
        for (int r = k + 1; r < KMAX; r++)
          if (r < K) w += A[r][k] * c[r];
        w *= tau;
        c[k] -= w;
#pragma unroll
        for (int r = k + 1; r < KMAX; r++)
          if (r < K) c[r] -= w * A[r][k];
      }
    }
  }
  double y[3] = {0, 0, 0};
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    if (i < nonzero) {
      double sacc = c[i];
#pragma unroll
      for (int j = i + 1; j < 3; j++)
        if (j < nonzero) sacc -= A[i][j] * y[j];
      y[i] = sacc / A[i][i];
    }
  }
  double abc[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (i < nonzero) {
#pragma unroll
      for (int t = 0; t < 3; t++)
        if (perm[i] == t) abc[t] = y[i];
    }
  }
  const double nn = sqrt(abc[0] * abc[0] + abc[1] * abc[1] + abc[2] * abc[2]);
  nrm = V3{abc[0] / nn, abc[1] / nn, abc[2] / nn};
  dist = 1.0 / nn;
  double sum = 0;
#pragma unroll
  for (int k = 0; k < KMAX; k++)
    if (k < K) sum += (P[k][0] * nrm.x + P[k][1] * nrm.y + P[k][2] * nrm.z) - dist;
  return sum / (double)K;
}

// ============================================================================ association (K4 + K5)

// K4: transform every source feature of every active pair by the current estimate (registration.cpp:34,75) and find
// its k nearest target features.  Threads walk the source set in ITS OWN Morton order (the source scan's NN
// structure holds a Morton-sorted copy), so the lanes of a warp carry neighbouring queries: they traverse the same
// nodes (loads coalesce, caches hit) and need similar numbers of leaves.  Results land at the original feature index.
// 8 resident CTAs of 128 threads = 64 registers per thread: measured 17.2 -> 13.2 ms/step against the unconstrained
// 72-80 registers (the traversal is latency-bound; the extra warps hide node-load latency better than the few
// spilled words cost)
#ifndef KNN_MINBLOCKS
#define KNN_MINBLOCKS 8
#endif
#ifndef KNN_THREADS
#define KNN_THREADS 128
#endif
constexpr int kKnnThreads = KNN_THREADS;
// Pairs still iterating: entry 0 of `active` is their count, the indices follow (null = all n_pairs, in order).
// Rows of the grid stride over that list, so late outer iterations (few or no active pairs) can be launched with a
// small grid instead of tens of thousands of CTAs that only discover they have nothing to do.
__device__ __forceinline__ uint32_t active_count(const uint32_t* active, uint32_t n_pairs) {
  return active ? active[0] : n_pairs;
}
__device__ __forceinline__ uint32_t active_pair(const uint32_t* active, uint32_t i) { return active ? active[1 + i] : i; }

// Shared by both association kernels: one source feature (Morton position `m` of its kind in the source set) of one
// pair.  `search` runs the k-NN of the transformed point in the target's structure.
template <int K, typename Search>
__device__ __forceinline__ void assoc_knn_query(const AssocArgs& a, int outer_iter, uint32_t pair, const double* est,
                                                uint32_t src_slot, bool is_plane, uint32_t m, Search&& search) {
  const BvhSetArrays& gs = is_plane ? a.gp : a.ge;
  const uint32_t src_set = pair + (uint32_t)a.src_offset;
  double4 sp;
  if (a.morton_queries) {
    sp = gs.sorted[(size_t)src_set * gs.pt_cap + m];
  } else {  // A/B switch: source features in their original order
    sp = is_plane ? a.planar_pts[(size_t)src_slot * a.capP_scan + m] : a.edge_pts[(size_t)src_slot * a.capE_scan + m];
    sp.w = __longlong_as_double((long long)m);
  }
  const V3 q = pose_act(est, V3{sp.x, sp.y, sp.z});
  const int k = is_plane ? a.rp.kp : a.rp.ke;
  const double md = is_plane ? a.rp.rp : a.rp.re;
  const size_t cap_src = (size_t)a.capE_scan + a.capP_scan;
  const size_t rec = (size_t)pair * cap_src + (is_plane ? a.capE_scan + m : m);  // query order (kernels.h: AssocArgs)
  uint32_t* out = a.nn_idx + rec * (size_t)a.nn_stride;
  // From the second outer iteration on, the previous neighbours give a bound before the search starts: k target
  // points lie within max_j |q - p_j|, so the k-th nearest distance cannot exceed it (the estimate moved by
  // millimetres, the bound is nearly tight and most of the tree is pruned on the way down).
  double d2_hint = CUDART_INF;
  if (outer_iter > 0 && (int)a.nn_cnt[rec] == k) {
    const double4* tp = a.ext_target ? (is_plane ? a.tp_pts : a.te_pts)
                        : is_plane   ? a.planar_pts + (size_t)((a.pair0 + pair) % a.n_slots) * a.capP_scan
                                     : a.edge_pts + (size_t)((a.pair0 + pair) % a.n_slots) * a.capE_scan;
    double worst = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) {
      if (j < k) {
        const double4 t = tp[out[j]];
        worst = fmax(worst, sqdist(q.x, q.y, q.z, t.x, t.y, t.z));
      }
    }
    d2_hint = worst;
  }
  TopK<K> tk;
  search(q, k, md, d2_hint, tk);
  const int cnt = radius_count(tk, k, md);
  a.rec_p[rec] = make_double4(q.x, q.y, q.z, sp.w);  // w: original index of the source feature (for the fit kernel)
  a.nn_cnt[rec] = (uint32_t)cnt;
#pragma unroll
  for (int j = 0; j < K; j++)
    if (j < k) out[j] = tk.id[j];
}

// General kernel: one query per thread, walks the target's node records in global memory (any target: this context's
// sets, device-resident maps, neighbour counts up to 32).  Also takes the pairs the shared-memory kernel left over.
template <int K>
__device__ __forceinline__ void assoc_knn_pair(const AssocArgs& a, int outer_iter, uint32_t pair) {
  const PairState* ps = a.state + pair;
  if (ps->status != -1) return;
  const uint32_t src_slot = (uint32_t)((a.pair0 + pair + a.src_offset) % a.n_slots);
  const uint32_t nE = a.feat_counts[src_slot * 2], nP = a.feat_counts[src_slot * 2 + 1];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nE + nP) return;
  const bool is_plane = i >= nE;
  double est[7];
#pragma unroll
  for (int j = 0; j < 7; j++) est[j] = ps->est[j];
  const BvhSetArrays& gt = a.ext_target ? (is_plane ? a.tp : a.te) : (is_plane ? a.gp : a.ge);  // the target's structure
  const uint32_t tset = a.ext_target ? 0u : pair;
  const BvhHdr g = gt.hdr[tset];
  assoc_knn_query<K>(a, outer_iter, pair, est, src_slot, is_plane, is_plane ? i - nE : i,
                     [&](const V3& q, int k, double md, double d2_hint, TopK<K>& tk) {
                       knn_bvh<K>(g, gt.nodes + (size_t)tset * gt.pt_cap, gt.sorted + (size_t)tset * gt.pt_cap, q.x, q.y,
                                  q.z, k, md, tk, d2_hint);
                     });
}

template <int K>
__global__ void __launch_bounds__(kKnnThreads, KNN_MINBLOCKS) assoc_knn_kernel(AssocArgs a, int outer_iter,
                                                                              const uint32_t* list) {
  // `list` = a.active (pairs still iterating) or a.leftover (pairs the shared-memory kernel skipped); null = all
  const uint32_t n_act = active_count(list, a.n_pairs);
  for (uint32_t i = blockIdx.y; i < n_act; i += gridDim.y) assoc_knn_pair<K>(a, outer_iter, active_pair(list, i));
}

// Batched kernel (sequence odometry, explicit batches): ONE persistent 1024-thread CTA per SM takes (pair, slice)
// items; for each it bulk-copies the compact records of the pair's target sets (edge, planar: ~120 KB for a 64x1024
// scan) into shared memory once, then its 32 warps pull 32-query chunks of the pair's source features (Morton order)
// from a shared counter and walk the records there (bvh.cuh: knn_compact).  Pairs whose records do not fit, or whose
// sets have none, go to a.leftover and are done by the general kernel right after.
#ifndef KNN_CTA_THREADS
#define KNN_CTA_THREADS 1024
#endif
constexpr int kKnnCtaThreads = KNN_CTA_THREADS;
template <int K>
__global__ void __launch_bounds__(kKnnCtaThreads, 1) assoc_knn_smem_kernel(AssocArgs a, int outer_iter, uint32_t n_slices,
                                                                          uint32_t max_recs) {
  extern __shared__ __align__(128) unsigned char knn_smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_next;
  __shared__ BvhQuant s_q[2];
  __shared__ double s_est[7];
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  // dynamic shared memory: [traversal-stack ring: KNN_SMEM_STACK x 1024 x 8 bytes][compact records]
  const uint32_t s_stack = smem_u32(knn_smem) + tid * 8u;
  unsigned char* const rec_smem = knn_smem + (size_t)KNN_SMEM_STACK * kKnnStackPitch;
  const uint32_t s_base = smem_u32(rec_smem);
  const uint32_t n_act = active_count(a.active, a.n_pairs);
  const uint32_t n_items = n_act * n_slices;
  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t pair = active_pair(a.active, item / n_slices), slice = item % n_slices;
    const PairState* ps = a.state + pair;
    if (ps->status != -1) continue;  // (CTA-uniform, like every exit below)
    const uint32_t src_slot = (uint32_t)((a.pair0 + pair + a.src_offset) % a.n_slots);
    const uint32_t nE = a.feat_counts[src_slot * 2], nP = a.feat_counts[src_slot * 2 + 1];
    const uint32_t recE = a.ge.quant[pair].n_rec, recP = a.gp.quant[pair].n_rec;
    if (recE == kNoRecs || recP == kNoRecs || recE + recP > max_recs) {
      if (slice == 0 && tid == 0) a.leftover[1 + atomicAdd(a.leftover, 1u)] = pair;
      continue;
    }
    // chunks of 32 queries: edge chunks first, then planar chunks; this item's share [c_lo, c_hi)
    const uint32_t chE = (nE + 31) / 32, chAll = chE + (nP + 31) / 32;
    const uint32_t per = (chAll + n_slices - 1) / n_slices;
    const uint32_t c_lo = min(slice * per, chAll), c_hi = min(c_lo + per, chAll);
    if (c_lo >= c_hi) continue;
    const bool need_e = c_lo < chE && recE > 0, need_p = c_hi > chE && recP > 0;
    __syncthreads();  // everyone is done with the previous item's records, grids and counter
    if (tid == 0) {
      s_next = c_lo;
      const uint32_t bytes = (need_e ? recE : 0u) * (uint32_t)sizeof(BvhRec) + (need_p ? recP : 0u) * (uint32_t)sizeof(BvhRec);
      if (bytes) {
        mbar_expect_tx(&s_bar, bytes);
        const BvhRec* src[2] = {bvh_recs(a.ge, pair), bvh_recs(a.gp, pair)};
        const uint32_t cnt[2] = {need_e ? recE : 0u, need_p ? recP : 0u}, off[2] = {0u, recE};
        for (int kind = 0; kind < 2; kind++)  // (pieces of at most 32 KB per bulk copy)
          for (uint32_t r = 0; r < cnt[kind]; r += 1024) {
            const uint32_t nr = min(1024u, cnt[kind] - r);
            bulk_g2s(rec_smem + (size_t)(off[kind] + r) * sizeof(BvhRec), src[kind] + r, nr * (uint32_t)sizeof(BvhRec),
                     &s_bar);
          }
      }
    }
    if (tid < 2) s_q[tid] = (tid ? a.gp : a.ge).quant[pair];
    if (tid >= 32 && tid < 39) s_est[tid - 32] = ps->est[tid - 32];
    __syncthreads();
    if (need_e || need_p) {
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    const uint32_t nTe = a.ge.hdr[pair].n, nTp = a.gp.hdr[pair].n;  // points of the target sets
    for (;;) {
      uint32_t c = 0;
      if (lane == 0) c = atomicAdd(&s_next, 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= c_hi) break;
      const bool is_plane = c >= chE;
      const uint32_t m = (is_plane ? c - chE : c) * 32 + lane;
      if (m >= (is_plane ? nP : nE)) continue;
      const BvhSetArrays& gt = is_plane ? a.gp : a.ge;
      const uint32_t s_recs = s_base + (is_plane ? recE : 0u) * (uint32_t)sizeof(BvhRec);
      const uint32_t n_pts = is_plane ? nTp : nTe;
      const BvhQuant* Q = s_q + (is_plane ? 1 : 0);
      const double4* sorted = gt.sorted + (size_t)pair * gt.pt_cap;
      assoc_knn_query<K>(a, outer_iter, pair, s_est, src_slot, is_plane, m,
                         [&](const V3& q, int k, double md, double d2_hint, TopK<K>& tk) {
                           knn_compact<K>(s_recs, s_stack, n_pts, Q, sorted, q.x, q.y, q.z, k, md, tk, d2_hint);
                         });
    }
  }
}

// K5: line / plane fit + guards for every source feature (associateEdges/associatePlanes, registration.cpp:39-57,
// 80-98), in source-index order.  Writes the residual records the LM kernel consumes.
template <int KMAX>
__device__ __forceinline__ void assoc_fit_pair(const AssocArgs& a, int outer_iter, uint32_t pair) {
  PairState* ps = a.state + pair;
  if (ps->status != -1) return;
  const uint32_t tgt_slot = (uint32_t)((a.pair0 + pair) % a.n_slots);
  const uint32_t src_slot = (uint32_t)((a.pair0 + pair + a.src_offset) % a.n_slots);
  const uint32_t nE = a.feat_counts[src_slot * 2], nP = a.feat_counts[src_slot * 2 + 1];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= nE + nP) return;  // whole block beyond the source features
  const bool active = i < nE + nP;
  const bool is_plane = active && i >= nE;
  bool ok = false;
  if (active) {
    const uint32_t pos = is_plane ? i - nE : i;  // query order (kernels.h: AssocArgs)
    const size_t cap_src = (size_t)a.capE_scan + a.capP_scan;
    const size_t rec = (size_t)pair * cap_src + (is_plane ? a.capE_scan + pos : pos);
    // the k-NN kernel left the feature's original index in rec_p.w; this kernel replaces it by the residual kind
    double* rec_w = reinterpret_cast<double*>(a.rec_p + rec) + 3;
    const uint32_t li = (uint32_t)__double_as_longlong(*rec_w);
    const int m = (int)a.nn_cnt[rec];
    const uint32_t* nn_g = a.nn_idx + rec * (size_t)a.nn_stride;
    // the neighbour numbers are read before the count is known (slots beyond it hold stale numbers and are not used):
    // count -> numbers -> points was three dependent round trips per feature, 45 % of the kernel's stall samples
    uint32_t nn[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; j++) nn[j] = j < (int)a.nn_stride ? nn_g[j] : 0u;
    const int need = is_plane ? a.rp.min_plane : a.rp.min_line;
    if (m >= need && m > 0) {
      const double4* tp = a.ext_target ? (is_plane ? a.tp_pts : a.te_pts)
                          : is_plane   ? a.planar_pts + (size_t)tgt_slot * a.capP_scan
                                       : a.edge_pts + (size_t)tgt_slot * a.capE_scan;
      double N[KMAX][3];
#pragma unroll
      for (int j = 0; j < KMAX; j++) {
        if (j < m) {
          const double4 t = tp[nn[j]];
          N[j][0] = t.x;
          N[j][1] = t.y;
          N[j][2] = t.z;
        }
      }
      if (!is_plane) {
        V3 la, lb;
        if (KMAX <= kKnnRegMax)
          fit_line_reg<KMAX>(N, m, la, lb);
        else
          fit_line<KMAX>(N, m, la, lb);
        // registration.cpp:49: condition_number is always DBL_MAX in the reference (geometry.cpp:55-56)
        if (!(1.7976931348623157e308 < a.rp.min_cond)) {
          ok = true;
          a.rec_a[rec] = make_double4(la.x, la.y, la.z, 0.0);
          a.rec_b[(size_t)pair * a.capE_scan + pos] = make_double4(lb.x, lb.y, lb.z, 0.0);
        }
      } else {
        V3 nrm;
        double dist;
        const double avg = KMAX <= kKnnRegMax ? fit_plane_reg<KMAX>(N, m, nrm, dist) : fit_plane<KMAX>(N, m, nrm, dist);
        if (!(avg > a.rp.max_avg)) {  // registration.cpp:90
          ok = true;
          a.rec_a[rec] = make_double4(nrm.x, nrm.y, nrm.z, dist);
        }
      }
    }
    *rec_w = ok ? (is_plane ? 2.0 : 1.0) : 0.0;  // w: 0 invalid / 1 edge / 2 plane
    if (a.nearest) a.nearest[((size_t)outer_iter * a.n_pairs + pair) * cap_src + (is_plane ? a.capE_scan + li : li)] =
        ok ? (int32_t)nn[0] : -1;
  }
  const unsigned be = __ballot_sync(0xffffffffu, ok && !is_plane);
  const unsigned bp = __ballot_sync(0xffffffffu, ok && is_plane);
  if ((threadIdx.x & 31) == 0) {
    if (be) atomicAdd(&ps->n_edge_assoc, (uint32_t)__popc(be));
    if (bp) atomicAdd(&ps->n_plane_assoc, (uint32_t)__popc(bp));
  }
}

#ifndef FIT_MINBLOCKS
#define FIT_MINBLOCKS 7
#endif
template <int KMAX>
__global__ void __launch_bounds__(kAssocThreads, FIT_MINBLOCKS) assoc_fit_kernel(AssocArgs a, int outer_iter) {
  const uint32_t n_act = active_count(a.active, a.n_pairs);
  for (uint32_t i = blockIdx.y; i < n_act; i += gridDim.y) assoc_fit_pair<KMAX>(a, outer_iter, active_pair(a.active, i));
}

// ============================================================================ LM solve + ICF update (K6/K7)

struct Eval {
  double H[21];  // upper triangle of J^T J (tangent 6x6, loss-corrected, unscaled)
  double g[6];   // J^T r
  double cost;   // 1/2 sum rho(s)
};

// ceres::QuaternionManifold (w-first) applied literally to Eigen's (x,y,z,w) memory — SURVEY §8a-notes
__device__ __forceinline__ void manifold_plus(const double* x, const double* delta, double* out) {
  const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd == 0.0) {
    out[0] = x[0];
    out[1] = x[1];
    out[2] = x[2];
    out[3] = x[3];
  } else {
    const double sbd = sin(nd) / nd;
    const double q0 = cos(nd), q1 = sbd * delta[0], q2 = sbd * delta[1], q3 = sbd * delta[2];
    out[0] = q0 * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
    out[1] = q0 * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
    out[2] = q0 * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
    out[3] = q0 * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
  }
  out[4] = x[4] + delta[3];
  out[5] = x[5] + delta[4];
  out[6] = x[6] + delta[5];
}

// Per-evaluation linear maps.  For a fixed iterate x = (u, w, t) the transformed point and its tangent Jacobian are
// linear in the source point p (Eigen's un-normalised rotation v + w 2(u x v) + u x 2(u x v), registration-inl.h:94):
//     pt          = M p + t                  M   = (1 - 2|u|^2) I + 2 u u^T + 2 w [u]x
//     d pt / d d_j = B_j p                    B_j = 2w [c_j]x + 2 (u c_j^T + c_j u^T) - 4 (u.c_j) I + 2 PJ[3][j] [u]x
// with c_j = (PJ[0][j], PJ[1][j], PJ[2][j]) and PJ the 4x3 plus-Jacobian of the (w-first) QuaternionManifold applied to
// Eigen's (x,y,z,w) memory.  Building the 39 numbers once per evaluation replaces ~130 fp64
// instructions per residual (three pairs of cross products and the 4x3 chain) by 36 FMAs; the kernel was fp64-pipe
// bound (61 % busy, profiles/r1_kernels_full_v9.md).  Layout in shared memory: M[9] t[3] B0[9] B1[9] B2[9].
constexpr int kLinDoubles = 40;

__device__ __forceinline__ void build_linear_maps(const double* x, double* s_lin) {
  const int j = threadIdx.x;
  if (j > 3) return;
  const double u0 = x[0], u1 = x[1], u2 = x[2], w = x[3];
  if (j == 3) {
    const double uu = u0 * u0 + u1 * u1 + u2 * u2;
    const double dg = 1.0 - 2.0 * uu;
    s_lin[0] = dg + 2.0 * u0 * u0;      s_lin[1] = 2.0 * u0 * u1 - 2.0 * w * u2;  s_lin[2] = 2.0 * u0 * u2 + 2.0 * w * u1;
    s_lin[3] = 2.0 * u1 * u0 + 2.0 * w * u2;  s_lin[4] = dg + 2.0 * u1 * u1;      s_lin[5] = 2.0 * u1 * u2 - 2.0 * w * u0;
    s_lin[6] = 2.0 * u2 * u0 - 2.0 * w * u1;  s_lin[7] = 2.0 * u2 * u1 + 2.0 * w * u0;  s_lin[8] = dg + 2.0 * u2 * u2;
    s_lin[9] = x[4];
    s_lin[10] = x[5];
    s_lin[11] = x[6];
    return;
  }
  // plus-Jacobian column j: rows = memory slots (x,y,z,w)
  double c0, c1, c2, c3;
  if (j == 0) {
    c0 = -x[1]; c1 = x[0]; c2 = -x[3]; c3 = x[2];
  } else if (j == 1) {
    c0 = -x[2]; c1 = x[3]; c2 = x[0]; c3 = -x[1];
  } else {
    c0 = -x[3]; c1 = -x[2]; c2 = x[1]; c3 = x[0];
  }
  const double uc = u0 * c0 + u1 * c1 + u2 * c2;
  const double dg = -4.0 * uc;
  double* B = s_lin + 12 + 9 * j;
  // 2w [c]x + 2 c3 [u]x  =  [2w c + 2 c3 u]x
  const double k0 = 2.0 * (w * c0 + c3 * u0), k1 = 2.0 * (w * c1 + c3 * u1), k2 = 2.0 * (w * c2 + c3 * u2);
  B[0] = dg + 4.0 * u0 * c0;              B[1] = 2.0 * (u0 * c1 + c0 * u1) - k2;  B[2] = 2.0 * (u0 * c2 + c0 * u2) + k1;
  B[3] = 2.0 * (u1 * c0 + c1 * u0) + k2;  B[4] = dg + 4.0 * u1 * c1;              B[5] = 2.0 * (u1 * c2 + c1 * u2) - k0;
  B[6] = 2.0 * (u2 * c0 + c2 * u0) - k1;  B[7] = 2.0 * (u2 * c1 + c2 * u1) + k0;  B[8] = dg + 4.0 * u2 * c2;
}

#ifndef LM_RPT
#define LM_RPT 2
#endif
constexpr int kLmRpt = LM_RPT;                // residuals per thread per tile: independent dependency chains (ILP)
constexpr int kLmTile = kLmRpt * kLmThreads;  // records per tile

// R residuals of one class at once, straight-line code so the R dependency chains interleave (the kernel runs at
// 16 warps/SM; with one residual per thread it stalled on fixed-latency fp64 dependencies and shared-memory loads).
// valid[r] == false contributes exactly nothing (selects, never arithmetic on possibly uninitialised records).
template <int R, bool kEdge>
__device__ __forceinline__ void accumulate_residuals(const double4 (&rp)[R], const double4 (&ra)[R],
                                                     const double4 (&rb)[R], const bool (&valid)[R],
                                                     const double* __restrict__ s_lin, Eval& e) {
  double r[R], gx[R], gy[R], gz[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    const double px = rp[q].x, py = rp[q].y, pz = rp[q].z;
    const double ptx = fma(s_lin[0], px, fma(s_lin[1], py, fma(s_lin[2], pz, s_lin[9])));
    const double pty = fma(s_lin[3], px, fma(s_lin[4], py, fma(s_lin[5], pz, s_lin[10])));
    const double ptz = fma(s_lin[6], px, fma(s_lin[7], py, fma(s_lin[8], pz, s_lin[11])));
    if (kEdge) {  // point-to-line, geometry-inl.h:21-27 :  |(pt-a) x (pt-b)| / |a-b|
      const double d1x = ptx - ra[q].x, d1y = pty - ra[q].y, d1z = ptz - ra[q].z;
      const double d2x = ptx - rb[q].x, d2y = pty - rb[q].y, d2z = ptz - rb[q].z;
      const double abx = ra[q].x - rb[q].x, aby = ra[q].y - rb[q].y, abz = ra[q].z - rb[q].z;
      const double cx = fma(d1y, d2z, -(d1z * d2y)), cy = fma(d1z, d2x, -(d1x * d2z)), cz = fma(d1x, d2y, -(d1y * d2x));
      const double num = sqrt(fma(cx, cx, fma(cy, cy, cz * cz)));
      const double inv_den = 1.0 / sqrt(fma(abx, abx, fma(aby, aby, abz * abz)));
      // d r / d pt = (ab x c/|c|) / |ab| ; a point exactly on the line has no gradient (ceres::Jet norm at 0)
      const double sc = num > 0 ? inv_den / num : 0.0;
      r[q] = valid[q] ? num * inv_den : 0.0;
      gx[q] = valid[q] ? fma(aby, cz, -(abz * cy)) * sc : 0.0;
      gy[q] = valid[q] ? fma(abz, cx, -(abx * cz)) * sc : 0.0;
      gz[q] = valid[q] ? fma(abx, cy, -(aby * cx)) * sc : 0.0;
    } else {  // point-to-plane, geometry-inl.h:30-33 :  |n . pt - d| , Jet abs: derivative sign = sign of the value
      const double sd = fma(ra[q].x, ptx, fma(ra[q].y, pty, ra[q].z * ptz)) - ra[q].w;
      const double sg = copysign(1.0, sd);
      r[q] = valid[q] ? fabs(sd) : 0.0;
      gx[q] = valid[q] ? sg * ra[q].x : 0.0;
      gy[q] = valid[q] ? sg * ra[q].y : 0.0;
      gz[q] = valid[q] ? sg * ra[q].z : 0.0;
    }
  }
  // Huber(1.0) + corrector (rho'' <= 0 branch): residual and Jacobian row scaled by sqrt(rho'); rare (r > 1 m)
  double rho0[R], rc[R];
  bool any_big = false;
#pragma unroll
  for (int q = 0; q < R; q++) {
    rho0[q] = r[q] * r[q];
    rc[q] = r[q];
    any_big |= rho0[q] > 1.0;
  }
  if (any_big) {
#pragma unroll
    for (int q = 0; q < R; q++) {
      if (rho0[q] > 1.0) {
        const double rr = sqrt(rho0[q]);
        const double sr1 = sqrt(fmax(2.2250738585072014e-308, 1.0 / rr));
        rho0[q] = 2.0 * rr - 1.0;
        rc[q] *= sr1;
        gx[q] *= sr1;
        gy[q] *= sr1;
        gz[q] *= sr1;
      }
    }
  }
  // rotation part of the tangent Jacobian: each B_j is read from shared memory once and applied to the R points
  double J[R][3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double* B = s_lin + 12 + 9 * j;
    const double b0 = B[0], b1 = B[1], b2 = B[2], b3 = B[3], b4 = B[4], b5 = B[5], b6 = B[6], b7 = B[7], b8 = B[8];
#pragma unroll
    for (int q = 0; q < R; q++) {
      const double px = rp[q].x, py = rp[q].y, pz = rp[q].z;
      const double vx = fma(b0, px, fma(b1, py, b2 * pz));
      const double vy = fma(b3, px, fma(b4, py, b5 * pz));
      const double vz = fma(b6, px, fma(b7, py, b8 * pz));
      J[q][j] = fma(gx[q], vx, fma(gy[q], vy, gz[q] * vz));
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) {
    const double Jr[6] = {J[q][0], J[q][1], J[q][2], gx[q], gy[q], gz[q]};
    e.cost = fma(0.5, rho0[q], e.cost);
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      // explicit fma: the library is built -fmad=false for the index-deciding arithmetic; these well-conditioned sums
      // only have to agree with the CPU restatement to rounding
      e.g[i] = fma(Jr[i], rc[q], e.g[i]);
#pragma unroll
      for (int j = i; j < 6; j++, t++) e.H[t] = fma(Jr[i], Jr[j], e.H[t]);
    }
  }
}

// Residual records of one pair are streamed HBM -> shared memory by TMA bulk copies, kLmStages tiles of kLmTile
// records ahead of the arithmetic, so the bytes in flight do not depend on registers.  Edge residuals (records
// [0, nE): p, a, b) and plane residuals (records [capE, capE + nP): p, a) are tiled separately: a tile holds one class.
struct LmStage {
  double4 p[kLmTile];  // transformed source point, w = kind (0 = no residual for this source feature)
  double4 a[kLmTile];  // line point a / plane normal + d
  double4 b[kLmTile];  // line point b (edge tiles only)
};
#ifndef LM_STAGES
#define LM_STAGES 2
#endif
constexpr int kLmStages = LM_STAGES;

struct LmPipe {
  LmStage* stages;   // [kLmStages]
  uint64_t* full;    // [kLmStages] mbarriers, armed by the producer thread with the tile's byte count
  uint32_t tile_no;  // tiles consumed so far by this CTA (same value in every thread): stage = tile_no % kLmStages
  // A pair may be shared by the CTAs of a thread-block cluster: CTA `rank` of `size` takes the tiles rank, rank + size,
  // ...; partial sums are exchanged through distributed shared memory (size 1 = plain one-CTA-per-pair operation).
  uint32_t rank, size;
  double* partial;   // [2][28] this CTA's partial sums, double-buffered by evaluation parity
  uint32_t eval_no;  // evaluations done by this cluster (same value in every thread of every CTA)
};

// Where an evaluation streams its residual records from: the pair's edge records and a list of plane records (all of
// the pair's, or the short list of planes that are not covered by the moment sums below).
struct LmSrc {
  const double4* edge_p;
  const double4* edge_a;
  const double4* edge_b;
  uint32_t nE;
  const double4* plane_p;
  const double4* plane_a;
  uint32_t nP;
};

// thread 0: arm the stage's barrier and issue the bulk copies of tile `tile` (edge tiles first, then plane tiles)
__device__ __forceinline__ void lm_issue_tile(const LmSrc& src, uint32_t tilesE, uint32_t tile, LmStage* st, uint64_t* bar) {
  if (tile < tilesE) {
    const uint32_t lo = tile * kLmTile, n = min((uint32_t)kLmTile, src.nE - lo);
    mbar_expect_tx(bar, n * 96u);
    bulk_g2s(st->p, src.edge_p + lo, n * 32u, bar);
    bulk_g2s(st->a, src.edge_a + lo, n * 32u, bar);
    bulk_g2s(st->b, src.edge_b + lo, n * 32u, bar);
  } else {
    const uint32_t lo = (tile - tilesE) * kLmTile, n = min((uint32_t)kLmTile, src.nP - lo);
    mbar_expect_tx(bar, n * 64u);
    bulk_g2s(st->p, src.plane_p + lo, n * 32u, bar);
    bulk_g2s(st->a, src.plane_a + lo, n * 32u, bar);
  }
}

// Evaluate the streamed part of this pair's problem at x; deterministic fixed-order reduction
// (per-thread strided partial -> warp shuffle tree -> per-warp shared partials summed in warp order).
template <bool kClustered>
__device__ void evaluate_problem(const LmSrc& src, const double* x, double* s_part /*[nwarps][28]*/,
                                 double* s_tot /*[28]: H[21] g[6] cost*/, double* s_lin /*[kLinDoubles]*/, LmPipe& pipe) {
  const uint32_t nE = src.nE, nP = src.nP;
  const uint32_t tilesE = (nE + kLmTile - 1) / kLmTile, tilesP = (nP + kLmTile - 1) / kLmTile;
  const uint32_t tiles_all = tilesE + tilesP;
  // this CTA's tiles: global tile rank + j * size, j = 0 .. tiles - 1
  const uint32_t tiles = tiles_all > pipe.rank ? (tiles_all - pipe.rank + pipe.size - 1) / pipe.size : 0u;
  const uint32_t g0 = pipe.tile_no;
  if (threadIdx.x == 0) {  // every stage is free here: the previous evaluation consumed all the tiles it issued
    for (uint32_t j = 0; j < min(tiles, (uint32_t)kLmStages); j++)
      lm_issue_tile(src, tilesE, pipe.rank + j * pipe.size, pipe.stages + (g0 + j) % kLmStages,
                    pipe.full + (g0 + j) % kLmStages);
  }
  __syncthreads();  // the previous evaluation's readers of s_lin are done
  build_linear_maps(x, s_lin);
  __syncthreads();
  Eval e;
#pragma unroll
  for (int i = 0; i < 21; i++) e.H[i] = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) e.g[i] = 0;
  e.cost = 0;
  for (uint32_t j = 0; j < tiles; j++) {
    const uint32_t k = pipe.rank + j * pipe.size;  // global tile
    const uint32_t g = g0 + j, sidx = g % kLmStages;
    LmStage* st = pipe.stages + sidx;
    const bool edge = k < tilesE;
    const uint32_t n_tile = edge ? min((uint32_t)kLmTile, nE - k * kLmTile) : min((uint32_t)kLmTile, nP - (k - tilesE) * kLmTile);
    mbar_wait(pipe.full + sidx, (g / kLmStages) & 1u);
    double4 rp[kLmRpt], ra[kLmRpt], rb[kLmRpt];
    bool valid[kLmRpt];
#pragma unroll
    for (int q = 0; q < kLmRpt; q++) {
      const uint32_t slot = q * kLmThreads + threadIdx.x;
      const bool in = slot < n_tile;
      rp[q] = in ? st->p[slot] : make_double4(0, 0, 0, 0);
      valid[q] = rp[q].w != 0.0;
      ra[q] = st->a[slot];  // (beyond n_tile: stale shared memory, masked by valid)
      rb[q] = ra[q];
    }
    if (edge) {  // CTA-uniform
#pragma unroll
      for (int q = 0; q < kLmRpt; q++) rb[q] = st->b[q * kLmThreads + threadIdx.x];
      accumulate_residuals<kLmRpt, true>(rp, ra, rb, valid, s_lin, e);
    } else {
      accumulate_residuals<kLmRpt, false>(rp, ra, rb, valid, s_lin, e);
    }
    __syncthreads();  // everyone has read the stage before it is refilled
    if (threadIdx.x == 0 && j + kLmStages < tiles)
      lm_issue_tile(src, tilesE, pipe.rank + (j + kLmStages) * pipe.size, st, pipe.full + sidx);
  }
  pipe.tile_no = g0 + tiles;
  double v[28];
#pragma unroll
  for (int i = 0; i < 21; i++) v[i] = e.H[i];
#pragma unroll
  for (int i = 0; i < 6; i++) v[21 + i] = e.g[i];
  v[27] = e.cost;
#pragma unroll
  for (int i = 0; i < 28; i++) {
    double t = v[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    v[i] = t;
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 28; i++) s_part[warp * 28 + i] = v[i];
  }
  __syncthreads();
  if (!kClustered) {
    if (threadIdx.x < 28) {
      double t = 0;
      for (int wv = 0; wv < nw; wv++) t += s_part[wv * 28 + threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
  } else {
    // this CTA's partial -> its own buffer of this evaluation's parity; after the cluster barrier every CTA adds the
    // partials of all ranks in rank order, so all CTAs of the cluster hold bit-identical totals and run the
    // controller in lock step.  The buffer of parity p is rewritten two evaluations later, i.e. after another
    // cluster barrier that every reader of this evaluation has passed.
    double* mine = pipe.partial + 28 * (pipe.eval_no & 1u);
    if (threadIdx.x < 28) {
      double t = 0;
      for (int wv = 0; wv < nw; wv++) t += s_part[wv * 28 + threadIdx.x];
      mine[threadIdx.x] = t;
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (threadIdx.x < 28) {
      double t = 0;
      for (uint32_t r = 0; r < pipe.size; r++) t += cluster.map_shared_rank(mine, r)[threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
  }
  pipe.eval_no++;
  __syncthreads();
}

// ---------------------------------------------------------------------------- moment sums of the inlier planes
// The LM kernel was HBM-bound on records it re-read for every evaluation (12.3 MB per pair against 1.6 MB algorithmic,
// profiles/r1_kernels_full_v29.md).  Most residuals are point-to-plane residuals far inside Huber's quadratic region,
// and for those the sums an evaluation needs are quadratic forms in the iterate with coefficients that do not depend
// on it:  r_i(x) = | s_i + a_i . phi(x) |  with  s_i = n_i.p_i - d_i  (the signed residual at the identity update),
// a_i = n_i (x) (p_i, 1)  (12 numbers) and  phi(x) = (rows of M(x) - I | t(x))  (§6: pt = M p + t).  With
//   A = sum a a^T (12x12; (n n^T) (x) (p~ p~^T): 60 distinct sums),  v = sum a s,  S0 = sum s^2   over such records:
//   cost = 1/2 (S0 + 2 v.phi + phi^T A phi),   J^T r = D^T (v + A phi),   J^T J = D^T A D,   D = d phi / d delta
// (the sign of r_i cancels in both products; Huber's corrector is the identity for r^2 <= 1).  A record may be covered
// only if it stays an inlier at every iterate the solve evaluates: |r_i(x)| <= |s_i| + |M - I|_F |p_i| + |t|, so a
// plane with |s_i| + kCoreAlpha |p_i| + kCoreTau <= 1 is covered ("core") and an evaluation uses the sums whenever
// |M - I|_F <= kCoreAlpha and |t| <= kCoreTau (a few centimetres / hundredths of a radian: every LM candidate of a
// scan-to-scan update); otherwise it streams all records as before.  The first evaluation of a solve (identity)
// streams every record once, accumulates the sums (split over the two halves of the CTA: 42 + 31 accumulators per
// thread) and writes the few planes that are not core into a compact list; later evaluations stream the edge records
// and that list only.  The sums are reduced in a fixed order (thread -> shuffle tree -> warps in order): results are
// run-to-run identical; they differ from the streamed evaluation by rounding (different summation order).
constexpr int kMomN = 73;         // A[60] (nn6 index * 10 + pp10 index), v[12] (4 r + c), S0
constexpr int kMomHalf0 = 42;     // half 0 of the CTA: nn rows 0-2 (30 sums) + v (12); half 1: nn rows 3-5 (30) + S0
constexpr double kCoreAlpha = 0.02, kCoreTau = 0.2;

__device__ __forceinline__ int mom_nn(int r, int q) {  // index into (n0n0, n0n1, n0n2, n1n1, n1n2, n2n2)
  const int a = r < q ? r : q, b = r < q ? q : r;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);
}
__device__ __forceinline__ int mom_pp(int c, int e) {  // index into (00,01,02,03,11,12,13,22,23,33), p~3 = 1
  const int a = c < e ? c : e, b = c < e ? e : c;
  return a == 0 ? b : (a == 1 ? 3 + b : (a == 2 ? 5 + b : 9));
}
__device__ __forceinline__ double mom_A(const double* s_mom, int m, int mp) {  // A[m][m'], m = 4 r + c
  return s_mom[mom_nn(m >> 2, mp >> 2) * 10 + mom_pp(m & 3, mp & 3)];
}

// |M(x) - I|_F and |t(x)| of an iterate (M = (1 - 2|u|^2) I + 2 u u^T + 2 w [u]x)
__device__ __forceinline__ bool moments_valid_at(const double* x) {
  const double u0 = x[0], u1 = x[1], u2 = x[2], w = x[3];
  const double dg = -2.0 * (u0 * u0 + u1 * u1 + u2 * u2);
  const double m[9] = {dg + 2.0 * u0 * u0,           2.0 * u0 * u1 - 2.0 * w * u2, 2.0 * u0 * u2 + 2.0 * w * u1,
                       2.0 * u1 * u0 + 2.0 * w * u2, dg + 2.0 * u1 * u1,           2.0 * u1 * u2 - 2.0 * w * u0,
                       2.0 * u2 * u0 - 2.0 * w * u1, 2.0 * u2 * u1 + 2.0 * w * u0, dg + 2.0 * u2 * u2};
  // (Eigen's rotation formula is v + 2w (u x v) + 2 u x (u x v) whatever |q| is: M - I has no constant term)
  double f = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) f += m[i] * m[i];
  const double t2 = x[4] * x[4] + x[5] * x[5] + x[6] * x[6];
  return f <= kCoreAlpha * kCoreAlpha * 0.98 && t2 <= kCoreTau * kCoreTau * 0.98;  // (2 % slack for the roundings)
}

// First pass of a solve over ALL plane records of the pair (the update is the identity: pt = p): accumulates the moment
// sums of the core planes into s_mom and writes the other valid planes, in record order, to nc_p / nc_a.  Returns the
// length of that list, or 0xFFFFFFFF when it does not fit (the solve then streams everything).
__device__ uint32_t plane_moments_pass(const LmSrc& all, double4* __restrict__ nc_p, double4* __restrict__ nc_a,
                                       uint32_t nc_cap, double* s_mpart /*[8][kMomHalf0]*/, double* s_mom /*[kMomN]*/,
                                       uint32_t* s_ncw /*[17]*/, LmPipe& pipe) {
  LmSrc planes = all;
  planes.nE = 0;
  const uint32_t tiles = (all.nP + kLmTile - 1) / kLmTile;
  const uint32_t g0 = pipe.tile_no;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t half = tid >> 7, u = tid & 127;  // both halves visit every record; each keeps its share of the sums
  if (tid == 0) {
    for (uint32_t j = 0; j < min(tiles, (uint32_t)kLmStages); j++)
      lm_issue_tile(planes, 0, j, pipe.stages + (g0 + j) % kLmStages, pipe.full + (g0 + j) % kLmStages);
  }
  double acc[kMomHalf0];
#pragma unroll
  for (int i = 0; i < kMomHalf0; i++) acc[i] = 0.0;
  uint32_t nc_base = 0;
  bool overflow = false;
  for (uint32_t j = 0; j < tiles; j++) {
    const uint32_t g = g0 + j, sidx = g % kLmStages;
    LmStage* st = pipe.stages + sidx;
    const uint32_t n_tile = min((uint32_t)kLmTile, all.nP - j * kLmTile);
    mbar_wait(pipe.full + sidx, (g / kLmStages) & 1u);
    unsigned ncb[kLmTile / 128];  // per quarter of the tile: this warp's ballot of non-core planes
#pragma unroll
    for (int q = 0; q < kLmTile / 128; q++) {
      const uint32_t slot = q * 128 + u;
      const bool in = slot < n_tile;
      const double4 p = in ? st->p[slot] : make_double4(0, 0, 0, 0);
      const double4 n = st->a[slot];  // (beyond n_tile: stale shared memory, masked by `valid`)
      const bool valid = p.w != 0.0;
      const double sd = fma(n.x, p.x, fma(n.y, p.y, n.z * p.z)) - n.w;
      // |s| + alpha |p| + tau <= 1  without the square root:  alpha^2 |p|^2 <= (1 - tau - |s|)^2, right side >= 0
      const double room = (1.0 - kCoreTau) - fabs(sd);
      const double p2 = fma(p.x, p.x, fma(p.y, p.y, p.z * p.z));
      const bool core = valid && room >= 0.0 && (kCoreAlpha * kCoreAlpha) * p2 <= room * room;
      ncb[q] = __ballot_sync(0xffffffffu, valid && !core);
      if (core) {
        const double pp[10] = {p.x * p.x, p.x * p.y, p.x * p.z, p.x, p.y * p.y, p.y * p.z, p.y, p.z * p.z, p.z, 1.0};
        if (half == 0) {
          const double nn[3] = {n.x * n.x, n.x * n.y, n.x * n.z};
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 10; k++) acc[i * 10 + k] = fma(nn[i], pp[k], acc[i * 10 + k]);
          const double ns[3] = {n.x * sd, n.y * sd, n.z * sd};
          const double pt4[4] = {p.x, p.y, p.z, 1.0};
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[30 + 4 * r + c] = fma(ns[r], pt4[c], acc[30 + 4 * r + c]);
        } else {
          const double nn[3] = {n.y * n.y, n.y * n.z, n.z * n.z};
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 10; k++) acc[i * 10 + k] = fma(nn[i], pp[k], acc[i * 10 + k]);
          acc[30] = fma(sd, sd, acc[30]);
        }
      }
    }
    // compact list of the non-core planes, in record order: quarter-major, then the four warps of half 0, then lanes
    if (half == 0 && lane == 0) {
#pragma unroll
      for (int q = 0; q < kLmTile / 128; q++) s_ncw[q * 4 + warp] = (uint32_t)__popc(ncb[q]);
    }
    __syncthreads();
    {
      uint32_t total = 0;
      for (int k = 0; k < 4 * (kLmTile / 128); k++) total += s_ncw[k];
      if (half == 0 && !overflow) {
        uint32_t before = 0;
#pragma unroll
        for (int q = 0; q < kLmTile / 128; q++) {
          uint32_t mine = before;
          for (uint32_t w = 0; w < warp; w++) mine += s_ncw[q * 4 + w];
          if (ncb[q] & (1u << lane)) {
            const uint32_t pos = nc_base + mine + (uint32_t)__popc(ncb[q] & ((1u << lane) - 1u));
            if (pos < nc_cap) {
              const uint32_t slot = q * 128 + u;
              nc_p[pos] = st->p[slot];
              nc_a[pos] = st->a[slot];
            }
          }
          for (uint32_t w = 0; w < 4; w++) before += s_ncw[q * 4 + w];
        }
      }
      nc_base += total;
      overflow = overflow || nc_base > nc_cap;
    }
    __syncthreads();  // everyone has read the stage (and the counts) before it is refilled
    if (tid == 0 && j + kLmStages < tiles) lm_issue_tile(planes, 0, j + kLmStages, st, pipe.full + sidx);
  }
  pipe.tile_no = g0 + tiles;
  // reduce: shuffle tree per warp, then the four warps of each half in order
  const int n_mine = half == 0 ? kMomHalf0 : 31;
#pragma unroll
  for (int i = 0; i < kMomHalf0; i++) {
    double t = acc[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0 && i < n_mine) s_mpart[warp * kMomHalf0 + i] = t;
  }
  // the compact list was written with plain stores and is read next by bulk copies (async proxy, through L2)
  __threadfence();
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  if (tid < kMomN) {
    // s_mom index -> (half, local index): A rows 0-2 and v live in half 0, A rows 3-5 and S0 in half 1
    const int h = (tid < 30 || (tid >= 60 && tid < 72)) ? 0 : 1;
    const int l = tid < 30 ? tid : (tid < 60 ? tid - 30 : (tid < 72 ? 30 + (tid - 60) : 30));
    double t = 0;
    for (int w = 0; w < 4; w++) t += s_mpart[(4 * h + w) * kMomHalf0 + l];
    s_mom[tid] = t;
  }
  __syncthreads();
  return overflow ? 0xFFFFFFFFu : nc_base;
}

// Adds the covered planes' share of the sums at the iterate whose linear maps are in s_lin (M[9] t[3] B0[9] B1[9] B2[9]).
__device__ void moment_contrib(const double* __restrict__ s_lin, const double* __restrict__ s_mom,
                               double* s_scr /*[12 phi][12 w][72 AD]*/, double* s_tot /*[28] +=*/) {
  const int tid = threadIdx.x;
  double* phi = s_scr;
  double* wv = s_scr + 12;
  double* AD = s_scr + 24;
  auto D = [&](int m, int j) -> double {  // d phi_m / d delta_j
    const int r = m >> 2, c = m & 3;
    if (j < 3) return c < 3 ? s_lin[12 + 9 * j + 3 * r + c] : 0.0;
    return (c == 3 && r == j - 3) ? 1.0 : 0.0;
  };
  if (tid < 12) {
    const int r = tid >> 2, c = tid & 3;
    phi[tid] = c < 3 ? s_lin[3 * r + c] - (r == c ? 1.0 : 0.0) : s_lin[9 + r];
  }
  __syncthreads();
  if (tid < 72) {
    const int m = tid / 6, k = tid % 6;
    double t = 0;
    for (int mp = 0; mp < 12; mp++) t = fma(mom_A(s_mom, m, mp), D(mp, k), t);
    AD[tid] = t;
  } else if (tid < 84) {
    const int m = tid - 72;
    double t = s_mom[60 + m];
    for (int mp = 0; mp < 12; mp++) t = fma(mom_A(s_mom, m, mp), phi[mp], t);
    wv[m] = t;
  }
  __syncthreads();
  if (tid < 21) {
    int i = 0, rem = tid;
    while (rem >= 6 - i) {
      rem -= 6 - i;
      i++;
    }
    const int j = i + rem;
    double t = 0;
    for (int m = 0; m < 12; m++) t = fma(D(m, i), AD[m * 6 + j], t);
    s_tot[tid] += t;
  } else if (tid < 27) {
    const int j = tid - 21;
    double t = 0;
    for (int m = 0; m < 12; m++) t = fma(D(m, j), wv[m], t);
    s_tot[tid] += t;
  } else if (tid == 27) {
    double t = s_mom[72];
    for (int m = 0; m < 12; m++) t = fma(phi[m], s_mom[60 + m] + wv[m], t);
    s_tot[27] += 0.5 * t;
  }
  __syncthreads();
}

__device__ __forceinline__ int hidx(int i, int j) {  // upper-triangle index, i <= j
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

// Solve (A) y = b for symmetric positive definite 6x6 by Cholesky; returns false if not SPD / not finite.
// One reciprocal per pivot (the serial chain of 27 fp64 divisions + 6 square roots was a visible part of the
// ~16 us every CTA spent per LM iteration outside the residual evaluations).
__device__ __forceinline__ bool chol6_solve(const double (&A)[6][6], const double (&b)[6], double (&y)[6]) {
  double L[6][6], inv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double s = A[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (!(s > 0.0)) return false;
    const double d = sqrt(s);
    inv[j] = 1.0 / d;
    L[j][j] = d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double t = A[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t * inv[j];
    }
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i][k] * z[k];
    z[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * y[k];
    y[i] = s * inv[i];
  }
  bool fin = true;
#pragma unroll
  for (int i = 0; i < 6; i++) fin = fin && isfinite(y[i]);
  return fin;
}

__device__ __forceinline__ double grad_max_norm(const double* x, const double* g) {
  double ng[6], proj[7];
#pragma unroll
  for (int j = 0; j < 6; j++) ng[j] = -g[j];
  manifold_plus(x, ng, proj);
  double m = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) m = fmax(m, fabs(x[i] - proj[i]));
  return m;
}

__device__ __forceinline__ double norm7(const double* x) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) s += x[i] * x[i];
  return sqrt(s);
}

// One CTA per pair.  Every thread runs the (tiny, uniform) controller redundantly on the reduced sums, so
// no broadcast of the step is needed between evaluations.
struct LmMomentScratch {  // shared memory of the moment path (one-CTA-per-pair kernel only)
  double mom[kMomN];
  double part[8 * kMomHalf0];
  double scr[96];
  uint32_t ncw[17];
};

template <bool kClustered>
__device__ void lm_pair(const LmArgs& a, uint32_t pair, double* s_part, double* s_sum /*[2][28]*/, double* s_lin,
                        LmPipe& pipe, LmMomentScratch* ms) {
  PairState* ps = a.state + pair;
  if (ps->status != -1) return;
  const uint32_t src_slot = (uint32_t)((a.pair0 + pair + a.src_offset) % a.n_slots);
  const uint32_t nE = a.feat_counts[src_slot * 2], nP = a.feat_counts[src_slot * 2 + 1];
  const uint32_t n_ea = ps->n_edge_assoc, n_pa = ps->n_plane_assoc;
  // everyone (every CTA of the cluster) has read the counters before thread 0 of rank 0 resets them
  if (!kClustered)
    __syncthreads();
  else
    cg::this_cluster().sync();
  const bool writer = threadIdx.x == 0 && pipe.rank == 0;
  if (writer) {
    ps->n_edge_assoc = 0;
    ps->n_plane_assoc = 0;
  }
  if ((uint64_t)n_ea + n_pa < a.rp.min_assoc) {  // registration-inl.h:45-48
    if (writer) ps->status = 2;
    return;
  }
  const size_t cap_src = (size_t)a.capE_scan + a.capP_scan;
  const double4* rec_p = a.rec_p + (size_t)pair * cap_src;
  const double4* rec_a = a.rec_a + (size_t)pair * cap_src;
  const double4* rec_b = a.rec_b + (size_t)pair * a.capE_scan;
  const LmSrc all{rec_p, rec_a, rec_b, nE, rec_p + a.capE_scan, rec_a + a.capE_scan, nP};
  LmSrc rest = all;      // what an evaluation streams when the moment sums cover the core planes
  bool use_mom = false;  // (CTA-uniform)

  // ---- Ceres TrustRegionMinimizer, LEVENBERG_MARQUARDT, max_num_iterations = 4, defaults otherwise
  const int max_num_iterations = 4;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;

  double x[7] = {0, 0, 0, 1, 0, 0, 0};
  double x_norm = norm7(x);
  // reduced sums (H[21], g[6], cost) of the accepted point and of the candidate live in two shared-memory buffers
  // that swap roles when a step is accepted: nothing of that size is copied or kept in per-thread local memory
  const double* cur = s_sum;
  double* other = s_sum + 28;
  double scale[6], diag[6];
  uint32_t lm_iterations = 0;
  double cost0 = 0, x_cost = 0;
  if (n_ea + n_pa > 0) {
#ifdef LM_TIMING
    long long t_eval = 0, t_all0 = clock64();
    { const long long c0 = clock64();
#endif
    if (!kClustered && ms != nullptr && a.nc_p != nullptr && nP > 0) {
      double4* nc_p = a.nc_p + (size_t)pair * a.nc_cap;
      double4* nc_a = a.nc_a + (size_t)pair * a.nc_cap;
      const uint32_t n_nc = plane_moments_pass(all, nc_p, nc_a, a.nc_cap, ms->part, ms->mom, ms->ncw, pipe);
      if (n_nc != 0xFFFFFFFFu) {
        use_mom = true;
        rest.plane_p = nc_p;
        rest.plane_a = nc_a;
        rest.nP = n_nc;
      }
    }
    // every evaluation of this solve: the streamed part, plus the closed form of the covered planes
    auto evaluate = [&](const double* at, double* out) {
      const bool mom = use_mom && moments_valid_at(at);
      evaluate_problem<kClustered>(mom ? rest : all, at, s_part, out, s_lin, pipe);
      if (mom) moment_contrib(s_lin, ms->mom, ms->scr, out);
    };
    evaluate(x, s_sum);
#ifdef LM_TIMING
    t_eval += clock64() - c0; }
#endif
    x_cost = cur[27];
    cost0 = x_cost;
#pragma unroll
    for (int j = 0; j < 6; j++) scale[j] = 1.0 / (1.0 + sqrt(cur[hidx(j, j)]));
    double gmax = grad_max_norm(x, cur + 21);
    bool step_successful = true, armed = false;
    int iteration = 0;
    for (;;) {
      if (iteration >= max_num_iterations) break;
      if (step_successful && gmax <= gradient_tolerance) break;
      if (radius <= min_radius) break;
      iteration++;
      // scaled normal equations: Hs = S H S, gs = S g
      double Hs[6][6], gs[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        gs[i] = cur[21 + i] * scale[i];
#pragma unroll
        for (int j = i; j < 6; j++) {
          const double v = cur[hidx(i, j)] * scale[i] * scale[j];
          Hs[i][j] = v;
          Hs[j][i] = v;
        }
      }
      if (!reuse_diagonal) {
#pragma unroll
        for (int j = 0; j < 6; j++) diag[j] = fmin(fmax(Hs[j][j], min_diag), max_diag);
      }
      double A[6][6], y[6], step[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j < 6; j++) A[i][j] = Hs[i][j];
        A[i][i] += diag[i] / radius;  // (sqrt(diag/radius))^2
      }
      const bool solved = chol6_solve(A, gs, y);
      reuse_diagonal = true;
      bool valid = false;
      double model_cost_change = 0;
      if (solved) {
        double sg = 0, shs = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
          step[i] = -y[i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
          sg += step[i] * gs[i];
          double t = 0;
#pragma unroll
          for (int j = 0; j < 6; j++) t += Hs[i][j] * step[j];
          shs += step[i] * t;
        }
        model_cost_change = -(sg + 0.5 * shs);
        valid = model_cost_change > 0.0;
      }
      if (!valid) {  // HandleInvalidStep
        radius *= 0.5;
        reuse_diagonal = true;
        step_successful = false;
        continue;
      }
      double delta[6], cand[7];
#pragma unroll
      for (int j = 0; j < 6; j++) delta[j] = step[j] * scale[j];
      manifold_plus(x, delta, cand);
#ifdef LM_TIMING
      const long long c1 = clock64();
#endif
      evaluate(cand, other);
#ifdef LM_TIMING
      t_eval += clock64() - c1;
#endif
      const double cand_cost = other[27];
      if (armed) {
        double dn = 0;
#pragma unroll
        for (int i = 0; i < 7; i++) dn += (x[i] - cand[i]) * (x[i] - cand[i]);
        if (sqrt(dn) <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
        if (fabs(x_cost - cand_cost) <= function_tolerance * x_cost) break;
      }
      const double rel = (x_cost - cand_cost) / model_cost_change;
      if (rel > min_relative_decrease) {  // HandleSuccessfulStep (evaluation at cand already holds J, g)
#pragma unroll
        for (int i = 0; i < 7; i++) x[i] = cand[i];
        x_norm = norm7(x);
        {
          double* t = const_cast<double*>(cur);
          cur = other;
          other = t;
        }
        x_cost = cand_cost;
        gmax = grad_max_norm(x, cur + 21);
        step_successful = true;
        const double q = 2.0 * rel - 1.0;
        radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
        radius = fmin(max_radius, radius);
        decrease_factor = 2.0;
        reuse_diagonal = false;
        armed = true;
      } else {
        step_successful = false;
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = true;
      }
    }
    lm_iterations = (uint32_t)iteration;
#ifdef LM_TIMING
    if (threadIdx.x == 0 && (pair % 64) == 0)
      printf("LMT pair %u outer %d: total %lld cycles, in evaluations %lld, lm iterations %d, residuals %u\n", pair,
             a.outer_iter, clock64() - t_all0, t_eval, iteration, nE + nP);
#endif
  }

  // ---- ICF update (registration-inl.h:59-73)
  if (writer) {
    double est[7], nxt[7];
#pragma unroll
    for (int j = 0; j < 7; j++) est[j] = ps->est[j];
    if (a.d_iter_est && pair == 0) {
      for (int j = 0; j < 7; j++) {
        a.d_iter_est[7 * a.outer_iter + j] = est[j];
        a.d_iter_update[7 * a.outer_iter + j] = x[j];
      }
      a.d_assoc_n[2 * a.outer_iter] = n_ea;
      a.d_assoc_n[2 * a.outer_iter + 1] = n_pa;
      a.d_lm_iters[a.outer_iter] = lm_iterations;
      a.d_lm_cost[2 * a.outer_iter] = cost0;
      a.d_lm_cost[2 * a.outer_iter + 1] = x_cost;
    }
    quat_mul(x, est, nxt);  // left composition: est = update.compose(est), geometry.cpp:16-18
    const V3 rt = quat_rotate(x, V3{est[4], est[5], est[6]});
    nxt[4] = x[4] + rt.x;
    nxt[5] = x[5] + rt.y;
    nxt[6] = x[6] + rt.z;
#pragma unroll
    for (int j = 0; j < 7; j++) ps->est[j] = nxt[j];
    ps->iters = ps->iters + 1;
    // angularDistance(update.q, Identity) = 2 atan2(|vec|, |w|) ; |t|
    const double ang = 2.0 * atan2(sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), fabs(x[3]));
    const double pos = sqrt(x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
    if (ang < a.rp.rot_thr && pos < a.rp.pos_thr) {
      ps->status = 0;
    } else if (a.outer_iter + 1 >= a.rp.max_iterations) {
      ps->status = 1;
    }
  }
}

// kClustered = false: one CTA per pair (sequence odometry).  kClustered = true: launched with a cluster dimension, the
// CTAs of a cluster share one pair (single registrations).
template <bool kClustered>
__global__ void __launch_bounds__(kLmThreads, kLmMinBlocks) lm_kernel(LmArgs a) {
  extern __shared__ __align__(128) unsigned char lm_smem[];  // kLmStages x LmStage
  __shared__ double s_part[(kLmThreads / 32) * 28];
  __shared__ double s_sum[2 * 28];
  __shared__ double s_partial[kClustered ? 2 * 28 : 1];
  __shared__ __align__(16) double s_lin[kLinDoubles];
  __shared__ __align__(8) uint64_t s_full[kLmStages];
  __shared__ LmMomentScratch s_ms[1];  // (unused by the clustered kernel: 4 KB)
  static_assert(kLmThreads == 256 && kLmTile == 512, "the moment pass splits a 256-thread CTA in two halves of four warps");
  LmPipe pipe;
  pipe.stages = reinterpret_cast<LmStage*>(lm_smem);
  pipe.full = s_full;
  pipe.tile_no = 0;
  pipe.rank = 0;
  pipe.size = 1;
  if (kClustered) {
    cg::cluster_group cluster = cg::this_cluster();
    pipe.rank = cluster.block_rank();
    pipe.size = cluster.num_blocks();
  }
  pipe.partial = s_partial;
  pipe.eval_no = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLmStages; i++) mbar_init(s_full + i, 1);
  __syncthreads();
  const uint32_t n_act = active_count(a.active, a.n_pairs);
  const uint32_t cluster_id = blockIdx.x / pipe.size, n_clusters = gridDim.x / pipe.size;
  for (uint32_t i = cluster_id; i < n_act; i += n_clusters) {
    lm_pair<kClustered>(a, active_pair(a.active, i), s_part, s_sum, s_lin, pipe, kClustered ? nullptr : s_ms);
    __syncthreads();
  }
  if (kClustered) cg::this_cluster().sync();  // nobody leaves while a peer may still read its partial sums
}

// TEST HOOK (loamgpu_debug_problem_eval): the sums one evaluation of the LM kernel produces for pair 0's records at an
// arbitrary iterate — through the streamed evaluation (mode 0) or through the moment path (mode 1: first pass over the
// planes at the identity, then the hybrid evaluation; out[28] = 1 if the moment sums were used at x, out[29] = planes
// left uncovered).  tests/test_gpu_jacobians.py compares them with torch autograd of the reference's functors.
__global__ void __launch_bounds__(kLmThreads, kLmMinBlocks) lm_debug_eval_kernel(LmArgs a, const double* x_dev, int mode,
                                                                                double* out) {
  extern __shared__ __align__(128) unsigned char lm_smem[];
  __shared__ double s_part[(kLmThreads / 32) * 28];
  __shared__ double s_sum[28];
  __shared__ __align__(16) double s_lin[kLinDoubles];
  __shared__ __align__(8) uint64_t s_full[kLmStages];
  __shared__ LmMomentScratch s_ms;
  LmPipe pipe;
  pipe.stages = reinterpret_cast<LmStage*>(lm_smem);
  pipe.full = s_full;
  pipe.tile_no = 0;
  pipe.rank = 0;
  pipe.size = 1;
  pipe.partial = nullptr;
  pipe.eval_no = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLmStages; i++) mbar_init(s_full + i, 1);
  __syncthreads();
  const uint32_t nE = a.feat_counts[0], nP = a.feat_counts[1];
  const LmSrc all{a.rec_p, a.rec_a, a.rec_b, nE, a.rec_p + a.capE_scan, a.rec_a + a.capE_scan, nP};
  LmSrc rest = all;
  double x[7];
#pragma unroll
  for (int i = 0; i < 7; i++) x[i] = x_dev[i];
  bool mom = false;
  uint32_t n_nc = 0;
  if (mode == 1 && nP > 0) {
    n_nc = plane_moments_pass(all, a.nc_p, a.nc_a, a.nc_cap, s_ms.part, s_ms.mom, s_ms.ncw, pipe);
    if (n_nc != 0xFFFFFFFFu && moments_valid_at(x)) {
      mom = true;
      rest.plane_p = a.nc_p;
      rest.plane_a = a.nc_a;
      rest.nP = n_nc;
    }
  }
  evaluate_problem<false>(mom ? rest : all, x, s_part, s_sum, s_lin, pipe);
  if (mom) moment_contrib(s_lin, s_ms.mom, s_ms.scr, s_sum);
  if (threadIdx.x < 28) out[threadIdx.x] = s_sum[threadIdx.x];
  if (threadIdx.x == 0) {
    out[28] = mom ? 1.0 : 0.0;
    out[29] = (double)n_nc;
  }
}

// Rebuilds the list of pairs still iterating (ascending pair index) after an LM launch.
__global__ void __launch_bounds__(1024) compact_active_kernel(const PairState* st, uint32_t n_pairs, uint32_t* active) {
  __shared__ uint32_t s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (n_pairs + blockDim.x - 1) / blockDim.x;
  const uint32_t b = min(tid * per, n_pairs), e = min(b + per, n_pairs);
  uint32_t mine = 0;
  for (uint32_t i = b; i < e; i++) mine += st[i].status == -1 ? 1u : 0u;
  uint32_t inc = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t off = 0, total = 0;
  for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
    if (w < warp) off += s_warp[w];
    total += s_warp[w];
  }
  uint32_t pos = off + inc - mine;
  for (uint32_t i = b; i < e; i++)
    if (st[i].status == -1) active[1 + pos++] = i;
  if (tid == 0) active[0] = total;
}

__global__ void init_pairs_kernel(PairState* st, uint32_t n, const double* init, uint32_t init_stride) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PairState s;
  const double ident[7] = {0, 0, 0, 1, 0, 0, 0};
  for (int j = 0; j < 7; j++) s.est[j] = init ? init[(size_t)i * init_stride + j] : ident[j];
  s.status = -1;
  s.iters = 0;
  s.n_edge_assoc = 0;
  s.n_plane_assoc = 0;
  st[i] = s;
}

__global__ void transform_points_kernel(double4* pts, uint32_t n, const double* pose) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ps[7];
#pragma unroll
  for (int j = 0; j < 7; j++) ps[j] = pose[j];
  const double4 p = pts[i];
  const V3 q = pose_act(ps, V3{p.x, p.y, p.z});
  pts[i] = make_double4(q.x, q.y, q.z, 0.0);
}

__global__ void widen_kernel(WidenArgs a) {
  const uint32_t seg = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n[seg]) return;
  const double* s = a.src[seg] + 3 * (size_t)i;
  a.dst[seg][i] = make_double4(s[0], s[1], s[2], 0.0);
}

__global__ void finish_pairs_kernel(const PairState* st, uint32_t n, double* poses, int32_t* term, uint32_t* iters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PairState s = st[i];
  if (poses)
    for (int j = 0; j < 7; j++) poses[7 * (size_t)i + j] = s.est[j];
  if (term) term[i] = s.status == -1 ? 1 : s.status;  // never stopped early => MAX_ITER (registration-inl.h:27)
  if (iters) iters[i] = s.iters;
}

}  // namespace

// ============================================================================ host launchers

static bool sb_fits(const BvhBuildArgs& a, int optin) {
  static const bool legacy = []() { const char* e = getenv("LOAMGPU_BUILD_LEGACY"); return e && atoi(e) != 0; }();
  // sets whose intermediates fit in shared memory (16-bit point numbers, 8 bytes per point + counters)
  return !legacy && a.g.pt_cap <= kSbMaxPoints && sb_smem_bytes(a.g.pt_cap) + 1024 <= (size_t)optin;
}

static cudaError_t launch_sb(const BvhBuildArgs& a0, const BvhBuildArgs& a1, uint32_t n_sets, uint32_t kinds, cudaStream_t st,
                             int optin) {
  // one CTA per SM whatever it asks for: take all the shared memory there is — what the sort and the tree arrays of a
  // set leave free holds the boxes of its big nodes (bvh.cuh)
  const size_t need = std::max(sb_smem_bytes(a0.g.pt_cap), kinds > 1 ? sb_smem_bytes(a1.g.pt_cap) : (size_t)0);
  const size_t smem = std::max(need, (size_t)optin - 1024);
  cudaError_t err = cudaFuncSetAttribute(bvh_build_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  // $LOAMGPU_BUILD_GLOBAL_BOXES=1: tell the kernel it has no room for the boxes (A/B and test of its other path)
  static const bool global_boxes = []() { const char* e = getenv("LOAMGPU_BUILD_GLOBAL_BOXES"); return e && atoi(e) != 0; }();
  bvh_build_smem_kernel<<<dim3(n_sets, kinds), kSbThreads, smem, st>>>(a0, a1, global_boxes ? 0u : (uint32_t)smem);
  return cudaGetLastError();
}

cudaError_t launch_bvh_build(const BvhBuildArgs& a_in, uint32_t n_sets, cudaStream_t st) {
  if (n_sets == 0) return cudaSuccess;
  BvhBuildArgs a = a_in;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (sb_fits(a, optin)) return launch_sb(a, a, n_sets, 1, st, optin);
  const size_t sort_bytes = (size_t)kRadixBins * kBuildThreads * sizeof(uint32_t);
  // radix counters, reused after the sort for the sorted codes; + one readiness byte per node
  const size_t tree_bytes = std::max(sort_bytes, (size_t)a.g.pt_cap * 4) + a.g.pt_cap + 16;
  a.smem_tree = tree_bytes + 4096 <= (size_t)optin;  // (the kernel also has ~2.4 KB of static shared memory)
  const size_t smem = a.smem_tree ? tree_bytes : sort_bytes;
  cudaError_t err = cudaFuncSetAttribute(bvh_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  bvh_build_kernel<<<n_sets, kBuildThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_bvh_build2(const BvhBuildArgs& edge, const BvhBuildArgs& planar, uint32_t n_sets, cudaStream_t st,
                              uint64_t* launches) {
  if (n_sets == 0) return cudaSuccess;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (sb_fits(edge, optin) && sb_fits(planar, optin)) {
    if (launches) *launches += 1;
    return launch_sb(edge, planar, n_sets, 2, st, optin);
  }
  if (launches) *launches += 2;
  cudaError_t err = launch_bvh_build(edge, n_sets, st);
  if (err != cudaSuccess) return err;
  return launch_bvh_build(planar, n_sets, st);
}

// Rows of pair-indexed grids: every pair in the first two outer iterations (nearly all are still active), a short
// strided grid afterwards (typically nothing is left after 2-3 iterations).
static uint32_t pair_rows(uint32_t n_pairs, int outer_iter, bool has_list, uint32_t late_rows) {
  return (outer_iter < 2 || !has_list) ? n_pairs : std::min(n_pairs, late_rows);
}

template <int K>
static cudaError_t launch_knn_general(const AssocArgs& a, dim3 grid, int outer_iter, const uint32_t* list, cudaStream_t st) {
  assoc_knn_kernel<K><<<grid, kKnnThreads, 0, st>>>(a, outer_iter, list);
  return cudaGetLastError();
}

// Shared-memory budget of the batched k-NN kernel (bytes of compact records per CTA).  What is not given to shared
// memory stays L1 (leaf points, traversal stacks).  $LOAMGPU_KNN_SMEM_KB overrides; 0 switches the kernel off.
static uint32_t knn_smem_budget(int optin) {
  static const long env = []() { const char* e = getenv("LOAMGPU_KNN_SMEM_KB"); return e ? atol(e) : -1L; }();
  const long kb = env >= 0 ? env : 128;
  return (uint32_t)std::min<long>(kb * 1024, (long)optin - 2048);
}

cudaError_t launch_assoc_knn(const AssocArgs& a, uint32_t n_pairs, int outer_iter, cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  const uint32_t cap = a.capE_scan + a.capP_scan;
  const uint32_t grid_x = (cap + kKnnThreads - 1) / kKnnThreads;
  const int kmax = a.rp.ke > a.rp.kp ? a.rp.ke : a.rp.kp;
  static int n_sm = 0, optin = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  static const uint32_t min_pairs = []() { const char* e = getenv("LOAMGPU_KNN_SMEM_MIN_PAIRS"); return e ? (uint32_t)atol(e) : 16u; }();
  const uint32_t budget = knn_smem_budget(optin);
  // The batched kernel needs this context's own sets (compact records), Morton-ordered queries, register-resident
  // top-k lists and enough pairs to fill the SMs; everything else takes the general kernel alone.
  const bool batched = !a.ext_target && a.morton_queries && a.ge.quant && a.gp.quant && a.leftover && kmax <= kKnnRegMax &&
                       n_pairs >= min_pairs && budget >= 32 * 1024;
  const uint32_t* list = a.active;
  uint32_t rows = pair_rows(n_pairs, outer_iter, a.active != nullptr, 16);
  if (batched) {
    cudaError_t err = cudaMemsetAsync(a.leftover, 0, sizeof(uint32_t), st);
    if (err != cudaSuccess) return err;
    // items per launch ~ 6 per SM: whole waves of similar items, still few copies of a pair's records
    const uint32_t n_slices = std::min<uint32_t>(16u, std::max<uint32_t>(1u, (6u * (uint32_t)n_sm + n_pairs - 1) / n_pairs));
    const uint32_t grid = std::min<uint32_t>((uint32_t)n_sm, rows * n_slices);
    const uint32_t ring = (uint32_t)KNN_SMEM_STACK * kKnnStackPitch;  // traversal-stack ring in front of the records
    if (budget <= ring + 32 * 1024) return cudaErrorInvalidConfiguration;
    const uint32_t max_recs = (budget - ring) / (uint32_t)sizeof(BvhRec);
    if (kmax <= kKnnSmall) {
      err = cudaFuncSetAttribute(assoc_knn_smem_kernel<kKnnSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
      if (err != cudaSuccess) return err;
      assoc_knn_smem_kernel<kKnnSmall><<<grid, kKnnCtaThreads, budget, st>>>(a, outer_iter, n_slices, max_recs);
    } else {
      err = cudaFuncSetAttribute(assoc_knn_smem_kernel<kKnnRegMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
      if (err != cudaSuccess) return err;
      assoc_knn_smem_kernel<kKnnRegMax><<<grid, kKnnCtaThreads, budget, st>>>(a, outer_iter, n_slices, max_recs);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    // whatever it left over (normally nothing): a short strided grid of the general kernel
    list = a.leftover;
    rows = std::min<uint32_t>(n_pairs, 32u);
  }
  const dim3 grid(grid_x, rows);
  if (kmax <= kKnnSmall) return launch_knn_general<kKnnSmall>(a, grid, outer_iter, list, st);
  if (kmax <= kKnnRegMax) return launch_knn_general<kKnnRegMax>(a, grid, outer_iter, list, st);
  return launch_knn_general<kKnnMax>(a, grid, outer_iter, list, st);
}

cudaError_t launch_assoc_fit(const AssocArgs& a, uint32_t n_pairs, int outer_iter, cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  const uint32_t cap = a.capE_scan + a.capP_scan;
  dim3 grid((cap + kAssocThreads - 1) / kAssocThreads, pair_rows(n_pairs, outer_iter, a.active != nullptr, 16));
  const int kmax = a.rp.ke > a.rp.kp ? a.rp.ke : a.rp.kp;
  if (kmax <= kKnnSmall)
    assoc_fit_kernel<kKnnSmall><<<grid, kAssocThreads, 0, st>>>(a, outer_iter);
  else if (kmax <= kKnnRegMax)
    assoc_fit_kernel<kKnnRegMax><<<grid, kAssocThreads, 0, st>>>(a, outer_iter);
  else
    assoc_fit_kernel<kKnnMax><<<grid, kAssocThreads, 0, st>>>(a, outer_iter);
  return cudaGetLastError();
}

cudaError_t launch_lm(const LmArgs& a, uint32_t n_pairs, cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  const size_t smem = (size_t)kLmStages * sizeof(LmStage);
  const uint32_t rows = pair_rows(n_pairs, a.outer_iter, a.active != nullptr, 296);
  const uint32_t csize = a.cluster > 1 ? a.cluster : 1u;
  if (csize == 1) {
    cudaError_t err = cudaFuncSetAttribute(lm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    lm_kernel<false><<<rows, kLmThreads, smem, st>>>(a);
    return cudaGetLastError();
  }
  cudaError_t err = cudaFuncSetAttribute(lm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(rows * csize);
  cfg.blockDim = dim3(kLmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lm_kernel<true>, a);
}

cudaError_t launch_lm_debug_eval(const LmArgs& a, const double* x_dev, int mode, double* out_dev, cudaStream_t st) {
  const size_t smem = (size_t)kLmStages * sizeof(LmStage);
  cudaError_t err = cudaFuncSetAttribute(lm_debug_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  lm_debug_eval_kernel<<<1, kLmThreads, smem, st>>>(a, x_dev, mode, out_dev);
  return cudaGetLastError();
}

cudaError_t launch_compact_active(const PairState* s, uint32_t n_pairs, uint32_t* active, cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  compact_active_kernel<<<1, 1024, 0, st>>>(s, n_pairs, active);
  return cudaGetLastError();
}

cudaError_t launch_knn(const KnnArgs& a, cudaStream_t st) {
  if (a.n_queries == 0) return cudaSuccess;
  const uint32_t blocks = (uint32_t)((a.n_queries + 127) / 128);
  if (a.k <= kKnnSmall)
    knn_kernel<kKnnSmall><<<blocks, 128, 0, st>>>(a);
  else if (a.k <= kKnnRegMax)
    knn_kernel<kKnnRegMax><<<blocks, 128, 0, st>>>(a);
  else
    knn_kernel<kKnnMax><<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_widen(const WidenArgs& a, cudaStream_t st) {
  const uint32_t nmax = std::max(std::max(a.n[0], a.n[1]), std::max(a.n[2], a.n[3]));
  if (nmax == 0) return cudaSuccess;
  widen_kernel<<<dim3((nmax + 255) / 256, 4), 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_transform_points(double4* pts, uint32_t n, const double* pose_dev, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  transform_points_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, pose_dev);
  return cudaGetLastError();
}

cudaError_t launch_init_pairs(PairState* s, uint32_t n_pairs, const double* init_pose_or_null, uint32_t init_stride,
                              cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  init_pairs_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(s, n_pairs, init_pose_or_null, init_stride);
  return cudaGetLastError();
}

cudaError_t launch_finish_pairs(const PairState* s, uint32_t n_pairs, double* poses, int32_t* term, uint32_t* iters,
                                cudaStream_t st) {
  if (n_pairs == 0) return cudaSuccess;
  finish_pairs_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(s, n_pairs, poses, term, iters);
  return cudaGetLastError();
}

}  // namespace loamgpu
