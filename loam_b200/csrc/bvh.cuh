// GPU nearest-neighbour structure over one target feature set (replaces the nanoflann KD-tree of the
// reference: kdtree.h:24-41, kdtree.cpp:10-28, built at registration-inl.h:20-23).
//
//   K3  bvh_build_kernel  one CTA per set: bounding box -> 30-bit Morton codes -> in-CTA stable LSD radix sort
//                         (4-bit digits, counters in shared memory) -> Morton-ordered point copy -> binary radix
//                         tree over the sorted codes (every internal node finds its key range and split with
//                         count-leading-zeros binary searches, so all nodes are built in parallel) -> boxes
//                         bottom-up in barrier-separated passes (a node merges its children once both are ready)
//   K4  knn_bvh()         exact k-NN for one query per thread: near-child-first traversal with a short stack,
//                         subtrees of <= 8 points are scanned as leaves; float32 box lower bounds (boxes rounded
//                         outward, query rounded both ways, arithmetic rounded down => never above the true fp64
//                         distance), fp64 point distances in the reference's operation order, branch-free sorted
//                         insertion on (d2, index)
//
// Why a tree and not a uniform grid: LiDAR feature density falls with 1/range^2; a uniform cell sized for the
// average makes near-sensor queries scan hundreds of candidates (measured: profiles/r1_assoc_v1_hotlines.txt).
// Why a radix tree and not fixed 8-point runs of the Morton order: runs that straddle a jump of the Z curve get
// huge boxes; splitting at Morton-prefix boundaries visits ~3 leaves / ~18 points per query instead of ~10 / ~84
// (CPU simulation on a 64x1024 synthetic pair).
//
// Tie-break (documented, deterministic): equal squared distances resolve by ascending target index
// (nanoflann resolves them by tree-traversal order, i.e. unpinned in the reference).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace loamgpu {
namespace {

// ============================================================================ build (K3)

__device__ double block_reduce_minmax(double v, bool is_max, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, t) : fmin(v, t);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {  // warp 0 folds the per-warp partials, everyone reads the result
    double r = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : (is_max ? -CUDART_INF : CUDART_INF);
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, r, o);
      r = is_max ? fmax(r, t) : fmin(r, t);
    }
    if (threadIdx.x == 0) s_red[32] = r;
  }
  __syncthreads();
  return s_red[32];
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

constexpr int kRadixBits = 4;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kMortonBits = 30;

// -DBUILD_TIMING: set 0 prints the clock64 cycles of each phase (one line per launch)
#ifdef BUILD_TIMING
#define BT_MARK(i) \
  do {             \
    __syncthreads(); \
    if (threadIdx.x == 0) bt[i] = clock64(); \
  } while (0)
#else
#define BT_MARK(i)
#endif
__global__ void __launch_bounds__(kBuildThreads, BUILD_MINBLOCKS) bvh_build_kernel(BvhBuildArgs a) {
#ifdef BUILD_TIMING
  long long bt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  BT_MARK(0);
#endif
  // dynamic shared memory: [0, kRadixBins*kBuildThreads) radix counters during the sort; afterwards, when the set
  // fits (a.smem_tree), the sorted Morton codes (n words) and behind them one readiness byte per node
  extern __shared__ uint32_t s_cnt[];
  __shared__ double s_red[33];
  __shared__ uint32_t s_scan[kBuildThreads / 32];
  __shared__ uint32_t s_wtot[kRadixBins * (kBuildThreads / 32)];

  const uint32_t set = blockIdx.x;
  const uint32_t slot = (uint32_t)((a.slot0 + set) % a.n_slots);
  const uint32_t n = a.counts[slot * 2 + a.kind];
  const double4* pts = a.pts + (size_t)slot * a.pt_stride;
  BvhHdr* hdr = a.g.hdr + set;
  BvhNode* nodes = a.g.nodes + (size_t)set * a.g.pt_cap;
  double4* sorted = a.g.sorted + (size_t)set * a.g.pt_cap;
  uint2* keyA = a.g.keys + (size_t)set * 2 * a.g.pt_cap;
  uint2* keyB = keyA + a.g.pt_cap;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;

  if (tid == 0) {
    BvhHdr h;
    h.n = n;
    h.pad[0] = h.pad[1] = h.pad[2] = 0;
    *hdr = h;
    if (a.g.quant && n == 0) {
      BvhQuant q;
      q.org[0] = q.org[1] = q.org[2] = 0.0;
      q.inv_cell = q.inv_cell2 = 1.0;
      q.n_rec = 0;
      q.pad = 0;
      a.g.quant[set] = q;
    }
  }
  if (n == 0) return;

  // ---- bounding box
  double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
  for (uint32_t i = tid; i < n; i += nthr) {
    const double4 p = pts[i];
    lo[0] = fmin(lo[0], p.x);
    lo[1] = fmin(lo[1], p.y);
    lo[2] = fmin(lo[2], p.z);
    hi[0] = fmax(hi[0], p.x);
    hi[1] = fmax(hi[1], p.y);
    hi[2] = fmax(hi[2], p.z);
  }
  double emax = 0;
  for (int d = 0; d < 3; d++) {
    lo[d] = block_reduce_minmax(lo[d], false, s_red);
    hi[d] = block_reduce_minmax(hi[d], true, s_red);
    emax = fmax(emax, hi[d] - lo[d]);
  }
  // one scale for all axes (cubic cells); the ordering only has to be spatially coherent, not exact
  const double scale = emax > 0 ? 1023.999 / emax : 0.0;
  // grid of the compact records: origin a margin below the box (float-rounded box corners stay inside),
  // kRecCells cells across the longest axis plus the margins
  if (tid == 0 && a.g.quant) {
    double amax = emax;
    for (int d = 0; d < 3; d++) amax = fmax(amax, fmax(fabs(lo[d]), fabs(hi[d])));
    const double qmargin = 1e-6 * amax + 1e-9;
    const double qinv = kRecCells / (emax + 2.0 * qmargin);
    BvhQuant q;
    q.org[0] = lo[0] - qmargin;
    q.org[1] = lo[1] - qmargin;
    q.org[2] = lo[2] - qmargin;
    q.inv_cell = qinv;
    q.inv_cell2 = qinv * qinv * (1.0 + 1e-12);
    q.n_rec = 0;  // (sets of at most kBvhLeaf points keep 0; larger ones get their count at the end)
    q.pad = 0;
    a.g.quant[set] = q;
  }

  BT_MARK(1);
  // ---- Morton keys (thread-contiguous chunks: the radix passes below need a fixed item -> thread map)
  const uint32_t chunk = (n + nthr - 1) / nthr;
  const uint32_t c0 = min(tid * chunk, n), c1 = min(c0 + chunk, n);
  for (uint32_t i = c0; i < c1; i++) {
    const double4 p = pts[i];
    const uint32_t ix = (uint32_t)fmin(fmax((p.x - lo[0]) * scale, 0.0), 1023.0);
    const uint32_t iy = (uint32_t)fmin(fmax((p.y - lo[1]) * scale, 0.0), 1023.0);
    const uint32_t iz = (uint32_t)fmin(fmax((p.z - lo[2]) * scale, 0.0), 1023.0);
    keyA[i] = make_uint2(spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2), i);
  }
  __syncthreads();

  BT_MARK(2);
  // ---- stable LSD radix sort, 4 bits per pass.  Rank of an item = items with a smaller digit + items with the
  // same digit owned by earlier threads + earlier items of the same digit in this thread's own chunk.
  uint2* src = keyA;
  uint2* dst = keyB;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  // Leaves stop at 8 points, so a set of n points needs only ~log2(n)/2 + 1 bits per axis to separate them (CPU
  // simulation, DESIGN.md §5: 7 bits/axis give the same tree quality as 10 on a 14k-point set); fewer bits = fewer passes.
  int bits_axis = 6;
  while (bits_axis < 10 && (1u << (2 * (bits_axis - 1))) < n) bits_axis++;
  const int key_shift = 3 * (10 - bits_axis);  // drop the finest levels of the 30-bit code
  const int sort_bits = 3 * bits_axis;
  for (int shift = key_shift; shift < key_shift + sort_bits; shift += kRadixBits) {
#pragma unroll
    for (int b = 0; b < kRadixBins; b++) s_cnt[b * kBuildThreads + tid] = 0;
    for (uint32_t i = c0; i < c1; i++) s_cnt[((src[i].x >> shift) & (kRadixBins - 1)) * kBuildThreads + tid]++;
    // per-bin inclusive scan across the warp's 32 threads; warp totals -> s_wtot[bin][warp]
    uint32_t cnt[kRadixBins], inc[kRadixBins];
#pragma unroll
    for (int b = 0; b < kRadixBins; b++) {
      cnt[b] = s_cnt[b * kBuildThreads + tid];
      uint32_t v = cnt[b];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += t;
      }
      inc[b] = v;
      if (lane == 31) s_wtot[b * (kBuildThreads / 32) + warp] = v;
    }
    __syncthreads();
    // exclusive scan of the 16 x 32 warp totals (flattened bin-major) by the first 512 threads
    if (tid < kRadixBins * (kBuildThreads / 32)) {
      const uint32_t mine = s_wtot[tid];
      uint32_t v = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += t;
      }
      if (lane == 31) s_scan[warp] = v;
      s_wtot[tid] = v - mine;  // exclusive within this group of 32
    }
    __syncthreads();
    if (tid < kRadixBins * (kBuildThreads / 32)) {
      uint32_t off = 0;
      for (uint32_t w = 0; w < warp; w++) off += s_scan[w];
      s_wtot[tid] += off;
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kRadixBins; b++)
      s_cnt[b * kBuildThreads + tid] = s_wtot[b * (kBuildThreads / 32) + warp] + inc[b] - cnt[b];
    // (each thread reads back only its own column below, so no barrier is needed here)
    for (uint32_t i = c0; i < c1; i++) {
      const uint2 kv = src[i];
      const uint32_t pos = s_cnt[((kv.x >> shift) & (kRadixBins - 1)) * kBuildThreads + tid]++;
      dst[pos] = kv;
    }
    __syncthreads();
    uint2* t = src;
    src = dst;
    dst = t;
  }

  BT_MARK(3);
  // ---- Morton-ordered point copy
  for (uint32_t i = tid; i < n; i += nthr) {
    const uint32_t id = src[i].y;
    const double4 p = pts[id];
    sorted[i] = make_double4(p.x, p.y, p.z, __longlong_as_double((long long)id));
  }
  __syncthreads();

  if (n < 2) return;  // a single point has no internal node; knn_bvh scans it directly

  BT_MARK(4);
  const bool in_smem = a.smem_tree != 0;
  uint32_t* s_codes = s_cnt;
  uint8_t* s_ready = reinterpret_cast<uint8_t*>(s_cnt + max((uint32_t)(kRadixBins * kBuildThreads), a.g.pt_cap));
  if (in_smem) {
    for (uint32_t i = tid; i < n; i += nthr) {
      s_codes[i] = src[i].x >> key_shift;
      s_ready[i] = 0;
    }
    __syncthreads();
  }

  // ---- binary radix tree over the sorted (code, position) keys.  Internal node i covers the key range [first, last]
  // with i == first or i == last; its children are node/leaf `split` (range [first, split]) and `split + 1`
  // (range [split + 1, last]).
  const uint2* key = src;
  int* arrived = a.g.aux + (size_t)set * a.g.pt_cap;  // per internal node: 0 pending, 2 merged this pass, 1 ready
  auto delta = [&](int i, int j) -> int {  // common-prefix length of keys i and j, -1 outside the array
    if (j < 0 || j >= (int)n) return -1;
    const uint32_t ci = in_smem ? s_codes[i] : key[i].x >> key_shift, cj = in_smem ? s_codes[j] : key[j].x >> key_shift;
    return ci != cj ? __clz(ci ^ cj) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
  };
  for (uint32_t t = tid; t + 1 < n; t += nthr) {
    const int i = (int)t;
    const int d = delta(i, i + 1) - delta(i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(i, i - d);
    int lmax = 2;
    while (delta(i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int st = lmax >> 1; st >= 1; st >>= 1)
      if (delta(i, i + (l + st) * d) > dmin) l += st;
    const int j = i + l * d;
    const int dnode = delta(i, j);
    int sp = 0;
    for (int div = 2;; div <<= 1) {
      const int st = (l + div - 1) / div;
      if (delta(i, i + (sp + st) * d) > dnode) sp += st;
      if (st <= 1) break;
    }
    const int split = i + sp * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    uint32_t w = (uint32_t)split;
    if (first == split) w |= kLeftLeaf;
    if (last == split + 1) w |= kRightLeaf;
    nodes[i].split = w;
    nodes[i].pad = (uint32_t)j;  // the other end of the node's key range [min(i, j), max(i, j)]
    arrived[i] = 0;
  }
  __syncthreads();

  BT_MARK(5);
  // ---- boxes, bottom-up in passes: a node is merged from its children once both were merged in an EARLIER pass
  // (ready[] holds the 1-based pass in which a node was merged, 0 = not yet), so one barrier per pass separates a
  // box from its readers.  A thread tracks its own pending nodes in a register bit mask and only revisits those.
  // The number of passes is the tree height (~log2 n for scattered points, at most 62).
  int* ready = arrived;
  auto merged_pass = [&](uint32_t i) -> uint32_t { return in_smem ? (uint32_t)s_ready[i] : (uint32_t)__ldcg(ready + i); };
  auto try_merge = [&](uint32_t i, uint32_t pass) -> bool {  // true: node i got its box in this pass
    const uint32_t w = nodes[i].split;  // written by this same thread above
    const uint32_t sp = w & kSplitMask;
    // flags and boxes written by other threads of this CTA are read with ld.cg (L2), never from a possibly stale L1 line
    if (!(w & kLeftLeaf)) {
      const uint32_t m = merged_pass(sp);
      if (m == 0 || m >= pass) return false;
    }
    if (!(w & kRightLeaf)) {
      const uint32_t m = merged_pass(sp + 1);
      if (m == 0 || m >= pass) return false;
    }
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      float clo[3], chi[3];
      if (w & (c == 0 ? kLeftLeaf : kRightLeaf)) {
        const double4 pt = sorted[sp + c];
        clo[0] = __double2float_rd(pt.x); chi[0] = __double2float_ru(pt.x);
        clo[1] = __double2float_rd(pt.y); chi[1] = __double2float_ru(pt.y);
        clo[2] = __double2float_rd(pt.z); chi[2] = __double2float_ru(pt.z);
      } else {
        const float4* f = reinterpret_cast<const float4*>(nodes + sp + c);
        const float4 va = __ldcg(f), vb = __ldcg(f + 1);
        clo[0] = va.x; clo[1] = va.y; clo[2] = va.z;
        chi[0] = vb.x; chi[1] = vb.y; chi[2] = vb.z;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        lo[k] = c == 0 ? clo[k] : fminf(lo[k], clo[k]);
        hi[k] = c == 0 ? chi[k] : fmaxf(hi[k], chi[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      nodes[i].lo[k] = lo[k];
      nodes[i].hi[k] = hi[k];
    }
    if (in_smem)
      s_ready[i] = (uint8_t)pass;
    else
      __stcg(ready + i, (int)pass);
    return true;
  };
  const uint32_t n_int = n - 1;
  const uint32_t mine = tid < n_int ? (n_int - tid + nthr - 1) / nthr : 0u;  // nodes tid, tid + nthr, ...
  if (mine <= 64) {
    unsigned long long pend = mine == 64 ? ~0ull : ((1ull << mine) - 1ull);
    for (uint32_t pass = 1;; pass++) {
      unsigned long long m = pend;
      while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (try_merge(tid + (uint32_t)k * nthr, pass)) pend &= ~(1ull << k);
      }
      if (!__syncthreads_or(pend != 0)) break;
    }
  } else {  // very large sets in one CTA (the multi-CTA build normally takes those): rescan
    for (uint32_t pass = 1;; pass++) {
      bool pending = false;
      for (uint32_t i = tid; i < n_int; i += nthr) {
        if (merged_pass(i) != 0) continue;
        if (!try_merge(i, pass)) pending = true;
      }
      if (!__syncthreads_or(pending)) break;
    }
  }

  // ---- compact traversal records (common.cuh: BvhRec).  Internal node i is "big" when it covers more than kBvhLeaf
  // points; big nodes are numbered in index order (the root, node 0, gets record 0) and each writes one record with its
  // children's boxes on the records' grid.  The records overlay this set's sort scratch, which is dead by now.
  BT_MARK(6);
  if (a.g.quant == nullptr || n <= (uint32_t)kBvhLeaf) return;
  __syncthreads();
  uint32_t* cid = reinterpret_cast<uint32_t*>(arrived);  // readiness flags are dead: record number per big node
  BvhRec* recs = reinterpret_cast<BvhRec*>(keyA);
  const uint32_t rec_cap = a.g.pt_cap / 2;
  const uint32_t per = (n_int + nthr - 1) / nthr;
  const uint32_t b0 = min(tid * per, n_int), b1 = min(b0 + per, n_int);
  auto node_range = [&](uint32_t i, uint32_t& first, uint32_t& last) {
    const uint32_t other = __ldcg(&nodes[i].pad);
    first = min(i, other);
    last = max(i, other);
  };
  uint32_t cnt_big = 0;
  for (uint32_t i = b0; i < b1; i++) {
    uint32_t f, l;
    node_range(i, f, l);
    cnt_big += (l - f >= (uint32_t)kBvhLeaf) ? 1u : 0u;
  }
  uint32_t incl = cnt_big;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  uint32_t base = 0, total = 0;
  for (uint32_t w = 0; w < (nthr >> 5); w++) {
    if (w < warp) base += s_scan[w];
    total += s_scan[w];
  }
  if (total > rec_cap) {  // degenerate tree (long chains): this set keeps the general traversal only
    if (tid == 0) a.g.quant[set].n_rec = kNoRecs;
    return;
  }
  uint32_t run = base + incl - cnt_big;
  for (uint32_t i = b0; i < b1; i++) {
    uint32_t f, l;
    node_range(i, f, l);
    if (l - f >= (uint32_t)kBvhLeaf) __stcg(cid + i, run++);
  }
  __syncthreads();
  // (the grid was written by thread 0 before the sort; barriers since)
  const double qorg[3] = {__ldcg(&a.g.quant[set].org[0]), __ldcg(&a.g.quant[set].org[1]), __ldcg(&a.g.quant[set].org[2])};
  const double qinv = __ldcg(&a.g.quant[set].inv_cell);
  auto quant_lo = [&](float v, int d) -> uint32_t {
    const double t = ((double)v - qorg[d]) * qinv;
    return (uint32_t)fmin(fmax(floor(t - 1e-6), 0.0), (double)kRecCellMax);
  };
  auto quant_hi = [&](float v, int d) -> uint32_t {
    const double t = ((double)v - qorg[d]) * qinv;
    return (uint32_t)fmin(fmax(ceil(t + 1e-6), 0.0), (double)kRecCellMax);
  };
  for (uint32_t i = b0; i < b1; i++) {
    uint32_t f, l;
    node_range(i, f, l);
    if (l - f < (uint32_t)kBvhLeaf) continue;
    const uint32_t sp = __ldcg(&nodes[i].split) & kSplitMask;
    uint32_t cbox[2][3], cref[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint32_t cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
      float clo[3], chi[3];
      if (cf == cl) {  // a single point has no node record: its own coordinates, rounded outward
        const double4 pt = sorted[cf];
        clo[0] = __double2float_rd(pt.x); chi[0] = __double2float_ru(pt.x);
        clo[1] = __double2float_rd(pt.y); chi[1] = __double2float_ru(pt.y);
        clo[2] = __double2float_rd(pt.z); chi[2] = __double2float_ru(pt.z);
      } else {
        const float4* q = reinterpret_cast<const float4*>(nodes + sp + c);
        const float4 va = __ldcg(q), vb = __ldcg(q + 1);
        clo[0] = va.x; clo[1] = va.y; clo[2] = va.z;
        chi[0] = vb.x; chi[1] = vb.y; chi[2] = vb.z;
      }
#pragma unroll
      for (int d = 0; d < 3; d++) cbox[c][d] = quant_lo(clo[d], d) | (quant_hi(chi[d], d) << 16);
      cref[c] = (cl - cf < (uint32_t)kBvhLeaf) ? (kRefLeaf | ((cl - cf) << 24) | cf) : __ldcg(cid + sp + c);
    }
    uint32_t w6[6];
    rec_pack_boxes(cbox[0], cbox[1], w6);
    uint4* dst = reinterpret_cast<uint4*>(recs + __ldcg(cid + i));
    dst[0] = make_uint4(w6[0], w6[1], w6[2], w6[3]);
    dst[1] = make_uint4(w6[4], w6[5], cref[0], cref[1]);
  }
  if (tid == 0) a.g.quant[set].n_rec = total;
#ifdef BUILD_TIMING
  BT_MARK(7);
  if (tid == 0 && set == 0)
    printf("BT n %u: bbox %lld morton %lld sort %lld gather %lld codes+topology %lld boxes %lld compact %lld total %lld\n", n,
           bt[1] - bt[0], bt[2] - bt[1], bt[3] - bt[2], bt[4] - bt[3], bt[5] - bt[4], bt[6] - bt[5], bt[7] - bt[6],
           bt[7] - bt[0]);
#endif
}

// ============================================================================ build, shared-memory version (K3)
// The same structure (Morton-sorted copy, node split words + boxes for the general walk, compact records for the batched
// walk) for sets of up to kSbMaxPoints points, with every intermediate kept in SHARED memory.  Measured on the first
// kernel (-DBUILD_TIMING, 14.3 k points, 763 k cycles): sort 28 % (4-bit passes, (key, index) pairs ping-ponged through
// global memory), boxes 28 % (~28 level-synchronous passes over all n nodes with L2 round trips), topology 17 %,
// compaction 13 %.  Here:
//   * codes stay in shared memory indexed by point; the sort permutes 16-bit point numbers only, 8 bits per pass
//     (3 passes for 24-bit codes): a warp owns a contiguous block of the current order and ranks 32 keys at a time with
//     __match_any_sync, per-(digit, warp) counters in shared memory, one block-wide scan per pass — stable, so the
//     order equals the first kernel's;
//   * node ranges / splits live in shared memory; nodes INSIDE a subtree of at most kBvhLeaf points get neither a
//     split nor a box (no walk ever reads them: such subtrees are scanned as leaves);
//   * boxes are merged bottom-up over the BIG nodes only (~n / 4.6 of them, ~14 levels instead of ~28): a big node
//     whose child is a leaf-sized subtree takes that child's box straight from its <= kBvhLeaf points (and stores it
//     for the general walk), big children are waited for pass by pass with readiness bytes in shared memory;
//   * the compact records are written in the same sweep.
#ifndef BUILD_L2HINT
#define BUILD_L2HINT 0
#endif
#ifndef BUILD_GATHER_UNROLL
#define BUILD_GATHER_UNROLL 1
#endif
// point loads with an L2 eviction policy (-DBUILD_L2HINT=1): the build reads its points three times with ~100 k cycles
// in between while 148 CTAs stream ~1 MB each through the L2
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
  uint64_t p;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double4 ld_point(const double4* p, uint64_t pol) {
#if BUILD_L2HINT
  double4 v;
  asm("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  asm("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2+16], %3;" : "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
  return v;
#else
  return *p;
#endif
}
constexpr int kSbThreads = 1024;
constexpr int kSbMaxPer = 20;                               // sorted positions a thread carries in registers
constexpr uint32_t kSbMaxPoints = kSbMaxPer * kSbThreads;   // 20,480
__host__ __device__ inline size_t sb_smem_bytes(uint32_t cap) {
  const size_t capA = (cap + 15) & ~(size_t)15;
  // codes (later: node -> number) | two 16-bit arrays (order; later node ranges, splits) | sort counters (later: the
  // list of big nodes + their readiness bytes)
  return 8 * capA + (3 * capA > 32768 ? 3 * capA : 32768) + 64;
}

// grid (sets, kinds): row 0 builds `a0`'s sets, row 1 (when launched) `a1`'s — the edge and the planar sets of the same
// scans in one launch (short edge CTAs fill in around the long planar ones; one launch less per chunk / call)
__global__ void __launch_bounds__(kSbThreads, 1) bvh_build_smem_kernel(BvhBuildArgs a0, BvhBuildArgs a1, uint32_t smem_bytes) {
  const BvhBuildArgs& a = blockIdx.y ? a1 : a0;
  extern __shared__ __align__(16) unsigned char sb_smem[];
  __shared__ double s_red[33];
  __shared__ uint32_t s_scan[kSbThreads / 32];
  const uint32_t set = blockIdx.x;
  const uint32_t slot = (uint32_t)((a.slot0 + set) % a.n_slots);
  const uint32_t n = a.counts[slot * 2 + a.kind];
  const double4* pts = a.pts + (size_t)slot * a.pt_stride;
  BvhNode* nodes = a.g.nodes + (size_t)set * a.g.pt_cap;
  double4* sorted = a.g.sorted + (size_t)set * a.g.pt_cap;
  const uint32_t tid = threadIdx.x, nthr = kSbThreads, lane = tid & 31, warp = tid >> 5;
  // (arrays are laid out for the set's own size, not its capacity: what is left of the launch's shared memory holds the
  // boxes of the big nodes, below)
  const size_t capA = ((size_t)n + 15) & ~(size_t)15;
  uint32_t* s_code = reinterpret_cast<uint32_t*>(sb_smem);              // [capA] codes: by point, after the sort by position
  uint16_t* s_p0 = reinterpret_cast<uint16_t*>(s_code + capA);          // [capA] sort ping / node: other end of its range
  uint16_t* s_p1 = s_p0 + capA;                                         // [capA] sort pong / node: split position
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_p1 + capA);          // [32 warps][256 digits] sort counters
#ifdef BUILD_TIMING
  long long bt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  BT_MARK(0);
#endif
  if (tid == 0) {
    BvhHdr h;
    h.n = n;
    h.pad[0] = h.pad[1] = h.pad[2] = 0;
    a.g.hdr[set] = h;
    if (a.g.quant && n == 0) {
      BvhQuant q;
      q.org[0] = q.org[1] = q.org[2] = 0.0;
      q.inv_cell = q.inv_cell2 = 1.0;
      q.n_rec = 0;
      q.pad = 0;
      a.g.quant[set] = q;
    }
  }
  if (n == 0) return;

  // ---- bounding box, Morton scale, grid of the compact records (as in bvh_build_kernel)
  double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
  const uint64_t pol_keep = l2_policy(true), pol_last = l2_policy(false);
#pragma unroll 4
  for (uint32_t i = tid; i < n; i += nthr) {
    const double4 p = ld_point(pts + i, pol_keep);
    lo[0] = fmin(lo[0], p.x);
    lo[1] = fmin(lo[1], p.y);
    lo[2] = fmin(lo[2], p.z);
    hi[0] = fmax(hi[0], p.x);
    hi[1] = fmax(hi[1], p.y);
    hi[2] = fmax(hi[2], p.z);
  }
  double emax = 0;
  for (int d = 0; d < 3; d++) {
    lo[d] = block_reduce_minmax(lo[d], false, s_red);
    hi[d] = block_reduce_minmax(hi[d], true, s_red);
    emax = fmax(emax, hi[d] - lo[d]);
  }
  const double scale = emax > 0 ? 1023.999 / emax : 0.0;
  if (tid == 0 && a.g.quant) {
    double amax = emax;
    for (int d = 0; d < 3; d++) amax = fmax(amax, fmax(fabs(lo[d]), fabs(hi[d])));
    const double qmargin = 1e-6 * amax + 1e-9;
    const double qinv = kRecCells / (emax + 2.0 * qmargin);
    BvhQuant q;
    q.org[0] = lo[0] - qmargin;
    q.org[1] = lo[1] - qmargin;
    q.org[2] = lo[2] - qmargin;
    q.inv_cell = qinv;
    q.inv_cell2 = qinv * qinv * (1.0 + 1e-12);
    q.n_rec = 0;
    q.pad = 0;
    a.g.quant[set] = q;
  }
  BT_MARK(1);

  // ---- Morton codes (only as many bits as a set of n points needs, see bvh_build_kernel), identity order
  int bits_axis = 6;
  while (bits_axis < 10 && (1u << (2 * (bits_axis - 1))) < n) bits_axis++;
  const int key_shift = 3 * (10 - bits_axis);
  const int sort_bits = 3 * bits_axis;
#pragma unroll 4
  for (uint32_t i = tid; i < n; i += nthr) {
    const double4 p = ld_point(pts + i, pol_keep);
    const uint32_t ix = (uint32_t)fmin(fmax((p.x - lo[0]) * scale, 0.0), 1023.0);
    const uint32_t iy = (uint32_t)fmin(fmax((p.y - lo[1]) * scale, 0.0), 1023.0);
    const uint32_t iz = (uint32_t)fmin(fmax((p.z - lo[2]) * scale, 0.0), 1023.0);
    s_code[i] = (spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2)) >> key_shift;
    s_p0[i] = (uint16_t)i;
  }
  __syncthreads();
  BT_MARK(2);

  // ---- stable LSD radix sort of the order, 8 bits per pass.  Counters are laid out [warp][digit] (a warp's group
  // leaders hit different banks); lanes with the same digit find each other with 8 ballots (__match_any_sync measured
  // ~10x slower here: its latency grows with the number of distinct values, up to 32 per batch of 8-bit digits).
  {
    uint16_t* src = s_p0;
    uint16_t* dst = s_p1;
    const uint32_t blk = (((n + 31) / 32) + 31) & ~31u;  // positions per warp: whole batches of 32
    const uint32_t w_lo = min(warp * blk, n), w_hi = min(w_lo + blk, n);
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t* my_hist = s_hist + warp * 256;
    auto same_digit = [&](uint32_t dg, bool in) -> unsigned {  // lanes of this batch whose digit equals mine
      unsigned peers = __ballot_sync(0xffffffffu, in);
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const bool bit = (dg >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
      }
      return peers;
    };
    for (int shift = 0; shift < sort_bits; shift += 8) {
#pragma unroll
      for (int j = 0; j < 2; j++) reinterpret_cast<uint4*>(s_hist)[tid + j * kSbThreads] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      // counting needs no order: shared-memory atomics (the kernel is issue-bound here: 14 k keys x instructions per key)
      for (uint32_t pos = w_lo + lane; pos < w_hi; pos += 32) atomicAdd(&my_hist[(s_code[src[pos]] >> shift) & 255u], 1u);
      __syncthreads();
      {  // exclusive scan of the counters in (digit, warp) order: thread t takes digit t / 4, warps 8 (t % 4) .. + 7
        const uint32_t dgt = tid >> 2, w0 = (tid & 3u) * 8u;
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          c[j] = s_hist[(w0 + j) * 256 + dgt];
          tot += c[j];
        }
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if ((int)lane >= o) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t run = inc - tot;
        for (uint32_t w = 0; w < warp; w++) run += s_scan[w];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s_hist[(w0 + j) * 256 + dgt] = run;
          run += c[j];
        }
      }
      __syncthreads();
      // scattering must be stable: rank inside the batch by lane order among equal digits, batches in order
      for (uint32_t b = w_lo; b < w_hi; b += 32) {
        const uint32_t pos = b + lane;
        const bool in = pos < w_hi;
        const uint16_t e = in ? src[pos] : (uint16_t)0;
        const uint32_t dg = in ? (s_code[e] >> shift) & 255u : 0u;
        const unsigned peers = same_digit(dg, in);
        if (in) dst[my_hist[dg] + (uint32_t)__popc(peers & lt_mask)] = e;
        __syncwarp();  // every lane has read its group's base
        if (in && (peers & lt_mask) == 0u) my_hist[dg] += (uint32_t)__popc(peers);
        __syncwarp();
      }
      __syncthreads();
      uint16_t* t = src;
      src = dst;
      dst = t;
    }
    if (src != s_p0) {  // final order into s_p0
      for (uint32_t i = tid; i < n; i += nthr) s_p0[i] = src[i];
      __syncthreads();
    }
  }
  BT_MARK(3);

  // ---- Morton-ordered point copy; codes permuted in place (through registers) into sorted order
  {
    uint32_t c[kSbMaxPer];
#pragma unroll
    for (int j = 0; j < kSbMaxPer; j++) {
      const uint32_t pos = tid + (uint32_t)j * nthr;
      c[j] = pos < n ? s_code[s_p0[pos]] : 0u;
    }
    for (uint32_t pos0 = tid; pos0 < n; pos0 += BUILD_GATHER_UNROLL * nthr) {  // independent gathers in flight per thread
      double4 p[BUILD_GATHER_UNROLL];
      uint32_t id[BUILD_GATHER_UNROLL];
#pragma unroll
      for (int u = 0; u < BUILD_GATHER_UNROLL; u++) {
        const uint32_t pos = pos0 + (uint32_t)u * nthr;
        id[u] = s_p0[min(pos, n - 1)];
        p[u] = ld_point(pts + id[u], pol_last);
      }
#pragma unroll
      for (int u = 0; u < BUILD_GATHER_UNROLL; u++) {
        const uint32_t pos = pos0 + (uint32_t)u * nthr;
        if (pos < n) sorted[pos] = make_double4(p[u].x, p[u].y, p[u].z, __longlong_as_double((long long)id[u]));
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSbMaxPer; j++) {
      const uint32_t pos = tid + (uint32_t)j * nthr;
      if (pos < n) s_code[pos] = c[j];
    }
  }
  __syncthreads();
  BT_MARK(4);
  if (n <= (uint32_t)kBvhLeaf) return;  // scanned directly by every walk: no nodes, no records (n_rec stays 0)

  // ---- binary radix tree (Karras).  Nodes of at most kBvhLeaf points need no split and no box (no walk expands them),
  // and they are 4 of 5 nodes: a first sweep over ALL nodes only finds each node's range, giving up after three probes
  // once the range is known to exceed kBvhLeaf points (a warp pays for its slowest lane: with the full searches in this
  // sweep every warp paid for its biggest node, 118 k cycles on 14.3 k points).  The big nodes are then numbered and
  // finish their searches densely packed, one thread per big node.
  uint16_t* s_other = s_p0;  // (the order is dead: both 16-bit arrays now describe nodes) other end of the range
  uint16_t* s_split = s_p1;  // split position (big nodes)
  uint16_t* s_big = reinterpret_cast<uint16_t*>(s_hist);     // (the counters are dead) number -> node   [capA]
  uint8_t* s_ready = reinterpret_cast<uint8_t*>(s_big + capA);  // per number: pass in which its box was completed
  constexpr uint16_t kBigPending = 0xFFFFu;
  auto delta = [&](int i, int j) -> int {  // common-prefix length of keys i and j, -1 outside the array
    if (j < 0 || j >= (int)n) return -1;
    const uint32_t ci = s_code[i], cj = s_code[j];
    return ci != cj ? __clz(ci ^ cj) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
  };
  const uint32_t n_int = n - 1;
  for (uint32_t t = tid; t < n_int; t += nthr) {
    const int i = (int)t;
    const int d = delta(i, i + 1) - delta(i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(i, i - d);
    int lmax = 2;
    while (lmax < 2 * kBvhLeaf && delta(i, i + lmax * d) > dmin) lmax <<= 1;
    uint16_t other = kBigPending;
    if (lmax < 2 * kBvhLeaf) {  // the range ends within lmax - 1 < kBvhLeaf steps: a leaf-sized node
      int l = 0;
      for (int st = lmax >> 1; st >= 1; st >>= 1)
        if (delta(i, i + (l + st) * d) > dmin) l += st;
      if (l < kBvhLeaf) other = (uint16_t)(i + l * d);  // (else: big after all — leaf sizes that are no power of two)
    }
    s_other[i] = other;
  }
  __syncthreads();

  BT_MARK(9);
  // ---- number the big nodes in index order (the root, node 0, is number 0) and list them
  const uint32_t per = (n_int + nthr - 1) / nthr;
  const uint32_t b0 = min(tid * per, n_int), b1 = min(b0 + per, n_int);
  uint32_t cnt_big = 0;
  for (uint32_t i = b0; i < b1; i++) cnt_big += s_other[i] == kBigPending ? 1u : 0u;
  uint32_t incl = cnt_big;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  uint32_t base = 0, n_big = 0;
  for (uint32_t w = 0; w < (nthr >> 5); w++) {
    if (w < warp) base += s_scan[w];
    n_big += s_scan[w];
  }
  {
    uint32_t run = base + incl - cnt_big;
    for (uint32_t i = b0; i < b1; i++)
      if (s_other[i] == kBigPending) s_big[run++] = (uint16_t)i;
  }
  __syncthreads();

  // Boxes of the big nodes in shared memory (below) when the launch's shared memory has room for 12 bytes per big node
  // behind the arrays above and the set gets compact records at all; else the level passes through the node records.
  const size_t box_off = 8 * capA + (3 * capA > 32768 ? 3 * capA : 32768);
  const bool smem_boxes = a.g.quant != nullptr && n_big <= a.g.pt_cap / 2 && n_big <= (uint32_t)(capA / 2) &&
                          box_off + 12 * (size_t)n_big <= smem_bytes;

  // ---- full range + split searches of the big nodes, one thread per node
  for (uint32_t num = tid; num < n_big; num += nthr) {
    const int i = (int)s_big[num];
    const int d = delta(i, i + 1) - delta(i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(i, i - d);
    int lmax = 2;
    while (delta(i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int st = lmax >> 1; st >= 1; st >>= 1)
      if (delta(i, i + (l + st) * d) > dmin) l += st;
    const int j = i + l * d;
    const int dnode = delta(i, j);
    int sp = 0;
    for (int div = 2;; div <<= 1) {
      const int st = (l + div - 1) / div;
      if (delta(i, i + (sp + st) * d) > dnode) sp += st;
      if (st <= 1) break;
    }
    const int split = i + sp * d + min(d, 0);
    uint32_t w = (uint32_t)split;
    if (min(i, j) == split) w |= kLeftLeaf;
    if (max(i, j) == split + 1) w |= kRightLeaf;
    s_other[i] = (uint16_t)j;
    s_split[i] = (uint16_t)split;
    if (!smem_boxes) {  // (the shared-memory box path writes whole node records at its end)
      nodes[i].split = w;
      nodes[i].pad = (uint32_t)j;
    }
  }
  __syncthreads();
  BT_MARK(5);

  // ---- boxes + compact records, shared-memory version.  The level passes of the version below cost an L2 round trip
  // each (store, barrier, load: ~4.6 k cycles x 20 passes on 14.3 k points, as much as the sort) and the record pass
  // reads every child box back.  A box on the records' grid (two 16-bit cell numbers per axis) is 12 bytes: the ~n / 5 big nodes' boxes fit behind
  // the other arrays, the passes run on shared memory alone, and the records come out bit-identical (quantisation is
  // monotonic: the union of quantised boxes is the quantised union).  The float boxes the general walk reads are the
  // grid boxes converted back (outward): at most a cell (extent / 32766) wider than before — boxes only ever steer the
  // walk, they never decide a result (§5).
  if (smem_boxes) {
    uint16_t* s_cid = reinterpret_cast<uint16_t*>(s_code);  // (codes are dead) node -> number
    uint32_t* s_box = reinterpret_cast<uint32_t*>(sb_smem + box_off);  // [number][axis] lo | hi << 16
    for (uint32_t num = tid; num < n_big; num += nthr) s_cid[s_big[num]] = (uint16_t)num;
    BvhRec* recs = reinterpret_cast<BvhRec*>(a.g.keys + (size_t)set * 2 * a.g.pt_cap);
    const double qorg[3] = {__ldcg(&a.g.quant[set].org[0]), __ldcg(&a.g.quant[set].org[1]), __ldcg(&a.g.quant[set].org[2])};
    const double qinv = __ldcg(&a.g.quant[set].inv_cell);
    const double qcell = 1.0 / qinv;
    __syncthreads();
    // Per big node (by number): parent's number, sibling's number (0xFFFF: the sibling is leaf-sized) and how many big
    // children it still waits for.  (s_need overlays the upper half of the dead codes, s_parent / s_sib the readiness
    // bytes and the tail of the node list: n_big <= n / 2.)
    uint32_t* s_need = reinterpret_cast<uint32_t*>(s_cid + capA);        // [n_big]
    uint16_t* s_parent = reinterpret_cast<uint16_t*>(s_ready);           // [n_big]
    uint16_t* s_sib = s_big + (capA >> 1);                               // [n_big]
    const uint32_t mine = tid < n_big ? (n_big - tid + nthr - 1) / nthr : 0u;  // (at most 20: n_big < n <= 20,480)
    uint32_t done = 0;
    for (uint32_t k = 0; k < mine; k++) {  // pass 1: leaf-sized children straight from their points
      const uint32_t num = tid + k * nthr;
      const uint32_t i = s_big[num];
      const uint32_t o = s_other[i], sp = s_split[i];
      const uint32_t f = min(i, o), l = max(i, o);
      uint32_t acc[3] = {0x0000FFFFu, 0x0000FFFFu, 0x0000FFFFu};  // empty: lo above every cell, hi 0
      uint32_t cbox[2][3] = {{0, 0, 0}, {0, 0, 0}}, ref[2];
      uint32_t n_bigc = 0;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
        if (cl - cf < (uint32_t)kBvhLeaf) {
          double plo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, phi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
#pragma unroll
          for (int q = 0; q < kBvhLeaf; q++) {
            const double4 pt = sorted[min(cf + (uint32_t)q, cl)];
            plo[0] = fmin(plo[0], pt.x); phi[0] = fmax(phi[0], pt.x);
            plo[1] = fmin(plo[1], pt.y); phi[1] = fmax(phi[1], pt.y);
            plo[2] = fmin(plo[2], pt.z); phi[2] = fmax(phi[2], pt.z);
          }
          float blo[3], bhi[3];
#pragma unroll
          for (int d = 0; d < 3; d++) {
            blo[d] = __double2float_rd(plo[d]);
            bhi[d] = __double2float_ru(phi[d]);
            const double tl = ((double)blo[d] - qorg[d]) * qinv, th = ((double)bhi[d] - qorg[d]) * qinv;
            const uint32_t ql = (uint32_t)fmin(fmax(floor(tl - 1e-6), 0.0), (double)kRecCellMax);
            const uint32_t qh = (uint32_t)fmin(fmax(ceil(th + 1e-6), 0.0), (double)kRecCellMax);
            cbox[c][d] = ql | (qh << 16);
            acc[d] = min(acc[d] & 0xFFFFu, ql) | (max(acc[d] >> 16, qh) << 16);
          }
          if (cf != cl) {  // its node record, for the general walk (split word 0: never expanded)
            float4* q = reinterpret_cast<float4*>(nodes + sp + c);
            q[0] = make_float4(blo[0], blo[1], blo[2], 0.f);
            q[1] = make_float4(bhi[0], bhi[1], bhi[2], 0.f);
          }
          ref[c] = kRefLeaf | ((cl - cf) << 24) | cf;
        } else {
          ref[c] = (uint32_t)s_cid[sp + c];
          n_bigc++;
        }
      }
      uint32_t box[6];
      rec_pack_boxes(cbox[0], cbox[1], box);
      uint4* dst = reinterpret_cast<uint4*>(recs + num);  // (the halves of big children are filled in at the end)
      dst[0] = make_uint4(box[0], box[1], box[2], box[3]);
      dst[1] = make_uint4(box[4], box[5], ref[0], ref[1]);
#pragma unroll
      for (int d = 0; d < 3; d++) s_box[3 * num + d] = acc[d];
      s_need[num] = n_bigc;
#pragma unroll
      for (int c = 0; c < 2; c++)
        if (!(ref[c] & kRefLeaf)) {
          s_parent[ref[c]] = (uint16_t)num;
          s_sib[ref[c]] = (ref[c ^ 1] & kRefLeaf) ? (uint16_t)0xFFFFu : (uint16_t)ref[c ^ 1];
        }
      if (n_bigc == 0) done |= 1u << k;
    }
    __syncthreads();
    BT_MARK(8);
    // Walk up: the thread that completes a node's last big child merges the children into the node and goes on to its
    // parent (shared-memory atomics; the serial chain is the depth of the tree, ~20 steps of ~150 cycles, where
    // level-synchronous passes paid a barrier and a dozen dependent shared-memory reads per level).  Unions of boxes
    // are order-independent, so the result does not depend on which thread arrives last.
    while (done) {
      const int k = __ffs((int)done) - 1;
      done &= done - 1;
      uint32_t cur = tid + (uint32_t)k * nthr;
      uint32_t b[3] = {s_box[3 * cur], s_box[3 * cur + 1], s_box[3 * cur + 2]};
      while (cur != 0) {  // (number 0 is the root)
        const uint32_t par = s_parent[cur], sib = s_sib[cur];
        __threadfence_block();  // my box before my arrival
        const uint32_t before = atomicSub(&s_need[par], 1u);
        if (before != 1u) break;  // the sibling is still open: its thread takes the parent
        __threadfence_block();
#pragma unroll
        for (int d = 0; d < 3; d++) {
          uint32_t pb = s_box[3 * par + d];  // the parent's leaf-sized child, if any (pass 1)
          pb = min(pb & 0xFFFFu, b[d] & 0xFFFFu) | (max(pb >> 16, b[d] >> 16) << 16);
          if (sib != 0xFFFFu) {
            const uint32_t sb = reinterpret_cast<volatile uint32_t*>(s_box)[3 * sib + d];
            pb = min(pb & 0xFFFFu, sb & 0xFFFFu) | (max(pb >> 16, sb >> 16) << 16);
          }
          b[d] = pb;
          s_box[3 * par + d] = pb;
        }
        cur = par;
      }
    }
    __syncthreads();
    BT_MARK(6);
    // node records of the big nodes (split word, range end, float box) and the big children's halves of the records
    for (uint32_t num = tid; num < n_big; num += nthr) {
      const uint32_t i = s_big[num];
      const uint32_t o = s_other[i], sp = s_split[i];
      const uint32_t f = min(i, o), l = max(i, o);
      uint32_t w = sp;
      if (f == sp) w |= kLeftLeaf;
      if (l == sp + 1) w |= kRightLeaf;
      float blo[3], bhi[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const uint32_t b = s_box[3 * num + d];
        blo[d] = __double2float_rd(fma((double)(b & 0xFFFFu), qcell, qorg[d]));
        bhi[d] = __double2float_ru(fma((double)(b >> 16), qcell, qorg[d]));
      }
      float4* q = reinterpret_cast<float4*>(nodes + i);
      q[0] = make_float4(blo[0], blo[1], blo[2], __uint_as_float(w));
      q[1] = make_float4(bhi[0], bhi[1], bhi[2], __uint_as_float(o));
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
        if (cl - cf >= (uint32_t)kBvhLeaf) {
          const uint32_t cn = s_cid[sp + c];
          uint16_t* half = reinterpret_cast<uint16_t*>(recs + num) + c;  // this child's half of each box word
#pragma unroll
          for (int d = 0; d < 3; d++) {
            const uint32_t b = s_box[3 * cn + d];
            half[2 * d] = (uint16_t)(b & 0xFFFFu);
            half[6 + 2 * d] = (uint16_t)(0u - (b >> 16));
          }
        }
      }
    }
    if (tid == 0) a.g.quant[set].n_rec = n_big;
#ifdef BUILD_TIMING
    BT_MARK(7);
    if (tid == 0 && set == 0)
      printf("BT n %u: bbox %lld morton %lld sort %lld gather %lld topo_sweep %lld topo_big %lld boxes_leaf %lld boxes_up %lld "
             "n_big %lld compact %lld total %lld\n", n, bt[1] - bt[0], bt[2] - bt[1], bt[3] - bt[2], bt[4] - bt[3],
             bt[9] - bt[4], bt[5] - bt[9], bt[8] - bt[5], bt[6] - bt[8], (long long)n_big, bt[7] - bt[6], bt[7] - bt[0]);
#endif
    return;
  }

  // ---- boxes.  Thread t owns the big nodes numbered t, t + 1024, ...  Pass 1: every big node takes the boxes of its
  // LEAF-SIZED children straight from their points (all loads independent; stored in the child's node record for the
  // general walk) — a node without big children is complete, the others keep the partial union in their record.
  // Later passes: a pending node whose big children were completed in an EARLIER pass merges them in (one L2 round trip
  // per pass, over the ~half of the big nodes that have a big child).
  uint16_t* s_cid = reinterpret_cast<uint16_t*>(s_code);  // (codes are dead) node -> number
  auto small_box = [&](uint32_t cf, uint32_t cl, uint32_t cnode, float* blo, float* bhi) {
    double plo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, phi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
#pragma unroll
    for (int k = 0; k < kBvhLeaf; k++) {
      const double4 pt = sorted[min(cf + (uint32_t)k, cl)];
      plo[0] = fmin(plo[0], pt.x); phi[0] = fmax(phi[0], pt.x);
      plo[1] = fmin(plo[1], pt.y); phi[1] = fmax(phi[1], pt.y);
      plo[2] = fmin(plo[2], pt.z); phi[2] = fmax(phi[2], pt.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      blo[k] = __double2float_rd(plo[k]);
      bhi[k] = __double2float_ru(phi[k]);
    }
    if (cf != cl) {  // (a single point has no node record; split word 0: never expanded)
      float4* q = reinterpret_cast<float4*>(nodes + cnode);
      q[0] = make_float4(blo[0], blo[1], blo[2], 0.f);
      q[1] = make_float4(bhi[0], bhi[1], bhi[2], 0.f);
    }
  };
  auto load_box = [&](uint32_t node, float* blo, float* bhi) {
    const float4* q = reinterpret_cast<const float4*>(nodes + node);
    const float4 va = __ldcg(q), vb = __ldcg(q + 1);
    blo[0] = va.x; blo[1] = va.y; blo[2] = va.z;
    bhi[0] = vb.x; bhi[1] = vb.y; bhi[2] = vb.z;
  };
  auto store_box = [&](uint32_t node, const float* blo, const float* bhi) {  // (keeps the split word and the range end)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      nodes[node].lo[k] = blo[k];
      nodes[node].hi[k] = bhi[k];
    }
  };
  const uint32_t mine = tid < n_big ? (n_big - tid + nthr - 1) / nthr : 0u;  // (at most 20: n_big < n <= 20,480)
  uint32_t pend = 0;
  for (uint32_t k = 0; k < mine; k++) {  // pass 1
    const uint32_t num = tid + k * nthr;
    const uint32_t i = s_big[num];
    const uint32_t o = s_other[i], sp = s_split[i];
    const uint32_t f = min(i, o), l = max(i, o);
    s_cid[i] = (uint16_t)num;
    const bool lbig = sp - f >= (uint32_t)kBvhLeaf, rbig = l - (sp + 1) >= (uint32_t)kBvhLeaf;
    float blo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, bhi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    float clo[3], chi[3];
    if (!lbig) {
      small_box(f, sp, sp, clo, chi);
#pragma unroll
      for (int q = 0; q < 3; q++) {
        blo[q] = fminf(blo[q], clo[q]);
        bhi[q] = fmaxf(bhi[q], chi[q]);
      }
    }
    if (!rbig) {
      small_box(sp + 1, l, sp + 1, clo, chi);
#pragma unroll
      for (int q = 0; q < 3; q++) {
        blo[q] = fminf(blo[q], clo[q]);
        bhi[q] = fmaxf(bhi[q], chi[q]);
      }
    }
    store_box(i, blo, bhi);
    const bool complete = !lbig && !rbig;
    s_ready[num] = complete ? 1 : 0;
    if (!complete) pend |= 1u << k;
  }
  __syncthreads();
  BT_MARK(8);
#ifdef BUILD_TIMING
  uint32_t bt_passes = 0;
#endif
  for (uint32_t pass = 2; __syncthreads_or(pend != 0); pass++) {
#ifdef BUILD_TIMING
    bt_passes = pass;
#endif
    uint32_t m = pend;
    while (m) {
      const int k = __ffs((int)m) - 1;
      m &= m - 1;
      const uint32_t num = tid + (uint32_t)k * nthr;
      const uint32_t i = s_big[num];
      const uint32_t o = s_other[i], sp = s_split[i];
      const uint32_t f = min(i, o), l = max(i, o);
      const bool lbig = sp - f >= (uint32_t)kBvhLeaf, rbig = l - (sp + 1) >= (uint32_t)kBvhLeaf;
      bool ok = true;
      if (lbig) {
        const uint32_t r = s_ready[s_cid[sp]];
        ok = ok && r != 0 && r < pass;
      }
      if (rbig) {
        const uint32_t r = s_ready[s_cid[sp + 1]];
        ok = ok && r != 0 && r < pass;
      }
      if (!ok) continue;
      float blo[3], bhi[3], clo[3], chi[3];
      load_box(i, blo, bhi);  // the partial union of pass 1
      if (lbig) {
        load_box(sp, clo, chi);
#pragma unroll
        for (int q = 0; q < 3; q++) {
          blo[q] = fminf(blo[q], clo[q]);
          bhi[q] = fmaxf(bhi[q], chi[q]);
        }
      }
      if (rbig) {
        load_box(sp + 1, clo, chi);
#pragma unroll
        for (int q = 0; q < 3; q++) {
          blo[q] = fminf(blo[q], clo[q]);
          bhi[q] = fmaxf(bhi[q], chi[q]);
        }
      }
      store_box(i, blo, bhi);
      s_ready[num] = (uint8_t)min(pass, 255u);
      pend &= ~(1u << k);
    }
  }
  BT_MARK(6);

  // ---- compact records (common.cuh: BvhRec), one per big node, numbered as above: the children's boxes are in their
  // node records by now (single points: the point itself)
  if (a.g.quant == nullptr) return;
  const uint32_t rec_cap = a.g.pt_cap / 2;
  if (n_big > rec_cap) {  // degenerate tree (long chains): this set keeps the general walk only
    if (tid == 0) a.g.quant[set].n_rec = kNoRecs;
    return;
  }
  BvhRec* recs = reinterpret_cast<BvhRec*>(a.g.keys + (size_t)set * 2 * a.g.pt_cap);
  const double qorg[3] = {__ldcg(&a.g.quant[set].org[0]), __ldcg(&a.g.quant[set].org[1]), __ldcg(&a.g.quant[set].org[2])};
  const double qinv = __ldcg(&a.g.quant[set].inv_cell);
  for (uint32_t num = tid; num < n_big; num += nthr) {
    const uint32_t i = s_big[num];
    const uint32_t o = s_other[i], sp = s_split[i];
    const uint32_t f = min(i, o), l = max(i, o);
    uint32_t cbox[2][3], ref[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint32_t cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
      float blo[3], bhi[3];
      if (cf == cl) {
        const double4 pt = sorted[cf];
        blo[0] = __double2float_rd(pt.x); bhi[0] = __double2float_ru(pt.x);
        blo[1] = __double2float_rd(pt.y); bhi[1] = __double2float_ru(pt.y);
        blo[2] = __double2float_rd(pt.z); bhi[2] = __double2float_ru(pt.z);
      } else {
        load_box(sp + (uint32_t)c, blo, bhi);
      }
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double tl = ((double)blo[d] - qorg[d]) * qinv, th = ((double)bhi[d] - qorg[d]) * qinv;
        const uint32_t ql = (uint32_t)fmin(fmax(floor(tl - 1e-6), 0.0), (double)kRecCellMax);
        const uint32_t qh = (uint32_t)fmin(fmax(ceil(th + 1e-6), 0.0), (double)kRecCellMax);
        cbox[c][d] = ql | (qh << 16);
      }
      ref[c] = (cl - cf < (uint32_t)kBvhLeaf) ? (kRefLeaf | ((cl - cf) << 24) | cf) : (uint32_t)s_cid[sp + c];
    }
    uint32_t box[6];
    rec_pack_boxes(cbox[0], cbox[1], box);
    uint4* dst = reinterpret_cast<uint4*>(recs + num);
    dst[0] = make_uint4(box[0], box[1], box[2], box[3]);
    dst[1] = make_uint4(box[4], box[5], ref[0], ref[1]);
  }
  if (tid == 0) a.g.quant[set].n_rec = n_big;
#ifdef BUILD_TIMING
  BT_MARK(7);
  if (tid == 0 && set == 0)
    printf("BT n %u: bbox %lld morton %lld sort %lld gather %lld topology %lld boxes_leaf %lld boxes_levels %lld passes %lld "
           "n_big %lld compact %lld total %lld\n", n, bt[1] - bt[0], bt[2] - bt[1], bt[3] - bt[2], bt[4] - bt[3],
           bt[5] - bt[4], bt[8] - bt[5], bt[6] - bt[8], (long long)bt_passes, (long long)n_big, bt[7] - bt[6], bt[7] - bt[0]);
#endif
}

// ============================================================================ exact k-NN (K4)

#ifndef KNN_PTX_INSERT
#define KNN_PTX_INSERT 1
#endif
#ifndef KNN_LEAF_LEAN
#define KNN_LEAF_LEAN 1
#endif
// K best (d2, id) pairs, sorted ascending, in registers.  insert() is a branch-free shifting network:
// every lane of a warp executes the same instructions whatever its data.
template <int K>
struct TopK {
  double d[K];
  uint32_t id[K];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) {
      d[i] = CUDART_INF;
      id[i] = 0xFFFFFFFFu;
    }
  }
  static __device__ __forceinline__ bool lt(double da, uint32_t ia, double db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
  }
  __device__ __forceinline__ void insert(double dn, uint32_t in) {
#if KNN_PTX_INSERT
    if constexpr (K == 5) {
      // The same network with the comparison chained through predicates: (in < id) -> (dn == d) AND that -> (dn < d) OR
      // that, three instructions per slot where the compiler's form takes four (a predicate initialisation, the
      // ordered compare and two predicated compares); the shifts are predicated selects as in the compiler's code.
      asm("{\n"
          ".reg .pred q, t, c0, c1, c2, c3, c4;\n"
          "setp.lt.u32 q, %11, %5;  setp.eq.and.f64 t, %10, %0, q;  setp.lt.or.f64 c0, %10, %0, t;\n"
          "setp.lt.u32 q, %11, %6;  setp.eq.and.f64 t, %10, %1, q;  setp.lt.or.f64 c1, %10, %1, t;\n"
          "setp.lt.u32 q, %11, %7;  setp.eq.and.f64 t, %10, %2, q;  setp.lt.or.f64 c2, %10, %2, t;\n"
          "setp.lt.u32 q, %11, %8;  setp.eq.and.f64 t, %10, %3, q;  setp.lt.or.f64 c3, %10, %3, t;\n"
          "setp.lt.u32 q, %11, %9;  setp.eq.and.f64 t, %10, %4, q;  setp.lt.or.f64 c4, %10, %4, t;\n"
          "@c4 selp.f64 %4, %3, %10, c3;  @c4 selp.u32 %9, %8, %11, c3;\n"
          "@c3 selp.f64 %3, %2, %10, c2;  @c3 selp.u32 %8, %7, %11, c2;\n"
          "@c2 selp.f64 %2, %1, %10, c1;  @c2 selp.u32 %7, %6, %11, c1;\n"
          "@c1 selp.f64 %1, %0, %10, c0;  @c1 selp.u32 %6, %5, %11, c0;\n"
          "@c0 mov.f64 %0, %10;  @c0 mov.u32 %5, %11;\n"
          "}"
          : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]), "+d"(d[4]), "+r"(id[0]), "+r"(id[1]), "+r"(id[2]), "+r"(id[3]),
            "+r"(id[4])
          : "d"(dn), "r"(in));
      return;
    }
#endif
    bool c[K];
#pragma unroll
    for (int i = 0; i < K; i++) c[i] = lt(dn, in, d[i], id[i]);
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      // c[i-1] implies c[i] (the array is sorted)
      d[i] = c[i] ? (c[i - 1] ? d[i - 1] : dn) : d[i];
      id[i] = c[i] ? (c[i - 1] ? id[i - 1] : in) : id[i];
    }
    d[0] = c[0] ? dn : d[0];
    id[0] = c[0] ? in : id[0];
  }
  // k-th best distance (k <= K, runtime)
  __device__ __forceinline__ double kth(int k) const {
    double v = d[K - 1];
#pragma unroll
    for (int i = 0; i < K - 1; i++)
      if (i == k - 1) v = d[i];
    return v;
  }
};

struct QueryF {  // the query rounded down / up to float, for conservative box tests
  float lo[3], hi[3];
};

// max(a, b, 0) in one instruction (3-input FMNMX3 of sm_100)
__device__ __forceinline__ float fmax3_nonneg(float a, float b) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(0.f));
  return r;
}

// Lower bound (never above the true value) of the squared distance from the query to any point inside the box.
__device__ __forceinline__ float box_lower_bound(const BvhNode& b, const QueryF& q) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const float below = __fsub_rd(b.lo[d], q.hi[d]);  // > 0 when the query is below the box
    const float above = __fsub_rd(q.lo[d], b.hi[d]);  // > 0 when the query is above the box
    const float g = fmax3_nonneg(below, above);
    s = __fmaf_rd(g, g, s);  // one rounding, downward: still never above the exact g*g + s
  }
  return s;  // empty boxes (lo = +inf, hi = -inf) give +inf
}

#ifndef KNN_LD256
#define KNN_LD256 1
#endif
#ifndef KNN_STACK4
#define KNN_STACK4 1
#endif
// 32-byte records (node, point) are fetched with ONE 256-bit load (sm_100: ld.global.nc.v8.b32 / .v4.f64, SASS
// LDG.E.256) instead of two 128-bit ones: half the load instructions of the traversal.  Both record arrays are
// 32-byte aligned (cudaMalloc base + index * 32).
__device__ __forceinline__ BvhNode load_node(const BvhNode* __restrict__ p) {
  BvhNode n;
#if KNN_LD256
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
      : "l"(p));
  n.lo[0] = __uint_as_float(r0); n.lo[1] = __uint_as_float(r1); n.lo[2] = __uint_as_float(r2); n.split = r3;
  n.hi[0] = __uint_as_float(r4); n.hi[1] = __uint_as_float(r5); n.hi[2] = __uint_as_float(r6); n.pad = 0;
#else
  const float4* f = reinterpret_cast<const float4*>(p);
  const float4 a = __ldg(f), b = __ldg(f + 1);
  n.lo[0] = a.x; n.lo[1] = a.y; n.lo[2] = a.z; n.split = __float_as_uint(a.w);
  n.hi[0] = b.x; n.hi[1] = b.y; n.hi[2] = b.z; n.pad = 0;
#endif
  return n;
}

__device__ __forceinline__ double4 load_point(const double4* __restrict__ p) {
#if KNN_LD256
  double4 t;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(t.x), "=d"(t.y), "=d"(t.z), "=d"(t.w) : "l"(p));
  return t;
#else
  return *p;
#endif
}

// Leaf scan shared by the global and the shared-memory walk: the points [first, last] (at most kBvhLeaf) of the sorted copy.
template <int K>
__device__ __forceinline__ void scan_leaf(const double4* __restrict__ sorted, uint32_t first, uint32_t last, double qx,
                                          double qy, double qz, double d2_cut, TopK<K>& tk) {
  // ---- leaf: fp64 distances in nanoflann's L2_Simple order, branch-free insertion
#if KNN_LEAF_LEAN
  // All kBvhLeaf records behind `first` are loaded from one base address (immediate offsets; slots beyond `last` read
  // the following points of the sorted copy, or the padding behind the last set — capi.cu reserves it — and are
  // discarded).  Candidates beyond d2_cut are not filtered here: they can only sit behind the true neighbours in the
  // list (the cut is a distance within which k points exist, or the radius that radius_count() applies afterwards),
  // and the pruning bound is min(k-th, d2_cut) either way — same result, two instructions less per slot.
  (void)d2_cut;
  const double4* base = sorted + first;
  const uint32_t cnt = last - first;
#pragma unroll
  for (int j = 0; j < kBvhLeaf; j++) {
    const double4 t = load_point(base + j);
    double d2 = sqdist(qx, qy, qz, t.x, t.y, t.z);
    uint32_t id = (uint32_t)__double_as_longlong(t.w);
    const bool ok = (uint32_t)j <= cnt;
    d2 = ok ? d2 : CUDART_INF;
    id = ok ? id : 0xFFFFFFFFu;
    tk.insert(d2, id);
  }
#else
#pragma unroll
  for (int j = 0; j < kBvhLeaf; j++) {
    const uint32_t p = first + j;
    const double4 t = load_point(sorted + min(p, last));
    double d2 = sqdist(qx, qy, qz, t.x, t.y, t.z);
    uint32_t id = (uint32_t)__double_as_longlong(t.w);
    const bool ok = p <= last && d2 <= d2_cut;
    d2 = ok ? d2 : CUDART_INF;
    id = ok ? id : 0xFFFFFFFFu;
    tk.insert(d2, id);
  }
#endif
}

// Exact k nearest neighbours of (qx,qy,qz) among the points of one set, restricted to candidates that can pass the
// strict radius filter of kdtree.cpp:24-26 (d2 <= d2_cut is a superset of sqrt(d2) < max_dist; radius_count() applies
// the exact test afterwards).  A subtree is skipped only when its box lower bound is STRICTLY above the current k-th
// best (or the radius cut), so candidates that tie the k-th distance are still seen and resolved by index.
constexpr int kBvhStack = 64;  // >= tree depth: 30 Morton bits + 32 position bits for duplicate codes

template <int K>
__device__ __forceinline__ void knn_bvh(const BvhHdr& h, const BvhNode* __restrict__ nodes,
                                        const double4* __restrict__ sorted, double qx, double qy, double qz, int k,
                                        double max_dist, TopK<K>& tk, double d2_hint) {
  // d2_hint: a squared distance within which at least k points are known to lie (prunes only; inf = none)
  tk.init();
  if (h.n == 0) return;
  // one cap for both the radius cut and the caller's hint: a candidate beyond the hint lies strictly beyond the k-th
  // nearest (k points are known within it), so it can be dropped like one beyond the radius
  const double d2_cut = fmin(max_dist > 0 ? max_dist * max_dist * (1.0 + 1e-12) : CUDART_INF, d2_hint);
  QueryF q;
  q.lo[0] = __double2float_rd(qx); q.hi[0] = __double2float_ru(qx);
  q.lo[1] = __double2float_rd(qy); q.hi[1] = __double2float_ru(qy);
  q.lo[2] = __double2float_rd(qz); q.hi[2] = __double2float_ru(qz);
  float bound = __double2float_ru(d2_cut);  // a subtree is pruned when its lower bound > bound

  // pending subtrees: (node index, key range, lower bound); a subtree of <= kBvhLeaf points is scanned as a leaf
  // (the node index itself is not needed: a node's record is only read to test it, and that read also yields
  // its split word, which is all the descent needs)
#if KNN_STACK4
  uint4 st[kBvhStack];  // (split word, first, last, lower bound): one 128-bit local store / load per push / pop
#else
  uint32_t st_split[kBvhStack], st_first[kBvhStack], st_last[kBvhStack];
  float st_lb[kBvhStack];
#endif
  int sp = 0;
  uint32_t first = 0, last = h.n - 1, cur_w = 0;
  bool have = true, done = false;
  if (h.n > (uint32_t)kBvhLeaf) {
    const BvhNode root = load_node(nodes);
    if (box_lower_bound(root, q) > bound) return;
    cur_w = root.split;
  }

  // All lanes advance their own traversal one node per step until each stands on a leaf it must scan (or is done);
  // then the warp scans the leaves together — the scan is ~10x the cost of a step, so it must run converged.
  while (true) {
    bool at_leaf = false;
    while (!done && !at_leaf) {
      if (!have) {
        if (sp == 0) {
          done = true;
        } else {
          --sp;
#if KNN_STACK4
          const uint4 e = st[sp];
          if (__uint_as_float(e.w) <= bound) {
            cur_w = e.x;
            first = e.y;
            last = e.z;
            have = true;
          }
#else
          if (st_lb[sp] <= bound) {
            cur_w = st_split[sp];
            first = st_first[sp];
            last = st_last[sp];
            have = true;
          }
#endif
        }
      } else if (last - first < (uint32_t)kBvhLeaf) {
        at_leaf = true;
      } else {
        const uint32_t w = cur_w;
        const uint32_t s = w & kSplitMask;
        // children are adjacent records; a single-point child has no record of its own: bound 0 (it is scanned as
        // a one-point leaf and its split word is never used)
        const BvhNode cl = load_node(nodes + s), cr = load_node(nodes + s + 1);
        const float dl = (w & kLeftLeaf) ? 0.f : box_lower_bound(cl, q);
        const float dr = (w & kRightLeaf) ? 0.f : box_lower_bound(cr, q);
        const bool right_first = dr < dl;
        const float dn = right_first ? dr : dl, df = right_first ? dl : dr;
        const uint32_t l_first = first, l_last = s, r_first = s + 1, r_last = last;
        if (df <= bound) {  // far child stays pending
#if KNN_STACK4
          st[sp] = make_uint4(right_first ? cl.split : cr.split, right_first ? l_first : r_first,
                              right_first ? l_last : r_last, __float_as_uint(df));
#else
          st_split[sp] = right_first ? cl.split : cr.split;
          st_first[sp] = right_first ? l_first : r_first;
          st_last[sp] = right_first ? l_last : r_last;
          st_lb[sp] = df;
#endif
          sp++;
        }
        if (dn <= bound) {
          cur_w = right_first ? cr.split : cl.split;
          first = right_first ? r_first : l_first;
          last = right_first ? r_last : l_last;
        } else {
          have = false;
        }
      }
    }
    if (!at_leaf) break;  // done
    scan_leaf<K>(sorted, first, last, qx, qy, qz, d2_cut, tk);
    bound = __double2float_ru(fmin(tk.kth(k), d2_cut));
    have = false;
  }
}

// ---------------------------------------------------------------------------- the same search over compact records
// The batched association kernel keeps a target set's compact records (common.cuh: BvhRec) in SHARED memory, so a
// node step costs two LDS.128 (29 cycles) instead of two dependent global loads (L1 hit 32, L2 hit ~250 cycles; 37 % of
// them missed L1 in the global walk: profiles/r1_kernels_full_v29.md) and the state of a walk is one 32-bit child
// reference.  Box tests run on the set's 15-bit grid in integer arithmetic, both children of a record at once: the
// query's cell coordinates are rounded outward (below / above) and clamped onto the grid, a record word holds the same
// box corner of both children in its two halves, and per axis
//     u   = (-hi) + qlo                         VIADDMNMX.S16x2  (max with -32768: a plain 16x2 add)
//     gap = max(lo + (-qhi), u, 0)              VIADDMNMX.S16x2.RELU
// give both children's gaps in cells (every difference of two 15-bit cells fits an int16).  gap^2 is summed in 32-bit
// integers (3 * 32767^2 < 2^32) and compared with the pruning bound converted to cells^2 and rounded up: a subtree is
// skipped only when its true distance is strictly above the bound, exactly as in knn_bvh — the results are identical,
// whatever the tree looks like.  (Round 2 first decoded the corners as floats 2^23 + c and took exact float differences:
// 13 instructions per axis for the two children, 25 % of the kernel's instructions; this form takes 6.)
struct QueryG {  // per axis, replicated in both halves: lower cell of the query / minus its upper cell
  uint32_t lo2[3], nhi2[3];
};

__device__ __forceinline__ void query_grid(const BvhQuant* __restrict__ Q, double qx, double qy, double qz, QueryG& g) {
  const double q[3] = {qx, qy, qz};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double t = (q[d] - Q->org[d]) * Q->inv_cell;
    const double e = 1e-6 + fabs(t) * 1e-12;  // covers the rounding of t itself
    // clamping is conservative: a clamped coordinate lies between the true one and every box (cells 0 .. kRecCellMax)
    const uint32_t lo = (uint32_t)fmin(fmax(floor(t - e), 0.0), (double)kRecCellMax);
    const uint32_t hi = (uint32_t)fmin(fmax(ceil(t + e), 0.0), (double)kRecCellMax);
    g.lo2[d] = lo | (lo << 16);
    const uint32_t nh = (0u - hi) & 0xFFFFu;
    g.nhi2[d] = nh | (nh << 16);
  }
}

// lower bounds, in cells^2, of the squared distances from the query to the boxes of both children of a record
__device__ __forceinline__ void rec_lower_bounds(uint32_t lo_x, uint32_t lo_y, uint32_t lo_z, uint32_t nhi_x, uint32_t nhi_y,
                                                 uint32_t nhi_z, const QueryG& g, uint32_t& sl, uint32_t& sr) {
  const uint32_t lo[3] = {lo_x, lo_y, lo_z}, nhi[3] = {nhi_x, nhi_y, nhi_z};
  sl = 0;
  sr = 0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const uint32_t u = __viaddmax_s16x2(nhi[d], g.lo2[d], 0x80008000u);   // qlo - hi
    const uint32_t gap = __viaddmax_s16x2_relu(lo[d], g.nhi2[d], u);      // max(lo - qhi, qlo - hi, 0)
    const uint32_t gl = gap & 0xFFFFu, gr = gap >> 16;
    sl += gl * gl;
    sr += gr * gr;
  }
}

__device__ __forceinline__ void load_rec_shared(uint32_t saddr, uint4& r0, uint4& r1) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(saddr));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(saddr + 16u));
}

// Top of the traversal stack in shared memory (KNN_SMEM_STACK = entries per thread, a power of two; 0 = local memory
// only).  Every push also goes to the thread's local-memory stack (a store nobody waits for); the newest
// KNN_SMEM_STACK entries additionally sit in a ring in shared memory, where a pop costs an LDS instead of a
// local-memory load that missed L1 (8 % of the stall samples of the first shared-memory version).  `s_stack` is the
// shared address of this thread's slot 0; slot e lives kKnnStackPitch bytes further (conflict-free across a warp).
#ifndef KNN_SMEM_STACK
#define KNN_SMEM_STACK 0
#endif
#ifndef KNN_PREFETCH
#define KNN_PREFETCH 0
#endif
constexpr uint32_t kKnnStackPitch = 1024u * 8u;

// s_recs: shared-memory address of the set's record 0; n_pts: points of the set; Q: its grid (shared or global memory)
template <int K>
__device__ __forceinline__ void knn_compact(uint32_t s_recs, uint32_t s_stack, uint32_t n_pts,
                                            const BvhQuant* __restrict__ Q, const double4* __restrict__ sorted,
                                            double qx, double qy, double qz, int k, double max_dist, TopK<K>& tk,
                                            double d2_hint) {
  tk.init();
  if (n_pts == 0) return;
  const double d2_cut = fmin(max_dist > 0 ? max_dist * max_dist * (1.0 + 1e-12) : CUDART_INF, d2_hint);
  QueryG g;
  query_grid(Q, qx, qy, qz, g);
  const double inv2 = Q->inv_cell2;
  // cells^2, rounded up and saturated (the conversion maps +inf and anything above 2^32 - 1 to 2^32 - 1, which no sum of
  // three squared gaps exceeds); a subtree is pruned when its lower bound > bound
  uint32_t bound = __double2uint_ru(d2_cut * inv2);
  uint2 st[kBvhStack];                             // pending subtrees: (child reference, lower bound)
  int sp = 0;
  (void)s_stack;  // (the shared-memory stack ring of §10b was measured out; the parameter stays for the flag's layout)
  // The state of the walk is ONE word: a record number, a leaf reference (kRefLeaf set), kWalkPop (take the next
  // pending subtree) or kWalkDone (kRefLeaf set too, so the node loop ends on it like on a leaf).  (Separate
  // have / done / at-leaf flags cost a handful of byte-juggling instructions per step of the node loop.)
  constexpr uint32_t kWalkPop = 0x7FFFFFFFu, kWalkDone = 0xFFFFFFFFu;
  uint32_t cur = n_pts > (uint32_t)kBvhLeaf ? 0u : (kRefLeaf | ((n_pts - 1u) << 24));
  while (true) {
    while (!(cur & kRefLeaf)) {
      if (cur == kWalkPop) {
        if (sp == 0) {
          cur = kWalkDone;
        } else {
          --sp;
          const uint2 e = st[sp];
          if (e.y <= bound) cur = e.x;
        }
      } else {
        uint4 r0, r1;  // (lo.x lo.y lo.z nhi.x) (nhi.y nhi.z refL refR)
        load_rec_shared(s_recs + cur * (uint32_t)sizeof(BvhRec), r0, r1);
        uint32_t dl, dr;
        rec_lower_bounds(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, g, dl, dr);
        const bool right_first = dr < dl;
        const uint32_t dn = right_first ? dr : dl, df = right_first ? dl : dr;
        if (df <= bound) {  // far child stays pending
          st[sp] = make_uint2(right_first ? r1.z : r1.w, df);
          sp++;
        }
        cur = dn <= bound ? (right_first ? r1.w : r1.z) : kWalkPop;
      }
    }
    if (cur == kWalkDone) break;
    const uint32_t first = cur & 0x00FFFFFFu;
    scan_leaf<K>(sorted, first, first + ((cur >> 24) & 15u), qx, qy, qz, d2_cut, tk);
    bound = __double2uint_ru(fmin(tk.kth(k), d2_cut) * inv2);
    cur = kWalkPop;
  }
}

// kdtree.cpp:24-26 : keep neighbours with max_dist <= 0 || sqrt(d2) < max_dist (strict). Sorted => prefix.
template <int K>
__device__ __forceinline__ int radius_count(const TopK<K>& tk, int k, double max_dist) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < K; i++) {
    if (i < k && tk.id[i] != 0xFFFFFFFFu && (max_dist <= 0 || sqrt(tk.d[i]) < max_dist)) m++;
  }
  return m;
}

template <int K>
__global__ void __launch_bounds__(128) knn_kernel(KnnArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_queries) return;
  const BvhHdr h = a.g.hdr[0];
  TopK<K> tk;
  knn_bvh<K>(h, a.g.nodes, a.g.sorted, a.queries[3 * i], a.queries[3 * i + 1], a.queries[3 * i + 2], a.k,
             a.max_dist, tk, CUDART_INF);
  const int m = radius_count(tk, a.k, a.max_dist);
  a.count_out[i] = (uint32_t)m;
#pragma unroll
  for (int j = 0; j < K; j++)
    if (j < a.k) a.idx_out[i * a.k + j] = j < m ? tk.id[j] : 0xFFFFFFFFu;
}

}  // namespace
}  // namespace loamgpu
