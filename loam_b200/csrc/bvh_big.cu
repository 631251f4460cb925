// Multi-CTA build of the LBVH over ONE large point set (the "local map" target of scan-to-map registration:
// ~1M accumulated feature points, BASELINE.json config 5).  Produces exactly the layout bvh.cuh's single-CTA
// build produces (BvhHdr, BvhNode[n-1], Morton-sorted double4 copy), so knn_bvh() traverses either.  Replaces the
// nanoflann KD-tree constructor over the target set (reference: kdtree.h:24-41, registration-inl.h:20-23).
//
//   big_bbox_kernel      grid-stride min/max, warp shuffle + atomics on order-preserving u64 encodings
//   big_morton_kernel    30-bit Morton key + original index per point
//   8 x { big_hist_kernel, big_scan_kernel, big_scatter_kernel }   stable LSD radix sort, 4-bit digits:
//                        a thread owns 8 consecutive keys, its 16 digit counters live in ONE 64-bit register
//                        (4 bits each), block prefixes come from warp shuffles over 4 x u64 of 16-bit fields,
//                        the (bin, block) base offsets from a single-CTA exclusive scan
//   big_gather_kernel    Morton-ordered point copy + sorted code array
//   big_topology_kernel  binary radix tree (one thread per internal node, count-leading-zeros binary searches),
//                        parent links
//   big_boxes_kernel     bottom-up box merge: the second child to arrive at a node (atomic counter) continues
//                        upward; boxes are float32 rounded outward exactly as in the single-CTA build
//
// Tree SHAPE never affects k-NN results (bvh.cuh: a subtree is skipped only on a conservative lower bound), so
// this build needs no bit-exactness argument beyond "every point is in exactly one leaf range and every box
// contains its points".
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace loamgpu {
namespace {

constexpr int kBigThreads = 256;
constexpr int kBigItems = 8;
constexpr int kBigTile = kBigThreads * kBigItems;
constexpr int kBigBins = 16;

__device__ __forceinline__ unsigned long long enc_d(double v) {  // order-preserving double -> u64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_d(unsigned long long e) {
  const unsigned long long b = (e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ uint32_t spread10_big(uint32_t v) {
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// bbox[0..2] = min x,y,z ; bbox[3..5] = max x,y,z  (encoded; initialised by the launcher: min = ~0, max = 0)
__global__ void __launch_bounds__(256) big_bbox_kernel(const double4* __restrict__ pts, uint32_t n,
                                                       unsigned long long* bbox) {
  double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 p = pts[i];
    lo[0] = fmin(lo[0], p.x); hi[0] = fmax(hi[0], p.x);
    lo[1] = fmin(lo[1], p.y); hi[1] = fmax(hi[1], p.y);
    lo[2] = fmin(lo[2], p.z); hi[2] = fmax(hi[2], p.z);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      atomicMin(bbox + d, enc_d(lo[d]));
      atomicMax(bbox + 3 + d, enc_d(hi[d]));
    }
  }
}

__global__ void __launch_bounds__(256) big_morton_kernel(const double4* __restrict__ pts, uint32_t n,
                                                         const unsigned long long* __restrict__ bbox, uint2* keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lx = dec_d(bbox[0]), ly = dec_d(bbox[1]), lz = dec_d(bbox[2]);
  const double emax = fmax(fmax(dec_d(bbox[3]) - lx, dec_d(bbox[4]) - ly), dec_d(bbox[5]) - lz);
  const double scale = emax > 0 ? 1023.999 / emax : 0.0;  // cubic cells, as in the single-CTA build
  const double4 p = pts[i];
  const uint32_t ix = (uint32_t)fmin(fmax((p.x - lx) * scale, 0.0), 1023.0);
  const uint32_t iy = (uint32_t)fmin(fmax((p.y - ly) * scale, 0.0), 1023.0);
  const uint32_t iz = (uint32_t)fmin(fmax((p.z - lz) * scale, 0.0), 1023.0);
  keys[i] = make_uint2(spread10_big(ix) | (spread10_big(iy) << 1) | (spread10_big(iz) << 2), i);
}

// Digit counters of this thread's 8 keys, 4 bits per bin in one register.
__device__ __forceinline__ unsigned long long count_digits(const uint2* __restrict__ src, uint32_t n, uint32_t base,
                                                           int shift, uint2 (&kv)[kBigItems]) {
  unsigned long long packed = 0;
#pragma unroll
  for (int j = 0; j < kBigItems; j++) {
    const uint32_t i = base + j;
    if (i < n) {
      kv[j] = src[i];
      packed += 1ull << (4 * ((kv[j].x >> shift) & (kBigBins - 1)));
    }
  }
  return packed;
}

// 16 x 4-bit fields -> 4 registers of 4 x 16-bit fields (bins 4g .. 4g+3 in register g)
__device__ __forceinline__ void widen_counts(unsigned long long packed, unsigned long long (&w)[4]) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t q = (uint32_t)(packed >> (16 * g)) & 0xFFFFu;
    w[g] = (unsigned long long)(q & 0xFu) | ((unsigned long long)((q >> 4) & 0xFu) << 16) |
           ((unsigned long long)((q >> 8) & 0xFu) << 32) | ((unsigned long long)((q >> 12) & 0xFu) << 48);
  }
}

// hist[bin * n_blocks + block] = keys of this block's tile whose digit is `bin`
__global__ void __launch_bounds__(kBigThreads) big_hist_kernel(const uint2* __restrict__ src, uint32_t n, int shift,
                                                               uint32_t* hist, uint32_t n_blocks) {
  __shared__ unsigned long long s_w[kBigThreads / 32][4];
  uint2 kv[kBigItems];
  const uint32_t base = blockIdx.x * kBigTile + threadIdx.x * kBigItems;
  unsigned long long w[4];
  widen_counts(count_digits(src, n, base, shift, kv), w);
#pragma unroll
  for (int g = 0; g < 4; g++)
    for (int o = 16; o > 0; o >>= 1) w[g] += __shfl_xor_sync(0xffffffffu, w[g], o);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int g = 0; g < 4; g++) s_w[threadIdx.x >> 5][g] = w[g];
  }
  __syncthreads();
  if (threadIdx.x < kBigBins) {
    const int g = threadIdx.x >> 2, f = threadIdx.x & 3;
    uint32_t t = 0;
    for (int wv = 0; wv < kBigThreads / 32; wv++) t += (uint32_t)(s_w[wv][g] >> (16 * f)) & 0xFFFFu;
    hist[threadIdx.x * n_blocks + blockIdx.x] = t;
  }
}

// in-place exclusive scan of `m` words by one CTA
__global__ void __launch_bounds__(1024) big_scan_kernel(uint32_t* v, uint32_t m) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t b = 0; b < m; b += 1024) {
    const uint32_t i = b + tid;
    const uint32_t mine = i < m ? v[i] : 0;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t wt = s_warp[lane];
      uint32_t winc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if ((int)lane >= o) winc += t;
      }
      s_warp[lane] = winc - wt;  // exclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    if (i < m) v[i] = carry + s_warp[warp] + inc - mine;
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_warp[31] + inc;
    __syncthreads();
  }
}

// Stable scatter: position of a key = base of (its digit, this block) + keys of the same digit owned by earlier
// threads of the block + earlier keys of the same digit among this thread's own 8.
__global__ void __launch_bounds__(kBigThreads) big_scatter_kernel(const uint2* __restrict__ src, uint2* dst, uint32_t n,
                                                                  int shift, const uint32_t* __restrict__ hist_scanned,
                                                                  uint32_t n_blocks) {
  __shared__ unsigned long long s_w[kBigThreads / 32][4];
  __shared__ uint32_t s_off[kBigBins * kBigThreads];  // running output position per (bin, thread)
  uint2 kv[kBigItems];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t base = blockIdx.x * kBigTile + tid * kBigItems;
  unsigned long long mine[4], inc[4];
  widen_counts(count_digits(src, n, base, shift, kv), mine);
#pragma unroll
  for (int g = 0; g < 4; g++) {
    unsigned long long v = mine[g];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
      if ((int)lane >= o) v += t;
    }
    inc[g] = v;
    if (lane == 31) s_w[warp][g] = v;
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < 4; g++) {
    unsigned long long before = 0;
    for (uint32_t wv = 0; wv < warp; wv++) before += s_w[wv][g];
    const unsigned long long excl = before + inc[g] - mine[g];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int bin = 4 * g + f;
      s_off[bin * kBigThreads + tid] = hist_scanned[bin * n_blocks + blockIdx.x] + ((uint32_t)(excl >> (16 * f)) & 0xFFFFu);
    }
  }
  // (a thread reads back only its own column of s_off)
#pragma unroll
  for (int j = 0; j < kBigItems; j++) {
    if (base + j < n) {
      const uint32_t d = (kv[j].x >> shift) & (kBigBins - 1);
      const uint32_t pos = s_off[d * kBigThreads + tid]++;
      dst[pos] = kv[j];
    }
  }
}

__global__ void __launch_bounds__(256) big_gather_kernel(const double4* __restrict__ pts, const uint2* __restrict__ keys,
                                                         uint32_t n, double4* sorted, uint32_t* codes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 kv = keys[i];
  const double4 p = pts[kv.y];
  sorted[i] = make_double4(p.x, p.y, p.z, __longlong_as_double((long long)kv.y));
  codes[i] = kv.x;
}

// One thread per internal node (Karras 2012): range and split from common-prefix lengths of the sorted
// (code, position) keys; same conventions as bvh.cuh (children of node i are records split / split+1).
__global__ void __launch_bounds__(256) big_topology_kernel(const uint32_t* __restrict__ codes, uint32_t n, BvhNode* nodes,
                                                           uint32_t* parent, uint32_t* counter) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n) return;
  auto delta = [&](int i, int j) -> int {
    if (j < 0 || j >= (int)n) return -1;
    const uint32_t ci = __ldg(codes + i), cj = __ldg(codes + j);
    return ci != cj ? __clz(ci ^ cj) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
  };
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
  if (first == split) w |= kLeftLeaf; else parent[split] = (uint32_t)i;
  if (last == split + 1) w |= kRightLeaf; else parent[split + 1] = (uint32_t)i;
  nodes[i].split = w;
  nodes[i].pad = (uint32_t)j;  // the other end of the node's key range
  counter[i] = 0;
  if (i == 0) parent[0] = 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) big_boxes_kernel(const double4* __restrict__ sorted, uint32_t n, BvhNode* nodes,
                                                        const uint32_t* __restrict__ parent, uint32_t* counter) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n) return;
  uint32_t cur = t;
  {
    const uint32_t w = nodes[cur].split;
    if (!((w & kLeftLeaf) && (w & kRightLeaf))) return;  // only nodes with two point children start a walk
  }
  for (;;) {
    const uint32_t w = nodes[cur].split;  // written by the previous kernel, never modified here
    const uint32_t sp = w & kSplitMask;
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      float clo[3], chi[3];
      if (w & (c == 0 ? kLeftLeaf : kRightLeaf)) {
        const double4 pt = sorted[sp + c];
        clo[0] = __double2float_rd(pt.x); chi[0] = __double2float_ru(pt.x);
        clo[1] = __double2float_rd(pt.y); chi[1] = __double2float_ru(pt.y);
        clo[2] = __double2float_rd(pt.z); chi[2] = __double2float_ru(pt.z);
      } else {  // merged by another thread: read through L2
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
      nodes[cur].lo[k] = lo[k];
      nodes[cur].hi[k] = hi[k];
    }
    if (cur == 0) break;
    __threadfence();  // the box is visible before the arrival is counted
    const uint32_t p = parent[cur];
    const uint32_t pw = nodes[p].split;
    const uint32_t needed = 2u - (((pw & kLeftLeaf) ? 1u : 0u) + ((pw & kRightLeaf) ? 1u : 0u));
    const uint32_t arrived = atomicAdd(counter + p, 1u) + 1u;
    if (arrived < needed) break;  // the sibling subtree is not finished: its last thread will continue upward
    __threadfence();
    cur = p;
  }
}

__global__ void big_header_kernel(BvhHdr* hdr, uint32_t n) {
  BvhHdr h;
  h.n = n;
  h.pad[0] = h.pad[1] = h.pad[2] = 0;
  *hdr = h;
}

}  // namespace

size_t bvh_big_scratch_bytes(uint32_t n) {
  const size_t n_blocks = ((size_t)n + kBigTile - 1) / kBigTile;
  // bbox (6 u64, padded) + codes + parent + counter + histogram
  return 64 + 3 * (size_t)std::max<uint32_t>(n, 1) * 4 + (size_t)kBigBins * std::max<size_t>(n_blocks, 1) * 4;
}

cudaError_t launch_bvh_build_big(const double4* pts, uint32_t n, const BvhSetArrays& g, void* scratch, cudaStream_t st,
                                 uint64_t* launches) {
  uint64_t nl = 0;
  big_header_kernel<<<1, 1, 0, st>>>(g.hdr, n);
  nl++;
  if (n == 0) {
    if (launches) *launches += nl;
    return cudaGetLastError();
  }
  unsigned long long* bbox = reinterpret_cast<unsigned long long*>(scratch);
  uint32_t* codes = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(scratch) + 64);
  uint32_t* parent = codes + n;
  uint32_t* counter = parent + n;
  uint32_t* hist = counter + n;
  const uint32_t n_blocks = (n + kBigTile - 1) / kBigTile;
  const uint32_t g256 = (n + 255) / 256;
  cudaError_t e = cudaMemsetAsync(bbox, 0xFF, 24, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(bbox + 3, 0, 24, st);
  if (e != cudaSuccess) return e;
  big_bbox_kernel<<<std::min<uint32_t>(g256, 1184), 256, 0, st>>>(pts, n, bbox);
  uint2* src = g.keys;
  uint2* dst = g.keys + g.pt_cap;
  big_morton_kernel<<<g256, 256, 0, st>>>(pts, n, bbox, src);
  nl += 2;
  for (int shift = 0; shift < 30; shift += 4) {  // bits 30/31 of the last digit are zero
    big_hist_kernel<<<n_blocks, kBigThreads, 0, st>>>(src, n, shift, hist, n_blocks);
    big_scan_kernel<<<1, 1024, 0, st>>>(hist, kBigBins * n_blocks);
    big_scatter_kernel<<<n_blocks, kBigThreads, 0, st>>>(src, dst, n, shift, hist, n_blocks);
    std::swap(src, dst);
    nl += 3;
  }
  big_gather_kernel<<<g256, 256, 0, st>>>(pts, src, n, g.sorted, codes);
  nl++;
  if (n >= 2) {
    big_topology_kernel<<<g256, 256, 0, st>>>(codes, n, g.nodes, parent, counter);
    big_boxes_kernel<<<g256, 256, 0, st>>>(g.sorted, n, g.nodes, parent, counter);
    nl += 2;
  }
  if (launches) *launches += nl;
  return cudaGetLastError();
}

}  // namespace loamgpu
