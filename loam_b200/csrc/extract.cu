// Feature extraction kernels (replaces loam/features-inl.h + src/features.cpp of the reference).
//
// K1+K2 fused: one CTA per scan ring.  The ring is staged once into shared memory (a single
// TMA bulk copy when the layout allows it), then range / curvature / validity mask / per-sector
// ordering / greedy selection all run out of shared memory; the scan is read from HBM exactly once.
//
//   phase A  stage ring (cp.async.bulk + mbarrier, or strided loads for odd layouts)
//   phase B  range r[j] = sqrt((x^2+y^2)+z^2)                         common.h:81-86
//   phase C  curvature stencil (2N+1 taps, fp64, unfused)             features-inl.h:53-87
//            validity mask (4 checks, idempotent "set false" scatter) features-inl.h:90-124, features.cpp:20-68
//   phase D  selection: per (sector, edge|planar) walk in the reference's order, the greedy pick-and-suppress
//            recursion is resolved in parallel rounds over the sector's columns, picks are ranked by
//            (curvature, index) and truncated to max+1                         features-inl.h:27-48,137-180
//
// Tie-break (documented): the reference's std::sort is unstable and compares curvature only (features.h:91,
// features-inl.h:38); here equal curvatures order by ascending index in the sorted sector, i.e. the edge walk
// (which runs from the end) prefers the larger index and the planar walk the smaller one.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

#ifndef EXTRACT_NB_PER_SWEEP
#define EXTRACT_NB_PER_SWEEP 1
#endif

namespace loamgpu {

namespace {

// packed staging record in shared memory: kE scalars of type T per point (x, y, z first)
template <typename T, int kE>
struct Rec {
  static constexpr int kBytes = (int)sizeof(T) * kE;  // float x 4: sensor-driver float4 records; float x 3: packed xyz;
  static constexpr int kElems = kE;                   // double x 3: Eigen::Vector3d-shaped
};

// Shared memory of one ring: [staging records | P doubles | P mask bytes | mbarrier].  The doubles hold the ranges
// while the mask is derived and the curvature afterwards; once the curvature is known the staged points are dead and
// their space holds the selection's walk state (a uint32 neighbour-priority word and a 64-bit pick slot per column,
// a uint16 pick column, state / validity / candidate / threshold bits per 32 columns: 14.75 bytes per column).
// (Round 1 kept ranges and curvature in separate arrays: 33 KB per 1024-column ring and 6 resident rings per SM; this
// layout needs 24 KB with packed xyz records (the walk state is larger than the staged ring), 25 KB with float4 records.)
__host__ __device__ inline uint32_t walk_bytes(uint32_t P) {  // prio | pick keys | state words | four bit arrays | pick columns
  return 4 * ((P + 1) & ~1u) + 8 * P + 24 * ((P + 31) >> 5) + 2 * P;
}
__host__ __device__ inline size_t stage_bytes(uint32_t P, uint32_t rec) {
  const size_t a = (size_t)P * rec, b = walk_bytes(P);
  return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// 0xFFFFFFFF if (ka, ia) < (kb, ib) lexicographically, else 0: the borrow of the 96-bit subtraction (ka:ia) - (kb:ib),
// three carry-chained subtractions instead of a 64-bit compare, an equality test and an index compare.
__device__ __forceinline__ uint32_t lex_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
  uint32_t r;
  asm("{\n.reg .u32 t;\n"
      "sub.cc.u32 t, %1, %2;\n"
      "subc.cc.u32 t, %3, %4;\n"
      "subc.cc.u32 t, %5, %6;\n"
      "subc.u32 %0, 0, 0;\n}"
      : "=r"(r)
      : "r"(ia), "r"(ib), "r"((uint32_t)ka), "r"((uint32_t)kb), "r"((uint32_t)(ka >> 32)), "r"((uint32_t)(kb >> 32)));
  return r;
}

// T = scalar type of the staged ring (what the arithmetic widens from), TIn = scalar type of the caller's records.
// They differ only for de-warping float input: the moved points are not float-representable, so the ring is staged
// as doubles.
template <typename T, typename TIn, bool kDewarp, int kE>
__device__ __forceinline__ void extract_ring_body(const ExtractArgs& a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t P = a.P, N = a.N, S = a.S;
  const uint32_t ring = blockIdx.x, scan = blockIdx.y;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;

  // ---- shared memory carve-up (see extract_smem_bytes) ----
  T* stage = reinterpret_cast<T*>(smem);                                       // P * Rec<T, kE>::kBytes; later the walk state
  double* cur = reinterpret_cast<double*>(smem + stage_bytes(P, Rec<T, kE>::kBytes));  // ranges, then curvature
  uint8_t* mask = reinterpret_cast<uint8_t*>(cur + P);                         // P bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(mask + ((P + 15) & ~15u));       // 8-byte aligned mbarrier

  const unsigned char* ring_src =
      a.pts + (size_t)scan * a.scan_stride_bytes + (size_t)ring * P * a.stride;

  // ---- phase A: stage the ring ----
  if (a.use_bulk) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = P * Rec<T, kE>::kBytes;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(stage, ring_src, bytes, bar);
    }
    mbar_wait(bar, 0);
  } else {
    double mq[4] = {0.0, 0.0, 0.0, 1.0}, mt[3] = {0.0, 0.0, 0.0};
    if constexpr (kDewarp) {
      double mo[7];
#pragma unroll
      for (int i = 0; i < 7; i++) mo[i] = a.motions ? a.motions[(size_t)scan * 7 + i] : a.motion[i];
      dewarp_hemisphere(mo, mq);
      mt[0] = mo[4];
      mt[1] = mo[5];
      mt[2] = mo[6];
    }
    for (uint32_t j = tid; j < P; j += nthr) {
      const TIn* src = reinterpret_cast<const TIn*>(ring_src + (size_t)j * a.stride);
      T x = (T)src[0], y = (T)src[1], z = (T)src[2];
      if constexpr (kDewarp) {
        const V3 m = dewarp_point(mq, mt, j, P, V3{(double)x, (double)y, (double)z});
        x = (T)m.x;
        y = (T)m.y;
        z = (T)m.z;
        if (a.dewarp_out) {
          double* o = a.dewarp_out + ((size_t)scan * a.R * P + (size_t)ring * P + j) * 3;
          o[0] = m.x;
          o[1] = m.y;
          o[2] = m.z;
        }
      }
      stage[j * Rec<T, kE>::kElems + 0] = x;
      stage[j * Rec<T, kE>::kElems + 1] = y;
      stage[j * Rec<T, kE>::kElems + 2] = z;
    }
    __syncthreads();
  }

  // ---- phase B: ranges (into the curvature buffer), ring-edge part of the mask ----
  // P - N wraps exactly like the reference's size_t arithmetic when P < N (features-inl.h:66-67)
  const uint64_t hi_edge = (uint64_t)P - (uint64_t)N;
  for (uint32_t j = tid; j < P; j += nthr) {
    const double x = (double)stage[j * Rec<T, kE>::kElems + 0];
    const double y = (double)stage[j * Rec<T, kE>::kElems + 1];
    const double z = (double)stage[j * Rec<T, kE>::kElems + 2];
    cur[j] = point_range(x, y, z);
    mask[j] = ((j < N) || ((uint64_t)j >= hi_edge)) ? 0 : 1;  // CHECK 1
  }
  __syncthreads();

  // ---- phase C: validity mask from the ranges, then the curvature over them ----
  for (uint32_t j = tid; j < P; j += nthr) {
    if ((j < N) || ((uint64_t)j >= hi_edge)) continue;  // CHECK 1 handled above
    const double r = cur[j], rn = cur[j + 1], rp = cur[j - 1];
    if (r < a.min_range || r > a.max_range) {  // CHECK 2
      mask[j] = 0;
      for (uint32_t k = 1; k <= N; k++) {
        mask[j + k] = 0;
        mask[j - k] = 0;
      }
    } else if (dsub(rn, r) > a.occ) {  // CHECK 3 case 1
      for (uint32_t k = 1; k <= N; k++) mask[j + k] = 0;
    } else if (dsub(r, rn) > a.occ) {  // CHECK 3 case 2
      for (uint32_t k = 0; k < N; k++) mask[j - k] = 0;
    } else {  // CHECK 4
      const double diff_next = fabs(dsub(rp, r));
      const double diff_prev = fabs(dsub(rn, r));
      const double lim = dmul(a.par, r);
      if (diff_next > lim && diff_prev > lim) mask[j] = 0;
    }
  }
  __syncthreads();  // the ranges are dead: the curvature takes their place
  const double m2n = -(2.0 * (double)N);
  for (uint32_t j = tid; j < P; j += nthr) {
    const bool ring_edge = (j < N) || ((uint64_t)j >= hi_edge);
    double c = -1.0;
    if (!ring_edge) {
      double dx = dmul(m2n, (double)stage[j * Rec<T, kE>::kElems + 0]);
      double dy = dmul(m2n, (double)stage[j * Rec<T, kE>::kElems + 1]);
      double dz = dmul(m2n, (double)stage[j * Rec<T, kE>::kElems + 2]);
      for (uint32_t k = 1; k <= N; k++) {
        const T* lo = stage + (size_t)(j - k) * Rec<T, kE>::kElems;
        const T* hi = stage + (size_t)(j + k) * Rec<T, kE>::kElems;
        dx = dadd(dadd(dx, (double)lo[0]), (double)hi[0]);
        dy = dadd(dadd(dy, (double)lo[1]), (double)hi[1]);
        dz = dadd(dadd(dz, (double)lo[2]), (double)hi[2]);
      }
      c = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    }
    cur[j] = c;
  }
  __syncthreads();  // (also: the staged points are dead from here on)

  if (a.curv_out != nullptr || a.mask_out != nullptr) {  // secondary entry points stop here
    const size_t base = (size_t)scan * a.R * P + (size_t)ring * P;
    for (uint32_t j = tid; j < P; j += nthr) {
      if (a.curv_out) a.curv_out[base + j] = cur[j];
      if (a.mask_out) a.mask_out[base + j] = mask[j];
    }
    return;
  }

  // ---- phase D: selection.  The reference sorts each sector by curvature and walks it greedily (edge walk from the
  // largest curvature down, then planar walk from the smallest up; a pick invalidates its +-(N-1) neighbours; a walk
  // stops after max+1 picks).  The outcome of one walk is fully determined by local comparisons: a candidate is
  // picked iff no higher-priority candidate within +-(N-1) columns is picked.  That recursion is resolved here in
  // parallel rounds (decided states are final, so in-place updates are safe), the picks are then ranked by priority
  // (= the reference's selection order), truncated to max+1, and only the accepted picks invalidate neighbours.
  // Walks run in the reference's order (sector-major, edge before planar) because each sees the mask left by the
  // previous ones, including suppression that spills across a sector boundary (features-inl.h:148-151).
  // Bit-parallel form (round 2; the byte-per-column form of round 1 spent 30 % of the kernel's instructions in the rounds at
  // 6-8 active lanes and 12 % on the neighbour-priority bits of every walk).  Column states live in 32-column words —
  // one 64-bit word holds the "open" bits (low half) and the "picked" bits (high half) of 32 columns, so a reader
  // always sees a consistent pair — and a column's view of its +-(N-1) neighbours is a window of at most 31 bits cut
  // out of three adjacent words by a funnel shift (bit k of a window = column j - (N-1) + k).  Which neighbours
  // precede a column in the EDGE order is computed once per ring (prio); the planar order is its complement.  A
  // round is then two ANDs per column, decided for 32 columns at once and published with one ballot.
  const uint32_t reach = N - 1;
  const size_t ring_id = (size_t)scan * a.R + ring;
  __shared__ uint32_t s_m, s_base[2];
  if (reach > 15) {  // only reachable with N >= P (capi.cu: plan_extract), where CHECK 1 leaves no valid column
    if (tid == 0) a.ring_counts[ring_id * 2 + 0] = a.ring_counts[ring_id * 2 + 1] = 0;
    return;
  }
  const uint32_t W = (P + 31) >> 5;
  uint32_t* prio = reinterpret_cast<uint32_t*>(stage);      // [P] neighbours that precede j in the edge order
  uint64_t* pkey = reinterpret_cast<uint64_t*>(prio + ((P + 1) & ~1u));  // [P] picks of the walk: bits of their curvature
  uint64_t* state = pkey + P;                               // [W] open | picked << 32
  uint32_t* vbits = reinterpret_cast<uint32_t*>(state + W); // [W] validity mask (features.cpp:20-68 + accepted picks)
  uint32_t* cbits = vbits + W;                              // [W] candidates of the current walk
  uint32_t* ebits = cbits + W;                              // [W] curvature above the edge threshold
  uint32_t* fbits = ebits + W;                              // [W] curvature below the planar threshold
  uint16_t* plist = reinterpret_cast<uint16_t*>(fbits + W); // [P] picks of the walk: their columns
  const uint32_t lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const uint32_t wmask = (2u << (2 * reach)) - 1u;  // 2 (N-1) + 1 window bits
  const uint32_t self = 1u << reach;
  auto window = [&](uint32_t prev, uint32_t own, uint32_t next) -> uint32_t {
    const int pos = (int)lane - (int)reach;
    return (pos >= 0 ? __funnelshift_r(own, next, pos) : __funnelshift_r(prev, own, 32 + pos)) & wmask;
  };
  if (tid == 0) s_base[0] = s_base[1] = 0;
  for (uint32_t w = warp; w < W; w += nwarps) {  // per-ring bit words: a walk's candidates are then three ANDs per word
    const uint32_t j = 32 * w + lane;
    const double c = j < P ? cur[j] : 0.0;
    const uint32_t vb = __ballot_sync(0xffffffffu, j < P && mask[j] != 0);
    const uint32_t eb = __ballot_sync(0xffffffffu, j < P && c > a.edge_thr);
    const uint32_t fb = __ballot_sync(0xffffffffu, j < P && c < a.planar_thr);
    if (lane == 0) {
      vbits[w] = vb;
      ebits[w] = eb;
      fbits[w] = fb;
    }
  }
  for (uint32_t j = tid; j < P; j += nthr) {
    const double cj = cur[j];
    uint32_t bits = 0;
    for (uint32_t n = 1; n <= reach; n++) {
      if (j >= n && cur[j - n] > cj) bits |= 1u << (reach - n);    // (a left neighbour loses ties: larger index first)
      if (j + n < P && cur[j + n] >= cj) bits |= 1u << (reach + n);
    }
    prio[j] = bits;
  }
  __syncthreads();
  const uint32_t pps = P / S;
  uint32_t* ge = a.ring_edge + ring_id * a.capE_ring;
  uint32_t* gp = a.ring_planar + ring_id * a.capP_ring;
  volatile uint64_t* vstate = state;
  for (uint32_t walk = 0; walk < 2 * S; walk++) {
    const uint32_t sec = walk >> 1;
    const bool planar = (walk & 1u) != 0;
    const uint32_t b = sec * pps, e = (sec == S - 1) ? P : b + pps;
    const uint32_t cap = planar ? a.maxP : a.maxE;  // the walk accepts cap + 1 picks
    if (b >= e) continue;  // (more sectors than columns)
    const uint32_t w0 = b >> 5, w1 = (e - 1) >> 5;
    bool any = false;
    for (uint32_t w = w0 + tid; w <= w1; w += nthr) {  // candidates: valid, inside the sector, past the walk's threshold
      const uint32_t f = max(b, 32 * w) - 32 * w, l = min(e - 1, 32 * w + 31) - 32 * w;
      const uint32_t word = vbits[w] & (planar ? fbits[w] : ebits[w]) & (0xFFFFFFFFu >> (31 - l)) & (0xFFFFFFFFu << f);
      cbits[w] = word;
      state[w] = (uint64_t)word;
      any |= word != 0;
    }
    if (tid == 0) s_m = 0;
    if (!__syncthreads_or(any)) continue;  // no candidate in this walk (uniform: every thread sees the same result)
    // A warp owns whole state words, and a candidate only waits on columns within +-(N-1): most dependency chains never
    // leave the word.  Those are resolved in warp-local sweeps on the word in registers (the neighbouring words are
    // read once per CTA round) and only chains that cross words cost a CTA barrier.  Any interleaving gives the same
    // result: a column's state changes once (open -> picked / dropped) and a stale "open" seen across warps only
    // delays a decision.
    for (;;) {
      bool open_left = false;
      for (uint32_t w = w0 + warp; w <= w1; w += nwarps) {
        const uint64_t sw = vstate[w];
        uint32_t ow = (uint32_t)sw, pw = (uint32_t)(sw >> 32);
        if (ow == 0) continue;
        const uint32_t j = 32 * w + lane;
        // neighbours that are candidates of this walk and precede j in its order (priority: edge = larger curvature
        // first, ties by larger index — the ascending (c, idx) order walked from the end; planar = the complement)
        uint32_t h = 0;
        if ((ow >> lane) & 1u) {
          const uint32_t cw = window(w > w0 ? cbits[w - 1] : 0u, cbits[w], w < w1 ? cbits[w + 1] : 0u);
          h = cw & (planar ? ~prio[j] : prio[j]) & ~self;
        }
#if !EXTRACT_NB_PER_SWEEP
        const uint64_t sp = w > w0 ? vstate[w - 1] : 0ull, sn = w < w1 ? vstate[w + 1] : 0ull;
#endif
        for (;;) {
#if EXTRACT_NB_PER_SWEEP
          const uint64_t sp = w > w0 ? vstate[w - 1] : 0ull, sn = w < w1 ? vstate[w + 1] : 0ull;
#endif
          const bool my = ((ow >> lane) & 1u) != 0;
          const uint32_t openwin = window((uint32_t)sp, ow, (uint32_t)sn);
          const uint32_t pickwin = window((uint32_t)(sp >> 32), pw, (uint32_t)(sn >> 32));
          const bool drop = my && (pickwin & h) != 0;
          const bool pick = my && !drop && (openwin & h) == 0;
          const uint32_t dmask = __ballot_sync(0xffffffffu, drop), kmask = __ballot_sync(0xffffffffu, pick);
          if ((dmask | kmask) == 0) break;
          ow &= ~(dmask | kmask);
          pw |= kmask;
#if EXTRACT_NB_PER_SWEEP
          if (lane == 0) vstate[w] = (uint64_t)ow | ((uint64_t)pw << 32);
#endif
          if (kmask) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(&s_m, (uint32_t)__popc(kmask));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (pick) {
              const uint32_t slot = at + (uint32_t)__popc(kmask & ((1u << lane) - 1u));
              pkey[slot] = (uint64_t)__double_as_longlong(cur[j]);
              plist[slot] = (uint16_t)j;
            }
          }
          if (ow == 0) break;
        }
#if !EXTRACT_NB_PER_SWEEP
        if (lane == 0) vstate[w] = (uint64_t)ow | ((uint64_t)pw << 32);
#endif
        open_left |= ow != 0;
      }
      if (!__syncthreads_or(open_left)) break;
    }
    // rank the picks by priority (one thread per pick, taken from the compact list); accept the first cap + 1; only
    // those invalidate their neighbours.  Curvatures are >= +0 (a sum of squares; NaN is never a candidate), so their
    // bit patterns order like the doubles: the comparison is integer work, ties go to the index as in `prio`.
    const uint32_t m = s_m;
    const uint32_t base = s_base[planar ? 1 : 0];
    for (uint32_t t = tid; t < m; t += nthr) {
      const uint64_t kj = pkey[t];
      const uint32_t j = plist[t];
      uint32_t rank = 0;
      if (planar) {
        for (uint32_t i = 0; i < m; i++) rank -= lex_less(pkey[i], plist[i], kj, j);
      } else {
        for (uint32_t i = 0; i < m; i++) rank -= lex_less(kj, j, pkey[i], plist[i]);
      }
      if (rank > cap) continue;
      (planar ? gp : ge)[base + rank] = ring * P + j;
      const uint32_t lo = j >= reach ? j - reach : 0u, hi = min(j + reach, P - 1);
      for (uint32_t w = lo >> 5; w <= (hi >> 5); w++) {
        const uint32_t f = max(lo, 32 * w) - 32 * w, l = min(hi, 32 * w + 31) - 32 * w;  // bits f .. l of word w
        atomicAnd(&vbits[w], ~((0xFFFFFFFFu >> (31 - l)) & (0xFFFFFFFFu << f)));
      }
    }
    __syncthreads();
    if (tid == 0) s_base[planar ? 1 : 0] = base + min(m, cap + 1);
  }
  __syncthreads();
  if (tid == 0) {
    a.ring_counts[ring_id * 2 + 0] = s_base[0];
    a.ring_counts[ring_id * 2 + 1] = s_base[1];
  }
}

// The plain kernels (what every throughput path runs) and the de-warping ones are separate entry points so that the
// register budget of one never shapes the other: plain 37-40 registers / 6 CTAs per SM; de-warping (24-byte staging
// records, 41 KB of shared memory per 1024-column ring) capped for 5.
// float records: 8 resident rings per SM (32 registers; 24-25 KB of shared memory per 1024-column ring) — measured
// 4.04 -> 3.70 ms per 1024 float4 scans against the 40 registers / 6 rings the compiler picks unconstrained; a minimum
// of 1 lets it take far more registers and runs at 6.4 ms
template <typename T, int kE>
__global__ void __launch_bounds__(kExtractThreads, sizeof(T) == 4 ? 8 : 6) extract_ring_kernel(ExtractArgs a) {
  extract_ring_body<T, T, false, kE>(a);
}
template <typename TIn>
__global__ void __launch_bounds__(kExtractThreads, 5) extract_ring_dewarp_kernel(ExtractArgs a) {
  extract_ring_body<double, TIn, true, 3>(a);
}

// Pack per-ring pick lists into the scan-level feature arrays (reference output order:
// line-major, sector-major, selection order) and gather the widened feature points
// (featuresToEigen, features.h:188-198).  One CTA per scan.
template <bool kDewarp>
__global__ void __launch_bounds__(256) pack_features_kernel(PackArgs a) {
  extern __shared__ uint32_t offs[];  // [R+1][2]
  const uint32_t scan = blockIdx.x, R = a.R;
  const uint32_t slot = (uint32_t)((a.scan0 + scan) % a.n_slots);
  if (threadIdx.x < 32) {  // exclusive prefix of the per-ring counts: warp 0, 32 rings per step
    const uint32_t lane = threadIdx.x;
    uint32_t e0 = 0, p0 = 0;
    for (uint32_t r0 = 0; r0 < R; r0 += 32) {
      const uint32_t r = r0 + lane;
      const uint32_t ce = r < R ? a.ring_counts[((size_t)scan * R + r) * 2] : 0u;
      const uint32_t cp = r < R ? a.ring_counts[((size_t)scan * R + r) * 2 + 1] : 0u;
      uint32_t ie = ce, ip = cp;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t te = __shfl_up_sync(0xffffffffu, ie, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
        if ((int)lane >= o) {
          ie += te;
          ip += tp;
        }
      }
      if (r < R) {
        offs[2 * r] = e0 + ie - ce;
        offs[2 * r + 1] = p0 + ip - cp;
      }
      e0 += __shfl_sync(0xffffffffu, ie, 31);
      p0 += __shfl_sync(0xffffffffu, ip, 31);
    }
    if (lane == 0) {
      offs[2 * R] = e0;
      offs[2 * R + 1] = p0;
      if (blockIdx.y == 0) {
        a.feat_counts[slot * 2] = e0;
        a.feat_counts[slot * 2 + 1] = p0;
        if (a.n_edge_out) a.n_edge_out[a.scan0 + scan] = e0;
        if (a.n_planar_out) a.n_planar_out[a.scan0 + scan] = p0;
      }
    }
  }
  __syncthreads();
  const unsigned char* base = a.pts + (size_t)scan * a.scan_stride_bytes;
  double mq[4] = {0.0, 0.0, 0.0, 1.0}, mt[3] = {0.0, 0.0, 0.0};
  if constexpr (kDewarp) {
    double mo[7];
#pragma unroll
    for (int i = 0; i < 7; i++) mo[i] = a.motions ? a.motions[(size_t)scan * 7 + i] : a.motion[i];
    dewarp_hemisphere(mo, mq);
    mt[0] = mo[4];
    mt[1] = mo[5];
    mt[2] = mo[6];
  }
  for (uint32_t r = blockIdx.y; r < R; r += gridDim.y) {  // (grid.y > 1: few scans, rings spread over CTAs)
    for (int kind = 0; kind < 2; kind++) {
      const uint32_t n = offs[2 * (r + 1) + kind] - offs[2 * r + kind];
      const uint32_t cap_ring = kind ? a.capP_ring : a.capE_ring;
      const uint32_t cap_scan = kind ? a.capP_scan : a.capE_scan;
      const uint32_t* src = (kind ? a.ring_planar : a.ring_edge) + ((size_t)scan * R + r) * cap_ring;
      uint32_t* dst_idx = (kind ? a.planar_idx : a.edge_idx) + (size_t)slot * cap_scan + offs[2 * r + kind];
      double4* dst_pt = (kind ? a.planar_pts : a.edge_pts) + (size_t)slot * cap_scan + offs[2 * r + kind];
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t id = src[i];
        dst_idx[i] = id;
        const unsigned char* rec = base + (size_t)id * a.stride;
        double4 v;
        if (a.dtype == LOAMGPU_F32) {
          const float* f = reinterpret_cast<const float*>(rec);
          v = make_double4((double)f[0], (double)f[1], (double)f[2], 0.0);
        } else {
          const double* d = reinterpret_cast<const double*>(rec);
          v = make_double4(d[0], d[1], d[2], 0.0);
        }
        if constexpr (kDewarp) {  // the feature points of a de-warped extraction are the moved points
          const V3 m = dewarp_point(mq, mt, id % a.P, a.P, V3{v.x, v.y, v.z});
          v = make_double4(m.x, m.y, m.z, 0.0);
        }
        dst_pt[i] = v;
      }
    }
  }
}

}  // namespace

size_t extract_smem_bytes(int dtype, uint32_t P, uint32_t S, uint32_t rec_bytes) {
  (void)S;
  const uint32_t rec = rec_bytes ? rec_bytes : (dtype == LOAMGPU_F32 ? 16 : 24);
  const size_t b = stage_bytes(P, rec) + 8 * (size_t)P + ((P + 15) & ~15u) + 8 + 16;
  return (b + 127) & ~(size_t)127;
}

template <typename K>
static cudaError_t launch_extract_as(K kernel, const ExtractArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  kernel<<<grid, kExtractThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_extract(const ExtractArgs& a, uint32_t n_scans, cudaStream_t st) {
  const dim3 grid(a.R, n_scans);
  if (a.dewarp) {  // double staging whatever the input type; strided loads (use_bulk is off)
    const size_t smem = extract_smem_bytes(LOAMGPU_F64, a.P, a.S, 24);
    return a.dtype == LOAMGPU_F32 ? launch_extract_as(extract_ring_dewarp_kernel<float>, a, grid, smem, st)
                                  : launch_extract_as(extract_ring_dewarp_kernel<double>, a, grid, smem, st);
  }
  if (a.dtype == LOAMGPU_F32) {
    // packed xyz records (12 bytes) are staged as they are: one bulk copy of 12 P bytes, and three-float records
    // read without shared-memory bank conflicts; every other float layout is staged as float4
    if (a.stride == 12) return launch_extract_as(extract_ring_kernel<float, 3>, a, grid, extract_smem_bytes(a.dtype, a.P, a.S, 12), st);
    return launch_extract_as(extract_ring_kernel<float, 4>, a, grid, extract_smem_bytes(a.dtype, a.P, a.S, 16), st);
  }
  return launch_extract_as(extract_ring_kernel<double, 3>, a, grid, extract_smem_bytes(a.dtype, a.P, a.S, 24), st);
}

cudaError_t launch_pack(const PackArgs& a, uint32_t n_scans, cudaStream_t st) {
  // one CTA per scan when there are scans enough to fill the GPU; a single call's scan is spread ring-wise over CTAs
  // (one CTA walking 64 rings' dependent gathers took 83 us of a 0.23 ms loamgpu_extract call)
  const uint32_t split = n_scans >= 128 ? 1u : std::min<uint32_t>(std::min<uint32_t>(a.R, 65535u), std::max<uint32_t>(1u, 296u / std::max(n_scans, 1u)));
  const dim3 grid(n_scans, split);
  if (a.dewarp)
    pack_features_kernel<true><<<grid, 256, (size_t)(a.R + 1) * 2 * sizeof(uint32_t), st>>>(a);
  else
    pack_features_kernel<false><<<grid, 256, (size_t)(a.R + 1) * 2 * sizeof(uint32_t), st>>>(a);
  return cudaGetLastError();
}

}  // namespace loamgpu
