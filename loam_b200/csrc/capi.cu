// C-ABI entry points of libloamgpu.so (declared in include/loamgpu.h) and the host-side context:
// device buffers, streams, chunked sequence pipeline.  Host code only packs buffers and launches
// kernels; every number in the results is produced by the CUDA kernels in extract.cu / register.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost one pointer test when no tool is attached

#include "kernels.h"
#include "loamgpu.h"

using namespace loamgpu;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// page-locked host staging (results of the single-scan / single-pair calls come back through it in one piece)
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

thread_local std::string g_create_err;

// NVTX range around one C-ABI call (tracing): visible in Nsight Systems / Compute timelines
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define API_RANGE() NvtxRange _nvtx_range(__func__)

}  // namespace

struct loamgpu_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  std::string err;
  uint64_t launches = 0;
  uint32_t chunk_pairs = 0;  // pairs per internal chunk of the sequence calls; 0 = automatic (pick_chunk)
  int max_smem_optin = 0;
  int morton_queries = 1;  // LOAMGPU_QUERY_ORDER=original switches the k-NN kernel to source-index order (A/B)
  int lm_moments = 1;      // LOAMGPU_LM_MOMENTS=0: the LM kernel streams every record for every evaluation (round 1)
  uint32_t lm_cluster = 0;          // 0 = automatic (see run_register)
  uint64_t mem_budget = 0;          // bytes the automatic chunk size may plan with (a third of the memory free at creation, <= 24 GB)
  bool staging_unguarded = true;    // scan_in[] was last used outside the event-guarded odometry_host pipeline
  int stage_buf = 0;                // staging buffer the next chunk of a host sequence call copies into (alternates ACROSS calls)
  uint64_t big_target_min = 60000;  // targets at least this large get the multi-CTA NN build ($LOAMGPU_BIG_TARGET_MIN)

  DevBuf scan_in[2];                       // H2D staging of scans
  DevBuf ring_edge, ring_planar, ring_counts;
  DevBuf edge_idx, planar_idx, edge_pts, planar_pts, feat_counts;
  DevBuf ge_hdr, ge_nodes, ge_sorted, ge_keys, ge_aux, gp_hdr, gp_nodes, gp_sorted, gp_keys, gp_aux;
  DevBuf ge_quant, gp_quant, leftover, nc_p, nc_a;     // compact-record grids per set; pairs left to the general k-NN kernel
  DevBuf state, rec_p, rec_a, rec_b, nearest, nn_idx, nn_cnt, active;
  DevBuf big_scratch, misc, motions, out_pose, out_term, out_iters, out_ne, out_np;
  DevBuf det_est, det_upd, det_assoc_n, det_lm_iters, det_lm_cost, init_pose, reg_in;
  HostBuf pinned;                          // staging of the single-call entry points

  // optional per-kernel-class timing (CUDA events on the launching stream)
  struct ProfRec {
    int kind;
    cudaEvent_t a, b;
  };
  bool profiling = false;
  std::vector<ProfRec> prof_open;       // recorded, not yet read back
  std::vector<cudaEvent_t> prof_pool;   // free events
  double prof_ms[LOAMGPU_K_COUNT] = {};
  uint64_t prof_n[LOAMGPU_K_COUNT] = {};
};

// Device-resident registration target ("local map"): feature points in insertion order + their NN structures.
struct loamgpu_map {
  int device = 0;
  uint64_t n[2] = {0, 0};                    // edge / planar points
  DevBuf pts[2];                             // double4 (x,y,z,0), insertion order
  DevBuf hdr[2], nodes[2], sorted[2], keys[2], scratch;
  bool built = false;
};

namespace {

int fail(loamgpu_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(loamgpu_ctx* c, cudaError_t e, const char* what) {
  return fail(c, LOAMGPU_ERR_CUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define CU(call)                                              \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call);  \
  } while (0)

// Times one launch when profiling is on: events bracket the launch on the launching stream.
struct ProfScope {
  loamgpu_ctx* c;
  cudaEvent_t a = nullptr, b = nullptr;
  int kind;
  static cudaEvent_t get(loamgpu_ctx* c) {
    if (!c->prof_pool.empty()) {
      cudaEvent_t e = c->prof_pool.back();
      c->prof_pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  ProfScope(loamgpu_ctx* ctx, int k) : c(ctx), kind(k) {
    c->launches++;
    if (!c->profiling) return;
    a = get(c);
    b = get(c);
    cudaEventRecord(a, c->stream);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, c->stream);
    c->prof_open.push_back({kind, a, b});
  }
};
#define TIMED(kind, call)          \
  do {                             \
    ProfScope _ps(ctx, kind);      \
    CU(call);                      \
  } while (0)

// geometry of one extraction problem, validated
struct ExtractPlan {
  uint32_t R, P, N, S, maxE, maxP, capE_ring, capP_ring, capE_scan, capP_scan;
  size_t smem;
};

int plan_extract(loamgpu_ctx* ctx, int dtype, size_t stride, uint64_t n_points, const loamgpu_lidar_params* lp,
                 const loamgpu_fe_params* fe, ExtractPlan* pl) {
  if (!lp || !fe) return fail(ctx, LOAMGPU_ERR_INVALID, "null parameter struct");
  if (dtype != LOAMGPU_F32 && dtype != LOAMGPU_F64) return fail(ctx, LOAMGPU_ERR_INVALID, "unknown dtype");
  if (stride < (dtype == LOAMGPU_F32 ? 12u : 24u) || stride % (dtype == LOAMGPU_F32 ? 4u : 8u))
    return fail(ctx, LOAMGPU_ERR_INVALID, "stride_bytes too small / misaligned for dtype");
  if (n_points != lp->scan_lines * lp->points_per_line) {
    // exact text of the reference's std::runtime_error (common.h:106-111)
    char buf[256];
    snprintf(buf, sizeof buf, "LOAM: provided lidar scan size ( %llu)  does not match provided lidar parameters (%llu x %llu)",
             (unsigned long long)n_points, (unsigned long long)lp->scan_lines,
             (unsigned long long)lp->points_per_line);
    return fail(ctx, LOAMGPU_ERR_SIZE_MISMATCH, buf);
  }
  if (n_points == 0) {
    memset(pl, 0, sizeof *pl);
    return LOAMGPU_OK;
  }
  if (fe->number_sectors == 0)
    return fail(ctx, LOAMGPU_ERR_INVALID, "number_sectors must be >= 1 (the reference divides by it)");
  if (fe->neighbor_points == 0)
    return fail(ctx, LOAMGPU_ERR_INVALID,
                "neighbor_points must be >= 1 (the reference indexes point idx-1 and throws std::out_of_range)");
  if (fe->neighbor_points > 16 && fe->neighbor_points < lp->points_per_line)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "neighbor_points > 16 not supported");
  if (lp->points_per_line > 65535 || lp->scan_lines > 0xFFFFFFu || n_points > 0xFFFFFFFFull)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "scan too large (points_per_line <= 65535, total points < 2^32)");
  const uint64_t P = lp->points_per_line;
  pl->R = (uint32_t)lp->scan_lines;
  pl->P = (uint32_t)P;
  pl->N = (uint32_t)std::min<uint64_t>(fe->neighbor_points, P);
  pl->S = (uint32_t)std::min<uint64_t>(fe->number_sectors, 0xFFFFFFFFull);
  if (fe->number_sectors > 65536) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "number_sectors > 65536");
  pl->maxE = (uint32_t)std::min<uint64_t>(fe->max_edge_feats_per_sector, P);
  pl->maxP = (uint32_t)std::min<uint64_t>(fe->max_planar_feats_per_sector, P);
  pl->capE_ring = (uint32_t)std::min<uint64_t>((uint64_t)pl->S * ((uint64_t)pl->maxE + 1), P);
  pl->capP_ring = (uint32_t)std::min<uint64_t>((uint64_t)pl->S * ((uint64_t)pl->maxP + 1), P);
  pl->capE_scan = pl->R * pl->capE_ring;
  pl->capP_scan = pl->R * pl->capP_ring;
  pl->smem = extract_smem_bytes(dtype, pl->P, pl->S, dtype == LOAMGPU_F32 && stride == 12 ? 12 : 0);
  if (pl->smem > (size_t)ctx->max_smem_optin)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "points_per_line too large for one CTA's shared memory");
  return LOAMGPU_OK;
}

void fill_extract_args(ExtractArgs& a, const ExtractPlan& pl, const void* dev_pts, int dtype, size_t stride,
                       const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe) {
  memset(&a, 0, sizeof a);
  a.pts = (const unsigned char*)dev_pts;
  a.scan_stride_bytes = (uint64_t)pl.R * pl.P * stride;
  a.stride = (uint32_t)stride;
  a.dtype = dtype;
  const size_t rec = dtype == LOAMGPU_F32 ? (stride == 12 ? 12 : 16) : 24;  // staging record (launch_extract)
  a.use_bulk = (stride == rec) && (((size_t)pl.P * rec) % 16 == 0) && (((uintptr_t)dev_pts) % 16 == 0) &&
               (a.scan_stride_bytes % 16 == 0);
  a.R = pl.R;
  a.P = pl.P;
  a.N = pl.N;
  a.S = pl.S;
  a.maxE = pl.maxE;
  a.maxP = pl.maxP;
  a.edge_thr = fe->edge_feat_threshold;
  a.planar_thr = fe->planar_feat_threshold;
  a.occ = fe->occlusion_thresh;
  a.par = fe->parallel_thresh;
  a.min_range = lp->min_range;
  a.max_range = lp->max_range;
  a.capE_ring = pl.capE_ring;
  a.capP_ring = pl.capP_ring;
}

int reserve_extract(loamgpu_ctx* ctx, const ExtractPlan& pl, uint32_t n_scans, uint32_t n_slots) {
  CU(ctx->ring_edge.reserve((size_t)n_scans * pl.R * pl.capE_ring * 4));
  CU(ctx->ring_planar.reserve((size_t)n_scans * pl.R * pl.capP_ring * 4));
  CU(ctx->ring_counts.reserve((size_t)n_scans * pl.R * 8));
  CU(ctx->edge_idx.reserve((size_t)n_slots * pl.capE_scan * 4));
  CU(ctx->planar_idx.reserve((size_t)n_slots * pl.capP_scan * 4));
  CU(ctx->edge_pts.reserve((size_t)n_slots * pl.capE_scan * 32));
  CU(ctx->planar_pts.reserve((size_t)n_slots * pl.capP_scan * 32));
  CU(ctx->feat_counts.reserve((size_t)n_slots * 8));
  return LOAMGPU_OK;
}

// Extract `n_scans` device-resident scans into feature slots (scan0 + i) % n_slots.
int run_extract(loamgpu_ctx* ctx, const ExtractPlan& pl, const void* dev_pts, int dtype, size_t stride,
                const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, uint32_t n_scans, uint64_t scan0,
                uint32_t n_slots, uint32_t* n_edge_out, uint32_t* n_planar_out, const double* motion = nullptr,
                double* dewarp_out = nullptr, const double* motions_dev = nullptr) {
  ExtractArgs a;
  fill_extract_args(a, pl, dev_pts, dtype, stride, lp, fe);
  const bool dewarp = motion != nullptr || motions_dev != nullptr;
  if (dewarp) {  // de-warp fused into the staging loop (strided loads; the bulk copy cannot transform)
    a.dewarp = 1;
    a.use_bulk = 0;
    if (motion) memcpy(a.motion, motion, sizeof a.motion);
    a.motions = motions_dev;  // one motion per scan of this launch (device), or null: `motion` for all
    a.dewarp_out = dewarp_out;
  }
  a.ring_edge = ctx->ring_edge.as<uint32_t>();
  a.ring_planar = ctx->ring_planar.as<uint32_t>();
  a.ring_counts = ctx->ring_counts.as<uint32_t>();
  TIMED(LOAMGPU_K_EXTRACT, launch_extract(a, n_scans, ctx->stream));
  PackArgs p;
  memset(&p, 0, sizeof p);
  p.pts = a.pts;
  p.scan_stride_bytes = a.scan_stride_bytes;
  p.stride = a.stride;
  p.dtype = dtype;
  p.R = pl.R;
  p.capE_ring = pl.capE_ring;
  p.capP_ring = pl.capP_ring;
  p.capE_scan = pl.capE_scan;
  p.capP_scan = pl.capP_scan;
  p.ring_edge = a.ring_edge;
  p.ring_planar = a.ring_planar;
  p.ring_counts = a.ring_counts;
  p.scan0 = scan0;
  p.n_slots = n_slots;
  p.edge_idx = ctx->edge_idx.as<uint32_t>();
  p.planar_idx = ctx->planar_idx.as<uint32_t>();
  p.edge_pts = ctx->edge_pts.as<double4>();
  p.planar_pts = ctx->planar_pts.as<double4>();
  p.feat_counts = ctx->feat_counts.as<uint32_t>();
  p.n_edge_out = n_edge_out;
  p.n_planar_out = n_planar_out;
  if (dewarp) {
    p.dewarp = 1;
    p.P = pl.P;
    memcpy(p.motion, a.motion, sizeof p.motion);
    p.motions = motions_dev;
  }
  TIMED(LOAMGPU_K_PACK, launch_pack(p, n_scans, ctx->stream));
  return LOAMGPU_OK;
}

int make_regp(loamgpu_ctx* ctx, const loamgpu_reg_params* rp, RegP* out) {
  if (!rp) return fail(ctx, LOAMGPU_ERR_INVALID, "null registration params");
  if (rp->num_edge_neighbors == 0 || rp->num_plane_neighbors == 0)
    return fail(ctx, LOAMGPU_ERR_INVALID, "num_*_neighbors must be >= 1");
  if (rp->num_edge_neighbors > (uint64_t)kKnnMax || rp->num_plane_neighbors > (uint64_t)kKnnMax)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "num_*_neighbors > 32 not supported");
  out->ke = (int)rp->num_edge_neighbors;
  out->kp = (int)rp->num_plane_neighbors;
  out->re = rp->max_edge_neighbor_dist;
  out->rp = rp->max_plane_neighbor_dist;
  out->min_line = (int)std::min<uint64_t>(rp->min_line_fit_points, 1u << 30);
  out->min_plane = (int)std::min<uint64_t>(rp->min_plane_fit_points, 1u << 30);
  out->min_cond = rp->min_line_condition_number;
  out->max_avg = rp->max_avg_point_plane_dist;
  out->max_iterations = (int)std::min<uint64_t>(rp->max_iterations, 1u << 20);
  out->rot_thr = rp->rotation_convergence_thresh;
  out->pos_thr = rp->position_convergence_thresh;
  out->min_assoc = rp->min_associations;
  return LOAMGPU_OK;
}

int reserve_register(loamgpu_ctx* ctx, uint32_t n_pairs, uint32_t capE, uint32_t capP, uint32_t detail_iters,
                     uint32_t nn_stride, uint32_t n_sets_or_0 = 0) {
  // sequence: every pair's target + the source of the last pair; explicit pairs: a target and a source set per pair
  const size_t n_sets = n_sets_or_0 ? n_sets_or_0 : (size_t)n_pairs + 1;
  CU(ctx->ge_hdr.reserve(n_sets * sizeof(BvhHdr)));
  CU(ctx->gp_hdr.reserve(n_sets * sizeof(BvhHdr)));
  CU(ctx->ge_nodes.reserve(n_sets * capE * sizeof(BvhNode)));
  CU(ctx->gp_nodes.reserve(n_sets * capP * sizeof(BvhNode)));
  CU(ctx->ge_quant.reserve(n_sets * sizeof(BvhQuant)));
  CU(ctx->gp_quant.reserve(n_sets * sizeof(BvhQuant)));
  CU(ctx->leftover.reserve(((size_t)n_pairs + 1) * 4));
  if (ctx->lm_moments) {  // short lists of planes the LM kernel's moment sums do not cover
    CU(ctx->nc_p.reserve((size_t)n_pairs * kLmNcCap * 32));
    CU(ctx->nc_a.reserve((size_t)n_pairs * kLmNcCap * 32));
  }
  CU(ctx->ge_aux.reserve(n_sets * capE * 4));
  CU(ctx->gp_aux.reserve(n_sets * capP * 4));
  // (+ kBvhLeaf records: a leaf scan loads whole groups of kBvhLeaf records behind a leaf's first point)
  CU(ctx->ge_sorted.reserve(n_sets * capE * 32 + kBvhLeaf * 32));
  CU(ctx->gp_sorted.reserve(n_sets * capP * 32 + kBvhLeaf * 32));
  CU(ctx->ge_keys.reserve(n_sets * capE * 16));
  CU(ctx->gp_keys.reserve(n_sets * capP * 16));
  CU(ctx->nn_idx.reserve((size_t)n_pairs * (capE + capP) * nn_stride * 4));
  CU(ctx->nn_cnt.reserve((size_t)n_pairs * (capE + capP) * 4));
  CU(ctx->active.reserve(((size_t)n_pairs + 1) * 4));
  CU(ctx->state.reserve((size_t)n_pairs * sizeof(PairState)));
  CU(ctx->rec_p.reserve((size_t)n_pairs * (capE + capP) * 32));
  CU(ctx->rec_a.reserve((size_t)n_pairs * (capE + capP) * 32));
  CU(ctx->rec_b.reserve((size_t)n_pairs * capE * 32));
  if (detail_iters) CU(ctx->nearest.reserve((size_t)detail_iters * n_pairs * (capE + capP) * 4));
  return LOAMGPU_OK;
}

BvhSetArrays bvh_arrays(loamgpu_ctx* ctx, bool planar, uint32_t cap) {
  BvhSetArrays g;
  g.hdr = (planar ? ctx->gp_hdr : ctx->ge_hdr).as<BvhHdr>();
  g.nodes = (planar ? ctx->gp_nodes : ctx->ge_nodes).as<BvhNode>();
  g.sorted = (planar ? ctx->gp_sorted : ctx->ge_sorted).as<double4>();
  g.keys = (planar ? ctx->gp_keys : ctx->ge_keys).as<uint2>();
  g.aux = (planar ? ctx->gp_aux : ctx->ge_aux).as<int>();
  g.pt_cap = cap;
  g.quant = (planar ? ctx->gp_quant : ctx->ge_quant).as<BvhQuant>();
  return g;
}

BvhSetArrays map_arrays(const loamgpu_map* m, int kind) {
  BvhSetArrays g;
  g.hdr = m->hdr[kind].as<BvhHdr>();
  g.nodes = m->nodes[kind].as<BvhNode>();
  g.sorted = m->sorted[kind].as<double4>();
  g.keys = m->keys[kind].as<uint2>();
  g.aux = nullptr;
  g.quant = nullptr;
  g.pt_cap = (uint32_t)std::max<uint64_t>(m->n[kind], 1);
  return g;
}

// (Re)build both NN structures of a map with the multi-CTA build.
int build_map(loamgpu_ctx* ctx, loamgpu_map* m) {
  const uint32_t nmax = (uint32_t)std::max<uint64_t>(std::max(m->n[0], m->n[1]), 1);
  CU(m->scratch.reserve(bvh_big_scratch_bytes(nmax)));
  for (int kind = 0; kind < 2; kind++) {
    const size_t cap = (size_t)std::max<uint64_t>(m->n[kind], 1);
    CU(m->hdr[kind].reserve(sizeof(BvhHdr)));
    CU(m->nodes[kind].reserve(cap * sizeof(BvhNode)));
    CU(m->sorted[kind].reserve(cap * 32 + kBvhLeaf * 32));
    CU(m->keys[kind].reserve(cap * 16));
    ProfScope ps(ctx, LOAMGPU_K_GRID);
    ctx->launches--;  // the launcher counts its own kernels
    CU(launch_bvh_build_big(m->pts[kind].as<double4>(), (uint32_t)m->n[kind], map_arrays(m, kind), m->scratch.p,
                            ctx->stream, &ctx->launches));
  }
  m->built = true;
  // a map may be used next through another context (stream) of the same device: finish the build before returning
  CU(cudaStreamSynchronize(ctx->stream));
  return LOAMGPU_OK;
}

// Register n_pairs pairs whose features sit in slots: tgt = (pair0+p) % n_slots, src = (pair0+p+src_offset) % n_slots.
// With `map` every pair registers onto the map instead (src_offset must be 0: set p / slot p = source of pair p).
int run_register(loamgpu_ctx* ctx, const RegP& rp, uint32_t n_pairs, uint64_t pair0, uint32_t n_slots, int src_offset,
                 uint32_t capE, uint32_t capP, const double* init_pose_dev, bool detail,
                 const loamgpu_map* map = nullptr, bool single_call = false, uint32_t n_sets_or_0 = 0,
                 uint32_t init_stride = 0) {
  // sets to build = feature slots pair0 .. pair0 + n_sets - 1 (without a map: + the last pair's source set)
  const uint32_t n_sets = n_sets_or_0 ? n_sets_or_0 : n_pairs + (map ? 0u : 1u);
  BvhBuildArgs gb, gbp;
  memset(&gb, 0, sizeof gb);
  gb.counts = ctx->feat_counts.as<uint32_t>();
  gb.slot0 = pair0;
  gb.n_slots = n_slots;
  gbp = gb;
  gb.pts = ctx->edge_pts.as<double4>();
  gb.pt_stride = capE;
  gb.kind = 0;
  gb.g = bvh_arrays(ctx, false, capE);
  gbp.pts = ctx->planar_pts.as<double4>();
  gbp.pt_stride = capP;
  gbp.kind = 1;
  gbp.g = bvh_arrays(ctx, true, capP);
  {  // edge and planar sets in one launch when both take the shared-memory build
    ProfScope ps(ctx, LOAMGPU_K_GRID);
    ctx->launches--;  // (the launcher counts its own kernels)
    CU(launch_bvh_build2(gb, gbp, n_sets, ctx->stream, &ctx->launches));
  }
  TIMED(LOAMGPU_K_MISC, launch_init_pairs(ctx->state.as<PairState>(), n_pairs, init_pose_dev, init_stride, ctx->stream));

  AssocArgs aa;
  memset(&aa, 0, sizeof aa);
  aa.edge_pts = ctx->edge_pts.as<double4>();
  aa.planar_pts = ctx->planar_pts.as<double4>();
  aa.feat_counts = ctx->feat_counts.as<uint32_t>();
  aa.capE_scan = capE;
  aa.capP_scan = capP;
  aa.pair0 = pair0;
  aa.n_slots = n_slots;
  aa.src_offset = src_offset;
  aa.ge = bvh_arrays(ctx, false, capE);
  aa.gp = bvh_arrays(ctx, true, capP);
  aa.state = ctx->state.as<PairState>();
  aa.rec_p = ctx->rec_p.as<double4>();
  aa.rec_a = ctx->rec_a.as<double4>();
  aa.rec_b = ctx->rec_b.as<double4>();
  aa.nearest = detail ? ctx->nearest.as<int32_t>() : nullptr;
  aa.nn_idx = ctx->nn_idx.as<uint32_t>();
  aa.nn_cnt = ctx->nn_cnt.as<uint32_t>();
  aa.leftover = ctx->leftover.as<uint32_t>();
  aa.nn_stride = (uint32_t)std::max(rp.ke, rp.kp);
  aa.morton_queries = ctx->morton_queries;
  if (map) {
    aa.ext_target = 1;
    aa.te = map_arrays(map, 0);
    aa.tp = map_arrays(map, 1);
    aa.te_pts = map->pts[0].as<double4>();
    aa.tp_pts = map->pts[1].as<double4>();
  }
  aa.rp = rp;
  LmArgs la;
  memset(&la, 0, sizeof la);
  la.state = aa.state;
  la.rec_p = aa.rec_p;
  la.rec_a = aa.rec_a;
  la.rec_b = aa.rec_b;
  la.feat_counts = aa.feat_counts;
  la.capE_scan = capE;
  la.capP_scan = capP;
  la.pair0 = pair0;
  la.n_slots = n_slots;
  la.src_offset = src_offset;
  la.rp = rp;
  if (ctx->lm_moments) {
    la.nc_p = ctx->nc_p.as<double4>();
    la.nc_a = ctx->nc_a.as<double4>();
    la.nc_cap = kLmNcCap;
  }
  // CTAs per pair in the LM kernel.  Sequence odometry keeps 1 whatever the chunk size (sums are then formed in the
  // same order for every chunking: results do not depend on it; measured: clusters of 2 / 4 are slower on full
  // chunks, 1.04 / 1.34 vs 0.82 ms); explicit single registrations spread the pair over a cluster of 8 CTAs, which
  // cuts the latency of the one call the reference API is made of ($LOAMGPU_LM_CLUSTER overrides both).
  la.cluster = ctx->lm_cluster ? ctx->lm_cluster : (single_call ? 8u : 1u);
  if (detail) {
    la.d_iter_est = ctx->det_est.as<double>();
    la.d_iter_update = ctx->det_upd.as<double>();
    la.d_assoc_n = ctx->det_assoc_n.as<uint32_t>();
    la.d_lm_iters = ctx->det_lm_iters.as<uint32_t>();
    la.d_lm_cost = ctx->det_lm_cost.as<double>();
  }
  aa.n_pairs = la.n_pairs = n_pairs;
  for (int it = 0; it < rp.max_iterations; it++) {
    aa.active = la.active = (it == 0 || single_call) ? nullptr : ctx->active.as<uint32_t>();
    TIMED(LOAMGPU_K_ASSOC, launch_assoc_knn(aa, n_pairs, it, ctx->stream));
    TIMED(LOAMGPU_K_FIT, launch_assoc_fit(aa, n_pairs, it, ctx->stream));
    la.outer_iter = it;
    TIMED(LOAMGPU_K_LM, launch_lm(la, n_pairs, ctx->stream));
    if (it + 1 >= rp.max_iterations) break;
    if (single_call) {
      // one pair: ask the device whether it is still iterating instead of launching up to max_iterations x 3 kernels
      // that would find nothing to do (the wait costs less than the idle launches; batches keep launching ahead)
      CU(ctx->pinned.reserve(256));
      volatile int32_t* status = reinterpret_cast<volatile int32_t*>(ctx->pinned.as<unsigned char>() + 192);
      *status = -1;
      CU(cudaMemcpyAsync(const_cast<int32_t*>(status), &ctx->state.as<PairState>()->status, sizeof(int32_t),
                         cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if (*status != -1) break;
      continue;
    }
    TIMED(LOAMGPU_K_MISC, launch_compact_active(aa.state, n_pairs, ctx->active.as<uint32_t>(), ctx->stream));
    // Idle iterations cost four empty launches each; with the default of 10 that is noise, but a caller may set a
    // large max_iterations as "run until converged" (valid in the reference).  Beyond 16 iterations the host asks, at
    // iterations 8, 16, 32, ..., how many pairs still iterate and stops enqueuing when none does.
    if (rp.max_iterations > 16 && it + 1 >= 8 && ((it + 1) & it) == 0) {
      CU(ctx->pinned.reserve(256));
      volatile uint32_t* left = reinterpret_cast<volatile uint32_t*>(ctx->pinned.as<unsigned char>() + 200);
      CU(cudaMemcpyAsync(const_cast<uint32_t*>(left), ctx->active.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if (*left == 0) break;
    }
  }
  return LOAMGPU_OK;
}

}  // namespace

// =============================================================================================== C ABI

extern "C" {

int loamgpu_create(int device, loamgpu_ctx** out) {
  if (!out) return LOAMGPU_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                   "); libloamgpu has no CPU fallback";
    return LOAMGPU_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_err = "device index out of range";
    return LOAMGPU_ERR_INVALID;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return LOAMGPU_ERR_CUDA;
  }
  loamgpu_ctx* c = new loamgpu_ctx();
  c->device = device;
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "cudaStreamCreate failed";
    delete c;
    return LOAMGPU_ERR_CUDA;
  }
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming);
  }
  c->stream = c->own_stream;
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c->mem_budget = std::min<uint64_t>(free_b / 3, 24ull << 30);
  }
  if (const char* qo = getenv("LOAMGPU_QUERY_ORDER")) c->morton_queries = strcmp(qo, "original") != 0;
  if (const char* lmm = getenv("LOAMGPU_LM_MOMENTS")) c->lm_moments = atoi(lmm) != 0;
  if (const char* v = getenv("LOAMGPU_LM_CLUSTER")) {
    const unsigned long cs = strtoul(v, nullptr, 10);
    c->lm_cluster = (cs == 1 || cs == 2 || cs == 4 || cs == 8) ? (uint32_t)cs : 0u;
  }
  if (const char* bt = getenv("LOAMGPU_BIG_TARGET_MIN")) c->big_target_min = strtoull(bt, nullptr, 10);
  *out = c;
  return LOAMGPU_OK;
}

void loamgpu_destroy(loamgpu_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&c->scan_in[0], &c->scan_in[1], &c->ring_edge, &c->ring_planar, &c->ring_counts, &c->edge_idx,
                    &c->planar_idx, &c->edge_pts, &c->planar_pts, &c->feat_counts, &c->ge_hdr, &c->ge_nodes,
                    &c->ge_sorted, &c->ge_keys, &c->ge_aux, &c->gp_hdr, &c->gp_nodes, &c->gp_sorted, &c->gp_keys, &c->gp_aux, &c->ge_quant, &c->gp_quant, &c->leftover, &c->nc_p, &c->nc_a, &c->state,
                    &c->rec_p, &c->rec_a, &c->rec_b, &c->nearest, &c->nn_idx, &c->nn_cnt, &c->active, &c->misc, &c->motions, &c->out_pose, &c->out_term,
                    &c->out_iters, &c->out_ne, &c->out_np, &c->det_est, &c->det_upd, &c->det_assoc_n,
                    &c->det_lm_iters, &c->det_lm_cost, &c->init_pose, &c->big_scratch, &c->reg_in};
  for (DevBuf* b : bufs) b->release();
  c->pinned.release();
  for (int i = 0; i < 2; i++) {
    if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
  }
  for (auto& r : c->prof_open) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

const char* loamgpu_last_error(const loamgpu_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int loamgpu_set_stream(loamgpu_ctx* c, void* s) {
  if (!c) return LOAMGPU_ERR_INVALID;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return LOAMGPU_OK;
}

void loamgpu_default_fe_params(loamgpu_fe_params* p) {  // features.h:37-66
  p->neighbor_points = 3;
  p->number_sectors = 6;
  p->max_edge_feats_per_sector = 10;
  p->max_planar_feats_per_sector = 50;
  p->edge_feat_threshold = 100.0;
  p->planar_feat_threshold = 1.0;
  p->occlusion_thresh = 0.5;
  p->parallel_thresh = 1.0;
}

void loamgpu_default_reg_params(loamgpu_reg_params* p) {  // registration.h:40-75
  p->num_edge_neighbors = 5;
  p->max_edge_neighbor_dist = 1.0;
  p->min_line_fit_points = 3;
  p->min_line_condition_number = 10;
  p->num_plane_neighbors = 5;
  p->max_plane_neighbor_dist = 2.0;
  p->min_plane_fit_points = 4;
  p->max_avg_point_plane_dist = 0.1;
  p->max_iterations = 10;
  p->rotation_convergence_thresh = 1e-3;
  p->position_convergence_thresh = 1e-2;
  p->min_associations = 100;
}

uint64_t loamgpu_launch_count(const loamgpu_ctx* c) { return c ? c->launches : 0; }

int loamgpu_set_profiling(loamgpu_ctx* c, int on) {
  if (!c) return LOAMGPU_ERR_INVALID;
  c->profiling = on != 0;
  return LOAMGPU_OK;
}

int loamgpu_kernel_times(loamgpu_ctx* ctx, double* ms, uint64_t* launches) {
  if (!ctx) return LOAMGPU_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof_open) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ctx->prof_ms[r.kind] += (double)t;
      ctx->prof_n[r.kind]++;
    }
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_open.clear();
  for (int k = 0; k < LOAMGPU_K_COUNT; k++) {
    if (ms) ms[k] = ctx->prof_ms[k];
    if (launches) launches[k] = ctx->prof_n[k];
    ctx->prof_ms[k] = 0;
    ctx->prof_n[k] = 0;
  }
  return LOAMGPU_OK;
}

int loamgpu_set_chunk_pairs(loamgpu_ctx* c, uint32_t pairs) {
  if (!c) return LOAMGPU_ERR_INVALID;
  c->chunk_pairs = pairs;  // 0 = automatic
  return LOAMGPU_OK;
}

// ------------------------------------------------------------------------------------------ extraction

static int extract_one(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                       const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, const double* motion,
                       uint32_t* edge_idx, uint64_t edge_cap, uint64_t* n_edge, uint32_t* planar_idx,
                       uint64_t planar_cap, uint64_t* n_planar, double* dewarped_xyz) {
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!n_edge || !n_planar) return fail(ctx, LOAMGPU_ERR_INVALID, "null count pointer");
  CU(cudaSetDevice(ctx->device));
  ExtractPlan pl;
  int rc = plan_extract(ctx, dtype, stride, n_points, lp, fe, &pl);
  if (rc) return rc;
  *n_edge = 0;
  *n_planar = 0;
  if (n_points == 0) return LOAMGPU_OK;
  if (!pts) return fail(ctx, LOAMGPU_ERR_INVALID, "null point buffer");
  if (motion) {
    for (int i = 0; i < 7; i++)
      if (!std::isfinite(motion[i])) return fail(ctx, LOAMGPU_ERR_INVALID, "start_T_end is not finite");
    // the de-warped ring is staged as doubles whatever the input type
    if (extract_smem_bytes(LOAMGPU_F64, pl.P, pl.S) > (size_t)ctx->max_smem_optin)
      return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "points_per_line too large for one CTA's shared memory");
  }
  const size_t bytes = (size_t)n_points * stride;
  CU(ctx->scan_in[0].reserve(bytes));
  ctx->staging_unguarded = true;
  rc = reserve_extract(ctx, pl, 1, 1);
  if (rc) return rc;
  double* d_dewarped = nullptr;
  if (motion && dewarped_xyz) {
    CU(ctx->misc.reserve((size_t)n_points * 24));
    d_dewarped = ctx->misc.as<double>();
  }
  CU(cudaMemcpyAsync(ctx->scan_in[0].p, pts, bytes, cudaMemcpyHostToDevice, ctx->stream));
  rc = run_extract(ctx, pl, ctx->scan_in[0].p, dtype, stride, lp, fe, 1, 0, 1, nullptr, nullptr, motion, d_dewarped);
  if (rc) return rc;
  if (d_dewarped)
    CU(cudaMemcpyAsync(dewarped_xyz, d_dewarped, (size_t)n_points * 24, cudaMemcpyDeviceToHost, ctx->stream));
  // counts and both index lists (full capacity) come back through page-locked staging with ONE wait; the lists are a
  // few tens of KB, cheaper to over-copy than to wait for the counts first
  const size_t eb = (size_t)pl.capE_scan * 4, pb = (size_t)pl.capP_scan * 4;
  CU(ctx->pinned.reserve(16 + eb + pb));
  unsigned char* h = ctx->pinned.as<unsigned char>();
  CU(cudaMemcpyAsync(h, ctx->feat_counts.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h + 16, ctx->edge_idx.p, eb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h + 16 + eb, ctx->planar_idx.p, pb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const uint32_t* counts = reinterpret_cast<const uint32_t*>(h);
  if (counts[0] > edge_cap || counts[1] > planar_cap || (counts[0] && !edge_idx) || (counts[1] && !planar_idx))
    return fail(ctx, LOAMGPU_ERR_INVALID, "output index buffers too small");
  if (counts[0]) memcpy(edge_idx, h + 16, 4 * (size_t)counts[0]);
  if (counts[1]) memcpy(planar_idx, h + 16 + eb, 4 * (size_t)counts[1]);
  *n_edge = counts[0];
  *n_planar = counts[1];
  return LOAMGPU_OK;
}

int loamgpu_extract(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                    const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, uint32_t* edge_idx,
                    uint64_t edge_cap, uint64_t* n_edge, uint32_t* planar_idx, uint64_t planar_cap,
                    uint64_t* n_planar) {
  API_RANGE();
  return extract_one(ctx, pts, dtype, stride, n_points, lp, fe, nullptr, edge_idx, edge_cap, n_edge, planar_idx,
                     planar_cap, n_planar, nullptr);
}

int loamgpu_extract_dewarped(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                             const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, const double* start_T_end,
                             uint32_t* edge_idx, uint64_t edge_cap, uint64_t* n_edge, uint32_t* planar_idx,
                             uint64_t planar_cap, uint64_t* n_planar, double* dewarped_xyz) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!start_T_end) return fail(ctx, LOAMGPU_ERR_INVALID, "null start_T_end");
  return extract_one(ctx, pts, dtype, stride, n_points, lp, fe, start_T_end, edge_idx, edge_cap, n_edge, planar_idx,
                     planar_cap, n_planar, dewarped_xyz);
}

static int curvature_or_mask(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                             const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, double* curv,
                             uint8_t* mask) {
  if (!ctx) return LOAMGPU_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  ExtractPlan pl;
  int rc = plan_extract(ctx, dtype, stride, n_points, lp, fe, &pl);
  if (rc) return rc;
  if (n_points == 0) return LOAMGPU_OK;
  if (!pts || (!curv && !mask)) return fail(ctx, LOAMGPU_ERR_INVALID, "null buffer");
  const size_t bytes = (size_t)n_points * stride;
  CU(ctx->scan_in[0].reserve(bytes));
  ctx->staging_unguarded = true;
  CU(ctx->misc.reserve((size_t)n_points * 9));
  CU(cudaMemcpyAsync(ctx->scan_in[0].p, pts, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ExtractArgs a;
  fill_extract_args(a, pl, ctx->scan_in[0].p, dtype, stride, lp, fe);
  double* d_curv = ctx->misc.as<double>();
  uint8_t* d_mask = ctx->misc.as<uint8_t>() + (size_t)n_points * 8;
  a.curv_out = curv ? d_curv : nullptr;
  a.mask_out = mask ? d_mask : nullptr;
  TIMED(LOAMGPU_K_EXTRACT, launch_extract(a, 1, ctx->stream));
  if (curv) CU(cudaMemcpyAsync(curv, d_curv, (size_t)n_points * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (mask) CU(cudaMemcpyAsync(mask, d_mask, (size_t)n_points, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LOAMGPU_OK;
}

int loamgpu_curvature(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                      const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, double* curvature) {
  API_RANGE();
  return curvature_or_mask(ctx, pts, dtype, stride, n_points, lp, fe, curvature, nullptr);
}

int loamgpu_valid_mask(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_points,
                       const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, uint8_t* mask) {
  API_RANGE();
  return curvature_or_mask(ctx, pts, dtype, stride, n_points, lp, fe, nullptr, mask);
}

// ------------------------------------------------------------------------------------------ registration

static void widen(const double* src, uint64_t n, std::vector<double>& dst, size_t off_pts) {
  for (uint64_t i = 0; i < n; i++) {
    double* d = dst.data() + (off_pts + i) * 4;
    d[0] = src[3 * i];
    d[1] = src[3 * i + 1];
    d[2] = src[3 * i + 2];
    d[3] = 0.0;
  }
}

// Shared body of loamgpu_register / loamgpu_register_to_map.  With `map` the target arguments are ignored.
static int register_core(loamgpu_ctx* ctx, const double* src_edge, uint64_t n_se, const double* src_planar, uint64_t n_sp,
                         const double* tgt_edge, uint64_t n_te, const double* tgt_planar, uint64_t n_tp,
                         const loamgpu_map* map, const double init_pose[7], const loamgpu_reg_params* params,
                         double out_pose[7], loamgpu_detail* detail) {
  RegP rp;
  int rc = make_regp(ctx, params, &rp);
  if (rc) return rc;
  if (map) n_te = n_tp = 0;
  const uint32_t capE = (uint32_t)std::max<uint64_t>(std::max(n_se, n_te), 1);
  const uint32_t capP = (uint32_t)std::max<uint64_t>(std::max(n_sp, n_tp), 1);
  const bool want_detail = detail != nullptr;
  const uint32_t det_iters = want_detail ? (uint32_t)std::max(rp.max_iterations, 1) : 0;
  // slots: 0 = target, 1 = source   (with a map: 0 = source)
  const uint32_t n_slots = map ? 1 : 2;
  CU(ctx->edge_pts.reserve((size_t)n_slots * capE * 32));
  CU(ctx->planar_pts.reserve((size_t)n_slots * capP * 32));
  CU(ctx->feat_counts.reserve(16));
  rc = reserve_register(ctx, 1, capE, capP, det_iters, (uint32_t)std::max(rp.ke, rp.kp));
  if (rc) return rc;
  CU(ctx->init_pose.reserve(7 * 8));
  CU(ctx->out_pose.reserve(7 * 8));
  CU(ctx->out_term.reserve(4));
  CU(ctx->out_iters.reserve(4));
  if (want_detail) {
    CU(ctx->det_est.reserve((size_t)det_iters * 7 * 8));
    CU(ctx->det_upd.reserve((size_t)det_iters * 7 * 8));
    CU(ctx->det_assoc_n.reserve((size_t)det_iters * 8));
    CU(ctx->det_lm_iters.reserve((size_t)det_iters * 4));
    CU(ctx->det_lm_cost.reserve((size_t)det_iters * 16));
    CU(cudaMemsetAsync(ctx->det_assoc_n.p, 0, (size_t)det_iters * 8, ctx->stream));
  }
  // Inputs: the four packed n x 3 clouds go up as they are (no host-side widening pass, 25 % fewer bytes) into a scratch
  // buffer and one small kernel widens them into the double4 feature slots; counts and the initial pose travel in one
  // small page-locked block.
  const uint64_t n_in[4] = {map ? 0 : n_te, map ? 0 : n_tp, n_se, n_sp};
  const double* in[4] = {tgt_edge, tgt_planar, src_edge, src_planar};
  uint64_t off[4], total = 0;
  for (int k = 0; k < 4; k++) {
    off[k] = total;
    total += n_in[k];
  }
  CU(ctx->reg_in.reserve(std::max<uint64_t>(total, 1) * 24));
  const size_t cap_src = (size_t)capE + capP;
  const size_t head_bytes = 128, det_rows = want_detail ? det_iters : 0;
  // page-locked block: [0,16) counts, [16,72) init pose | results: [128 ..) pose 56, term 4, iters 4, status 4, then
  // the detail rows (est, update, assoc counts, lm iterations, lm costs) and the nearest-neighbour table
  const size_t res_off = head_bytes, det_off = res_off + 128;
  const size_t det_bytes = det_rows * (56 + 56 + 8 + 4 + 16);
  const size_t near_off = det_off + ((det_bytes + 15) & ~(size_t)15);
  CU(ctx->pinned.reserve(near_off + det_rows * cap_src * 4 + 16));
  unsigned char* hpin = ctx->pinned.as<unsigned char>();
  {
    uint32_t* counts = reinterpret_cast<uint32_t*>(hpin);
    counts[0] = (uint32_t)(map ? n_se : n_te);
    counts[1] = (uint32_t)(map ? n_sp : n_tp);
    counts[2] = (uint32_t)n_se;
    counts[3] = (uint32_t)n_sp;
    memcpy(hpin + 16, init_pose, 56);
  }
  for (int k = 0; k < 4; k++)
    if (n_in[k])
      CU(cudaMemcpyAsync(ctx->reg_in.as<double>() + 3 * off[k], in[k], n_in[k] * 24, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->feat_counts.p, hpin, 16, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->init_pose.p, hpin + 16, 56, cudaMemcpyHostToDevice, ctx->stream));
  {
    WidenArgs wa;
    memset(&wa, 0, sizeof wa);
    double4* de = ctx->edge_pts.as<double4>();
    double4* dp = ctx->planar_pts.as<double4>();
    double4* dst[4] = {de, dp, map ? de : de + capE, map ? dp : dp + capP};  // slot 0 = target, slot 1 = source
    for (int k = 0; k < 4; k++) {
      wa.src[k] = ctx->reg_in.as<double>() + 3 * off[k];
      wa.dst[k] = dst[k];
      wa.n[k] = (uint32_t)n_in[k];
    }
    TIMED(LOAMGPU_K_MISC, launch_widen(wa, ctx->stream));
  }
  rc = run_register(ctx, rp, 1, 0, n_slots, map ? 0 : 1, capE, capP, ctx->init_pose.as<double>(), want_detail, map,
                    /*single_call=*/true);
  if (rc) return rc;
  TIMED(LOAMGPU_K_MISC, launch_finish_pairs(ctx->state.as<PairState>(), 1, ctx->out_pose.as<double>(),
                                            ctx->out_term.as<int32_t>(), ctx->out_iters.as<uint32_t>(), ctx->stream));
  // results (and, if asked for, every detail row of capacity) through the page-locked block with ONE wait
  CU(cudaMemcpyAsync(hpin + res_off, ctx->out_pose.p, 56, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hpin + res_off + 56, ctx->out_term.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hpin + res_off + 60, ctx->out_iters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  unsigned char* hdet = hpin + det_off;
  if (want_detail) {
    CU(cudaMemcpyAsync(hdet, ctx->det_est.p, det_rows * 56, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hdet + det_rows * 56, ctx->det_upd.p, det_rows * 56, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hdet + det_rows * 112, ctx->det_assoc_n.p, det_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hdet + det_rows * 120, ctx->det_lm_iters.p, det_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hdet + det_rows * 124, ctx->det_lm_cost.p, det_rows * 16, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out_pose, hpin + res_off, 56);
  int32_t term = 1;
  uint32_t iters = 0;
  memcpy(&term, hpin + res_off + 56, 4);
  memcpy(&iters, hpin + res_off + 60, 4);
  if (want_detail) {
    detail->n_iters = iters;
    detail->termination = term;
    const uint32_t rows = std::min<uint32_t>(iters, detail->max_iters_cap);
    if (rows) {
      if (detail->iter_est) memcpy(detail->iter_est, hdet, (size_t)rows * 56);
      if (detail->iter_update) memcpy(detail->iter_update, hdet + det_rows * 56, (size_t)rows * 56);
      if (detail->lm_iters) memcpy(detail->lm_iters, hdet + det_rows * 120, (size_t)rows * 4);
      if (detail->lm_cost) memcpy(detail->lm_cost, hdet + det_rows * 124, (size_t)rows * 16);
      // nearest-neighbour table: only the rows that were recorded (second, usually much smaller, wait)
      int32_t* nearest = reinterpret_cast<int32_t*>(hpin + near_off);
      CU(cudaMemcpyAsync(nearest, ctx->nearest.p, (size_t)rows * cap_src * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      for (uint32_t it = 0; it < rows; it++) {
        // association lists in source-index order (registration.cpp:59,100)
        uint32_t ne = 0, np = 0;
        const int32_t* row = nearest + (size_t)it * cap_src;
        for (uint64_t i = 0; i < n_se; i++)
          if (row[i] >= 0) {
            if (detail->edge_assoc && ne < detail->n_src_edge) {
              uint32_t* o = detail->edge_assoc + ((size_t)it * detail->n_src_edge + ne) * 2;
              o[0] = (uint32_t)i;
              o[1] = (uint32_t)row[i];
            }
            ne++;
          }
        for (uint64_t i = 0; i < n_sp; i++)
          if (row[capE + i] >= 0) {
            if (detail->plane_assoc && np < detail->n_src_planar) {
              uint32_t* o = detail->plane_assoc + ((size_t)it * detail->n_src_planar + np) * 2;
              o[0] = (uint32_t)i;
              o[1] = (uint32_t)row[capE + i];
            }
            np++;
          }
        if (detail->n_edge_assoc) detail->n_edge_assoc[it] = ne;
        if (detail->n_plane_assoc) detail->n_plane_assoc[it] = np;
      }
    }
  }
  return LOAMGPU_OK;
}

// Upload n x 3 doubles as double4 records behind `at` existing records of a map's point array.
static int map_upload(loamgpu_ctx* ctx, loamgpu_map* m, int kind, const double* pts, uint64_t n, uint64_t at) {
  if (n == 0) return LOAMGPU_OK;
  std::vector<double> h((size_t)n * 4);
  widen(pts, n, h, 0);
  CU(cudaMemcpyAsync(m->pts[kind].as<double4>() + at, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // `h` is pageable and dies here
  return LOAMGPU_OK;
}

// Grow a map's point array to `want` records, keeping the first `keep` (cudaMalloc'd buffers do not realloc).
static int map_grow(loamgpu_ctx* ctx, loamgpu_map* m, int kind, uint64_t keep, uint64_t want) {
  if (want * 32 <= m->pts[kind].cap) return LOAMGPU_OK;
  DevBuf nb;
  CU(nb.reserve((size_t)(want + want / 2) * 32));
  if (keep) CU(cudaMemcpyAsync(nb.p, m->pts[kind].p, (size_t)keep * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  m->pts[kind].release();
  m->pts[kind] = nb;
  return LOAMGPU_OK;
}

int loamgpu_register(loamgpu_ctx* ctx, const double* src_edge, uint64_t n_se, const double* src_planar, uint64_t n_sp,
                     const double* tgt_edge, uint64_t n_te, const double* tgt_planar, uint64_t n_tp,
                     const double init_pose[7], const loamgpu_reg_params* params, double out_pose[7],
                     loamgpu_detail* detail) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!init_pose || !out_pose) return fail(ctx, LOAMGPU_ERR_INVALID, "null pose pointer");
  if ((n_se && !src_edge) || (n_sp && !src_planar) || (n_te && !tgt_edge) || (n_tp && !tgt_planar))
    return fail(ctx, LOAMGPU_ERR_INVALID, "null feature buffer");
  if (std::max(std::max(n_se, n_sp), std::max(n_te, n_tp)) > 0x3FFFFFFFull)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "feature set too large");
  CU(cudaSetDevice(ctx->device));
  if (std::max(n_te, n_tp) >= ctx->big_target_min) {
    // large target (scan-to-map): multi-CTA NN build over a temporary device-resident map
    loamgpu_map* m = nullptr;
    int rc = loamgpu_map_create(ctx, tgt_edge, n_te, tgt_planar, n_tp, &m);
    if (rc) return rc;
    rc = register_core(ctx, src_edge, n_se, src_planar, n_sp, nullptr, 0, nullptr, 0, m, init_pose, params, out_pose,
                       detail);
    loamgpu_map_destroy(ctx, m);
    return rc;
  }
  return register_core(ctx, src_edge, n_se, src_planar, n_sp, tgt_edge, n_te, tgt_planar, n_tp, nullptr, init_pose, params,
                       out_pose, detail);
}

// ------------------------------------------------------------------------------------ explicit batches

int loamgpu_extract_batch(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride, uint64_t n_scans,
                          uint64_t n_points_per_scan, const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                          uint32_t* edge_idx,
                          uint64_t edge_cap, uint32_t* n_edge, uint32_t* planar_idx, uint64_t planar_cap,
                          uint32_t* n_planar) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!lp || !fe) return fail(ctx, LOAMGPU_ERR_INVALID, "null parameter struct");
  if (n_scans == 0) return LOAMGPU_OK;
  if (!n_edge || !n_planar) return fail(ctx, LOAMGPU_ERR_INVALID, "null count pointer");
  CU(cudaSetDevice(ctx->device));
  const uint64_t n_per = n_points_per_scan;
  ExtractPlan pl;
  int rc = plan_extract(ctx, dtype, stride, n_per, lp, fe, &pl);  // validateLidarScan: size must be lines x points
  if (rc) return rc;
  if (n_per == 0) {
    for (uint64_t s = 0; s < n_scans; s++) n_edge[s] = n_planar[s] = 0;
    return LOAMGPU_OK;
  }
  if (!pts) return fail(ctx, LOAMGPU_ERR_INVALID, "null point buffer");
  if (pl.capE_scan > edge_cap || pl.capP_scan > planar_cap || !edge_idx || !planar_idx)
    return fail(ctx, LOAMGPU_ERR_INVALID,
                "output index buffers too small: rows of scan_lines*number_sectors*(max_*_feats_per_sector+1) needed");
  const size_t scan_bytes = (size_t)n_per * stride;
  const uint32_t chunk = (uint32_t)std::min<uint64_t>(n_scans, ctx->chunk_pairs ? ctx->chunk_pairs : 256u);
  CU(ctx->scan_in[0].reserve((size_t)chunk * scan_bytes));
  ctx->staging_unguarded = true;
  rc = reserve_extract(ctx, pl, chunk, chunk);
  if (rc) return rc;
  std::vector<uint32_t> counts((size_t)chunk * 2);
  for (uint64_t s0 = 0; s0 < n_scans; s0 += chunk) {
    const uint32_t ns = (uint32_t)std::min<uint64_t>(chunk, n_scans - s0);
    CU(cudaMemcpyAsync(ctx->scan_in[0].p, (const unsigned char*)pts + s0 * scan_bytes, (size_t)ns * scan_bytes,
                       cudaMemcpyHostToDevice, ctx->stream));
    rc = run_extract(ctx, pl, ctx->scan_in[0].p, dtype, stride, lp, fe, ns, 0, chunk, nullptr, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(counts.data(), ctx->feat_counts.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, ctx->stream));
    // index rows: device row pitch = per-scan capacity, host row pitch = the caller's capacities
    CU(cudaMemcpy2DAsync(edge_idx + s0 * edge_cap, edge_cap * 4, ctx->edge_idx.p, (size_t)pl.capE_scan * 4,
                         (size_t)pl.capE_scan * 4, ns, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpy2DAsync(planar_idx + s0 * planar_cap, planar_cap * 4, ctx->planar_idx.p, (size_t)pl.capP_scan * 4,
                         (size_t)pl.capP_scan * 4, ns, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < ns; i++) {
      n_edge[s0 + i] = counts[2 * i];
      n_planar[s0 + i] = counts[2 * i + 1];
    }
  }
  return LOAMGPU_OK;
}

int loamgpu_register_pairs(loamgpu_ctx* ctx, uint64_t n_pairs, const double* src_edge, const uint64_t* n_src_edge,
                           const double* src_planar, const uint64_t* n_src_planar, const double* tgt_edge,
                           const uint64_t* n_tgt_edge, const double* tgt_planar, const uint64_t* n_tgt_planar,
                           const double* init_poses, const loamgpu_reg_params* params, double* out_poses,
                           int32_t* termination, uint32_t* iterations) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (n_pairs == 0) return LOAMGPU_OK;
  if (!n_src_edge || !n_src_planar || !n_tgt_edge || !n_tgt_planar || !out_poses)
    return fail(ctx, LOAMGPU_ERR_INVALID, "null count / output pointer");
  CU(cudaSetDevice(ctx->device));
  RegP rp;
  int rc = make_regp(ctx, params, &rp);
  if (rc) return rc;
  const uint64_t* cnt[4] = {n_tgt_edge, n_tgt_planar, n_src_edge, n_src_planar};
  const double* data[4] = {tgt_edge, tgt_planar, src_edge, src_planar};
  uint64_t maxE = 1, maxP = 1, total[4] = {0, 0, 0, 0};
  for (uint64_t p = 0; p < n_pairs; p++) {
    maxE = std::max(maxE, std::max(n_tgt_edge[p], n_src_edge[p]));
    maxP = std::max(maxP, std::max(n_tgt_planar[p], n_src_planar[p]));
    for (int a = 0; a < 4; a++) total[a] += cnt[a][p];
  }
  for (int a = 0; a < 4; a++)
    if (total[a] && !data[a]) return fail(ctx, LOAMGPU_ERR_INVALID, "null feature buffer");
  if (std::max(maxE, maxP) >= ctx->big_target_min)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "feature sets this large go through loamgpu_register / loamgpu_map_*");
  const uint32_t capE = (uint32_t)maxE, capP = (uint32_t)maxP;
  // pairs per chunk: bounded by the configured chunk size and by ~2 GB of feature slots
  const uint64_t per_pair_bytes = 2 * ((uint64_t)capE + capP) * 32;
  const uint32_t chunk = (uint32_t)std::max<uint64_t>(
      1, std::min<uint64_t>(std::min<uint64_t>(n_pairs, ctx->chunk_pairs ? ctx->chunk_pairs : 256u), (2ull << 30) / per_pair_bytes));
  const uint32_t n_slots = 2 * chunk;  // slots [0, chunk) = targets, [chunk, 2 chunk) = sources
  CU(ctx->edge_pts.reserve((size_t)n_slots * capE * 32));
  CU(ctx->planar_pts.reserve((size_t)n_slots * capP * 32));
  CU(ctx->feat_counts.reserve((size_t)n_slots * 8));
  rc = reserve_register(ctx, chunk, capE, capP, 0, (uint32_t)std::max(rp.ke, rp.kp), n_slots);
  if (rc) return rc;
  CU(ctx->init_pose.reserve((size_t)chunk * 56));
  CU(ctx->out_pose.reserve((size_t)chunk * 56));
  CU(ctx->out_term.reserve((size_t)chunk * 4));
  CU(ctx->out_iters.reserve((size_t)chunk * 4));
  std::vector<double> he((size_t)n_slots * capE * 4), hp((size_t)n_slots * capP * 4);
  std::vector<uint32_t> hc((size_t)n_slots * 2);
  uint64_t off[4] = {0, 0, 0, 0};  // running point offsets into the four concatenated clouds
  for (uint64_t p0 = 0; p0 < n_pairs; p0 += chunk) {
    const uint32_t np = (uint32_t)std::min<uint64_t>(chunk, n_pairs - p0);
    for (uint32_t i = 0; i < np; i++) {
      const uint64_t p = p0 + i;
      widen(tgt_edge ? tgt_edge + 3 * off[0] : nullptr, n_tgt_edge[p], he, (size_t)i * capE);
      widen(tgt_planar ? tgt_planar + 3 * off[1] : nullptr, n_tgt_planar[p], hp, (size_t)i * capP);
      widen(src_edge ? src_edge + 3 * off[2] : nullptr, n_src_edge[p], he, (size_t)(chunk + i) * capE);
      widen(src_planar ? src_planar + 3 * off[3] : nullptr, n_src_planar[p], hp, (size_t)(chunk + i) * capP);
      hc[2 * i] = (uint32_t)n_tgt_edge[p];
      hc[2 * i + 1] = (uint32_t)n_tgt_planar[p];
      hc[2 * (chunk + i)] = (uint32_t)n_src_edge[p];
      hc[2 * (chunk + i) + 1] = (uint32_t)n_src_planar[p];
      for (int a = 0; a < 4; a++) off[a] += cnt[a][p];
    }
    for (uint32_t i = np; i < chunk; i++) hc[2 * i] = hc[2 * i + 1] = hc[2 * (chunk + i)] = hc[2 * (chunk + i) + 1] = 0;
    CU(cudaMemcpyAsync(ctx->edge_pts.p, he.data(), he.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->planar_pts.p, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->feat_counts.p, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (init_poses)
      CU(cudaMemcpyAsync(ctx->init_pose.p, init_poses + 7 * p0, (size_t)np * 56, cudaMemcpyHostToDevice, ctx->stream));
    rc = run_register(ctx, rp, np, 0, n_slots, (int)chunk, capE, capP, init_poses ? ctx->init_pose.as<double>() : nullptr,
                      false, nullptr, false, n_slots, 7);
    if (rc) return rc;
    TIMED(LOAMGPU_K_MISC, launch_finish_pairs(ctx->state.as<PairState>(), np, ctx->out_pose.as<double>(),
                                              ctx->out_term.as<int32_t>(), ctx->out_iters.as<uint32_t>(), ctx->stream));
    CU(cudaMemcpyAsync(out_poses + 7 * p0, ctx->out_pose.p, (size_t)np * 56, cudaMemcpyDeviceToHost, ctx->stream));
    if (termination) CU(cudaMemcpyAsync(termination + p0, ctx->out_term.p, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (iterations) CU(cudaMemcpyAsync(iterations + p0, ctx->out_iters.p, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // the host staging vectors are reused by the next chunk
  }
  return LOAMGPU_OK;
}

// ------------------------------------------------------------------------------------ device-resident local map

int loamgpu_map_create(loamgpu_ctx* ctx, const double* edge, uint64_t n_edge, const double* planar, uint64_t n_planar,
                       loamgpu_map** out) {
  API_RANGE();
  if (!ctx || !out) return LOAMGPU_ERR_INVALID;
  *out = nullptr;
  if ((n_edge && !edge) || (n_planar && !planar)) return fail(ctx, LOAMGPU_ERR_INVALID, "null feature buffer");
  if (std::max(n_edge, n_planar) > 0x3FFFFFFFull) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "feature set too large");
  CU(cudaSetDevice(ctx->device));
  loamgpu_map* m = new loamgpu_map();
  m->device = ctx->device;
  int rc = map_grow(ctx, m, 0, 0, std::max<uint64_t>(n_edge, 1));
  if (!rc) rc = map_grow(ctx, m, 1, 0, std::max<uint64_t>(n_planar, 1));
  if (!rc) rc = map_upload(ctx, m, 0, edge, n_edge, 0);
  if (!rc) rc = map_upload(ctx, m, 1, planar, n_planar, 0);
  if (!rc) {
    m->n[0] = n_edge;
    m->n[1] = n_planar;
    rc = build_map(ctx, m);
  }
  if (rc) {
    loamgpu_map_destroy(ctx, m);
    return rc;
  }
  *out = m;
  return LOAMGPU_OK;
}

void loamgpu_map_destroy(loamgpu_ctx* ctx, loamgpu_map* m) {
  if (!m) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  for (int k = 0; k < 2; k++) {
    m->pts[k].release();
    m->hdr[k].release();
    m->nodes[k].release();
    m->sorted[k].release();
    m->keys[k].release();
  }
  m->scratch.release();
  delete m;
}

int loamgpu_map_size(const loamgpu_map* m, uint64_t* n_edge, uint64_t* n_planar) {
  if (!m) return LOAMGPU_ERR_INVALID;
  if (n_edge) *n_edge = m->n[0];
  if (n_planar) *n_planar = m->n[1];
  return LOAMGPU_OK;
}

int loamgpu_map_update(loamgpu_ctx* ctx, loamgpu_map* m, const double* edge, uint64_t n_edge, const double* planar,
                       uint64_t n_planar, const double pose[7], uint64_t max_edge, uint64_t max_planar) {
  API_RANGE();
  if (!ctx || !m) return LOAMGPU_ERR_INVALID;
  if ((n_edge && !edge) || (n_planar && !planar)) return fail(ctx, LOAMGPU_ERR_INVALID, "null feature buffer");
  if (m->device != ctx->device) return fail(ctx, LOAMGPU_ERR_INVALID, "map belongs to another device");
  CU(cudaSetDevice(ctx->device));
  const double* add[2] = {edge, planar};
  const uint64_t n_add[2] = {n_edge, n_planar}, cap[2] = {max_edge, max_planar};
  // the point arrays change below: until build_map has succeeded again the old NN structures describe nothing
  // (a failure on the way leaves the map unusable for registration instead of walking a stale tree)
  m->built = false;
  if (pose) {
    CU(ctx->init_pose.reserve(56));
    CU(cudaMemcpyAsync(ctx->init_pose.p, pose, 56, cudaMemcpyHostToDevice, ctx->stream));
  }
  for (int k = 0; k < 2; k++) {
    const uint64_t total = m->n[k] + n_add[k];
    if (total > 0x3FFFFFFFull) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "feature set too large");
    int rc = map_grow(ctx, m, k, m->n[k], std::max<uint64_t>(total, 1));
    if (!rc) rc = map_upload(ctx, m, k, add[k], n_add[k], m->n[k]);
    if (rc) return rc;
    if (pose && n_add[k])
      TIMED(LOAMGPU_K_MISC, launch_transform_points(m->pts[k].as<double4>() + m->n[k], (uint32_t)n_add[k],
                                                    ctx->init_pose.as<double>(), ctx->stream));
    uint64_t n_new = total;
    if (cap[k] && total > cap[k]) {  // sliding window: the oldest points leave
      const uint64_t drop = total - cap[k];
      n_new = cap[k];
      // overlapping move toward the front: stage through the sort scratch of this kind
      CU(m->sorted[k].reserve((size_t)n_new * 32));
      CU(cudaMemcpyAsync(m->sorted[k].p, m->pts[k].as<double4>() + drop, (size_t)n_new * 32, cudaMemcpyDeviceToDevice,
                         ctx->stream));
      CU(cudaMemcpyAsync(m->pts[k].p, m->sorted[k].p, (size_t)n_new * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    m->n[k] = n_new;
  }
  return build_map(ctx, m);
}

int loamgpu_register_to_map(loamgpu_ctx* ctx, const loamgpu_map* map, const double* src_edge, uint64_t n_se,
                            const double* src_planar, uint64_t n_sp, const double init_pose[7],
                            const loamgpu_reg_params* params, double out_pose[7], loamgpu_detail* detail) {
  API_RANGE();
  if (!ctx || !map) return LOAMGPU_ERR_INVALID;
  if (!init_pose || !out_pose) return fail(ctx, LOAMGPU_ERR_INVALID, "null pose pointer");
  if ((n_se && !src_edge) || (n_sp && !src_planar)) return fail(ctx, LOAMGPU_ERR_INVALID, "null feature buffer");
  if (std::max(n_se, n_sp) > 0x3FFFFFFFull) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "feature set too large");
  if (map->device != ctx->device) return fail(ctx, LOAMGPU_ERR_INVALID, "map belongs to another device");
  if (!map->built) return fail(ctx, LOAMGPU_ERR_INVALID, "map has no NN structure");
  CU(cudaSetDevice(ctx->device));
  return register_core(ctx, src_edge, n_se, src_planar, n_sp, nullptr, 0, nullptr, 0, map, init_pose, params, out_pose,
                       detail);
}

int loamgpu_knn(loamgpu_ctx* ctx, const double* targets, uint64_t n_t, const double* queries, uint64_t n_q, uint32_t k,
                double max_dist, uint32_t* idx_out, uint32_t* count_out) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (k == 0 || k > (uint32_t)kKnnMax) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "k must be in [1, 32]");
  if ((n_t && !targets) || (n_q && (!queries || !idx_out || !count_out)))
    return fail(ctx, LOAMGPU_ERR_INVALID, "null buffer");
  if (n_t > 0x7FFFFFFFull) return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "target set too large");
  if (n_q == 0) return LOAMGPU_OK;
  CU(cudaSetDevice(ctx->device));
  const uint32_t cap = (uint32_t)std::max<uint64_t>(n_t, 1);
  CU(ctx->planar_pts.reserve((size_t)cap * 32));
  CU(ctx->feat_counts.reserve(8));
  int rc = reserve_register(ctx, 1, 1, cap, 0, 1);
  if (rc) return rc;
  CU(ctx->misc.reserve((size_t)n_q * (24 + 4 * (size_t)k + 4)));
  std::vector<double> hp((size_t)cap * 4);
  widen(targets, n_t, hp, 0);
  const uint32_t counts[2] = {0, (uint32_t)n_t};
  CU(cudaMemcpyAsync(ctx->planar_pts.p, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->feat_counts.p, counts, 8, cudaMemcpyHostToDevice, ctx->stream));
  BvhBuildArgs gb;
  memset(&gb, 0, sizeof gb);
  gb.counts = ctx->feat_counts.as<uint32_t>();
  gb.slot0 = 0;
  gb.n_slots = 1;
  gb.pts = ctx->planar_pts.as<double4>();
  gb.pt_stride = cap;
  gb.kind = 1;
  gb.g = bvh_arrays(ctx, true, cap);
  if (n_t >= ctx->big_target_min) {
    CU(ctx->big_scratch.reserve(bvh_big_scratch_bytes((uint32_t)n_t)));
    ProfScope ps(ctx, LOAMGPU_K_GRID);
    ctx->launches--;  // the launcher counts its own kernels
    CU(launch_bvh_build_big(gb.pts, (uint32_t)n_t, gb.g, ctx->big_scratch.p, ctx->stream, &ctx->launches));
  } else {
    TIMED(LOAMGPU_K_GRID, launch_bvh_build(gb, 1, ctx->stream));
  }
  double* dq = ctx->misc.as<double>();
  uint32_t* didx = reinterpret_cast<uint32_t*>(dq + 3 * n_q);
  uint32_t* dcnt = didx + (size_t)n_q * k;
  CU(cudaMemcpyAsync(dq, queries, (size_t)n_q * 24, cudaMemcpyHostToDevice, ctx->stream));
  KnnArgs ka;
  ka.queries = dq;
  ka.n_queries = n_q;
  ka.g = gb.g;
  ka.k = (int)k;
  ka.max_dist = max_dist;
  ka.idx_out = didx;
  ka.count_out = dcnt;
  TIMED(LOAMGPU_K_ASSOC, launch_knn(ka, ctx->stream));
  CU(cudaMemcpyAsync(idx_out, didx, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(count_out, dcnt, (size_t)n_q * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LOAMGPU_OK;
}

// TEST HOOK: one evaluation of the LM kernel on explicit residual blocks (what associateEdges / associatePlanes hand to
// Ceres, registration.cpp:52-57,93-98) at an arbitrary iterate: out[0..20] upper triangle of J^T J (tangent, loss-
// corrected), out[21..26] J^T r, out[27] cost, out[28] = 1 if the moment sums were used, out[29] = planes left to the
// streamed part.  mode 0: streamed evaluation; mode 1: moment path (lm_kernel's default).
int loamgpu_debug_problem_eval(loamgpu_ctx* ctx, uint64_t n_edge, const double* edge_p, const double* edge_a,
                               const double* edge_b, uint64_t n_plane, const double* plane_p, const double* plane_n,
                               const double* plane_d, const double x[7], int mode, double out[30]) {
  if (!ctx || !x || !out) return LOAMGPU_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  const uint32_t capE = (uint32_t)std::max<uint64_t>(n_edge, 1), capP = (uint32_t)std::max<uint64_t>(n_plane, 1);
  const size_t cap = (size_t)capE + capP;
  CU(ctx->rec_p.reserve(cap * 32));
  CU(ctx->rec_a.reserve(cap * 32));
  CU(ctx->rec_b.reserve((size_t)capE * 32));
  CU(ctx->nc_p.reserve((size_t)kLmNcCap * 32));
  CU(ctx->nc_a.reserve((size_t)kLmNcCap * 32));
  CU(ctx->feat_counts.reserve(16));
  CU(ctx->misc.reserve(64 * 8));
  std::vector<double> hp(cap * 4, 0.0), ha(cap * 4, 0.0), hb((size_t)capE * 4, 0.0);
  for (uint64_t i = 0; i < n_edge; i++) {
    for (int k = 0; k < 3; k++) {
      hp[4 * i + k] = edge_p[3 * i + k];
      ha[4 * i + k] = edge_a[3 * i + k];
      hb[4 * i + k] = edge_b[3 * i + k];
    }
    hp[4 * i + 3] = 1.0;
  }
  for (uint64_t i = 0; i < n_plane; i++) {
    const size_t r = capE + i;
    for (int k = 0; k < 3; k++) {
      hp[4 * r + k] = plane_p[3 * i + k];
      ha[4 * r + k] = plane_n[3 * i + k];
    }
    hp[4 * r + 3] = 2.0;
    ha[4 * r + 3] = plane_d[i];
  }
  const uint32_t counts[4] = {(uint32_t)n_edge, (uint32_t)n_plane, 0, 0};
  CU(cudaMemcpyAsync(ctx->rec_p.p, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->rec_a.p, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->rec_b.p, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->feat_counts.p, counts, 16, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->misc.p, x, 56, cudaMemcpyHostToDevice, ctx->stream));
  LmArgs la;
  memset(&la, 0, sizeof la);
  la.rec_p = ctx->rec_p.as<double4>();
  la.rec_a = ctx->rec_a.as<double4>();
  la.rec_b = ctx->rec_b.as<double4>();
  la.feat_counts = ctx->feat_counts.as<uint32_t>();
  la.capE_scan = capE;
  la.capP_scan = capP;
  la.nc_p = ctx->nc_p.as<double4>();
  la.nc_a = ctx->nc_a.as<double4>();
  la.nc_cap = kLmNcCap;
  TIMED(LOAMGPU_K_LM, launch_lm_debug_eval(la, ctx->misc.as<double>(), mode, ctx->misc.as<double>() + 8, ctx->stream));
  CU(cudaMemcpyAsync(out, ctx->misc.as<double>() + 8, 30 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LOAMGPU_OK;
}

// ------------------------------------------------------------------------------------------ sequence odometry

}  // extern "C"

// Core of both odometry entry points.  `fetch(c0, n, buf)` must make scans [c0, c0+n) available on the device and
// return their device pointer (the _device variant just offsets; the _host variant copies on the copy stream).
// Pairs per chunk of a sequence call.  Measured on 64x1024 (DESIGN.md "Host pipeline"): device-resident scans and
// asynchronous host calls gain from large chunks (the one-CTA-per-pair grids fill whole waves, fewer idle tail
// launches: 35.8k scans/s at 256 pairs, 38.0k at 592, 39.8k at 1023), a synchronous host call is fastest with 256
// (its first copy is exposed and grows with the chunk).  Bounded by a share of the free device memory.
enum OdometryMode { kResident = 0, kHostSync = 1, kHostAsync = 2 };

static uint32_t pick_chunk(loamgpu_ctx* ctx, OdometryMode mode, uint64_t n_scans, uint64_t n_per, uint32_t capE,
                           uint32_t capP, uint32_t nn_stride, size_t pt_stride) {
  const uint64_t n_pairs_max = std::max<uint64_t>(n_scans, 2) - 1;
  if (ctx->chunk_pairs) return (uint32_t)std::min<uint64_t>(ctx->chunk_pairs, n_pairs_max);
  // Asynchronous host calls: the staging buffers alternate ACROSS calls, so even a call made of one chunk copies into
  // the buffer the previous call is not reading and its copy runs under the previous call's kernels: full-size chunks
  // (measured, 12 calls of 1024 float4 scans: 44.7 k scans/s with 1024-pair chunks, 42.3 k with 512, 40.2 k with 256).
  // A synchronous host call exposes the copy of its first chunk and is fastest with 256.
  uint64_t want = mode == kHostSync ? 256 : 1024;
  const uint64_t cap = (uint64_t)capE + capP;
  // per pair: k-NN lists, residual records, two NN structures (nodes, sorted copy, sort keys, flags), feature slots
  // (indices + widened points), ring pick lists, and for host calls two staging copies of the scan
  const uint64_t per_pair = cap * ((uint64_t)nn_stride * 4 + 4 + 96 + 84 + 36 + 4 ) +
                            (mode == kResident ? 0 : 2 * n_per * pt_stride);
  if (ctx->mem_budget && per_pair)  // (queried once at context creation: cudaMemGetInfo per call is far too slow)
    want = std::min<uint64_t>(want, std::max<uint64_t>(32, ctx->mem_budget / per_pair));
  return (uint32_t)std::min<uint64_t>(want, n_pairs_max);
}

static int chunk_for(loamgpu_ctx* ctx, OdometryMode mode, uint64_t n_scans, const loamgpu_lidar_params* lp,
                     const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, size_t pt_stride, uint32_t* chunk) {
  ExtractPlan pl;
  RegP rp;
  const uint64_t n_per = lp->scan_lines * lp->points_per_line;
  int rc = plan_extract(ctx, LOAMGPU_F32, pt_stride, n_per, lp, fe, &pl);
  if (rc) return rc;
  rc = make_regp(ctx, reg, &rp);
  if (rc) return rc;
  *chunk = pick_chunk(ctx, mode, n_scans, n_per, pl.capE_scan, pl.capP_scan, (uint32_t)std::max(rp.ke, rp.kp), pt_stride);
  return LOAMGPU_OK;
}

template <typename Fetch>
static int odometry_core(loamgpu_ctx* ctx, uint64_t n_scans, const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                         const loamgpu_reg_params* reg, double* poses_dev, int32_t* term_dev, uint32_t* iters_dev,
                         uint32_t* ne_dev, uint32_t* np_dev, OdometryMode mode, uint32_t chunk, bool short_lead,
                         Fetch fetch, const double* motions_dev = nullptr, size_t pt_stride = 16) {
  ExtractPlan pl;
  const uint64_t n_per = lp->scan_lines * lp->points_per_line;
  int rc = plan_extract(ctx, LOAMGPU_F32, pt_stride, n_per, lp, fe, &pl);
  if (rc) return rc;
  if (n_per == 0) return fail(ctx, LOAMGPU_ERR_INVALID, "empty scans");
  if (motions_dev && extract_smem_bytes(LOAMGPU_F64, pl.P, pl.S) > (size_t)ctx->max_smem_optin)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "points_per_line too large for one CTA's shared memory");
  RegP rp;
  rc = make_regp(ctx, reg, &rp);
  if (rc) return rc;
  const uint32_t lead0 = []() { const char* e = getenv("LOAMGPU_LEAD"); return e ? (uint32_t)atoi(e) : 128u; }();
  // a host call whose first copy has nothing to hide behind starts with a shorter chunk (see below)
  (void)mode;
  const uint32_t lead = !short_lead ? chunk : std::max<uint32_t>(1, std::min<uint32_t>(chunk, lead0));
  const uint32_t n_slots = chunk + 1;
  rc = reserve_extract(ctx, pl, chunk + 1, n_slots);
  if (rc) return rc;
  rc = reserve_register(ctx, chunk, pl.capE_scan, pl.capP_scan, 0, (uint32_t)std::max(rp.ke, rp.kp));
  if (rc) return rc;
  if (n_scans == 1) {  // no pair: just the feature counts
    const float* d = nullptr;
    rc = fetch(0, 1, 0, &d);
    if (rc) return rc;
    return run_extract(ctx, pl, d, LOAMGPU_F32, pt_stride, lp, fe, 1, 0, n_slots, ne_dev, np_dev, nullptr, nullptr, motions_dev);
  }
  const uint64_t n_pairs = n_scans - 1;
  // the two staging buffers alternate across calls too: an asynchronous call made of ONE chunk then still copies into
  // the buffer the previous call is not reading, i.e. under the previous call's kernels
  int buf = ctx->stage_buf;
  // A synchronous host call starts with a shorter chunk (`lead` pairs, then x`ramp` up to the full size): the H2D copy
  // of the first chunk is the only one that cannot hide behind kernels.  Copying a scan takes ~0.6x the time of
  // processing it (55 GB/s measured), so longer ramps expose more than they save; 128 -> 256 measured best
  // ($LOAMGPU_LEAD / $LOAMGPU_RAMP override).  Device-resident and asynchronous host calls use full chunks
  // throughout (an asynchronous call's first copy hides behind the previous call's kernels).
  const double ramp = []() { const char* e = getenv("LOAMGPU_RAMP"); return e ? std::max(1.0, atof(e)) : 2.0; }();
  uint32_t np = 0, want = lead;
  for (uint64_t p0 = 0; p0 < n_pairs; p0 += np, buf ^= 1) {
    np = (uint32_t)std::min<uint64_t>(want, n_pairs - p0);
    want = std::min<uint32_t>(chunk, (uint32_t)std::ceil(want * ramp));
    // scans needed: p0 .. p0+np ; scan p0 is already in its slot except for the first chunk
    const uint64_t s0 = p0 == 0 ? 0 : p0 + 1;
    const uint32_t ns = (uint32_t)(p0 + np + 1 - s0);
    const float* d = nullptr;
    rc = fetch(s0, ns, buf, &d);
    if (rc) return rc;
    rc = run_extract(ctx, pl, d, LOAMGPU_F32, pt_stride, lp, fe, ns, s0, n_slots, ne_dev, np_dev, nullptr, nullptr,
                     motions_dev ? motions_dev + 7 * s0 : nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_consumed[buf], ctx->stream));
    rc = run_register(ctx, rp, np, p0, n_slots, 1, pl.capE_scan, pl.capP_scan, nullptr, false);
    if (rc) return rc;
    TIMED(LOAMGPU_K_MISC,
          launch_finish_pairs(ctx->state.as<PairState>(), np, poses_dev ? poses_dev + 7 * p0 : nullptr,
                              term_dev ? term_dev + p0 : nullptr, iters_dev ? iters_dev + p0 : nullptr, ctx->stream));
  }
  ctx->stage_buf = buf;
  return LOAMGPU_OK;
}

extern "C" {

static int odometry_device_impl(loamgpu_ctx* ctx, const void* scans_dev, size_t pt_stride, uint64_t n_scans,
                                const double* motions_dev, const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                const loamgpu_reg_params* reg, double* poses_dev, int32_t* term_dev, uint32_t* iters_dev,
                                uint32_t* ne_dev, uint32_t* np_dev) {
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!lp || !fe || !reg) return fail(ctx, LOAMGPU_ERR_INVALID, "null parameter struct");
  if (pt_stride != 12 && pt_stride != 16)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "sequence calls take float records of 12 (x y z) or 16 (x y z .) bytes");
  if (n_scans == 0) return LOAMGPU_OK;
  if (!scans_dev) return fail(ctx, LOAMGPU_ERR_INVALID, "null scan buffer");
  CU(cudaSetDevice(ctx->device));
  const uint64_t n_per = lp->scan_lines * lp->points_per_line;
  auto fetch = [&](uint64_t s0, uint32_t, int, const float** out) {
    *out = reinterpret_cast<const float*>(static_cast<const unsigned char*>(scans_dev) + s0 * n_per * pt_stride);
    return (int)LOAMGPU_OK;
  };
  uint32_t chunk = 0;
  const int rc = chunk_for(ctx, kResident, n_scans, lp, fe, reg, pt_stride, &chunk);
  if (rc) return rc;
  return odometry_core(ctx, n_scans, lp, fe, reg, poses_dev, term_dev, iters_dev, ne_dev, np_dev, kResident, chunk,
                       /*short_lead=*/false, fetch, motions_dev, pt_stride);
}

int loamgpu_odometry_device(loamgpu_ctx* ctx, const float* scans_dev, uint64_t n_scans, const loamgpu_lidar_params* lp,
                            const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses_dev,
                            int32_t* term_dev, uint32_t* iters_dev, uint32_t* ne_dev, uint32_t* np_dev) {
  API_RANGE();
  return odometry_device_impl(ctx, scans_dev, 16, n_scans, nullptr, lp, fe, reg, poses_dev, term_dev, iters_dev, ne_dev,
                              np_dev);
}

int loamgpu_odometry_device_strided(loamgpu_ctx* ctx, const void* scans_dev, size_t stride_bytes, uint64_t n_scans,
                                    const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                    const loamgpu_reg_params* reg, double* poses_dev, int32_t* term_dev,
                                    uint32_t* iters_dev, uint32_t* ne_dev, uint32_t* np_dev) {
  API_RANGE();
  return odometry_device_impl(ctx, scans_dev, stride_bytes, n_scans, nullptr, lp, fe, reg, poses_dev, term_dev, iters_dev,
                              ne_dev, np_dev);
}

int loamgpu_odometry_device_dewarped(loamgpu_ctx* ctx, const float* scans_dev, uint64_t n_scans,
                                     const double* start_T_end_dev, const loamgpu_lidar_params* lp,
                                     const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses_dev,
                                     int32_t* term_dev, uint32_t* iters_dev, uint32_t* ne_dev, uint32_t* np_dev) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (n_scans && !start_T_end_dev) return fail(ctx, LOAMGPU_ERR_INVALID, "null start_T_end");
  return odometry_device_impl(ctx, scans_dev, 16, n_scans, start_T_end_dev, lp, fe, reg, poses_dev, term_dev, iters_dev,
                              ne_dev, np_dev);
}

int loamgpu_synchronize(loamgpu_ctx* ctx) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LOAMGPU_OK;
}

}  // extern "C"

static int odometry_host_impl(loamgpu_ctx* ctx, OdometryMode mode, const void* scans, uint64_t n_scans,
                              const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, const loamgpu_reg_params* reg,
                              double* poses, int32_t* termination, uint32_t* iterations, uint32_t* n_edge,
                              uint32_t* n_planar, const double* motions = nullptr, size_t pt_stride = 16);

extern "C" {

int loamgpu_odometry_host(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans, const loamgpu_lidar_params* lp,
                          const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                          uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar) {
  API_RANGE();
  const int rc = odometry_host_impl(ctx, kHostSync, scans, n_scans, lp, fe, reg, poses, termination, iterations, n_edge,
                                    n_planar);
  if (rc) return rc;
  return loamgpu_synchronize(ctx);
}

int loamgpu_odometry_host_dewarped(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans, const double* start_T_end,
                                   const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                   const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                   uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar) {
  API_RANGE();
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (n_scans && !start_T_end) return fail(ctx, LOAMGPU_ERR_INVALID, "null start_T_end");
  const int rc = odometry_host_impl(ctx, kHostSync, scans, n_scans, lp, fe, reg, poses, termination, iterations, n_edge,
                                    n_planar, start_T_end);
  if (rc) return rc;
  return loamgpu_synchronize(ctx);
}

int loamgpu_odometry_host_async(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans, const loamgpu_lidar_params* lp,
                                const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses,
                                int32_t* termination, uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar) {
  API_RANGE();
  return odometry_host_impl(ctx, kHostAsync, scans, n_scans, lp, fe, reg, poses, termination, iterations, n_edge,
                            n_planar);
}

int loamgpu_odometry_host_strided(loamgpu_ctx* ctx, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                  const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                  const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                  uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar) {
  API_RANGE();
  const int rc = odometry_host_impl(ctx, kHostSync, scans, n_scans, lp, fe, reg, poses, termination, iterations, n_edge,
                                    n_planar, nullptr, stride_bytes);
  if (rc) return rc;
  return loamgpu_synchronize(ctx);
}

int loamgpu_odometry_host_async_strided(loamgpu_ctx* ctx, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                        const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                        const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                        uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar) {
  API_RANGE();
  return odometry_host_impl(ctx, kHostAsync, scans, n_scans, lp, fe, reg, poses, termination, iterations, n_edge,
                            n_planar, nullptr, stride_bytes);
}

}  // extern "C"

static int odometry_host_impl(loamgpu_ctx* ctx, OdometryMode mode, const void* scans_v, uint64_t n_scans,
                              const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe, const loamgpu_reg_params* reg,
                              double* poses, int32_t* termination, uint32_t* iterations, uint32_t* n_edge,
                              uint32_t* n_planar, const double* motions, size_t pt_stride) {
  if (!ctx) return LOAMGPU_ERR_INVALID;
  if (!lp || !fe || !reg) return fail(ctx, LOAMGPU_ERR_INVALID, "null parameter struct");
  if (pt_stride != 12 && pt_stride != 16)
    return fail(ctx, LOAMGPU_ERR_UNSUPPORTED, "sequence calls take float records of 12 (x y z) or 16 (x y z .) bytes");
  const unsigned char* scans = static_cast<const unsigned char*>(scans_v);
  if (n_scans == 0) return LOAMGPU_OK;
  if (!scans) return fail(ctx, LOAMGPU_ERR_INVALID, "null scan buffer");
  CU(cudaSetDevice(ctx->device));
  const uint64_t n_per = lp->scan_lines * lp->points_per_line;
  const size_t scan_bytes = (size_t)n_per * pt_stride;
  const uint64_t n_pairs = n_scans - 1;
  uint32_t chunk = 0;
  {
    const int rc = chunk_for(ctx, mode, n_scans, lp, fe, reg, pt_stride, &chunk);
    if (rc) return rc;
  }
  // The first copy of a call is exposed unless kernels of a previous asynchronous call are still running: a
  // synchronous call, and an asynchronous call that finds the compute stream idle, start with a short lead chunk.
  const bool short_lead = mode == kHostSync || cudaStreamQuery(ctx->stream) == cudaSuccess;
  for (int b = 0; b < 2; b++) CU(ctx->scan_in[b].reserve((size_t)(chunk + 1) * scan_bytes));
  CU(ctx->out_pose.reserve(std::max<uint64_t>(n_pairs, 1) * 56));
  CU(ctx->out_term.reserve(std::max<uint64_t>(n_pairs, 1) * 4));
  CU(ctx->out_iters.reserve(std::max<uint64_t>(n_pairs, 1) * 4));
  CU(ctx->out_ne.reserve(n_scans * 4));
  CU(ctx->out_np.reserve(n_scans * 4));
  const double* motions_dev = nullptr;
  if (motions) {  // per-sweep motions for the fused de-warp: small, copied ahead of the first extract
    for (uint64_t i = 0; i < 7 * n_scans; i++)
      if (!std::isfinite(motions[i])) return fail(ctx, LOAMGPU_ERR_INVALID, "start_T_end is not finite");
    CU(ctx->motions.reserve(n_scans * 56));
    CU(cudaMemcpyAsync(ctx->motions.p, motions, n_scans * 56, cudaMemcpyHostToDevice, ctx->stream));
    motions_dev = ctx->motions.as<double>();
  }
  // Staging buffers are guarded by per-buffer events recorded right after the extract that consumed them, so the first
  // copies of this call may overlap the registration kernels of a previous, still running call.  Other entry points
  // use the staging memory on the compute stream without those events: after one of them, wait for all of it.
  if (ctx->staging_unguarded) {
    CU(cudaEventRecord(ctx->ev_consumed[0], ctx->stream));
    CU(cudaEventRecord(ctx->ev_consumed[1], ctx->stream));
    ctx->staging_unguarded = false;
  }
  auto fetch = [&](uint64_t s0, uint32_t ns, int buf, const float** out) {
    // copy stream: wait until the previous user of this staging buffer is done, copy, signal the compute stream
    cudaError_t e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ctx->scan_in[buf].p, scans + s0 * scan_bytes, (size_t)ns * scan_bytes,
                          cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "odometry_host fetch");
    *out = ctx->scan_in[buf].as<float>();
    return (int)LOAMGPU_OK;
  };
  int rc = odometry_core(ctx, n_scans, lp, fe, reg, ctx->out_pose.as<double>(), ctx->out_term.as<int32_t>(),
                         ctx->out_iters.as<uint32_t>(), ctx->out_ne.as<uint32_t>(), ctx->out_np.as<uint32_t>(), mode, chunk,
                         short_lead, fetch, motions_dev, pt_stride);
  if (rc) return rc;
  if (poses && n_pairs) CU(cudaMemcpyAsync(poses, ctx->out_pose.p, n_pairs * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (termination && n_pairs)
    CU(cudaMemcpyAsync(termination, ctx->out_term.p, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (iterations && n_pairs)
    CU(cudaMemcpyAsync(iterations, ctx->out_iters.p, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_edge) CU(cudaMemcpyAsync(n_edge, ctx->out_ne.p, n_scans * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_planar) CU(cudaMemcpyAsync(n_planar, ctx->out_np.p, n_scans * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return LOAMGPU_OK;
}

// =============================================================================================== multi-GPU
// One sequence over several GPUs of one box: pair k depends only on scans k and k + 1, so the pair range is
// cut into contiguous blocks, one per device; each device extracts its own scans plus one halo scan (the first scan of
// the next block) and no feature set ever crosses GPUs.  One host thread + one context (own streams, own buffers) per
// device, no collective anywhere: the only cross-device data are the result rows each thread writes into its own slice
// of the caller's arrays.  Results are identical to the single-device call, whatever the split.
struct loamgpu_multi {
  std::vector<loamgpu_ctx*> ctx;  // one per entry of `devices` (a device may appear more than once)
  std::string err;
};

extern "C" {

int loamgpu_multi_create(const int* devices, int n_devices, loamgpu_multi** out) {
  if (!out) return LOAMGPU_ERR_INVALID;
  *out = nullptr;
  if (!devices || n_devices <= 0) {
    g_create_err = "loamgpu_multi_create: empty device list";
    return LOAMGPU_ERR_INVALID;
  }
  loamgpu_multi* m = new loamgpu_multi();
  for (int i = 0; i < n_devices; i++) {
    loamgpu_ctx* c = nullptr;
    const int rc = loamgpu_create(devices[i], &c);
    if (rc != LOAMGPU_OK) {
      for (loamgpu_ctx* p : m->ctx) loamgpu_destroy(p);
      delete m;
      return rc;  // (message in loamgpu_last_error(NULL))
    }
    m->ctx.push_back(c);
  }
  *out = m;
  return LOAMGPU_OK;
}

void loamgpu_multi_destroy(loamgpu_multi* m) {
  if (!m) return;
  for (loamgpu_ctx* c : m->ctx) loamgpu_destroy(c);
  delete m;
}

const char* loamgpu_multi_last_error(const loamgpu_multi* m) { return m ? m->err.c_str() : g_create_err.c_str(); }

int loamgpu_multi_device_count(const loamgpu_multi* m) { return m ? (int)m->ctx.size() : 0; }

int loamgpu_multi_odometry_host(loamgpu_multi* m, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                const loamgpu_lidar_params* lp, const loamgpu_fe_params* fe,
                                const loamgpu_reg_params* reg, double* poses, int32_t* termination, uint32_t* iterations,
                                uint32_t* n_edge, uint32_t* n_planar) {
  if (!m) return LOAMGPU_ERR_INVALID;
  NvtxRange range(__func__);
  m->err.clear();
  if (!lp || !fe || !reg) {
    m->err = "null parameter struct";
    return LOAMGPU_ERR_INVALID;
  }
  if (n_scans == 0) return LOAMGPU_OK;
  const uint64_t n_pairs = n_scans - 1;
  const uint64_t n_per = lp->scan_lines * lp->points_per_line;
  const size_t G = m->ctx.size();
  // contiguous pair blocks: device g takes pairs [g * per, (g + 1) * per) and scans [g * per, (g + 1) * per] (halo)
  const uint64_t per = n_pairs ? (n_pairs + G - 1) / G : 0;
  std::vector<int> rc(G, LOAMGPU_OK);
  std::vector<std::thread> workers;
  auto work = [&](size_t g) {
    const uint64_t p_lo = std::min<uint64_t>(g * per, n_pairs), p_hi = std::min<uint64_t>(p_lo + per, n_pairs);
    uint64_t s_lo = p_lo, ns = p_hi > p_lo ? p_hi - p_lo + 1 : 0;
    if (n_pairs == 0) {  // a single scan: device 0 extracts it
      s_lo = 0;
      ns = g == 0 ? 1 : 0;
    }
    if (ns == 0) return;
    const unsigned char* base = static_cast<const unsigned char*>(scans) + s_lo * n_per * stride_bytes;
    // feature counts of a halo scan are written by both neighbours (same values): each device writes its own scans
    // except the halo, the last device also its final scan
    const bool last = p_hi == n_pairs;
    std::vector<uint32_t> ne(ns), np(ns);
    rc[g] = loamgpu_odometry_host_strided(m->ctx[g], base, stride_bytes, ns, lp, fe, reg, poses ? poses + 7 * p_lo : nullptr,
                                          termination ? termination + p_lo : nullptr,
                                          iterations ? iterations + p_lo : nullptr, ne.data(), np.data());
    if (rc[g] != LOAMGPU_OK) return;
    const uint64_t own = last || n_pairs == 0 ? ns : ns - 1;
    if (n_edge) memcpy(n_edge + s_lo, ne.data(), own * sizeof(uint32_t));
    if (n_planar) memcpy(n_planar + s_lo, np.data(), own * sizeof(uint32_t));
  };
  for (size_t g = 1; g < G; g++) workers.emplace_back(work, g);
  work(0);
  for (std::thread& t : workers) t.join();
  for (size_t g = 0; g < G; g++)
    if (rc[g] != LOAMGPU_OK) {
      m->err = std::string("device slot ") + std::to_string(g) + ": " + loamgpu_last_error(m->ctx[g]);
      return rc[g];
    }
  return LOAMGPU_OK;
}

}  // extern "C"
