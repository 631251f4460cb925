/* loamgpu.h — C-ABI of the B200-native LOAM hot path (libloamgpu.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The host-side C++ templates in include/loam/ (same signatures as the reference's
 * loam/features.h:108-111,119-122,166-169 and loam/registration.h:128-131) call only
 * these entry points; INTEGRATION.md shows the binding a maintainer of the reference
 * would add.  Every entry point runs hand-written sm_100a CUDA kernels; there is no
 * CPU fallback — without a CUDA device loamgpu_create() fails with LOAMGPU_ERR_CUDA.
 *
 * Conventions
 *   - return value 0 = OK, otherwise a loamgpu_status; loamgpu_last_error() has the text.
 *   - poses are 7 doubles: qx qy qz qw tx ty tz  (Eigen coeffs() order, geometry.h:27-50).
 *   - point clouds: dtype LOAMGPU_F32 (x,y,z floats at byte offsets 0/4/8 of each
 *     `stride_bytes` record, e.g. 16 for a float4/PCL point) or LOAMGPU_F64 (x,y,z doubles
 *     at offsets 0/8/16, e.g. stride 24 for Eigen::Vector3d).  Arithmetic is always IEEE
 *     fp64 after widening, exactly as the reference's accessors do (common.h:55-78).
 *   - feature / correspondence indices are uint32.
 *   - a context is bound to one device and one stream; use one context per thread/GPU
 *     (the reference is stateless and re-entrant; contexts give the same property).
 *   - every entry point opens an NVTX range named after itself (Nsight timelines).
 *
 * Limits the reference does not have (LOAMGPU_ERR_UNSUPPORTED, never a silent fallback):
 *   - one ring is staged in one CTA's shared memory: points_per_line <= ~9,700 (packed xyz floats),
 *     ~9,200 (float4), ~7,000 (doubles) on a B200 (227 KB per SM), and <= 65,535 in any case;
 *   - num_edge_neighbors, num_plane_neighbors <= 32; neighbor_points <= 16; total points < 2^32;
 *   - the sequence calls take float records of 12 or 16 bytes.
 * Sizes that change the kernels used, not the results: feature sets of up to 20,480 points take the
 * shared-memory NN build, pairs whose two target sets hold up to ~18 k points the shared-memory k-NN
 * walk ($LOAMGPU_KNN_SMEM_KB, default 128); larger ones run the general kernels, targets of >= 60,000
 * points ($LOAMGPU_BIG_TARGET_MIN) the multi-CTA build.
 */
#ifndef LOAMGPU_H
#define LOAMGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct loamgpu_ctx loamgpu_ctx;

typedef enum {
  LOAMGPU_OK = 0,
  LOAMGPU_ERR_SIZE_MISMATCH = 1, /* scan size != scan_lines*points_per_line (common.h:104-113 -> std::runtime_error) */
  LOAMGPU_ERR_INVALID = 2,       /* bad argument (null pointer, number_sectors == 0, capacity too small ...) */
  LOAMGPU_ERR_CUDA = 3,          /* CUDA runtime failure / no device */
  LOAMGPU_ERR_UNSUPPORTED = 4    /* outside kernel limits (points_per_line or neighbour count too large) */
} loamgpu_status;

enum { LOAMGPU_F32 = 0, LOAMGPU_F64 = 1 };

/* replaces loam::LidarParams, common.h:29-41 */
typedef struct {
  uint64_t scan_lines;
  uint64_t points_per_line;
  double min_range;
  double max_range;
} loamgpu_lidar_params;

/* replaces loam::FeatureExtractionParams, features.h:37-66 (same defaults via loamgpu_default_fe_params) */
typedef struct {
  uint64_t neighbor_points;
  uint64_t number_sectors;
  uint64_t max_edge_feats_per_sector;
  uint64_t max_planar_feats_per_sector;
  double edge_feat_threshold;
  double planar_feat_threshold;
  double occlusion_thresh;
  double parallel_thresh;
} loamgpu_fe_params;

/* replaces loam::RegistrationParams, registration.h:40-75 */
typedef struct {
  uint64_t num_edge_neighbors;
  double max_edge_neighbor_dist;
  uint64_t min_line_fit_points;
  double min_line_condition_number;
  uint64_t num_plane_neighbors;
  double max_plane_neighbor_dist;
  uint64_t min_plane_fit_points;
  double max_avg_point_plane_dist;
  uint64_t max_iterations;
  double rotation_convergence_thresh;
  double position_convergence_thresh;
  uint64_t min_associations;
} loamgpu_reg_params;

/* replaces loam::RegistrationDetail, registration.h:79-109, flattened.  All buffers are
 * caller-allocated host memory; any pointer may be NULL to skip that output. */
typedef struct {
  uint32_t max_iters_cap;  /* in : rows available in the per-iteration buffers */
  uint32_t n_src_edge;     /* in : row stride (in pairs) of edge_assoc  */
  uint32_t n_src_planar;   /* in : row stride (in pairs) of plane_assoc */
  uint32_t n_iters;        /* out: iteration_info.size() */
  int32_t termination;     /* out: 0 CONVERGED, 1 MAX_ITER, 2 INSUFFICIENT_ASSOCIATIONS */
  double* iter_est;        /* [cap][7] target_T_source_init of each iteration */
  double* iter_update;     /* [cap][7] estimate_update */
  uint32_t* n_edge_assoc;  /* [cap] */
  uint32_t* n_plane_assoc; /* [cap] */
  uint32_t* edge_assoc;    /* [cap][n_src_edge][2]   (source idx, nearest target idx), source-idx order */
  uint32_t* plane_assoc;   /* [cap][n_src_planar][2] */
  uint32_t* lm_iters;      /* [cap] LM iterations of each inner solve */
  double* lm_cost;         /* [cap][2] initial / final cost of each inner solve */
} loamgpu_detail;

/* ---------------------------------------------------------------- lifecycle */
int loamgpu_create(int device, loamgpu_ctx** out);
void loamgpu_destroy(loamgpu_ctx* ctx);
/* text of the last error on this context (ctx == NULL: last loamgpu_create failure of this thread) */
const char* loamgpu_last_error(const loamgpu_ctx* ctx);
/* run subsequent calls on an externally owned cudaStream_t (NULL = the context's own stream) */
int loamgpu_set_stream(loamgpu_ctx* ctx, void* cuda_stream);
void loamgpu_default_fe_params(loamgpu_fe_params* p);
void loamgpu_default_reg_params(loamgpu_reg_params* p);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t loamgpu_launch_count(const loamgpu_ctx* ctx);

/* Per-kernel-class device time, measured with CUDA events on the launching stream
 * (bench.py's roofline.achieved).  set_profiling(1) brackets every launch with an event
 * pair; kernel_times() synchronises the stream, returns accumulated milliseconds and launch
 * counts per class since the previous call, and resets the accumulators. */
enum {
  LOAMGPU_K_EXTRACT = 0, /* extract_ring_kernel  (K1+K2) */
  LOAMGPU_K_PACK = 1,    /* pack_features_kernel */
  LOAMGPU_K_GRID = 2,    /* bvh_build_kernel     (K3) */
  LOAMGPU_K_ASSOC = 3,   /* assoc_knn_kernel / knn_kernel (K4) */
  LOAMGPU_K_LM = 4,      /* lm_kernel            (K6+K7) */
  LOAMGPU_K_MISC = 5,    /* init/finish pair state */
  LOAMGPU_K_FIT = 6,     /* assoc_fit_kernel     (K5) */
  LOAMGPU_K_COUNT = 7
};
int loamgpu_set_profiling(loamgpu_ctx* ctx, int on);
int loamgpu_kernel_times(loamgpu_ctx* ctx, double ms[LOAMGPU_K_COUNT], uint64_t launches[LOAMGPU_K_COUNT]);

/* ------------------------------------------------- feature extraction (host buffers) */
/* replaces loam::extractFeatures, features.h:108-111 / features-inl.h:11-50.  Writes the
 * indices (into the input scan) of the selected edge / planar points, in the reference's
 * output order (line-major, sector-major, selection order); the C++ wrapper gathers the
 * point copies.  Capacities needed: scan_lines*number_sectors*(max_*_feats_per_sector+1). */
int loamgpu_extract(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride_bytes, uint64_t n_points,
                    const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe, uint32_t* edge_idx,
                    uint64_t edge_cap, uint64_t* n_edge, uint32_t* planar_idx, uint64_t planar_cap,
                    uint64_t* n_planar);
/* replaces loam::computeCurvature, features.h:119-122 / features-inl.h:53-87 (curvature[i] for point i) */
int loamgpu_curvature(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride_bytes, uint64_t n_points,
                      const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe, double* curvature);
/* replaces loam::computeValidPoints, features.h:166-169 / features-inl.h:90-124 (1 = valid) */
int loamgpu_valid_mask(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride_bytes, uint64_t n_points,
                       const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe, uint8_t* mask);

/* EXTENSION — SURVEY §8f-3, the step the reference leaves to its caller immediately before this
 * path ("de-warp", README.md:63): loamgpu_extract of a motion-compensated scan, with the
 * compensation fused into the ring staging (no separate pass over the scan).  Column c of every
 * ring was measured at fraction s = c / points_per_line of the sweep; each point is moved into the
 * frame of the sweep start by T(s) = interp(Identity, start_T_end, s) (rotation: normalised linear
 * interpolation of the quaternion on Identity's hemisphere, translation: linear) before ranges,
 * curvature, mask and selection look at it.  start_T_end is a pose [qx qy qz qw tx ty tz].
 * dewarped_xyz (nullable, n_points x 3 doubles) receives the moved points — the feature points to
 * hand to loamgpu_register are dewarped_xyz[idx], not pts[idx].  The move uses IEEE + - * / sqrt
 * only, in a fixed order, so its results are bit-reproducible on a CPU; the
 * indices equal the reference's extraction run on the moved points. */
int loamgpu_extract_dewarped(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride_bytes, uint64_t n_points,
                             const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                             const double start_T_end[7], uint32_t* edge_idx, uint64_t edge_cap, uint64_t* n_edge,
                             uint32_t* planar_idx, uint64_t planar_cap, uint64_t* n_planar, double* dewarped_xyz);

/* ---------------------------------------------------- registration (host buffers) */
/* replaces loam::registerFeatures, registration.h:128-131 / registration-inl.h:11-78.
 * Feature clouds are n x 3 contiguous doubles (the reference widens with featuresToEigen,
 * features.h:188-198). */
int loamgpu_register(loamgpu_ctx* ctx, const double* src_edge, uint64_t n_src_edge, const double* src_planar,
                     uint64_t n_src_planar, const double* tgt_edge, uint64_t n_tgt_edge, const double* tgt_planar,
                     uint64_t n_tgt_planar, const double init_pose[7], const loamgpu_reg_params* params,
                     double out_pose[7], loamgpu_detail* detail);
/* replaces kdtree_internal::knnSearch, kdtree.cpp:10-28, batched: for each query the k nearest
 * targets (ascending squared distance, ties by ascending index) that lie strictly inside
 * max_dist (max_dist <= 0: unbounded).  idx_out is [n_queries][k] (unused slots 0xFFFFFFFF). */
int loamgpu_knn(loamgpu_ctx* ctx, const double* targets, uint64_t n_targets, const double* queries,
                uint64_t n_queries, uint32_t k, double max_dist, uint32_t* idx_out, uint32_t* count_out);

/* ---------------------------------------------------- explicit batches (host buffers) */
/* loamgpu_extract for n_scans organised scans of the same geometry stored back to back
 * (n_scans * n_points_per_scan records; n_points_per_scan must equal scan_lines*points_per_line,
 * checked like validateLidarScan, common.h:104-113).  Row s of edge_idx / planar_idx (row pitch
 * edge_cap / planar_cap, each at least scan_lines*number_sectors*(max_*_feats_per_sector+1))
 * receives the indices of scan s relative to that scan; n_edge[s] / n_planar[s] their counts. */
int loamgpu_extract_batch(loamgpu_ctx* ctx, const void* pts, int dtype, size_t stride_bytes, uint64_t n_scans,
                          uint64_t n_points_per_scan, const loamgpu_lidar_params* lidar,
                          const loamgpu_fe_params* fe, uint32_t* edge_idx,
                          uint64_t edge_cap, uint32_t* n_edge, uint32_t* planar_idx, uint64_t planar_cap,
                          uint32_t* n_planar);
/* loamgpu_register for n_pairs independent (source, target) pairs in one call (loop-closure
 * candidates, scan-to-keyframe sets ...).  The four feature clouds are concatenated in pair
 * order (n x 3 doubles each); n_*[p] are the per-pair point counts.  init_poses: [n_pairs][7]
 * or NULL (identity).  Outputs per pair: pose [7], termination, outer iterations (the last two
 * may be NULL).  Same results as n_pairs loamgpu_register calls up to the rounding of the
 * 6x6 normal-equation sums (a pair is reduced by one CTA here, by a cluster of 8 there). */
int loamgpu_register_pairs(loamgpu_ctx* ctx, uint64_t n_pairs, const double* src_edge, const uint64_t* n_src_edge,
                           const double* src_planar, const uint64_t* n_src_planar, const double* tgt_edge,
                           const uint64_t* n_tgt_edge, const double* tgt_planar, const uint64_t* n_tgt_planar,
                           const double* init_poses, const loamgpu_reg_params* params, double* out_poses,
                           int32_t* termination, uint32_t* iterations);

/* ------------------------------------------- device-resident local map (scan-to-map) */
/* The reference's README.md:63 leaves "maintain a local map of points" to the caller, who then
 * passes the accumulated map as the `target` of registerFeatures (registration.h:128-131) and
 * pays a KD-tree build over the whole map on every call (registration-inl.h:20-23).  A
 * loamgpu_map keeps the target's feature points and their NN structures on the device between
 * calls.  Point indices (loamgpu_detail associations) refer to the map's current point order:
 * insertion order, after eviction.  One map belongs to the device of the context that made it.
 *
 *   loamgpu_map_create      map from n_edge x 3 / n_planar x 3 doubles (either may be empty)
 *   loamgpu_map_update      append features (transformed into the map frame by `pose`, i.e.
 *                           Pose3d::act, when pose != NULL), then keep only the newest max_edge /
 *                           max_planar points (0 = unbounded) and rebuild the NN structures
 *   loamgpu_register_to_map registerFeatures(source, <map>, init, params, detail): identical
 *                           results to loamgpu_register with the map's points as the target
 * loamgpu_register itself switches to the same multi-CTA NN build for large targets. */
typedef struct loamgpu_map loamgpu_map;
int loamgpu_map_create(loamgpu_ctx* ctx, const double* edge, uint64_t n_edge, const double* planar,
                       uint64_t n_planar, loamgpu_map** out);
void loamgpu_map_destroy(loamgpu_ctx* ctx, loamgpu_map* map);
int loamgpu_map_size(const loamgpu_map* map, uint64_t* n_edge, uint64_t* n_planar);
int loamgpu_map_update(loamgpu_ctx* ctx, loamgpu_map* map, const double* edge, uint64_t n_edge,
                       const double* planar, uint64_t n_planar, const double pose[7], uint64_t max_edge,
                       uint64_t max_planar);
int loamgpu_register_to_map(loamgpu_ctx* ctx, const loamgpu_map* map, const double* src_edge, uint64_t n_src_edge,
                            const double* src_planar, uint64_t n_src_planar, const double init_pose[7],
                            const loamgpu_reg_params* params, double out_pose[7], loamgpu_detail* detail);

/* ------------------------------------------------------- sequence odometry (batched) */
/* extract + scan-to-scan register over a sequence of organised float4 scans
 * ({x,y,z,unused} floats, n_scans * scan_lines*points_per_line records): every scan is
 * extracted once; pair k registers source = scan k+1 onto target = scan k from an identity
 * initial estimate (the loop of the reference's README.md:46-59).  Features never leave the
 * device.  Outputs per pair: pose [7], termination, outer iterations run, and the feature
 * counts of every scan (n_edge/n_planar, may be NULL).
 *   _host  : scans/outputs in host memory (pinned or pageable); H2D/D2H inside the call.
 *   _device: scans/outputs already resident in device memory. */
int loamgpu_odometry_host(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans, const loamgpu_lidar_params* lidar,
                          const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses,
                          int32_t* termination, uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);
/* _host without the final wait: copies and kernels are only enqueued (input and output buffers
 * must stay valid, and should be page-locked, until loamgpu_synchronize returns).  Consecutive
 * calls pipeline: the host-to-device copies of a call overlap the kernels of the previous one,
 * which is how a long recording is streamed through in pieces. */
int loamgpu_odometry_host_async(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans,
                                const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);
/* wait for everything this context has enqueued */
int loamgpu_synchronize(loamgpu_ctx* ctx);
int loamgpu_odometry_device(loamgpu_ctx* ctx, const float* scans_dev, uint64_t n_scans,
                            const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                            const loamgpu_reg_params* reg, double* poses_dev, int32_t* termination_dev,
                            uint32_t* iterations_dev, uint32_t* n_edge_dev, uint32_t* n_planar_dev);
/* The same three calls on float records of `stride_bytes` = 12 (packed x y z: numpy (N,3) float32, PCL-style packed
 * clouds) or 16 (x y z + one unused float).  The fourth float of a sensor record is never read, so a host that can
 * hand over packed xyz moves 25 % fewer bytes through the host-to-device copy that bounds the host-buffer calls;
 * the extraction kernel stages 12-byte rings with one bulk copy just like float4 ones.  Results are identical. */
int loamgpu_odometry_host_strided(loamgpu_ctx* ctx, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                  const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                  const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                  uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);
int loamgpu_odometry_host_async_strided(loamgpu_ctx* ctx, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                        const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                        const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                        uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);
int loamgpu_odometry_device_strided(loamgpu_ctx* ctx, const void* scans_dev, size_t stride_bytes, uint64_t n_scans,
                                    const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                    const loamgpu_reg_params* reg, double* poses_dev, int32_t* termination_dev,
                                    uint32_t* iterations_dev, uint32_t* n_edge_dev, uint32_t* n_planar_dev);
/* EXTENSION: the sequence calls on sweeps that still need de-warping.  start_T_end is
 * [n_scans][7] (qx qy qz qw tx ty tz): the sensor motion during sweep s, e.g. the previous pair's
 * estimate under a constant-velocity model, or an IMU / wheel-odometry prediction.  Every scan is
 * de-warped inside the extraction kernel exactly as loamgpu_extract_dewarped does it; the feature
 * points that reach registration are the moved points.  Same outputs as the plain calls.  (All-identity
 * motions give bit-identical results to loamgpu_odometry_host / _device.) */
int loamgpu_odometry_host_dewarped(loamgpu_ctx* ctx, const float* scans, uint64_t n_scans, const double* start_T_end,
                                   const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                   const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                   uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);
int loamgpu_odometry_device_dewarped(loamgpu_ctx* ctx, const float* scans_dev, uint64_t n_scans,
                                     const double* start_T_end_dev, const loamgpu_lidar_params* lidar,
                                     const loamgpu_fe_params* fe, const loamgpu_reg_params* reg, double* poses_dev,
                                     int32_t* termination_dev, uint32_t* iterations_dev, uint32_t* n_edge_dev,
                                     uint32_t* n_planar_dev);

/* ------------------------------------------------------- one sequence over several GPUs of one box
 * ("one host thread + one loamgpu_ctx + one stream set per GPU"; the reference's README loop shards by scan
 * pair, README.md:46-59.)  Pair k depends only on scans k and k+1: the pair range is cut into contiguous blocks, one
 * per entry of `devices` (a device may be listed more than once); every block's device extracts its own scans plus
 * one halo scan.  No collective, no peer traffic; results are identical to loamgpu_odometry_host_strided on one
 * device, whatever the split.  Host buffers as for loamgpu_odometry_host_strided. */
typedef struct loamgpu_multi loamgpu_multi;
int loamgpu_multi_create(const int* devices, int n_devices, loamgpu_multi** out);
void loamgpu_multi_destroy(loamgpu_multi* m);
const char* loamgpu_multi_last_error(const loamgpu_multi* m);
int loamgpu_multi_device_count(const loamgpu_multi* m);
int loamgpu_multi_odometry_host(loamgpu_multi* m, const void* scans, size_t stride_bytes, uint64_t n_scans,
                                const loamgpu_lidar_params* lidar, const loamgpu_fe_params* fe,
                                const loamgpu_reg_params* reg, double* poses, int32_t* termination,
                                uint32_t* iterations, uint32_t* n_edge, uint32_t* n_planar);

/* TEST HOOK (tests/test_gpu_jacobians.py): the sums ONE evaluation of the on-device Levenberg-Marquardt solve forms for
 * explicit residual blocks — what associateEdges / associatePlanes hand to Ceres (registration.cpp:52-57, 93-98; functors
 * registration-inl.h:92-117) — at an arbitrary iterate x (qx qy qz qw tx ty tz of estimate_update): out[0..20] upper
 * triangle of J^T J (6x6 tangent, loss-corrected), out[21..26] J^T r, out[27] cost, out[28] = 1 if the moment sums of
 * the inlier planes were used, out[29] = planes left to the streamed part.  mode 0: every record streamed; 1: the
 * default path of the batched kernel. */
int loamgpu_debug_problem_eval(loamgpu_ctx* ctx, uint64_t n_edge, const double* edge_p, const double* edge_a,
                               const double* edge_b, uint64_t n_plane, const double* plane_p, const double* plane_n,
                               const double* plane_d, const double x[7], int mode, double out[30]);

/* pairs processed per internal chunk by the sequence / batch calls; 0 (default) = automatic: 1024 for
 * device-resident and asynchronous host calls, 256 for synchronous host calls and explicit batches,
 * bounded by a share of the free device memory */
int loamgpu_set_chunk_pairs(loamgpu_ctx* ctx, uint32_t pairs);

#ifdef __cplusplus
}
#endif
#endif /* LOAMGPU_H */
