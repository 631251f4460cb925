// TEST INFRASTRUCTURE ONLY — C-ABI shim around the REAL reference feature code.
//
// This file instantiates the reference's own templates (loam/features.h,
// loam/features-inl.h, loam/common.h + loam/src/features.cpp) where they lie
// under /root/reference; nothing from the reference is copied into this repo.
// It is compiled by oracle/Makefile into oracle/_ref/libloam_ref.so and is
// used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs, as the checker and the timed CPU baseline.
//
// The reference returns COPIES of the selected input points, not indices
// (features-inl.h:147,169).  To recover indices the point type carries its
// own index; FieldAccessor only touches .x/.y/.z (common.h:55-60).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "loam/features.h"

namespace {
struct PointD {
  double x, y, z;
  uint32_t idx;
};
struct PointF {  // 16-byte float4-shaped point (the benchmark layout)
  float x, y, z;
  uint32_t idx;
};
thread_local std::string g_err;

loam::FeatureExtractionParams make_fe(const uint64_t* u, const double* d) {
  loam::FeatureExtractionParams p;
  p.neighbor_points = u[0];
  p.number_sectors = u[1];
  p.max_edge_feats_per_sector = u[2];
  p.max_planar_feats_per_sector = u[3];
  p.edge_feat_threshold = d[0];
  p.planar_feat_threshold = d[1];
  p.occlusion_thresh = d[2];
  p.parallel_thresh = d[3];
  return p;
}

template <typename P, typename S>
std::vector<P> make_scan(const S* xyz, size_t n, size_t stride_elems) {
  std::vector<P> scan(n);
  for (size_t i = 0; i < n; i++) {
    scan[i].x = xyz[i * stride_elems + 0];
    scan[i].y = xyz[i * stride_elems + 1];
    scan[i].z = xyz[i * stride_elems + 2];
    scan[i].idx = (uint32_t)i;
  }
  return scan;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// lidar_u = {scan_lines, points_per_line}; lidar_d = {min_range, max_range}
// fe_u = {neighbor_points, number_sectors, max_edge, max_planar}
// fe_d = {edge_thr, planar_thr, occlusion_thresh, parallel_thresh}
int ref_extract_f64(const double* xyz, uint64_t n, const uint64_t* lidar_u, const double* lidar_d,
                    const uint64_t* fe_u, const double* fe_d, uint32_t* edge_idx, uint64_t* n_edge,
                    uint32_t* planar_idx, uint64_t* n_planar) {
  try {
    loam::LidarParams lp(lidar_u[0], lidar_u[1], lidar_d[0], lidar_d[1]);
    auto fe = make_fe(fe_u, fe_d);
    auto scan = make_scan<PointD>(xyz, n, 3);
    auto f = loam::extractFeatures(scan, lp, fe);
    *n_edge = f.edge_points.size();
    *n_planar = f.planar_points.size();
    for (size_t i = 0; i < f.edge_points.size(); i++) edge_idx[i] = f.edge_points[i].idx;
    for (size_t i = 0; i < f.planar_points.size(); i++) planar_idx[i] = f.planar_points[i].idx;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

int ref_curvature_f64(const double* xyz, uint64_t n, const uint64_t* lidar_u, const double* lidar_d,
                      const uint64_t* fe_u, const double* fe_d, double* curv, uint64_t* index) {
  try {
    loam::LidarParams lp(lidar_u[0], lidar_u[1], lidar_d[0], lidar_d[1]);
    auto fe = make_fe(fe_u, fe_d);
    auto scan = make_scan<PointD>(xyz, n, 3);
    auto c = loam::computeCurvature(scan, lp, fe);
    for (size_t i = 0; i < c.size(); i++) {
      curv[i] = c[i].curvature;
      index[i] = c[i].index;
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

int ref_valid_f64(const double* xyz, uint64_t n, const uint64_t* lidar_u, const double* lidar_d,
                  const uint64_t* fe_u, const double* fe_d, uint8_t* mask) {
  try {
    loam::LidarParams lp(lidar_u[0], lidar_u[1], lidar_d[0], lidar_d[1]);
    auto fe = make_fe(fe_u, fe_d);
    auto scan = make_scan<PointD>(xyz, n, 3);
    auto m = loam::computeValidPoints(scan, lp, fe);
    for (size_t i = 0; i < m.size(); i++) mask[i] = m[i] ? 1 : 0;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Timed run of the real reference extractFeatures on a float4-shaped scan
// (the 16-byte point layout of the benchmark).  Returns seconds per call
// (best and mean over `reps`) measured around extractFeatures only, and the
// feature indices of the last call so the caller can reuse them.
int ref_extract_f32x4_timed(const float* xyzw, uint64_t n, const uint64_t* lidar_u, const double* lidar_d,
                            const uint64_t* fe_u, const double* fe_d, int reps, double* best_s, double* mean_s,
                            uint32_t* edge_idx, uint64_t* n_edge, uint32_t* planar_idx, uint64_t* n_planar) {
  try {
    loam::LidarParams lp(lidar_u[0], lidar_u[1], lidar_d[0], lidar_d[1]);
    auto fe = make_fe(fe_u, fe_d);
    auto scan = make_scan<PointF>(xyzw, n, 4);
    double best = 1e300, sum = 0;
    for (int r = 0; r < reps; r++) {
      auto t0 = std::chrono::steady_clock::now();
      auto f = loam::extractFeatures(scan, lp, fe);
      auto t1 = std::chrono::steady_clock::now();
      double s = std::chrono::duration<double>(t1 - t0).count();
      best = s < best ? s : best;
      sum += s;
      if (r == reps - 1) {
        *n_edge = f.edge_points.size();
        *n_planar = f.planar_points.size();
        if (edge_idx)
          for (size_t i = 0; i < f.edge_points.size(); i++) edge_idx[i] = f.edge_points[i].idx;
        if (planar_idx)
          for (size_t i = 0; i < f.planar_points.size(); i++) planar_idx[i] = f.planar_points[i].idx;
      }
    }
    *best_s = best;
    *mean_s = sum / (reps > 0 ? reps : 1);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
