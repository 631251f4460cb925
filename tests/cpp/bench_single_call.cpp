// Latency of ONE loam::extractFeatures / loam::registerFeatures call through the C++ API of include/loam/*.h — the
// reference's own usage pattern (README.md:46-59 of the reference: per scan one extractFeatures, then registerFeatures
// against the previous scan's features), which is what its README quotes 3.5 ms + 13 ms for.
//
//   bench_single_call <scans.bin> <rings> <cols> <reps>
// scans.bin = two consecutive organised scans as float {x, y, z, pad} records (what bench.py writes).  Prints one JSON
// line with median / best milliseconds per call, host buffers in and out (every copy inside the timed call).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "loam/loam.h"

struct PointF {  // PCL-style point: FieldAccessor reads x / y / z, handed to the C-ABI without a copy
  float x, y, z, pad;
};

static double median(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main(int argc, char** argv) {
  if (argc != 5) {
    std::fprintf(stderr, "usage: bench_single_call <scans.bin> <rings> <cols> <reps>\n");
    return 2;
  }
  const size_t R = std::strtoul(argv[2], nullptr, 10), P = std::strtoul(argv[3], nullptr, 10);
  const int reps = std::atoi(argv[4]);
  std::vector<PointF> scan[2] = {std::vector<PointF>(R * P), std::vector<PointF>(R * P)};
  FILE* f = std::fopen(argv[1], "rb");
  if (!f || std::fread(scan[0].data(), sizeof(PointF), R * P, f) != R * P ||
      std::fread(scan[1].data(), sizeof(PointF), R * P, f) != R * P) {
    std::fprintf(stderr, "cannot read two %zux%zu scans from %s\n", R, P, argv[1]);
    return 2;
  }
  std::fclose(f);
  const loam::LidarParams lidar(R, P, 1.0, 120.0);
  using clock = std::chrono::steady_clock;
  auto ms = [](clock::time_point a, clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };

  loam::LoamFeatures<PointF> feat[2];
  std::vector<double> t_ext, t_reg, t_reg_detail;
  for (int i = 0; i < reps + 5; i++) {
    const auto t0 = clock::now();
    feat[i & 1] = loam::extractFeatures(scan[i & 1], lidar);
    const auto t1 = clock::now();
    if (i >= 5) t_ext.push_back(ms(t0, t1));
  }
  feat[0] = loam::extractFeatures(scan[0], lidar);
  feat[1] = loam::extractFeatures(scan[1], lidar);
  loam::Pose3d pose;
  for (int i = 0; i < reps + 5; i++) {
    const auto t0 = clock::now();
    pose = loam::registerFeatures(feat[1], feat[0], loam::Pose3d::Identity());
    const auto t1 = clock::now();
    if (i >= 5) t_reg.push_back(ms(t0, t1));
  }
  size_t iters = 0;
  for (int i = 0; i < reps / 2 + 2; i++) {
    auto detail = std::make_shared<loam::RegistrationDetail>();
    const auto t0 = clock::now();
    loam::registerFeatures(feat[1], feat[0], loam::Pose3d::Identity(), loam::RegistrationParams(), detail);
    const auto t1 = clock::now();
    if (i >= 2) t_reg_detail.push_back(ms(t0, t1));
    iters = detail->iteration_info.size();
  }
  std::printf(
      "{\"rings\": %zu, \"cols\": %zu, \"reps\": %d, \"extract_ms\": %.4f, \"extract_best_ms\": %.4f, \"register_ms\": %.4f, "
      "\"register_best_ms\": %.4f, \"register_with_detail_ms\": %.4f, \"outer_iterations\": %zu, \"edge_features\": %zu, "
      "\"planar_features\": %zu, \"pose\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g]}\n",
      R, P, reps, median(t_ext), *std::min_element(t_ext.begin(), t_ext.end()), median(t_reg),
      *std::min_element(t_reg.begin(), t_reg.end()), median(t_reg_detail), iters, feat[0].edge_points.size(),
      feat[0].planar_points.size(), pose.rotation.x(), pose.rotation.y(), pose.rotation.z(), pose.rotation.w(),
      pose.translation(0), pose.translation(1), pose.translation(2));
  return 0;
}
