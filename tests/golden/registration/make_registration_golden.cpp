// Golden-vector generator for the registration half — runs the REAL reference.
//
// Built against the unmodified DanMcGann/loam library (which pulls Ceres 2.2.0, nanoflann v1.5.5 and Eigen 3 through its
// own CMake, CMakeLists.txt:13,18-33) on a machine that has network access or those packages.  This image has none of
// the three, so the fixture cannot be produced here; the recipe is committed so that it can be produced anywhere else:
//
//     python tests/golden/registration/make_registration_inputs.py      # writes registration_inputs.txt (this repo)
//     cmake -S tests/golden/registration -B /tmp/rg -DLOAM_REFERENCE_DIR=/path/to/DanMcGann/loam && cmake --build /tmp/rg
//     /tmp/rg/make_registration_golden tests/golden/registration/registration_inputs.txt \
//                                      tests/golden/registration/registration_golden.txt
//     python tests/golden/registration/pack_registration_golden.py      # -> tests/golden/registration_golden.npz
//
// tests/test_registration_golden.py picks the .npz up when it exists (CPU: oracle vs golden; GPU: CUDA vs golden) and
// says "fixture absent" otherwise.  Until such a fixture is committed the registration half stays PARITY UNPINNED
// against the real Ceres.
//
// What is recorded per case: the returned pose, RegistrationDetail::termination_type and, for every outer iteration,
// IterationInfo{target_T_source_init, edge_associations, plane_associations, estimate_update}
// (reference: loam/include/loam/registration.h:79-131, registration-inl.h:26-77).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "Eigen/Dense"
#include "loam/loam.h"

using Cloud = std::vector<Eigen::Vector3d>;

static Cloud read_cloud(std::istream& in, const std::string& expect) {
  std::string tag;
  size_t n = 0;
  in >> tag >> n;
  if (tag != expect) {
    std::cerr << "expected '" << expect << "', found '" << tag << "'\n";
    std::exit(2);
  }
  Cloud c(n);
  for (size_t i = 0; i < n; i++) in >> c[i](0) >> c[i](1) >> c[i](2);
  return c;
}

static void write_pose(FILE* f, const char* tag, const loam::Pose3d& p) {
  // same memory order as everywhere in this repo: qx qy qz qw tx ty tz
  std::fprintf(f, "%s %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", tag, p.rotation.x(), p.rotation.y(), p.rotation.z(),
               p.rotation.w(), p.translation(0), p.translation(1), p.translation(2));
}

int main(int argc, char** argv) {
  if (argc != 3) {
    std::cerr << "usage: make_registration_golden <registration_inputs.txt> <registration_golden.txt>\n";
    return 2;
  }
  std::ifstream in(argv[1]);
  FILE* out = std::fopen(argv[2], "w");
  if (!in || !out) {
    std::cerr << "cannot open files\n";
    return 2;
  }
  size_t n_cases = 0;
  std::string tag;
  in >> tag >> n_cases;  // "cases N"
  std::fprintf(out, "cases %zu\n", n_cases);
  for (size_t c = 0; c < n_cases; c++) {
    std::string name;
    in >> tag >> name;  // "case <name>"
    loam::RegistrationParams rp;
    in >> tag >> rp.num_edge_neighbors >> rp.max_edge_neighbor_dist >> rp.min_line_fit_points >>
        rp.min_line_condition_number >> rp.num_plane_neighbors >> rp.max_plane_neighbor_dist >>
        rp.min_plane_fit_points >> rp.max_avg_point_plane_dist >> rp.max_iterations >> rp.rotation_convergence_thresh >>
        rp.position_convergence_thresh >> rp.min_associations;  // "params ..." in the struct's field order
    double q[4], t[3];
    in >> tag >> q[0] >> q[1] >> q[2] >> q[3] >> t[0] >> t[1] >> t[2];  // "init qx qy qz qw tx ty tz"
    const loam::Pose3d init(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(t[0], t[1], t[2]));
    loam::LoamFeatures<Eigen::Vector3d> source, target;
    source.edge_points = read_cloud(in, "source_edge");
    source.planar_points = read_cloud(in, "source_planar");
    target.edge_points = read_cloud(in, "target_edge");
    target.planar_points = read_cloud(in, "target_planar");

    auto detail = std::make_shared<loam::RegistrationDetail>();
    const loam::Pose3d result = loam::registerFeatures<loam::ParenAccessor>(source, target, init, rp, detail);

    std::fprintf(out, "case %s\n", name.c_str());
    write_pose(out, "result", result);
    std::fprintf(out, "termination %d\n", (int)detail->termination_type);
    std::fprintf(out, "iterations %zu\n", detail->iteration_info.size());
    for (const auto& it : detail->iteration_info) {
      write_pose(out, "est", it.target_T_source_init);
      write_pose(out, "update", it.estimate_update);
      std::fprintf(out, "edge_assoc %zu", it.edge_associations.size());
      for (const auto& a : it.edge_associations) std::fprintf(out, " %zu %zu", a.first, a.second);
      std::fprintf(out, "\nplane_assoc %zu", it.plane_associations.size());
      for (const auto& a : it.plane_associations) std::fprintf(out, " %zu %zu", a.first, a.second);
      std::fprintf(out, "\n");
    }
  }
  std::fclose(out);
  return 0;
}
