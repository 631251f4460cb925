// pybind11 module `loam_python` of the B200-native build — the same Python surface as the reference's
// python/loam_bindings.cpp (classes, field names, function names, argument names, defaults), over the C-ABI.
//
// Differences from the reference file, all additive or invisible to scripts written against it:
//   * FAST PATH: a scan / feature set given as ONE contiguous (N, >=3) float32 or float64 ndarray goes
//     to the C-ABI as it is (zero copy, one call).  The reference marshals one py::array_t<double> PER POINT
//     (loam_bindings.cpp:85-86: std::vector<py::array_t<double>>), 65,536 tiny arrays per 64x1024 scan.  Any other
//     sequence of points (what the reference accepts) is packed once with numpy and takes the same path.
//   * LoamFeatures.edge_points / planar_points are (M, C) ndarrays holding copies of the selected input rows; iterating
//     them yields one array per point, exactly what the reference's list of arrays yields.
//   * PointCurvature is registered (the reference forgot it, loam_bindings.cpp:88-92, so its computeCurvature cannot
//     return to Python); computeCurvatureArray / computeValidPointsArray return ndarrays without per-point objects.
//   * The GIL is released around every C-ABI call.
//   * Extensions: extractFeatureIndices, odometry (batched scan-to-scan sequence), setDevice.
//   * Eigen is not required: Pose3d.translation / act() exchange (3,) ndarrays, Quaterniond is bound with the
//     reference's constructor order (w, x, y, z).  With real Eigen + pybind11/eigen.h the reference's own
//     loam_bindings.cpp compiles unchanged against include/loam/*.h instead.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "loam/loam.h"

namespace py = pybind11;
using loam::gpu::check;
using loam::gpu::ThreadContext;

namespace {

// LoamFeatures of the Python surface: two (M, C) arrays
struct PyFeatures {
  py::array edge_points = py::array_t<double>(std::vector<py::ssize_t>{0, 3});
  py::array planar_points = py::array_t<double>(std::vector<py::ssize_t>{0, 3});
};

// A point cloud as the C-ABI wants it, borrowed from (or packed out of) a Python object.
struct Cloud {
  py::array arr;  // keeps the buffer alive
  const void* data = nullptr;
  int dtype = LOAMGPU_F64;
  size_t stride = 24;
  size_t n = 0;
};

Cloud as_cloud(const py::object& obj) {
  Cloud c;
  py::array a = py::array::ensure(obj);
  if (!a) throw std::invalid_argument("point cloud must be convertible to an (N, >=3) array");
  if (a.size() == 0) {
    c.arr = a;
    return c;
  }
  if (a.ndim() != 2 || a.shape(1) < 3) throw std::invalid_argument("point cloud must be an (N, >=3) array");
  const bool f32 = py::isinstance<py::array_t<float>>(a);
  if (!f32 && !py::isinstance<py::array_t<double>>(a)) a = py::array_t<double, py::array::c_style | py::array::forcecast>(a);
  // rows contiguous in x, y, z; any row pitch that is a multiple of the element size
  const py::ssize_t es = f32 ? 4 : 8;
  if (a.strides(1) != es || a.strides(0) < 3 * es || a.strides(0) % es != 0) {
    a = f32 ? py::array(py::array_t<float, py::array::c_style | py::array::forcecast>(a))
            : py::array(py::array_t<double, py::array::c_style | py::array::forcecast>(a));
  }
  c.arr = a;
  c.data = a.data();
  c.dtype = f32 ? LOAMGPU_F32 : LOAMGPU_F64;
  c.stride = (size_t)a.strides(0);
  c.n = (size_t)a.shape(0);
  return c;
}

// n x 3 doubles for the registration entry points (features are widened once, as featuresToEigen does)
py::array_t<double> as_xyz64(const py::object& obj) {
  py::array a = py::array::ensure(obj);
  if (!a) throw std::invalid_argument("feature points must be convertible to an (N, >=3) array");
  if (a.size() == 0) return py::array_t<double>(std::vector<py::ssize_t>{0, 3});
  if (a.ndim() != 2 || a.shape(1) < 3) throw std::invalid_argument("feature points must be an (N, >=3) array");
  py::array_t<double, py::array::c_style | py::array::forcecast> d(a);
  if (d.shape(1) == 3) return d;
  py::array_t<double> out({d.shape(0), (py::ssize_t)3});
  auto src = d.unchecked<2>();
  auto dst = out.mutable_unchecked<2>();
  for (py::ssize_t i = 0; i < d.shape(0); i++)
    for (int k = 0; k < 3; k++) dst(i, k) = src(i, k);
  return out;
}

py::array take_rows(const py::array& scan, const std::vector<uint32_t>& idx) {
  // copies of the selected input rows (reference: feature points are copies of input elements, features-inl.h:147,169)
  py::object np = py::module_::import("numpy");
  py::array_t<uint32_t> i((py::ssize_t)idx.size());
  if (!idx.empty()) std::memcpy(i.mutable_data(), idx.data(), idx.size() * sizeof(uint32_t));
  return np.attr("take")(scan, i, py::arg("axis") = 0);
}

std::pair<std::vector<uint32_t>, std::vector<uint32_t>> extract_indices(const Cloud& c, const loam::LidarParams& lp,
                                                                        const loam::FeatureExtractionParams& fp) {
  std::pair<std::vector<uint32_t>, std::vector<uint32_t>> out;
  const loamgpu_lidar_params clp = loam::gpu::toC(lp);
  const loamgpu_fe_params cfp = loam::gpu::toC(fp);
  out.first.resize(c.n);
  out.second.resize(c.n);
  uint64_t ne = 0, np_ = 0;
  {
    py::gil_scoped_release nogil;
    loamgpu_ctx* ctx = ThreadContext::get();
    check(ctx, loamgpu_extract(ctx, c.data, c.dtype, c.stride, c.n, &clp, &cfp, out.first.data(), out.first.size(), &ne,
                               out.second.data(), out.second.size(), &np_));
  }
  out.first.resize(ne);
  out.second.resize(np_);
  return out;
}

py::array_t<double> vec3(const Eigen::Vector3d& v) {
  py::array_t<double> a(3);
  for (int i = 0; i < 3; i++) a.mutable_at(i) = v(i);
  return a;
}
Eigen::Vector3d to_vec3(const py::object& o) {
  py::array_t<double, py::array::c_style | py::array::forcecast> a(o);
  if (a.size() != 3) throw std::invalid_argument("expected 3 numbers");
  return Eigen::Vector3d(a.data()[0], a.data()[1], a.data()[2]);
}

}  // namespace

PYBIND11_MODULE(loam_python, m) {
  m.doc() = "B200-native LOAM hot path (extractFeatures / registerFeatures) — drop-in for the reference's loam_python";

  // ------------------------------------------------------------------ common (loam_bindings.cpp:24-30)
  py::class_<loam::LidarParams>(m, "LidarParams")
      .def(py::init<size_t, size_t, double, double>(), py::arg("scan_lines"), py::arg("points_per_line"),
           py::arg("min_range"), py::arg("max_range"))
      .def_readonly("scan_lines", &loam::LidarParams::scan_lines)
      .def_readonly("points_per_line", &loam::LidarParams::points_per_line)
      .def_readonly("min_range", &loam::LidarParams::min_range)
      .def_readonly("max_range", &loam::LidarParams::max_range);

  // ------------------------------------------------------------------ geometry (loam_bindings.cpp:41-57)
  py::class_<Eigen::Quaterniond>(m, "Quaterniond")
      .def(py::init<double, double, double, double>(), py::arg("w"), py::arg("x"), py::arg("y"), py::arg("z"))
      .def("w", [](const Eigen::Quaterniond& q) { return q.w(); })
      .def("x", [](const Eigen::Quaterniond& q) { return q.x(); })
      .def("y", [](const Eigen::Quaterniond& q) { return q.y(); })
      .def("z", [](const Eigen::Quaterniond& q) { return q.z(); });

  py::class_<loam::Pose3d>(m, "Pose3d")
      .def(py::init([](const Eigen::Quaterniond& q, const py::object& t) { return loam::Pose3d(q, to_vec3(t)); }),
           py::arg("rotation"), py::arg("translation"))
      .def_static("Identity", &loam::Pose3d::Identity)
      .def("inverse", &loam::Pose3d::inverse)
      .def("compose", &loam::Pose3d::compose, py::arg("other"))
      .def("act", [](const loam::Pose3d& p, const py::object& pt) { return vec3(p.act(to_vec3(pt))); }, py::arg("point"))
      .def_readwrite("rotation", &loam::Pose3d::rotation)
      .def_property(
          "translation", [](const loam::Pose3d& p) { return vec3(p.translation); },
          [](loam::Pose3d& p, const py::object& t) { p.translation = to_vec3(t); });

  // ------------------------------------------------------------------ features (loam_bindings.cpp:69-92)
  py::class_<loam::FeatureExtractionParams>(m, "FeatureExtractionParams")
      .def(py::init<>())
      .def_readwrite("neighbor_points", &loam::FeatureExtractionParams::neighbor_points)
      .def_readwrite("number_sectors", &loam::FeatureExtractionParams::number_sectors)
      .def_readwrite("max_edge_feats_per_sector", &loam::FeatureExtractionParams::max_edge_feats_per_sector)
      .def_readwrite("max_planar_feats_per_sector", &loam::FeatureExtractionParams::max_planar_feats_per_sector)
      .def_readwrite("edge_feat_threshold", &loam::FeatureExtractionParams::edge_feat_threshold)
      .def_readwrite("planar_feat_threshold", &loam::FeatureExtractionParams::planar_feat_threshold)
      .def_readwrite("occlusion_thresh", &loam::FeatureExtractionParams::occlusion_thresh)
      .def_readwrite("parallel_thresh", &loam::FeatureExtractionParams::parallel_thresh);

  py::class_<PyFeatures>(m, "LoamFeatures")
      .def(py::init<>())
      .def(py::init([](const py::object& e, const py::object& p) {
             PyFeatures f;
             f.edge_points = py::array::ensure(e);
             f.planar_points = py::array::ensure(p);
             return f;
           }),
           py::arg("edge_points"), py::arg("planar_points"))
      .def_readwrite("edge_points", &PyFeatures::edge_points)
      .def_readwrite("planar_points", &PyFeatures::planar_points);

  py::class_<loam::PointCurvature>(m, "PointCurvature")
      .def(py::init<size_t, double>(), py::arg("index"), py::arg("curvature"))
      .def_readwrite("index", &loam::PointCurvature::index)
      .def_readwrite("curvature", &loam::PointCurvature::curvature);

  m.def(
      "extractFeatures",
      [](const py::object& input_scan, const loam::LidarParams& lp, const loam::FeatureExtractionParams& fp) {
        const Cloud c = as_cloud(input_scan);
        PyFeatures out;
        if (c.n != lp.scan_lines * lp.points_per_line) {  // the reference's std::runtime_error (common.h:104-113)
          std::vector<char> dummy(c.n);
          loam::validateLidarScan(dummy, lp);
        }
        if (c.n == 0) return out;
        const auto idx = extract_indices(c, lp, fp);
        out.edge_points = take_rows(c.arr, idx.first);
        out.planar_points = take_rows(c.arr, idx.second);
        return out;
      },
      py::arg("input_scan"), py::arg("lidar_params"), py::arg("params") = loam::FeatureExtractionParams());

  m.def(
      "extractFeatureIndices",
      [](const py::object& input_scan, const loam::LidarParams& lp, const loam::FeatureExtractionParams& fp) {
        const Cloud c = as_cloud(input_scan);
        if (c.n != lp.scan_lines * lp.points_per_line) {
          std::vector<char> dummy(c.n);
          loam::validateLidarScan(dummy, lp);
        }
        std::pair<std::vector<uint32_t>, std::vector<uint32_t>> idx;
        if (c.n) idx = extract_indices(c, lp, fp);
        py::array_t<uint32_t> e((py::ssize_t)idx.first.size()), p((py::ssize_t)idx.second.size());
        if (!idx.first.empty()) std::memcpy(e.mutable_data(), idx.first.data(), idx.first.size() * 4);
        if (!idx.second.empty()) std::memcpy(p.mutable_data(), idx.second.data(), idx.second.size() * 4);
        return py::make_tuple(e, p);
      },
      py::arg("input_scan"), py::arg("lidar_params"), py::arg("params") = loam::FeatureExtractionParams(),
      "EXTENSION: indices (into input_scan) of the edge / planar features in the reference's output order");

  auto curvature_array = [](const py::object& input_scan, const loam::LidarParams& lp,
                            const loam::FeatureExtractionParams& fp) {
    const Cloud c = as_cloud(input_scan);
    if (c.n != lp.scan_lines * lp.points_per_line) {
      std::vector<char> dummy(c.n);
      loam::validateLidarScan(dummy, lp);
    }
    py::array_t<double> out((py::ssize_t)c.n);
    if (c.n) {
      const loamgpu_lidar_params clp = loam::gpu::toC(lp);
      const loamgpu_fe_params cfp = loam::gpu::toC(fp);
      double* dst = out.mutable_data();
      py::gil_scoped_release nogil;
      loamgpu_ctx* ctx = ThreadContext::get();
      check(ctx, loamgpu_curvature(ctx, c.data, c.dtype, c.stride, c.n, &clp, &cfp, dst));
    }
    return out;
  };
  auto valid_array = [](const py::object& input_scan, const loam::LidarParams& lp,
                        const loam::FeatureExtractionParams& fp) {
    const Cloud c = as_cloud(input_scan);
    if (c.n != lp.scan_lines * lp.points_per_line) {
      std::vector<char> dummy(c.n);
      loam::validateLidarScan(dummy, lp);
    }
    py::array_t<bool> out((py::ssize_t)c.n);
    if (c.n) {
      const loamgpu_lidar_params clp = loam::gpu::toC(lp);
      const loamgpu_fe_params cfp = loam::gpu::toC(fp);
      uint8_t* dst = reinterpret_cast<uint8_t*>(out.mutable_data());
      py::gil_scoped_release nogil;
      loamgpu_ctx* ctx = ThreadContext::get();
      check(ctx, loamgpu_valid_mask(ctx, c.data, c.dtype, c.stride, c.n, &clp, &cfp, dst));
    }
    return out;
  };
  m.def(
      "computeCurvature",
      [curvature_array](const py::object& s, const loam::LidarParams& lp, const loam::FeatureExtractionParams& fp) {
        const py::array_t<double> c = curvature_array(s, lp, fp);
        std::vector<loam::PointCurvature> out;  // std::vector<PointCurvature>, as the reference returns (features.h:119-122)
        out.reserve((size_t)c.size());
        for (py::ssize_t i = 0; i < c.size(); i++) out.emplace_back((size_t)i, c.data()[i]);
        return out;
      },
      py::arg("input_scan"), py::arg("lidar_params"), py::arg("params") = loam::FeatureExtractionParams());
  m.def("computeCurvatureArray", curvature_array, py::arg("input_scan"), py::arg("lidar_params"),
        py::arg("params") = loam::FeatureExtractionParams(), "EXTENSION: curvature per point as one float64 ndarray");
  m.def(
      "computeValidPoints",
      [valid_array](const py::object& s, const loam::LidarParams& lp, const loam::FeatureExtractionParams& fp) {
        const py::array_t<bool> v = valid_array(s, lp, fp);
        return std::vector<bool>(v.data(), v.data() + v.size());  // std::vector<bool>, as the reference (features.h:166-169)
      },
      py::arg("input_scan"), py::arg("lidar_params"), py::arg("params") = loam::FeatureExtractionParams());
  m.def("computeValidPointsArray", valid_array, py::arg("input_scan"), py::arg("lidar_params"),
        py::arg("params") = loam::FeatureExtractionParams(), "EXTENSION: validity mask as one bool ndarray");

  // ------------------------------------------------------------------ registration (loam_bindings.cpp:104-144)
  py::class_<loam::RegistrationParams>(m, "RegistrationParams")
      .def(py::init<>())
      .def_readwrite("num_edge_neighbors", &loam::RegistrationParams::num_edge_neighbors)
      .def_readwrite("max_edge_neighbor_dist", &loam::RegistrationParams::max_edge_neighbor_dist)
      .def_readwrite("min_line_fit_points", &loam::RegistrationParams::min_line_fit_points)
      .def_readwrite("min_line_condition_number", &loam::RegistrationParams::min_line_condition_number)
      .def_readwrite("num_plane_neighbors", &loam::RegistrationParams::num_plane_neighbors)
      .def_readwrite("max_plane_neighbor_dist", &loam::RegistrationParams::max_plane_neighbor_dist)
      .def_readwrite("min_plane_fit_points", &loam::RegistrationParams::min_plane_fit_points)
      .def_readwrite("max_avg_point_plane_dist", &loam::RegistrationParams::max_avg_point_plane_dist)
      .def_readwrite("max_iterations", &loam::RegistrationParams::max_iterations)
      .def_readwrite("rotation_convergence_thresh", &loam::RegistrationParams::rotation_convergence_thresh)
      .def_readwrite("position_convergence_thresh", &loam::RegistrationParams::position_convergence_thresh)
      .def_readwrite("min_associations", &loam::RegistrationParams::min_associations);

  py::class_<loam::RegistrationDetail::IterationInfo>(m, "RegistrationIterationInfo")
      .def(py::init<const loam::Pose3d, const std::vector<std::pair<size_t, size_t>>,
                    const std::vector<std::pair<size_t, size_t>>, const loam::Pose3d>(),
           py::arg("target_T_source_init"), py::arg("edge_associations"), py::arg("plane_associations"),
           py::arg("estimate_update"))
      .def_readwrite("target_T_source_init", &loam::RegistrationDetail::IterationInfo::target_T_source_init)
      .def_readwrite("edge_associations", &loam::RegistrationDetail::IterationInfo::edge_associations)
      .def_readwrite("plane_associations", &loam::RegistrationDetail::IterationInfo::plane_associations)
      .def_readwrite("estimate_update", &loam::RegistrationDetail::IterationInfo::estimate_update);

  py::enum_<loam::RegistrationDetail::TerminationType>(m, "RegistrationTerminationType")
      .value("CONVERGED", loam::RegistrationDetail::TerminationType::CONVERGED)
      .value("MAX_ITER", loam::RegistrationDetail::TerminationType::MAX_ITER)
      .value("INSUFFICIENT_ASSOCIATIONS", loam::RegistrationDetail::TerminationType::INSUFFICIENT_ASSOCIATIONS)
      .export_values();

  py::class_<loam::RegistrationDetail, std::shared_ptr<loam::RegistrationDetail>>(m, "RegistrationDetail")
      .def(py::init<>())
      .def_readwrite("iteration_info", &loam::RegistrationDetail::iteration_info)
      .def_readwrite("termination_type", &loam::RegistrationDetail::termination_type);

  m.def(
      "registerFeatures",
      [](const PyFeatures& source, const PyFeatures& target, const loam::Pose3d& init, const loam::RegistrationParams& params,
         std::shared_ptr<loam::RegistrationDetail> detail) {
        const py::array_t<double> se = as_xyz64(source.edge_points), sp = as_xyz64(source.planar_points),
                                  te = as_xyz64(target.edge_points), tp = as_xyz64(target.planar_points);
        const loamgpu_reg_params rp = loam::gpu::toC(params);
        double in7[7], out7[7];
        loam::gpu::poseTo7(init, in7);
        std::unique_ptr<loam::gpu::DetailBuffers> buf;
        if (detail) buf.reset(new loam::gpu::DetailBuffers(params, (size_t)se.shape(0), (size_t)sp.shape(0)));
        {
          py::gil_scoped_release nogil;
          loamgpu_ctx* ctx = ThreadContext::get();
          check(ctx, loamgpu_register(ctx, se.data(), (uint64_t)se.shape(0), sp.data(), (uint64_t)sp.shape(0), te.data(),
                                      (uint64_t)te.shape(0), tp.data(), (uint64_t)tp.shape(0), in7, &rp, out7,
                                      buf ? &buf->d : nullptr));
        }
        if (detail) buf->appendTo(*detail);
        return loam::gpu::poseFrom7(out7);
      },
      py::arg("source"), py::arg("target"), py::arg("target_T_source_init"), py::arg("params") = loam::RegistrationParams(),
      py::arg("detail") = std::shared_ptr<loam::RegistrationDetail>());

  // ------------------------------------------------------------------ extensions
  m.def(
      "odometry",
      [](const py::array_t<float, py::array::c_style | py::array::forcecast>& scans, const loam::LidarParams& lp,
         const loam::FeatureExtractionParams& fp, const loam::RegistrationParams& params) {
        if (scans.ndim() != 3 || scans.shape(2) != 4)
          throw std::invalid_argument("scans must be a (S, N, 4) float32 array of {x, y, z, .} records");
        const uint64_t S = (uint64_t)scans.shape(0);
        if ((size_t)scans.shape(1) != lp.scan_lines * lp.points_per_line) {
          std::vector<char> dummy((size_t)scans.shape(1));
          loam::validateLidarScan(dummy, lp);
        }
        const py::ssize_t pairs = S > 0 ? (py::ssize_t)S - 1 : 0;
        py::array_t<double> poses({pairs, (py::ssize_t)7});
        py::array_t<int32_t> term(pairs);
        py::array_t<uint32_t> iters(pairs), ne((py::ssize_t)S), np_((py::ssize_t)S);
        const loamgpu_lidar_params clp = loam::gpu::toC(lp);
        const loamgpu_fe_params cfp = loam::gpu::toC(fp);
        const loamgpu_reg_params crp = loam::gpu::toC(params);
        if (S) {
          double* pp = poses.mutable_data();
          int32_t* pt = term.mutable_data();
          uint32_t *pi = iters.mutable_data(), *pe = ne.mutable_data(), *pn = np_.mutable_data();
          py::gil_scoped_release nogil;
          loamgpu_ctx* ctx = ThreadContext::get();
          check(ctx, loamgpu_odometry_host(ctx, scans.data(), S, &clp, &cfp, &crp, pp, pt, pi, pe, pn));
        }
        return py::make_tuple(poses, term, iters, ne, np_);
      },
      py::arg("scans"), py::arg("lidar_params"), py::arg("feature_params") = loam::FeatureExtractionParams(),
      py::arg("registration_params") = loam::RegistrationParams(),
      "EXTENSION: extract every scan once and register scan k+1 onto scan k from an identity initial estimate; returns "
      "(poses [S-1,7] qx qy qz qw tx ty tz, termination, outer iterations, edge counts, planar counts)");
  m.def("setDevice", &loam::gpu::setDevice, py::arg("device"), "EXTENSION: CUDA device of this thread's subsequent calls");
}
