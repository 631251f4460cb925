"""CPU: libloamgpu.so builds (nvcc cross-compiles sm_100a without a GPU), loads, exports every symbol that
include/loamgpu.h declares, and fails loudly — never falls back — when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from loam_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "loamgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(loamgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = _capi.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    assert sorted(_capi.SYMBOLS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(_capi.CLidarParams) == 32
    assert C.sizeof(_capi.CFeParams) == 64
    assert C.sizeof(_capi.CRegParams) == 96


def test_defaults_match_reference():
    fe = _capi.default_fe_params()  # features.h:37-66
    assert (fe.neighbor_points, fe.number_sectors, fe.max_edge_feats_per_sector, fe.max_planar_feats_per_sector) == (3, 6, 10, 50)
    assert (fe.edge_feat_threshold, fe.planar_feat_threshold, fe.occlusion_thresh, fe.parallel_thresh) == (100.0, 1.0, 0.5, 1.0)
    rp = _capi.default_reg_params()  # registration.h:40-75
    assert (rp.num_edge_neighbors, rp.max_edge_neighbor_dist, rp.min_line_fit_points, rp.min_line_condition_number) == (5, 1.0, 3, 10.0)
    assert (rp.num_plane_neighbors, rp.max_plane_neighbor_dist, rp.min_plane_fit_points, rp.max_avg_point_plane_dist) == (5, 2.0, 4, 0.1)
    assert (rp.max_iterations, rp.rotation_convergence_thresh, rp.position_convergence_thresh, rp.min_associations) == (10, 1e-3, 1e-2, 100)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(_capi.LoamGpuError) as ei:
        _capi.Context(0)
    assert ei.value.code == _capi.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
    import loam_b200
    import numpy as np
    with pytest.raises(_capi.LoamGpuError):
        loam_b200.extractFeatures(np.zeros((11, 3)), loam_b200.LidarParams(1, 11, 0.1, 10.0))


def test_product_package_never_imports_oracle():
    """The oracle is test infrastructure; nothing under loam_b200/ or include/ may reference it."""
    bad = []
    for base in (os.path.join(ROOT, "loam_b200"), os.path.join(ROOT, "include")):
        for dp, _, fs in os.walk(base):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|pyoracle|liboracle|loam_oracle|libloam_ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
