"""Build libloamgpu.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

-fmad=false : the reference is built for baseline x86-64 (no FMA); feature and correspondence
              indices must be bit-exact, so no multiply-add is ever contracted.
-lineinfo   : ncu source pages map to the .cu files.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("extract.cu", "register.cu", "bvh_big.cu", "capi.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "kernels.h", "bvh.cuh")] + [
    os.path.join(ROOT, "include", "loamgpu.h")]
LIB = os.path.join(HERE, "lib", "libloamgpu.so")


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(HERE, "csrc"), "-o", LIB] + SRC + ["-ldl"]
    cmd[1:1] = os.environ.get("LOAMGPU_NVCC_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libloamgpu.so")
    if verbose:
        print(r.stderr)
    return LIB


def python_module_path() -> str:
    import sysconfig
    return os.path.join(HERE, "loam_python" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_python_module(force: bool = False) -> str:
    """The pybind11 module `loam_python` (python/loam_b200_bindings.cpp) next to the package, linked against
    lib/libloamgpu.so through an $ORIGIN rpath.  Host code only (g++): the kernels live in libloamgpu.so."""
    import sysconfig

    import pybind11
    out = python_module_path()
    src = os.path.join(ROOT, "python", "loam_b200_bindings.cpp")
    deps = [src, LIB] + [os.path.join(ROOT, "include", "loam", f) for f in
                         ("common.h", "features.h", "registration.h", "geometry.h", "detail/gpu.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    # (plain g++, not $CXX: this image exports a wrapper there that links libstdc++ statically, and a module with its
    # own copy of the iostream machinery next to the shared one crashes on the first formatted exception message)
    cmd = [os.environ.get("LOAMGPU_CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "include", "loam_compat"), src, "-o", out, "-L", os.path.dirname(LIB), "-lloamgpu",
           "-Wl,-rpath,$ORIGIN/lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the loam_python module")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--python" in sys.argv:
        print(build_python_module(force="--force" in sys.argv))
