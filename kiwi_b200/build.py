"""Build libkiwi_b200.so (hand-written sm_100a kernels + C ABI) in-tree.

    python -m kiwi_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the shared library lands next to this file so that
it travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libkiwi_b200.so")
MINIMIZER = os.path.join(HERE, "kiwi_minimizer")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
CXX = os.environ.get("CXX", "g++")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# host arithmetic follows the reference statement by statement: no FMA contraction
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread",
             "-I" + os.path.join(CUDA_HOME, "include"), "-Wall", "-Wno-unused-function", "-Wno-misleading-indentation"]

CU_SOURCES = ["kernels.cu", "gulunay.cu", "eikonal.cu", "synth_exact.cu"]
# per-file extra flags: the interpolation kernels are written operation by operation (no FMA contraction)
EXTRA_NVCC_FLAGS = {"gulunay.cu": ["-fmad=false"], "eikonal.cu": ["-fmad=false"], "synth_exact.cu": ["-fmad=false"]}
CXX_SOURCES = ["engine.cpp", "host_math.cpp", "gfdb_host.cpp", "source_eikonal_host.cpp", "lm_host.cpp", "gfdb_hdf_host.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "kiwi_b200.h"))
    objs = []
    for src in CU_SOURCES + CXX_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src + ".o")
        objs.append(obj)
        if force or _newer(obj, [path] + headers):
            if src.endswith(".cu"):
                cmd = [NVCC] + NVCC_FLAGS + EXTRA_NVCC_FLAGS.get(src, []) + ["-c", path, "-o", obj]
            else:
                cmd = [CXX] + CXX_FLAGS + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if src.endswith(".cu"):
                with open(os.path.join(OBJ, src + ".ptxas.log"), "w") as f:
                    f.write(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("compilation failed: " + " ".join(cmd))
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed: " + " ".join(cmd))
    # the `minimizer`-compatible command front-end (text protocol of minimizer.f90:1676-1813)
    main_src = os.path.join(CSRC, "minimizer_main.cpp")
    if force or _newer(MINIMIZER, [main_src, LIB] + headers):
        cmd = [CXX, "-O2", "-std=c++17", main_src, "-o", MINIMIZER, "-L" + HERE, "-lkiwi_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
