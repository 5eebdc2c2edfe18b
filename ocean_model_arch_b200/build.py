"""Builds libswcuda.so (CUDA kernels for sm_100a + C ABI + C++ host init) in-tree with nvcc.

    python -m ocean_model_arch_b200.build [--force]

-fmad=false is REQUIRED: sw_formulas.cuh promises results bitwise equal to a strict IEEE
evaluation of the reference's Fortran expressions (no contracted multiply-adds).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libswcuda.so")
SOURCES = ["sw_api_level_a.cu", "sw_kernels_ref.cu", "sw_kernels_fused.cu", "sw_kernels_march.cu", "sw_init.cu", "sw_ctx.cu", "sw_host.cpp"]
HEADERS = ["sw_common.h", "sw_formulas.cuh", "sw_cells.cuh", "sw_fast.cuh", "sw_tables.h", "sw_fused.h", os.path.join("..", "..", "include", "swcuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), lib=None, objdir=None):
    """extra_flags / lib / objdir: tuning variants (e.g. -DSWCU_MW=3) built beside the product library."""
    lib = lib or LIB
    if lib == LIB and not force and not needs_build():
        return LIB
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    logs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc(), "-ccbin", ccbin, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + s)
        objs.append(o)
    cmd = [_nvcc(), "-ccbin", ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", lib, *objs, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
