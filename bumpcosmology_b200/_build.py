"""Build the in-tree CUDA library (sm_100a only).  Used by `__graft_entry__.build()` and `make`-less setups:

    python -m bumpcosmology_b200._build
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbump_b200.so")
PEAK = os.path.join(PKG, "bump_peak")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp"))]
    src.append(os.path.join(os.path.dirname(PKG), "include", "bump.h"))
    return src


def build(force=False, verbose=False):
    """Compile csrc/bump_lib.cu + csrc/bump_nuts.cpp -> libbump_b200.so and csrc/bump_peak.cu -> bump_peak (fp64 issue-rate probe)."""
    nvcc = _nvcc()
    src = sources()
    if force or _stale(LIB, src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-shared", "-o", LIB, os.path.join(CSRC, "bump_lib.cu"), os.path.join(CSRC, "bump_nuts.cpp"), "-ldl"]
        subprocess.run(cmd, check=True)
    peak_src = os.path.join(CSRC, "bump_peak.cu")
    if os.path.exists(peak_src) and (force or _stale(PEAK, [peak_src])):
        subprocess.run([nvcc] + NVCC_FLAGS + ["-o", PEAK, peak_src], check=True)
    return LIB


def build_xla_ffi():
    """Compile csrc/bump_xla_ffi.cc -> libbump_xla_ffi.so against jaxlib's headers (only where JAX is installed)."""
    import jaxlib   # ImportError here means: no JAX, no adaptor (the ctypes path does not need it)
    inc = os.path.join(os.path.dirname(jaxlib.__file__), "include")
    out = os.path.join(PKG, "libbump_xla_ffi.so")
    subprocess.run([_nvcc()] + NVCC_FLAGS + ["-shared", "-I", inc, "-o", out, os.path.join(CSRC, "bump_xla_ffi.cc"),
                                              "-L", PKG, "-l:libbump_b200.so", "-Xlinker", "-rpath=$ORIGIN"], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--xla-ffi" in sys.argv:
        print(build_xla_ffi())
