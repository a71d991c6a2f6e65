"""Build libaanet_b200.so (the C-ABI library of include/aanet_b200.h) in-tree with nvcc for sm_100a.

    python -m aanet_b200.build [--force]

The library is self-contained (static cudart) and has no torch dependency; the Python side loads it
with ctypes (aanet_b200/_lib.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libaanet_b200.so")
SOURCES = ["capi.cu", "softargmin.cu", "correlation.cu", "correlation_umma.cu", "correlation_tma.cu", "cost5d.cu", "csa_fuse.cu", "refine.cu", "mdcn_fwd.cu", "mdcn_bwd.cu", "mdcn_bwd_umma.cu",
           "mdcn_api.cu", "conv_umma.cu", "halo_engine.cu", "deform_halo.cu", "deform_tmem.cu", "conv_umma_m0.cu", "conv_umma_m1.cu", "conv_umma_m2.cu", "conv_umma_m3.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "aanet_b200.h"))
    return out


def build(force=False, verbose=False, variant=None, defs=()):
    """variant / defs: a second build of the same library with extra -D switches (e.g. the instrumented kernels,
    variant="prof", defs=["-DAANET_TMEM_PROF"]) written to lib/libaanet_b200_<variant>.so; load it with
    AANET_B200_LIB=<path>."""
    LIB = globals()["LIB"] if not variant else os.path.join(HERE, "lib", "libaanet_b200_%s.so" % variant)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(p) for p in _deps())):
        return LIB
    objs = []
    procs = []
    build_dir = os.path.join(HERE, "lib", "obj" if not variant else "obj_" + variant)
    os.makedirs(build_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("AANET_NVCC_DEFS", "").split() + list(defs) + \
              (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[nvcc %s]\n%s\n" % (src, out.decode()))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                   check=True)
    return LIB


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant="prof", defs=["-DAANET_TMEM_PROF"]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
