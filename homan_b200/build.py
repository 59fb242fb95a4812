"""Builds homan_b200/libhoman_b200.so (sm_100a) with nvcc. In-tree so that the .so travels to the GPU box.

    python -m homan_b200.build [--force]
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libhoman_b200.so")
OBJ_DIR = os.path.join(_HERE, "csrc", "_obj")

# file -> extra flags.  raster / sdf evaluate coverage, ownership and inside/outside predicates with
# individually rounded fp32 operations (no FMA contraction) to agree bit for bit with the CPU oracle.
SOURCES = {
    "api.cu": [],
    "raster.cu": ["-fmad=false"],
    "mano.cu": [],
    "geom.cu": [],
    "sdf.cu": ["-fmad=false"],
    "contact.cu": [],
}
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          ]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, "common.cuh"), os.path.join(_HERE, "..", "include", "homan_b200.h"),
            os.path.abspath(__file__)]
    objs, relink = [], force or not os.path.exists(OUT)
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s.replace(".cu", ".o"))
        stale = force or not os.path.exists(obj) or any(
            os.path.getmtime(d) > os.path.getmtime(obj) for d in [src] + deps)
        if stale:
            cmd = [_nvcc()] + COMMON + SOURCES[s] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
            relink = True
        objs.append(obj)
    if relink:
        subprocess.check_call([_nvcc(), "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
