"""ORACLE (test infrastructure). Builds / loads the plain-C part of the CPU oracle.

gcc -O2 -fopenmp -ffp-contract=off  oracle/csrc/*.c  ->  oracle/_build/liboracle.so
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = [os.path.join(_HERE, "csrc", f) for f in ("nmr_raster.c", "sdf_grid.c")]
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "liboracle.so")
_lib = None


def build(force=False):
    os.makedirs(_OUT_DIR, exist_ok=True)
    stale = force or not os.path.exists(_OUT) or any(
        os.path.getmtime(s) > os.path.getmtime(_OUT) for s in _SRC)
    if stale:
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-o", _OUT] + _SRC + ["-lm"]
        subprocess.check_call(cmd)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        c_f = ctypes.c_void_p
        _lib.nmr_face_index_map.argtypes = [c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_float, c_f, c_f]
        _lib.nmr_alpha_flip_pool.argtypes = [c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f, c_f]
        _lib.nmr_pixel_map_bwd.argtypes = [c_f, c_f, c_f, c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_float, c_f]
        _lib.sdf_grid.argtypes = [c_f, ctypes.c_int, c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f]
        for fn in (_lib.nmr_face_index_map, _lib.nmr_alpha_flip_pool, _lib.nmr_pixel_map_bwd, _lib.sdf_grid):
            fn.restype = None
    return _lib


if __name__ == "__main__":
    print(build(force=True))
