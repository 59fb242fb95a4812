"""ctypes binding of libhoman_b200.so (the C ABI declared in include/homan_b200.h).

There is no CPU fallback: if the shared library is missing the import of any product module that
needs a kernel raises, and every entry point raises on a non-zero return code.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HOMAN_B200_LIB") or os.path.join(_HERE, "libhoman_b200.so")   # env: an alternative build

_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float

HEADER_PATH = os.path.join(_HERE, "..", "include", "homan_b200.h")


def parse_header(path=HEADER_PATH):
    """{entry point: argument kinds} from the C header (p: pointer, i: int, f: float); all return int."""
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\bint\s+(hm_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        kinds = ""
        if args and args != "void":
            for a in args.split(","):
                kinds += "p" if "*" in a else ("f" if re.search(r"\bfloat\b", a) else "i")
        sigs[name] = kinds
    return sigs


SIGNATURES = parse_header()
_KIND = {"p": _P, "i": _I, "f": _F}


class HomanB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Loads the shared library (building is the job of `python -m homan_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HomanB200Error(
                f"{LIB_PATH} is missing: build it with `python -m homan_b200.build` "
                "(homan_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        handle.hm_version.restype = _I
        handle.hm_last_error.restype = ctypes.c_char_p
        for name, kinds in SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                continue  # reported by tests/test_abi.py; calling it raises below
            fn.argtypes = [_KIND[k] for k in kinds]
            fn.restype = _I
        _lib = handle
    return _lib


def call(name, *args):
    handle = lib()
    fn = getattr(handle, name, None)
    if fn is None:
        raise HomanB200Error(f"{name} is not exported by {LIB_PATH}")
    rc = fn(*args)
    if rc != 0:
        raise HomanB200Error(f"{name} failed (rc={rc}): {handle.hm_last_error().decode()}")


def ptr(t):
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise HomanB200Error("non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
