"""ctypes binding of libdexbotic_b200.so (the C-ABI declared in include/*.h).

The library is the product; there is NO CPU or PyTorch fallback behind these calls.  If the shared
object is missing (not built) or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_REPO = _PKG.parent
LIB_PATH = _PKG / "libdexbotic_b200.so"
INCLUDE_DIR = _REPO / "include"

BF16, F32 = 0, 1
ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU, ACT_SILU, ACT_RELU = range(6)
ACT_CODES = {
    None: ACT_NONE, "none": ACT_NONE, "gelu": ACT_GELU_ERF, "gelu_erf": ACT_GELU_ERF,
    "gelu_tanh": ACT_GELU_TANH, "gelu_pytorch_tanh": ACT_GELU_TANH, "quick_gelu": ACT_QUICK_GELU,
    "silu": ACT_SILU, "swish": ACT_SILU, "relu": ACT_RELU,
}


class GemmArgs(C.Structure):
    """Mirror of b200_gemm_args (include/dexbotic_b200.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("b2", C.c_void_p), ("d", C.c_void_p),
        ("ab_dtype", C.c_int32), ("d_dtype", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("a_ld", C.c_int64), ("a_s2", C.c_int64), ("a_s3", C.c_int64),
        ("b_ld", C.c_int64), ("b_s2", C.c_int64), ("b_s3", C.c_int64),
        ("d_ld", C.c_int64), ("d_s2", C.c_int64), ("d_s3", C.c_int64),
        ("a_z2", C.c_int32), ("b_z2", C.c_int32),
        ("z_lo", C.c_int32), ("z_hi", C.c_int32),
        ("a_div", C.c_int32), ("a_mul", C.c_int32), ("a_seg", C.c_int32),
        ("b_div", C.c_int32), ("b_mul", C.c_int32), ("b_seg", C.c_int32),
        ("k_segs", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("bias_dtype", C.c_int32),
        ("residual", C.c_void_p), ("res_dtype", C.c_int32),
        ("res_ld", C.c_int64), ("res_s2", C.c_int64), ("res_s3", C.c_int64),
        ("aux", C.c_void_p), ("aux2", C.c_void_p), ("aux_ld", C.c_int64),
        ("act", C.c_int32), ("dual_b", C.c_int32), ("block_n", C.c_int32),
        ("glu_bwd", C.c_int32), ("glu_g", C.c_void_p), ("glu_u", C.c_void_p), ("d2", C.c_void_p), ("glu_ld", C.c_int64),
    ]


_CTYPE = {
    "void": None, "int": C.c_int, "float": C.c_float, "int32_t": C.c_int32, "int64_t": C.c_int64,
    "uint64_t": C.c_uint64, "double": C.c_double,
}


def _parse_header_prototypes() -> dict[str, tuple]:
    """Every `int b200_xxx(...)` prototype in include/*.h -> (restype, [argtypes])."""
    protos: dict[str, tuple] = {}
    for hdr in sorted(INCLUDE_DIR.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", hdr.read_text(), flags=re.S)
        for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(b200_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
            ret, name, args = m.group(1), m.group(2), m.group(3).strip()
            argtypes = []
            if args and args != "void":
                for a in args.split(","):
                    a = a.strip()
                    if "*" in a:
                        argtypes.append(C.c_void_p)
                    else:
                        base = a.replace("const", "").split()[0]
                        argtypes.append(_CTYPE[base])
            restype = {"int": C.c_int, "int64_t": C.c_int64, "const char*": C.c_char_p}[ret]
            protos[name] = (restype, argtypes)
    return protos


EXPORTED = _parse_header_prototypes()
_lib = None


def lib_available() -> bool:
    return LIB_PATH.exists()


def load():
    """Load the shared object (lazily) and attach the prototypes declared in include/*.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C dexbotic_b200/csrc`). dexbotic_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in EXPORTED.items():
        fn = getattr(lib, name)  # AttributeError here = header declares a symbol the .so lacks
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"dexbotic_b200 {what} failed: {last_error()}")


def launch_count() -> int:
    return int(load().b200_launch_count())
