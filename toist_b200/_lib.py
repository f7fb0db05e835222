"""ctypes binding of libtoist_b200.so (the C ABI declared in include/toist_b200.h).

There is deliberately no fallback: if the shared library is missing or a kernel fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libtoist_b200.so"

MAX_TAPS = 49
GEMM_FWD, GEMM_DGRAD, GEMM_WGRAD = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3
BF16, F32 = 0, 1


class ToistError(RuntimeError):
    pass


class Tensor4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int64 * 4), ("stride", C.c_int64 * 4)]


class Tap(C.Structure):
    _fields_ = [("dx", C.c_int16), ("dy", C.c_int16), ("dn", C.c_int16), ("pad_", C.c_int16), ("col", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("a", Tensor4),
        ("b", Tensor4),
        ("ext_x", C.c_int32), ("ext_y", C.c_int32), ("ext_n", C.c_int32),
        ("tile_x", C.c_int32), ("tile_y", C.c_int32), ("tile_n", C.c_int32),
        ("stride_x", C.c_int32), ("stride_y", C.c_int32),
        ("n_cols", C.c_int32),
        ("m_rows", C.c_int32),
        ("k_per_tap", C.c_int32),
        ("n_taps", C.c_int32),
        ("taps", Tap * MAX_TAPS),
        ("b_batched", C.c_int32),
        ("batch_y", C.c_int32), ("batch_n", C.c_int32),
        ("splits", C.c_int32),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32),
        ("out_sx", C.c_int64), ("out_sy", C.c_int64), ("out_sn", C.c_int64),
        ("alpha", C.c_float),
        ("col_scale", C.c_void_p),
        ("col_shift", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("res", C.c_void_p),
        ("res_dtype", C.c_int32),
        ("mask", C.c_void_p),
        ("aux", C.c_void_p),
        ("act", C.c_int32),
        ("accumulate", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("lse", C.c_void_p),
        ("key_mask", C.c_void_p), ("seed", C.c_void_p),
        ("q_ss", C.c_int64), ("q_sb", C.c_int64), ("k_ss", C.c_int64), ("k_sb", C.c_int64),
        ("v_ss", C.c_int64), ("v_sb", C.c_int64), ("o_ss", C.c_int64), ("o_sb", C.c_int64),
        ("sq", C.c_int32), ("sk", C.c_int32), ("b", C.c_int32), ("h", C.c_int32), ("d", C.c_int32),
        ("p_drop", C.c_float), ("site", C.c_uint32), ("reserved", C.c_int32),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [
        ("fwd", AttnDesc),
        ("dout", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("workspace", C.c_void_p),
        ("do_ss", C.c_int64), ("do_sb", C.c_int64), ("dq_ss", C.c_int64), ("dq_sb", C.c_int64),
        ("dk_ss", C.c_int64), ("dk_sb", C.c_int64), ("dv_ss", C.c_int64), ("dv_sb", C.c_int64),
    ]


_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads the shared library once. Raises ToistError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ToistError(
            f"{_LIB_PATH} is missing: run `python -m toist_b200.build` (nvcc, sm_100a). "
            "toist_b200 has no CPU or library fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    _declare(lib)
    if lib.toist_abi_version() != 1:
        raise ToistError("libtoist_b200.so ABI version mismatch; rebuild")
    if lib.toist_sizeof_gemm_desc() != C.sizeof(GemmDesc):
        raise ToistError("toist_gemm_desc layout mismatch between _lib.py and libtoist_b200.so; rebuild")
    if lib.toist_sizeof_attn_desc() != C.sizeof(AttnDesc) or lib.toist_sizeof_attn_bwd_desc() != C.sizeof(AttnBwdDesc):
        raise ToistError("toist_attn_desc layout mismatch between _lib.py and libtoist_b200.so; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().toist_last_error().decode("utf-8", "replace")
        raise ToistError(f"toist_b200 kernel call failed (status {rc}): {msg}")


_HEADER = Path(__file__).resolve().parent.parent / "include" / "toist_b200.h"
_SCALARS = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "float": C.c_float, "size_t": C.c_size_t,
            "uint32_t": C.c_uint32}


def parse_header():
    """Returns {symbol: (restype, [argtypes])} for every prototype in include/toist_b200.h, so the binding cannot
    drift from the declared C ABI."""
    import re

    text = _HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int64_t|int|size_t|const char\*)\s+(toist_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = C.c_char_p if ret == "const char*" else _SCALARS[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(C.POINTER(GemmDesc) if "toist_gemm_desc" in a else C.c_void_p)
                else:
                    ctype = a.replace("const ", "").split()[0]
                    argtypes.append(_SCALARS[ctype])
        protos[name] = (restype, argtypes)
    return protos


def _declare(lib: C.CDLL) -> None:
    for name, (res, args) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError here means the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
