"""ctypes binding of libctagan.so (the C ABI declared in include/ctagan.h).

The signatures are parsed from the header itself, so the Python side can never drift from the ABI, and a missing
symbol or a missing library is a hard error: there is no CPU / PyTorch fallback for the product path.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTAGAN_LIB") or os.path.join(_HERE, "libctagan.so")       # (CTAGAN_LIB: developer A/B of two builds)
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "ctagan.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC, ENGINE_GENERIC = 0, 1, 2, 3


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "Hi", "Wi", "Ci", "Ho", "Wo", "Co", "KH", "KW", "stride", "dil", "pad_h", "pad_w", "act", "dtype", "gy_margin")]


MAX_GROUPS = 4


class ConvGroups(ctypes.Structure):
    _fields_ = [("groups", ctypes.c_int32), ("slot", ctypes.c_int32 * MAX_GROUPS)]


class PackItem(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("wp", ctypes.c_void_p), ("O", ctypes.c_int32), ("I", ctypes.c_int32), ("KH", ctypes.c_int32),
                ("KW", ctypes.c_int32), ("mode", ctypes.c_int32)]


class AdamItem(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p), ("wp0", ctypes.c_void_p),
                ("wp1", ctypes.c_void_p), ("O", ctypes.c_int32), ("I", ctypes.c_int32), ("KH", ctypes.c_int32), ("KW", ctypes.c_int32),
                ("g_packed", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def _ctype_of(decl: str):
    d = decl.strip()
    if "**" in d:
        return ctypes.POINTER(ctypes.c_void_p)
    if "*" in d:
        if "ctagan_conv_geom" in d:
            return ctypes.POINTER(ConvGeom)
        if "ctagan_conv_groups" in d:
            return ctypes.POINTER(ConvGroups)
        if "ctagan_pack_item" in d:
            return ctypes.POINTER(PackItem)
        if "ctagan_adam_item" in d:
            return ctypes.c_void_p          # host array or device pointer, depending on the entry point
        return ctypes.c_void_p
    base = d.split()[0] if d.split()[0] != "const" else d.split()[1]
    return {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
            "double": ctypes.c_double, "size_t": ctypes.c_size_t}[base]


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object]]]:
    """{symbol: (restype, [argtypes])} for every prototype in include/ctagan.h."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int|size_t)\s*(ctagan_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else (ctypes.c_size_t if ret == "size_t" else ctypes.c_int)
        args = args.strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (restype, argtypes)
    return protos


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python cta-gan_b200/build.py` (nvcc, sm_100a). "
            "The CTA-GAN B200 path has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libctagan: " + load().ctagan_last_error().decode())
