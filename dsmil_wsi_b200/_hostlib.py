"""ctypes binding of libdsmil_host.so (csrc_host/bagcsv.c): the native reader / writer of the reference's bag
feature CSV.  Host-only code; like the device library it has no silent fallback -- a missing build raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdsmil_host.so")

SIGNATURES = {
    "dsmil_host_abi_version": (C.c_int32, []),
    "dsmil_csv_format_bag": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64]),
    "dsmil_csv_write_bag": (C.c_int64, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32]),
    "dsmil_csv_shape": (C.c_int32, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "dsmil_csv_parse_bag": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "dsmil_jpeg_header_bytes": (C.c_int32, []),
    "dsmil_jpeg_parse": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "dsmil_jpeg_parse_batch": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "dsmil_files_offsets": (C.c_int64, [C.c_void_p, C.c_int32, C.c_void_p]),
    "dsmil_files_read": (C.c_int64, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
}
ERRORS = {-1: "bad argument", -2: "I/O error", -3: "ragged row (field count differs from the header)",
          -4: "field is not a number", -5: "output buffer too small"}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"libdsmil_host.so not found at {LIB_PATH}. Build it with `python -m dsmil_wsi_b200.build` "
                           "(gcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dsmil_host_abi_version() != 1:
        raise RuntimeError("libdsmil_host.so ABI version mismatch")
    _lib = lib
    return lib
