"""TEST INFRASTRUCTURE: builds and binds oracle/jpeg_host_check.c, a serial CPU build of the JPEG arithmetic the
device kernels use (dsmil_wsi_b200/csrc/jpeg_core.h).  The oracle of the JPEG loader is PIL itself (the decoder the
reference calls at compute_feats.py:28); this checker lets the `-m "not gpu"` suite prove the shared arithmetic
bit-exact against PIL without a GPU.  Never imported by the product."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "jpeg_host_check.c")
CORE = os.path.join(HERE, "..", "dsmil_wsi_b200", "csrc", "jpeg_core.h")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libjpegcheck.so")

_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for oracle/jpeg_host_check.c")
    r = subprocess.run([cc, "-O2", "-std=c11", "-Wall", "-Wextra", "-shared", "-fPIC", "-o", LIB, SRC],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the JPEG checker failed:\n" + r.stdout + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.jpegcheck_size.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.jpegcheck_size.restype = C.c_int
        lib.jpegcheck_decode.argtypes = [C.c_char_p, C.c_int64, C.c_void_p]
        lib.jpegcheck_decode.restype = C.c_int
        _lib = lib
    return _lib


def decode(data: bytes):
    """(status, RGB uint8 [H, W, 3] or None) of one JPEG file through the CPU build of jpeg_core.h."""
    lib = load()
    w, h, n = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.jpegcheck_size(data, len(data), w, h, n)
    if rc:
        return rc, None
    out = np.zeros((h.value, w.value, 3), np.uint8)
    rc = lib.jpegcheck_decode(data, len(data), out.ctypes.data)
    return rc, (out if rc == 0 else None)
