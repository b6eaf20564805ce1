"""oracle/warp_oracle.py -- TEST INFRASTRUCTURE ONLY (ctypes face of warp_oracle.c).

CPU restatement of ``cv2.warpPerspective`` as the reference calls it
(/root/reference/vis_homo.py:89,91; /root/reference/bev/tool/compo.py:38,46,47).
Pinned bit-for-bit against cv2 4.13.0.92 and tests/golden/ by tests/test_oracle_warp.py.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  bev_b200/ never does.
"""
import ctypes
import os
import subprocess

import numpy as np

INTER_NEAREST = 0
INTER_LINEAR = 1
WARP_INVERSE_MAP = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile warp_oracle.c -> liboracle (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "libbevoracle.so")
    src = os.path.join(_HERE, "warp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libbevoracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libbevoracle.so")
        if not os.path.exists(so):
            build()
        lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.bevo_invert3x3.argtypes = [dp, dp]
        lib.bevo_invert3x3.restype = ctypes.c_int
        lib.bevo_warp_perspective.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, dp]
        lib.bevo_warp_perspective.restype = ctypes.c_int
        lib.bevo_touched_pixels.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int,
            ctypes.POINTER(ctypes.c_int)]
        lib.bevo_touched_pixels.restype = ctypes.c_long
        _LIB = lib
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def invert3x3(H):
    """Adjugate inverse, bit-equal to cv2.invert on a 3x3 float64 matrix."""
    H = np.ascontiguousarray(H, dtype=np.float64).reshape(3, 3)
    M = np.empty((3, 3), np.float64)
    _lib().bevo_invert3x3(_dptr(H), _dptr(M))
    return M


def warp_perspective(src, M, dsize, flags=INTER_LINEAR, borderValue=0):
    """Same call shape as cv2.warpPerspective(src, M, dsize, flags=..., borderValue=...).

    src: HxW or HxWxC, uint8 / float32 / float16 (float16 goes through the float32 path and is
    rounded back, which is the oracle SURVEY.md 8c defines because cv2 rejects float16).
    """
    src = np.asarray(src)
    squeeze = src.ndim == 2
    s = src[:, :, None] if squeeze else src
    s = np.ascontiguousarray(s)
    half = s.dtype == np.float16
    if half:
        s = s.astype(np.float32)
    if s.dtype == np.uint8:
        dt = 0
    elif s.dtype == np.float32:
        dt = 2
    else:
        raise TypeError("oracle supports uint8/float16/float32, got %s" % src.dtype)
    h, w, c = s.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    out = np.empty((dh, dw, c), s.dtype)
    H = np.ascontiguousarray(M, dtype=np.float64).reshape(3, 3)
    bv = np.zeros(4, np.float64)
    b = np.atleast_1d(np.asarray(borderValue, dtype=np.float64))
    bv[:len(b)] = b[:4]
    rc = _lib().bevo_warp_perspective(s.ctypes.data, h, w, c, dt, out.ctypes.data, dh, dw,
                                      _dptr(H), int(flags), _dptr(bv))
    if rc != 0:
        raise ValueError("bevo_warp_perspective failed rc=%d" % rc)
    if half:
        out = out.astype(np.float16)
    return out[:, :, 0] if squeeze else out


def touched_pixels(ssize, dsize, M, flags=INTER_LINEAR):
    """(T, row_min, row_max): distinct in-bounds source pixels referenced."""
    sw, sh = int(ssize[0]), int(ssize[1])
    dw, dh = int(dsize[0]), int(dsize[1])
    H = np.ascontiguousarray(M, dtype=np.float64).reshape(3, 3)
    rr = (ctypes.c_int * 2)()
    t = _lib().bevo_touched_pixels(sh, sw, dh, dw, _dptr(H), int(flags), rr)
    if t < 0:
        raise ValueError("bevo_touched_pixels failed rc=%d" % t)
    return int(t), int(rr[0]), int(rr[1])


def algo_bytes(ssize, dsize, M, channels, elem_size, flags=INTER_LINEAR):
    """Algorithmic bytes per frame = (T + D) * C * es."""
    t, _, _ = touched_pixels(ssize, dsize, M, flags)
    return (t + int(dsize[0]) * int(dsize[1])) * channels * elem_size
