"""CPU: the C-ABI library loads without a GPU, exports every symbol include/bev_b200.h declares,
rejects bad arguments, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from bev_b200 import _native
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "bev_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bevk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    names = declared_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_native.SIGNATURES) == names  # python binding covers exactly the header


def test_version_and_host_invert():
    lib = _native.lib()
    assert lib.bevk_version() == 100
    H = np.array(util.load_json("homo_kat.json")["h_canon"])
    M = _native.invert3x3(H)
    assert np.allclose(M @ H, np.eye(3), atol=1e-9)
    assert not _native.invert3x3(np.zeros((3, 3))).any()


def _has_gpu():
    try:
        _native.device_info()
        return True
    except _native.NativeError:
        return False


def test_argument_errors_are_reported():
    lib = _native.lib()
    M = np.eye(3)
    dp = M.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    bad = [
        dict(n=1, sh=0, sw=4, dh=4, dw=4, c=3, dt=0, flags=1, bm=0),   # empty source
        dict(n=1, sh=4, sw=4, dh=4, dw=4, c=5, dt=0, flags=1, bm=0),   # channels
        dict(n=1, sh=4, sw=4, dh=4, dw=4, c=3, dt=3, flags=1, bm=0),   # float64 frames
        dict(n=1, sh=4, sw=4, dh=4, dw=4, c=3, dt=0, flags=2, bm=0),   # INTER_CUBIC
        dict(n=1, sh=4, sw=4, dh=4, dw=4, c=3, dt=0, flags=1, bm=1),   # BORDER_REPLICATE
        dict(n=1, sh=4, sw=40000, dh=4, dw=4, c=3, dt=0, flags=1, bm=0),
    ]
    buf = ctypes.create_string_buffer(64)
    for b in bad:
        rc = lib.bevk_warp_perspective(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(buf, ctypes.c_void_p),
                                       b["n"], b["sh"], b["sw"], b["dh"], b["dw"], b["c"], b["dt"], dp, 1,
                                       None, b["flags"], b["bm"], None, None)
        assert rc == -1, b
        assert lib.bevk_last_error()
    rc = lib.bevk_warp_perspective(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(buf, ctypes.c_void_p),
                                   3, 4, 4, 4, 4, 3, 0, dp, 2, None, 1, 0, None, None)
    assert rc == -1  # 2 matrices for 3 frames without mat_index


def test_no_cpu_fallback():
    import torch
    from bev_b200 import homo, rbox_torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        homo.warp_perspective(torch.zeros(8, 8, 3, dtype=torch.uint8), np.eye(3), (4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rbox_torch.xywhr2xyxy(torch.zeros(4, 5), "bev")
    with pytest.raises(AssertionError):
        rbox_torch.xywhr2xyxy(torch.zeros(4, 5), "image")
    if not _has_gpu():
        with pytest.raises(_native.NativeError, match="sm_100"):
            _native.device_info()
        with pytest.raises(_native.NativeError):
            _native.warp_perspective_host(np.zeros((8, 8, 3), np.uint8), np.eye(3), (4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bev_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "libbevoracle" not in text, fn


def test_torch_extension_loads_and_registers_the_operators():
    """The thin PyTorch C++ extension: loads on a machine without a GPU, registers
    torch.ops.bev_cuda.* with a CUDA implementation only (CPU tensors fail in the dispatcher)."""
    import pytest
    import torch
    from bev_b200 import torch_ops
    torch_ops.load()
    for name in ("warp_perspective", "project_points", "rbox_corners_project", "corners_to_rbox",
                 "rbox_similarity"):
        assert hasattr(torch.ops.bev_cuda, name), name
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.bev_cuda.project_points(torch.zeros((3, 2)), torch.eye(3, dtype=torch.float64))
