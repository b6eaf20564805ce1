"""torch.ops.bev_cuda -- the hot path as registered PyTorch operators.

``load()`` loads bev_b200/libbev_torch_ops.so (built by bev_b200/csrc/build_torch_ops.py, part of
``__graft_entry__.build()``): a thin C++ extension whose operators check tensors, allocate the
output with torch's caching allocator, take the current CUDA stream and call the extern "C" entry
points of libbev_b200.so.  After ``load()``:

    torch.ops.bev_cuda.warp_perspective(frames, M_cpu_f64, w, h, flags, 0, 0.0, mat_index_or_None)
    torch.ops.bev_cuda.project_points(pts, H_cpu_f64)
    torch.ops.bev_cuda.rbox_corners_project(xywhr, H_cpu_f64_or_None, mode)   # 0 "bev", 1 "world"
    torch.ops.bev_cuda.corners_to_rbox(xy8, H_cpu_f64_or_None, mode)
    torch.ops.bev_cuda.rbox_similarity(rbox, H_cpu_f64, src_mode)

Only a CUDA implementation is registered; there is no CPU fallback.  The Python modules
(bev_b200.homo, bev_b200.rbox_torch) reach the same entry points through ctypes; both routes run
the same kernels and give identical bytes (tests/test_torch_ops_gpu.py).
"""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
OPS_PATH = os.path.join(_HERE, "libbev_torch_ops.so")
_loaded = False


def load():
    """Register torch.ops.bev_cuda.* (idempotent).  Raises if the extension has not been built."""
    global _loaded
    if _loaded:
        return
    import torch
    if not os.path.exists(OPS_PATH):
        raise ImportError("bev_b200: %s is missing; build it with `python bev_b200/csrc/build_torch_ops.py` "
                          "(or __graft_entry__.build())" % OPS_PATH)
    torch.ops.load_library(OPS_PATH)
    _loaded = True
