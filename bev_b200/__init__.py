"""bev_b200 -- B200-native hot path of minghanz/bev: homography image<->BEV warp and batched
projection of points / rotated boxes, behind the reference's own Python surface
(``homo``, ``bev.BEVWorldSpec``, ``calib.Calib``, ``rbox_torch``).

    from bev_b200 import Calib, BEVWorldSpec, homo, rbox_torch
    bev = homo.warp_perspective(frames_cuda, H_bev_img, (bspec.u_size, bspec.v_size))

The compute runs in hand-written CUDA kernels for sm_100a in ``libbev_b200.so`` (C ABI in
include/bev_b200.h).  There is no CPU fallback.
"""
from . import compo, homo, rbox_torch
from .bev import BEVWorldSpec
from .calib import Calib
from .frozen_class import FrozenClass

__all__ = ["homo", "rbox_torch", "compo", "BEVWorldSpec", "Calib", "FrozenClass", "install_as_bev"]


def install_as_bev():
    """Register this package under the reference's import name so that unchanged caller code
    (``from bev.homo import ...``, ``from bev import Calib``) resolves to the B200 path."""
    import sys
    from . import bev as _bev, calib as _calib, frozen_class as _fc
    me = sys.modules[__name__]
    sys.modules["bev"] = me
    for name, mod in (("homo", homo), ("rbox_torch", rbox_torch), ("bev", _bev),
                      ("calib", _calib), ("frozen_class", _fc), ("tool.compo", compo)):
        sys.modules["bev." + name] = mod
    return me
