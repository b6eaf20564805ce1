"""Rotated-box and point projection on CUDA tensors.

Same function names, argument order and return shapes as /root/reference/bev/rbox_torch.py (so
``from bev_b200 import rbox_torch`` is a drop-in), plus the torch twins the reference only has in
numpy (``pts_world_bev``, ``xy82xywhr``, ``rbox_world_img`` -- bev/rbox.py:136,50,221) and the two
fused chains of BASELINE configs[2] (``xywhr_to_img_corners`` / ``img_corners_to_xywhr``).

Each call is ONE CUDA kernel that reads every box once and writes it once (the reference issues
~14-25 eager ATen ops with >= 10 temporaries per call).  Inputs must be float32/float64 CUDA
tensors; homographies are host float64 3x3 (numpy or CPU tensor).  Coordinate conventions
(reference rbox_torch.py:12-22):

  bev   : u right, v down; yaw 0 = +v, yaw = atan2(u, v); w along u, h along v
  world : right-handed x/y; yaw 0 = +x, yaw = atan2(y, x); h along x, w along y
"""
import math

from . import _native

_MODES = ("bev", "world")


def _mode(mode):
    assert mode in _MODES  # same AssertionError as the reference for a bad mode string
    return _native.MODE[mode]


def v2yaw(x, mode):
    """(N,2) direction vectors -> (N,) yaw (reference rbox_torch.py:24-31)."""
    return _native.rows_op("bevk_v2yaw", x, 2, (), _mode(mode))


def yaw2v(x, mode):
    """(N,) yaw -> (N,2) unit vectors (reference rbox_torch.py:33-40)."""
    return _native.rows_op("bevk_yaw2v", x, 1, (2,), _mode(mode))


def yaw2mat(x, mode):
    """(N,) yaw -> (N,2,2) rotation matrices (reference rbox_torch.py:42-50)."""
    return _native.rows_op("bevk_yaw2mat", x, 1, (2, 2), _mode(mode))


def xywhr2xyxy(x, mode, external_aa=False):
    """(N,5) [x,y,w,h,yaw] -> (N,8) corners tl,bl,br,tr (reference rbox_torch.py:52-99).

    ``external_aa=True`` is dead code in the reference (it raises IndexError there, SURVEY.md
    App. C) and is rejected here.
    """
    m = _mode(mode)
    if external_aa:
        raise NotImplementedError("external_aa=True is broken in the reference and not provided")
    return _native.rows_op("bevk_xywhr2xyxy", x, 5, (8,), m, H=None, has_H=True)


def xywhr2xyvec(xywhr, mode):
    """(N,5) -> (N,4) heading segment [x, y, x + h*dx, y + h*dy] (reference rbox_torch.py:101-112)."""
    return _native.rows_op("bevk_xywhr2xyvec", xywhr, 5, (4,), _mode(mode))


def xy82xyvec(xy8):
    """(N,8) corners -> (N,4) heading segment (reference rbox_torch.py:114-121)."""
    return _native.rows_op("bevk_xy82xyvec", xy8, 8, (4,))


def rbox_world_bev(rbox_src, H, src):
    """Similarity transform of (N,5) boxes between bev and world (reference rbox_torch.py:123-168).

    ``H`` must be affine and a similarity, else AssertionError as in the reference -- checked on
    the host copy of H before launch, so a CUDA tensor is never synchronised on.
    """
    return _native.rows_op("bevk_rbox_world_bev", rbox_src, 5, (5,), _mode(src), H=H, has_H=True)


# ---- torch twins of numpy-only reference functions ---------------------------------------------

def pts_world_bev(pts_src, H):
    """Homogeneous projection with divide; (N,2)->(N,2) or (N,3)->(N,3) (reference rbox.py:136-151)."""
    if pts_src.dim() == 1:
        pts_src = pts_src[None, :]
    dim = pts_src.shape[1]
    assert dim in (2, 3)
    return _native.rows_op("bevk_pts_project", pts_src, dim, (dim,), dim, H=H, has_H=True)


def dist_world_bev(dist_src, H):
    """Lengths through a similarity H: ``sqrt(H00^2 + H10^2) * dist_src``, any shape (reference
    rbox.py:153-160); AssertionError when H's two column norms differ, as there."""
    return _native.rows_op("bevk_dist_world_bev", dist_src, 1, (), H=H, has_H=True).reshape(dist_src.shape)


def angle_world_bev(angle_src, H, src):
    """Yaw angles (any shape, flattened like the reference) from ``src`` to the other system
    through H's upper-left 2x2 (reference rbox.py:162-171)."""
    return _native.rows_op("bevk_angle_world_bev", angle_src, 1, (), _mode(src), H=H, has_H=True)


def xy82xywhr(xy8, mode):
    """(N,8) corners -> (N,5) [x,y,w,h,yaw] (reference rbox.py:50-63)."""
    return _native.rows_op("bevk_xy82xywhr", xy8, 8, (5,), _mode(mode), H=None, has_H=True)


def rbox_world_img(rbox_world, H_img_world):
    """Box centres through a full homography (reference rbox.py:221-226)."""
    return pts_world_bev(rbox_world[:, :2], H_img_world)


# ---- fused chains (BASELINE configs[2]) ------------------------------------------------------------

def xywhr_to_img_corners(xywhr, H, mode):
    """xywhr2xyxy followed by perspective projection of the 4 corners, in one pass
    (the chain of reference bev/visualizer/rbox_vis.py:38-55)."""
    return _native.rows_op("bevk_xywhr2xyxy", xywhr, 5, (8,), _mode(mode), H=H, has_H=True)


def img_corners_to_xywhr(xy8, H, mode):
    """Project 4 corners with H, then xy82xywhr, in one pass (the way back of configs[2])."""
    return _native.rows_op("bevk_xy82xywhr", xy8, 8, (5,), _mode(mode), H=H, has_H=True)


# ---- 7-dof boxes: ground box + projected height tail (reference bev/rbox.py:228-314, numpy only) ----

def rboxtt_world_bev(rbox_src, H, src):
    """(N,7) [x,y,w,h,yaw,du,dv] between world and BEV through an affine similarity H
    (reference rbox.py:258-288; same asserts on H as ``rbox_world_bev``)."""
    return _native.rows_op("bevk_rboxtt_world_bev", rbox_src, 7, (7,), _mode(src), H=H, has_H=True)


def rbox_zt2tt_world(rboxzt, K, Rt):
    """(N,7) world boxes [x,y,w,h,yaw,z,t] -> [x',y',w,h,yaw,du,dv]: foot and top of the box seen
    by the camera K [R|t], dropped back onto the ground plane (reference rbox.py:228-256)."""
    return _native.rbox_zt2tt_world(rboxzt, K, Rt)


def rboxzt_world_bev(rbox_src, H, K, Rt, src):
    """World boxes with height -> BEV boxes with a height tail (reference rbox.py:291-314; like
    the reference only ``src == "world"`` is provided)."""
    assert src in _MODES
    if src != "world":
        raise NotImplementedError("rboxzt_world_bev only supports converting from world to bev")
    return rboxtt_world_bev(rbox_zt2tt_world(rbox_src, K, Rt), H, src)


# ---- rotated-box IoU: the tracker's association matrix -------------------------
def box2d_iou(boxes1, boxes2, method="rbox"):
    """Drop-in for the one d3d call of the reference, ``d3d.box.box2d_iou(boxes1, boxes2,
    method="rbox")`` (bev/tracker/rbox_tracker.py:92): (N, >=5) x (M, >=5) boxes [x, y, w, h, r]
    -> (N, M) IoU matrix, w along (cos r, sin r).  CUDA float32 / float64 tensors."""
    assert method == "rbox", "only the rotated-box IoU the reference uses is implemented"
    return _native.rbox_iou_matrix(boxes1, boxes2, 0.0)


def iou_batch_rbox(bb_test, bb_gt):
    """rbox_tracker.py:87-92: IoU matrix of detections (N, >=5) against trackers (M, >=5), both
    [x, y, w, h, r, ...], yaw shifted by pi/2 exactly as the reference does before calling d3d."""
    return _native.rbox_iou_matrix(bb_test, bb_gt, math.pi / 2)
