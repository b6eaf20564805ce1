"""Alpha compositing of foreground renders over a background, in image space or in BEV.

Mirror of /root/reference/bev/tool/compo.py (same function names, argument order and return
values) on CUDA tensors: the three ``cv2.warpPerspective`` calls of ``composite_bev_img``
(compo.py:38,46,47) become ONE batched launch of the warp kernel (three frame sets, two
homographies), and the float64 numpy blend of ``composite_reg_img`` (compo.py:16-23) becomes one
pass of ``bevk_composite_u8c3``.  Results are bit-identical to the reference on the same inputs.

The reference also accepts file names (``cv2.imread``); file IO is outside the hot path, so only
tensors are taken here: uint8 CUDA tensors of shape (H, W, 3) or batches (N, H, W, 3).
"""
import numpy as np

from . import _native
from .homo import homo_from_KRt, INTER_LINEAR


def _check_inputs(*tensors):
    for t in tensors:
        if isinstance(t, str):
            raise TypeError("bev_b200.compo takes CUDA tensors, not file names (image decoding is "
                            "outside the accelerated path)")


def _to_gray_bgr(fg):
    """cv2.cvtColor(cv2.cvtColor(fg, BGR2GRAY), GRAY2BGR) (compo.py:13-14) on the GPU."""
    import torch
    full = torch.full_like(fg, 255)
    return _native.composite_u8c3(fg, fg, full, bw_mode=True)


def composite_reg_img(bg, fg, fg_mask, bw_mode=False):
    """round(fg * mask/255 + bg * (1 - mask/255)) as uint8 (reference compo.py:5-24)."""
    _check_inputs(bg, fg, fg_mask)
    return _native.composite_u8c3(bg, fg, fg_mask, bw_mode=bw_mode)


def composite_bev_img(bg, fg, fg_mask, H_world2bev, H_img2world_fix, K, RT, x_size, y_size,
                      bw_mode=False):
    """Warp background (fixed camera) and foreground + mask (rendering camera K, RT) to the BEV of
    size (x_size, y_size) and blend them (reference compo.py:26-50).
    Returns (composite, H_world2img_cam) like the reference."""
    import torch
    _check_inputs(bg, fg, fg_mask)
    if bw_mode:
        fg = _to_gray_bgr(fg)
    H_world2bev = np.asarray(H_world2bev, np.float64)
    H_img2bev_fix = H_world2bev.dot(np.asarray(H_img2world_fix, np.float64))
    H_world2img_cam = homo_from_KRt(K, Rt_homo=RT)
    H_img2world_cam = np.linalg.inv(H_world2img_cam)
    H_img2bev_cam = H_world2bev.dot(H_img2world_cam)
    single = bg.dim() == 3
    b4, f4, m4 = (t[None] if single else t for t in (bg, fg, fg_mask))
    n = b4.shape[0]
    dsize = (int(x_size), int(y_size))
    if tuple(b4.shape) == tuple(f4.shape) == tuple(m4.shape):
        # one launch: [bg..., fg..., mask...] with two homographies
        frames = torch.cat([b4, f4, m4], 0)
        idx = np.concatenate([np.zeros(n, np.int32), np.ones(2 * n, np.int32)])
        warped = _native.warp_perspective(frames, np.stack([H_img2bev_fix, H_img2bev_cam]), dsize,
                                          flags=INTER_LINEAR, mat_index=idx)
        bg_bev, fg_bev, mask_bev = warped[:n], warped[n:2 * n], warped[2 * n:]
    else:  # background of another size than the renders
        bg_bev = _native.warp_perspective(b4, H_img2bev_fix, dsize, flags=INTER_LINEAR)
        fm = _native.warp_perspective(torch.cat([f4, m4], 0), H_img2bev_cam, dsize, flags=INTER_LINEAR)
        fg_bev, mask_bev = fm[:n], fm[n:]
    compo = _native.composite_u8c3(bg_bev, fg_bev, mask_bev)
    return (compo[0] if single else compo), H_world2img_cam
