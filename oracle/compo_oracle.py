"""oracle/compo_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's compositing (/root/reference/bev/tool/compo.py): the float64
blend of ``composite_reg_img`` (:5-24) and the three-warp BEV composite of ``composite_bev_img``
(:26-50), with the warps served by oracle/warp_oracle.py (pinned to cv2) and cv2's
BGR2GRAY -> GRAY2BGR (``bw_mode``, :13-14) restated as (3735 B + 19235 G + 9798 R + 16384) >> 15 (cv2 4.13, probed).

Pinned against the reference itself (imported from /root/reference in the build container, cv2
4.13) by oracle/gen_golden.py -> tests/golden/compo_kat.npz, checked in tests/test_oracle_compo.py.
Never imported by bev_b200/.
"""
import numpy as np

from oracle import warp_oracle


def gray_bgr(img):
    """cv2.cvtColor(cv2.cvtColor(img, COLOR_BGR2GRAY), COLOR_GRAY2BGR) for uint8 BGR."""
    i = img.astype(np.uint32)
    y = (3735 * i[..., 0] + 19235 * i[..., 1] + 9798 * i[..., 2] + 16384) >> 15
    return np.repeat(y[..., None].astype(np.uint8), 3, axis=-1)


def composite_reg_img(bg, fg, fg_mask, bw_mode=False):
    """compo.py:5-24."""
    if bw_mode:
        fg = gray_bgr(fg)
    bg = bg.astype(np.float64)
    fg = fg.astype(np.float64)
    fg_mask = fg_mask.astype(np.float64) / 255
    compo = fg * fg_mask + bg * (1 - fg_mask)
    compo = compo.round()
    compo[compo > 255] = 255
    return compo.astype(np.uint8)


def homo_from_KRt(K, Rt_homo):
    """bev/homo.py:6-26 for the Rt_homo form."""
    K = np.asarray(K)[:, :3]
    return K.dot(np.asarray(Rt_homo)[:3][:, [0, 1, 3]])


def composite_bev_img(bg, fg, fg_mask, H_world2bev, H_img2world_fix, K, RT, x_size, y_size,
                      bw_mode=False):
    """compo.py:26-50."""
    if bw_mode:
        fg = gray_bgr(fg)
    H_img2bev_fix = H_world2bev.dot(H_img2world_fix)
    bg_bev = warp_oracle.warp_perspective(bg, H_img2bev_fix, (x_size, y_size))
    H_world2img_cam = homo_from_KRt(K, RT)
    H_img2bev_cam = H_world2bev.dot(np.linalg.inv(H_world2img_cam))
    fg_bev = warp_oracle.warp_perspective(fg, H_img2bev_cam, (x_size, y_size))
    mask_bev = warp_oracle.warp_perspective(fg_mask, H_img2bev_cam, (x_size, y_size))
    return composite_reg_img(bg_bev, fg_bev, mask_bev), H_world2img_cam


def blend_integer(bg, fg, fg_mask):
    """The blend of compo.py:16-23 in integers: (fg*k + bg*(255-k) + 127) // 255, the quotient
    taken as ((n + 1) * 0x10101) >> 24 -- the form the CUDA kernels evaluate
    (bev_b200/csrc/compo.cu).  tests/test_oracle_compo.py proves it equal to composite_reg_img's
    float64 expression on all 2^24 (bg, fg, mask) byte triples."""
    k = fg_mask.astype(np.uint64)
    n = fg.astype(np.uint64) * k + bg.astype(np.uint64) * (255 - k) + 127
    return (((n + 1) * 65793) >> 24).astype(np.uint8)
