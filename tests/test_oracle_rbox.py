"""CPU: pins oracle/rbox_oracle.py (float64 restatement of bev/rbox.py) against outputs of the
reference itself, stored in tests/golden/rbox_kat.npz by oracle/gen_golden.py."""
import numpy as np
import pytest

from oracle import rbox_oracle as ro
from tests import util

K = util.load_npz("rbox_kat.npz")
BOX = K["box"].astype(np.float64)
TOL = 1e-12


def close(a, b, tol=TOL):
    return np.allclose(a, b, rtol=tol, atol=tol)


@pytest.mark.parametrize("mode", ["bev", "world"])
def test_box_functions(mode):
    assert close(ro.xywhr2xyxy(BOX, mode), K["xywhr2xyxy_" + mode])
    assert close(ro.xy82xywhr(K["xy8_in_" + mode], mode), K["xy82xywhr_" + mode])
    assert close(ro.xywhr2xyvec(BOX, mode), K["xywhr2xyvec_" + mode])
    assert close(ro.yaw2v(BOX[:, 4], mode), K["yaw2v_" + mode])
    assert close(ro.yaw2mat(BOX[:, 4], mode), K["yaw2mat_" + mode])
    assert close(ro.v2yaw(BOX[:, :2] - 512.0, mode), K["v2yaw_" + mode])


@pytest.mark.parametrize("mode", ["bev", "world"])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_rbox_world_bev(mode, tag):
    Hs = K["H_sim_" + tag]
    H = Hs if mode == "bev" else np.linalg.inv(Hs)
    assert close(ro.rbox_world_bev(BOX, H, mode), K["rbox_world_bev_%s_%s" % (tag, mode)], 1e-11)


def test_rbox_world_bev_asserts():
    with pytest.raises(AssertionError):
        ro.rbox_world_bev(BOX, K["H_canon"], "bev")  # perspective H is not affine
    with pytest.raises(AssertionError):
        ro.rbox_world_bev(BOX, np.diag([1.0, 2.0, 1.0]), "bev")  # anisotropic scale
    with pytest.raises(AssertionError):
        ro.xywhr2xyxy(BOX, "image")


@pytest.mark.parametrize("mode", ["bev", "world"])
def test_projection_chain(mode):
    Hc, Hi = K["H_canon"], K["H_canon_inv"]
    img = ro.xywhr_to_img_corners(BOX, Hi, mode)
    assert close(img, K["img_corners_" + mode], 1e-11)
    assert np.allclose(img, K["img_corners_cv2_" + mode], rtol=1e-9, atol=1e-9)  # cv2.perspectiveTransform
    back = ro.img_corners_to_xywhr(K["img_corners_in_" + mode], Hc, mode)
    assert close(back, K["back_xywhr_" + mode], 1e-10)
    # the round trip returns the box with yaw shifted by exactly pi
    assert util.rel_err(back[:, :4], BOX[:, :4]) < 1e-4
    assert util.yaw_err(back[:, 4], BOX[:, 4] + np.pi) < 1e-4


def test_points():
    assert close(ro.pts_world_bev(K["pts"], K["H_canon_inv"]), K["pts_proj"], 1e-11)
    assert close(ro.pts_world_bev(K["pts3"], K["H_canon_inv"]), K["pts3_proj"], 1e-11)
    assert close(ro.rbox_world_img(BOX, K["H_canon_inv"]), K["rbox_world_img"], 1e-11)
    assert close(ro.xy82xyvec(K["xy8_in_bev"]), K["xy82xyvec"])


def test_survey_appendix_b_vector():
    # SURVEY.md App. B first bullet
    c = ro.xywhr2xyxy(np.array([[100, 200, 20, 50, 0.3]]), "bev")[0]
    assert np.allclose(c, [83.05862994221, 179.071789838473, 97.834640275277, 226.838614294754,
                           116.94137005779, 220.928210161527, 102.165359724723, 173.161385705246],
                       atol=1e-9)
    r = ro.xy82xywhr(c[None], "bev")[0]
    assert np.allclose(r, [100, 200, 20, 50, 0.3 - np.pi], atol=1e-9)


def test_empty():
    e = np.zeros((0, 5))
    assert ro.rbox_world_bev(e, K["H_sim_a"], "bev").shape == (0, 5)
    assert ro.xywhr2xyxy(e, "bev").shape == (0, 8)


def test_reference_fp32_twin_is_looser_than_oracle():
    # documents why the float64 path is the oracle: the reference's own torch-fp32 twin is only
    # ~1e-6 .. 1e-5 relative from it on plain corners
    e = util.rel_err(K["xywhr2xyxy_t32_bev"], K["xywhr2xyxy_bev"])
    assert 0 < e < 1e-4


def test_seven_dof_boxes_match_the_reference():
    """oracle rbox_zt2tt_world / rboxtt_world_bev / rboxzt_world_bev vs the reference's own
    outputs (bev/rbox.py:228-314 run in the build container, tests/golden/rbox7_kat.npz)."""
    K7 = util.load_npz("rbox7_kat.npz")
    assert np.abs(ro.rbox_zt2tt_world(K7["zt"], K7["K"], K7["Rt"]) - K7["zt2tt"]).max() < 1e-12
    assert np.abs(ro.rboxtt_world_bev(K7["zt2tt"], K7["H"], "world") - K7["tt_bev"]).max() < 1e-12
    back = ro.rboxtt_world_bev(K7["tt_bev"], np.linalg.inv(K7["H"]), "bev")
    assert np.abs(back - K7["tt_back"]).max() < 1e-12
    assert np.abs(ro.rboxzt_world_bev(K7["zt"], K7["H"], K7["K"], K7["Rt"], "world") - K7["zt_bev"]).max() < 1e-12
    # the BEV -> world -> BEV round trip of the tail is the identity
    assert np.abs(back[:, [0, 1, 2, 3, 5, 6]] - K7["zt2tt"][:, [0, 1, 2, 3, 5, 6]]).max() < 1e-9
