"""GPU parity: bev_b200.rbox_torch (CUDA, through the C ABI) vs the float64 oracle
(oracle/rbox_oracle.py, pinned to the reference's bev/rbox.py) on the same float32 inputs.

Tolerance (north_star, made precise in SURVEY.md 8c):
    |out - ref64| <= 1e-5 * max(|ref64|, 1) element-wise; yaw compared modulo 2*pi."""
import numpy as np
import pytest
import torch

from bev_b200 import rbox_torch as rt
from oracle import rbox_oracle as ro
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
K = util.load_npz("rbox_kat.npz")


def cu(a, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(DEV)


def boxes(n, seed=0):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, 1024, n), rng.uniform(0, 1024, n), rng.uniform(4, 40, n),
                     rng.uniform(8, 120, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)


def check_box(out, ref, yaw_col=None):
    out = out.cpu().numpy()
    assert out.shape == ref.shape
    cols = [c for c in range(ref.shape[1]) if c != yaw_col]
    assert util.rel_err(out[:, cols], ref[:, cols]) <= TOL
    if yaw_col is not None:
        assert util.yaw_err(out[:, yaw_col], ref[:, yaw_col]) <= TOL


@pytest.mark.parametrize("mode", ["bev", "world"])
def test_golden_reference_outputs(mode):
    """Against numbers the reference itself produced (tests/golden/rbox_kat.npz)."""
    b = cu(K["box"])
    check_box(rt.xywhr2xyxy(b, mode), K["xywhr2xyxy_" + mode])
    check_box(rt.xy82xywhr(cu(K["xy8_in_" + mode]), mode), K["xy82xywhr_" + mode], yaw_col=4)
    check_box(rt.xywhr2xyvec(b, mode), K["xywhr2xyvec_" + mode])
    check_box(rt.yaw2v(b[:, 4], mode), K["yaw2v_" + mode])
    assert util.rel_err(rt.yaw2mat(b[:, 4], mode).cpu().numpy(), K["yaw2mat_" + mode]) <= TOL
    v = cu(K["box"][:, :2] - 512.0)
    assert util.yaw_err(rt.v2yaw(v, mode).cpu().numpy(), K["v2yaw_" + mode]) <= TOL
    for tag in ("a", "b"):
        Hs = K["H_sim_" + tag]
        H = Hs if mode == "bev" else np.linalg.inv(Hs)
        check_box(rt.rbox_world_bev(b, H, mode), K["rbox_world_bev_%s_%s" % (tag, mode)], yaw_col=4)
    check_box(rt.xywhr_to_img_corners(b, K["H_canon_inv"], mode), K["img_corners_" + mode])
    check_box(rt.img_corners_to_xywhr(cu(K["img_corners_in_" + mode]), K["H_canon"], mode),
              K["back_xywhr_" + mode], yaw_col=4)


def test_points_golden():
    check_box(rt.pts_world_bev(cu(K["pts"]), K["H_canon_inv"]), K["pts_proj"])
    check_box(rt.pts_world_bev(cu(K["pts3"]), K["H_canon_inv"]), K["pts3_proj"])
    check_box(rt.rbox_world_img(cu(K["box"]), K["H_canon_inv"]), K["rbox_world_img"])
    check_box(rt.xy82xyvec(cu(K["xy8_in_bev"])), K["xy82xyvec"])
    one = rt.pts_world_bev(cu(K["pts"][0]), K["H_canon_inv"])  # 1-D point is promoted (rbox.py:138-139)
    assert tuple(one.shape) == (1, 2)


@pytest.mark.parametrize("mode", ["bev", "world"])
@pytest.mark.parametrize("n", [1, 255, 256, 257, 100003])
def test_random_boxes_vs_oracle(mode, n):
    b = boxes(n, seed=n)
    Hc, Hi = util.h_canon(), np.linalg.inv(util.h_canon())
    check_box(rt.xywhr2xyxy(cu(b), mode), ro.xywhr2xyxy(b, mode))
    img = rt.xywhr_to_img_corners(cu(b), Hi, mode)
    check_box(img, ro.xywhr_to_img_corners(b, Hi, mode))
    img32 = img.cpu().numpy()
    check_box(rt.img_corners_to_xywhr(img, Hc, mode), ro.img_corners_to_xywhr(img32, Hc, mode), yaw_col=4)
    check_box(rt.xy82xywhr(img, mode), ro.xy82xywhr(img32, mode), yaw_col=4)


def test_cfg3_full_size_round_trip():
    """BASELINE configs[2] at full size: 10 M boxes -> image corners -> back.  Size-independent
    property: the round trip returns every box with yaw shifted by exactly pi."""
    n = 10_000_000
    g = torch.Generator(device=DEV).manual_seed(0)
    u = torch.rand((n, 5), generator=g, device=DEV)
    b = torch.stack([u[:, 0] * 1024, u[:, 1] * 1024, 4 + u[:, 2] * 36, 8 + u[:, 3] * 112,
                     (u[:, 4] * 2 - 1) * np.pi], 1).contiguous()
    Hc = util.h_canon()
    img = rt.xywhr_to_img_corners(b, np.linalg.inv(Hc), "bev")
    back = rt.img_corners_to_xywhr(img, Hc, "bev")
    assert tuple(img.shape) == (n, 8) and tuple(back.shape) == (n, 5)
    # float32 corner coordinates limit the round trip, not the kernels: the far field maps 1 BEV px
    # to ~0.1 image px, so compare in BEV pixels with a loose absolute bound
    d = (back[:, :4] - b[:, :4]).abs().max().item()
    assert d < 0.05, d
    dy = (back[:, 4] - b[:, 4]) % (2 * np.pi) - np.pi  # = back - (b + pi), wrapped to (-pi, pi]
    assert dy.abs().max().item() < 2e-2
    # and a slice of it against the oracle at the contract tolerance
    sl = slice(5_000_000, 5_000_512)
    check_box(img[sl], ro.xywhr_to_img_corners(b[sl].cpu().numpy(), np.linalg.inv(Hc), "bev"))


def test_float64_tensors():
    b = boxes(1000, 5).astype(np.float64)
    Hi = np.linalg.inv(util.h_canon())
    out = rt.xywhr_to_img_corners(cu(b, np.float64), Hi, "world")
    assert out.dtype == torch.float64
    assert util.rel_err(out.cpu().numpy(), ro.xywhr_to_img_corners(b, Hi, "world")) < 1e-12
    out = rt.rbox_world_bev(cu(b, np.float64), K["H_sim_a"], "bev")
    ref = ro.rbox_world_bev(b, K["H_sim_a"], "bev")
    assert util.rel_err(out.cpu().numpy()[:, :4], ref[:, :4]) < 1e-12
    assert util.yaw_err(out.cpu().numpy()[:, 4], ref[:, 4]) < 1e-12


def test_reference_asserts_and_empty():
    b = cu(boxes(8))
    with pytest.raises(AssertionError):
        rt.rbox_world_bev(b, util.h_canon(), "bev")          # not affine (rbox_torch.py:140)
    with pytest.raises(AssertionError):
        rt.rbox_world_bev(b, np.diag([1.0, 2.0, 1.0]), "bev")  # not a similarity (:161)
    with pytest.raises(AssertionError):
        rt.xywhr2xyxy(b, "img")
    with pytest.raises(NotImplementedError):
        rt.xywhr2xyxy(b, "bev", external_aa=True)
    e = torch.zeros((0, 5), device=DEV)
    assert tuple(rt.xywhr2xyxy(e, "bev").shape) == (0, 8)
    assert tuple(rt.rbox_world_bev(e, K["H_sim_a"], "world").shape) == (0, 5)
    H_t = torch.from_numpy(K["H_sim_a"])  # CPU-tensor H is accepted like a numpy H
    check_box(rt.rbox_world_bev(b, H_t, "bev"), ro.rbox_world_bev(b.cpu().numpy(), K["H_sim_a"], "bev"), yaw_col=4)


def test_large_angles_and_horizon_conditioning():
    b = boxes(4096, 9)
    b[:, 4] = np.random.default_rng(1).uniform(-50, 50, 4096).astype(np.float32)  # many turns
    check_box(rt.xywhr2xyxy(cu(b), "bev"), ro.xywhr2xyxy(b, "bev"))
    # points marching toward the horizon line of H_canon (w -> 0): float64 in-kernel keeps 1e-5
    Hc = util.h_canon()
    y_h = -Hc[2, 2] / Hc[2, 1]  # image row where w = 0
    pts = np.stack([np.linspace(100, 1800, 512), y_h + np.geomspace(0.5, 300, 512)], 1).astype(np.float32)
    check_box(rt.pts_world_bev(cu(pts), Hc), ro.pts_world_bev(pts, Hc))


def test_unaligned_views_and_all_row_widths():
    """A sliced tensor starts 20 bytes into its storage: the 16-byte aligned bulk-copy pipeline
    must hand it to the element-wise staging kernel.  Also runs every row width (1, 2, 3, 4, 5, 8
    words in or out) through blocks + tail."""
    b = boxes(3000, 11)
    whole = cu(b)
    Hi = np.linalg.inv(util.h_canon())
    check_box(rt.xywhr_to_img_corners(whole[1:], Hi, "bev"), ro.xywhr_to_img_corners(b[1:], Hi, "bev"))
    check_box(rt.xywhr2xyvec(whole, "world"), ro.xywhr2xyvec(b, "world"))
    check_box(rt.yaw2v(whole[:, 4].contiguous(), "bev"), ro.yaw2v(b[:, 4], "bev"))
    assert util.rel_err(rt.yaw2mat(whole[:, 4].contiguous(), "world").cpu().numpy(),
                        ro.yaw2mat(b[:, 4], "world")) <= TOL
    v = b[:, :2] - 512.0
    assert util.yaw_err(rt.v2yaw(cu(v), "bev").cpu().numpy(), ro.v2yaw(v, "bev")) <= TOL
    pts3 = np.concatenate([b[:, :2], np.ones((len(b), 1), np.float32)], 1)
    check_box(rt.pts_world_bev(cu(pts3), Hi), ro.pts_world_bev(pts3, Hi))
    check_box(rt.pts_world_bev(cu(b[:, :2]), Hi), ro.pts_world_bev(b[:, :2], Hi))
    xy8 = ro.xywhr2xyxy(b, "bev").astype(np.float32)
    check_box(rt.xy82xyvec(cu(xy8)), ro.xy82xyvec(xy8))


def test_seven_dof_boxes():
    """rbox_zt2tt_world / rboxtt_world_bev / rboxzt_world_bev (reference bev/rbox.py:228-314,
    numpy only there) against the reference's outputs and the float64 oracle."""
    K7 = util.load_npz("rbox7_kat.npz")
    zt = K7["zt"].astype(np.float32)
    ref_tt = ro.rbox_zt2tt_world(zt, K7["K"], K7["Rt"])
    tt = rt.rbox_zt2tt_world(cu(zt), K7["K"], K7["Rt"])
    check_box(tt, ref_tt, yaw_col=4)
    tt32 = tt.cpu().numpy()
    check_box(rt.rboxtt_world_bev(tt, K7["H"], "world"), ro.rboxtt_world_bev(tt32, K7["H"], "world"), yaw_col=4)
    check_box(rt.rboxzt_world_bev(cu(zt), K7["H"], K7["K"], K7["Rt"], "world"),
              ro.rboxzt_world_bev(zt, K7["H"], K7["K"], K7["Rt"], "world"), yaw_col=4)
    # float64 tensors reproduce the reference's own numbers
    out = rt.rboxzt_world_bev(cu(K7["zt"], np.float64), K7["H"], K7["K"], K7["Rt"], "world").cpu().numpy()
    assert util.rel_err(out[:, [0, 1, 2, 3, 5, 6]], K7["zt_bev"][:, [0, 1, 2, 3, 5, 6]]) < 1e-10
    assert util.yaw_err(out[:, 4], K7["zt_bev"][:, 4]) < 1e-10
    with pytest.raises(NotImplementedError):
        rt.rboxzt_world_bev(cu(zt), K7["H"], K7["K"], K7["Rt"], "bev")
    with pytest.raises(AssertionError):
        rt.rboxtt_world_bev(tt, util.h_canon(), "world")  # not affine
    assert tuple(rt.rboxtt_world_bev(torch.zeros((0, 7), device=DEV), K7["H"], "bev").shape) == (0, 7)


@pytest.mark.parametrize("src", ["bev", "world"])
def test_dist_and_angle_world_bev(src):
    """Torch twins of rbox.dist_world_bev / angle_world_bev (bev/rbox.py:153-171)."""
    rng = np.random.default_rng(21)
    H = K["H_sim_a"]
    d = rng.uniform(0.1, 80, (1000, 2)).astype(np.float32)
    out = rt.dist_world_bev(cu(d), H)
    assert tuple(out.shape) == (1000, 2)
    assert util.rel_err(out.cpu().numpy(), ro.dist_world_bev(d, H)) <= TOL
    yaw = rng.uniform(-7, 7, 3001).astype(np.float32)
    got = rt.angle_world_bev(cu(yaw), H, src).cpu().numpy()
    assert got.shape == (3001,)
    assert util.yaw_err(got, ro.angle_world_bev(yaw, H, src)) <= TOL
    # un-normalised H is used as given: a negative H22 keeps its sign, exactly as the numpy twin
    got = rt.angle_world_bev(cu(yaw), -2.5 * H, src).cpu().numpy()
    assert util.yaw_err(got, ro.angle_world_bev(yaw, -2.5 * H, src)) <= TOL
    with pytest.raises(AssertionError):
        rt.dist_world_bev(cu(d), np.diag([1.0, 2.0, 1.0]))
    with pytest.raises(AssertionError):
        rt.angle_world_bev(cu(yaw), H, "img")
    assert tuple(rt.angle_world_bev(torch.zeros(0, device=DEV), H, src).shape) == (0,)
