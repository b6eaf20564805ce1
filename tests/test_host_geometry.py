"""CPU: bev_b200 host geometry (homo / calib / bev) against known answers produced by the
reference itself (tests/golden/homo_kat.json, written by oracle/gen_golden.py)."""
import numpy as np
import pytest

from bev_b200 import BEVWorldSpec, Calib, FrozenClass, homo
from tests import util

KAT = util.load_json("homo_kat.json")
A = np.array


def close(a, b, tol=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.allclose(a, b, rtol=tol, atol=tol * max(1.0, np.abs(b).max()))


def test_homo_from_vps_and_back():
    k = KAT["homo_from_vps"]
    H = homo.homo_from_vps(A(k["vp1"]), A(k["vp2"]), k["height"], k["u_size"], k["v_size"])
    assert close(H, k["H_img_world"], 1e-12)
    pp = A([(1920 - 1) * 0.5, (1080 - 1) * 0.5])
    assert abs(homo.get_focal(A(k["vp1"]), A(k["vp2"]), pp) - k["focal"]) < 1e-9
    v1, v2 = homo.get_vps_from_homo(H)
    assert close(v1, k["vps_back"][0]) and close(v2, k["vps_back"][1])
    # SURVEY App. B literal
    assert np.allclose(H[0], [864.7402170393, -717.3571673556, 6943.074257693], rtol=1e-9)
    K, focal, R, t = homo.get_KRt_from_homo(H, pp)
    g = KAT["get_KRt_from_homo"]
    assert close(K, g["K"]) and close(R, g["R"], 1e-8) and close(t, g["t"], 1e-8)


def test_homo_from_KRt():
    k = KAT["homo_from_KRt"]
    K, Rt = A(k["K"]), A(k["Rt"])
    assert close(homo.homo_from_KRt(K, Rt_homo=Rt), k["H_Rt_homo"], 1e-13)
    assert close(homo.homo_from_KRt(K, R=Rt[:3, :3], t=Rt[:3, 3]), k["H_R_t"], 1e-13)
    K34 = np.concatenate([K, np.zeros((3, 1))], 1)
    assert close(homo.homo_from_KRt(K34, Rt_homo=Rt), k["H_Rt_homo"], 1e-13)
    with pytest.raises(AssertionError):
        homo.homo_from_KRt(K)


def test_h_canon_and_numpy_dlt():
    src = A([[700, 420], [1220, 420], [1900, 1060], [20, 1060]], dtype=float)
    dst = A([[200, 0], [824, 0], [824, 1024], [200, 1024]], dtype=float)
    assert close(homo.homo_from_pts(src, dst), KAT["h_canon"], 1e-10)
    assert close(homo.homo_from_pts_numpy(src, dst), KAT["h_canon"], 1e-9)
    assert close(homo.homo_from_pts_numpy(src * 2, dst * 2), KAT["h_canon_4k"], 1e-9)
    with pytest.raises(AssertionError):
        homo.homo_from_pts(src[:, :1], dst)


def test_numpy_dlt_overdetermined_matches_reference():
    # lturn preset: 16 noisy correspondences (least squares + reprojection refinement in cv2)
    p = KAT["presets"]["lturn_None"]
    H = homo.homo_from_pts_numpy(A(p["pts_image"]), A(p["pts_world"])[:, :2])
    ref = A(p["H_world_img"])
    assert close(H / H[2, 2], ref / ref[2, 2], 2e-5)


def test_calib_from_vps():
    k = KAT["homo_from_vps"]
    c = Calib(vp1=A(k["vp1"]), vp2=A(k["vp2"]), height=10, u_size=1920, v_size=1080)
    assert c.mode == "from_vps"
    assert close(c.gen_H_world_img(), KAT["calib_vps"]["H_world_img"], 1e-10)
    assert close(c.gen_center_in_world(), KAT["calib_vps"]["center_in_world"], 1e-10)
    assert np.allclose(c.gen_center_in_world(), [10.75968539, 2.6777536, 1], atol=1e-7)  # App. B
    for tag, c2 in (("scale_f", c.scale(align_corners=False, new_u=852, new_v=480)),
                    ("scale_t", c.scale(align_corners=True, new_u=852, new_v=480)),
                    ("pad", c.pad(10, 20, 30, 40)), ("flip", c.flip(lr=True, tb=True))):
        assert close(c2.gen_H_world_img(), KAT["calib_vps_" + tag]["H_world_img"], 1e-9), tag
    c3 = c.scale(align_corners=False, scale_ratio_u=852 / 1920, scale_ratio_v=480 / 1080)
    assert close(c3.gen_H_world_img(), KAT["calib_vps_scale_f"]["H_world_img"], 1e-9)


@pytest.mark.parametrize("name", ["KoPER_1", "KoPER_4", "lturn_None", "roundabout_None"])
def test_presets(name):
    p = KAT["presets"][name]
    if p["mode"] == "from_KRt":
        K = A(p["K"])
        c = Calib(fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], T=A(p["T"]), u_size=p["u_size"],
                  v_size=p["v_size"])
        variants = (("scale_f", c.scale(False, new_u=328, new_v=247)), ("pad", c.pad(3, 5, 7, 9)),
                    ("flip", c.flip(lr=True)))
        tol = 1e-5  # K is float32 in the reference (calib.py:74)
    else:
        c = Calib(pts_world=A(p["pts_world"], dtype=np.float32),
                  pts_image=A(p["pts_image"], dtype=np.float32), u_size=p["u_size"],
                  v_size=p["v_size"])
        variants = (("scale_f", c.scale(False, new_u=426, new_v=240)), ("pad", c.pad(3, 5, 7, 9)),
                    ("flip", c.flip(tb=True)))
        tol = 1e-7
    assert c.mode == p["mode"]
    Hwi = c.gen_H_world_img()
    assert close(Hwi, p["H_world_img"], tol)
    for tag, c2 in variants:
        assert close(c2.gen_H_world_img(), p["H_world_img_" + tag], tol * 10), tag

    b = BEVWorldSpec(**{k: v for k, v in p["bspec"].items() if v is not None and k not in ("x_max", "y_max")})
    Hwb = b.gen_H_world_bev()
    assert close(Hwb, p["H_world_bev"], 1e-9)
    assert close(np.linalg.inv(Hwb).dot(Hwi), p["H_bev_img"], max(tol, 1e-8) * 10)
    for tag, b2 in (("scale_f", b.scale(False, new_u=b.u_size // 2, new_v=b.v_size // 2)),
                    ("scale_t", b.scale(True, new_u=b.u_size // 2, new_v=b.v_size // 2)),
                    ("pad", b.pad(4, 8, 12, 16)), ("flip", b.flip(lr=True, tb=True))):
        assert close(b2.gen_H_world_bev(), p["H_world_bev_" + tag], 1e-9), tag
        for k, v in p["bspec_" + tag].items():
            got = getattr(b2, k)
            assert (got == v) or (got is not None and abs(got - v) < 1e-9), (tag, k)


def test_survey_appendix_b_koper1():
    p = KAT["presets"]["KoPER_1"]
    assert np.allclose(A(p["H_world_bev"]), [[-0.110294117647, 0, 45], [0, 0.110294117377, -30],
                                             [0, 0, 1]], atol=1e-9)
    assert np.allclose(A(p["rbox_bev"])[0],
                       [362.666666666667, 244.800000598614, 16.32, 40.8, -1.270796326105], atol=1e-6)


def test_axes_conventions():
    for key, ent in KAT["axes"].items():
        ua, va = key.split(",")
        b = BEVWorldSpec(u_size=320, v_size=200, u_axis=ua, v_axis=va, x_min=-3.0, x_size=40.0,
                         y_min=2.0, y_size=25.0)
        assert close(b.gen_bev_corners_in_world(), ent["corners"], 1e-13), key
        assert close(b.gen_H_world_bev(), ent["H_world_bev"], 1e-9), key
    with pytest.raises(AssertionError):
        BEVWorldSpec(u_size=4, v_size=4, u_axis="x", v_axis="-x", x_min=0, x_size=1, y_min=0, y_size=1)


def test_bspec_interval_logic():
    b = BEVWorldSpec(u_size=10, v_size=10, x_min=1.0, x_max=5.0, y_max=2.0, y_size=4.0)
    assert b.x_size == 4.0 and b.y_min == -2.0
    with pytest.raises(AssertionError):
        BEVWorldSpec(u_size=10, v_size=10, x_min=1.0, y_min=0.0, y_size=1.0)  # two of three missing
    with pytest.raises(AssertionError):
        BEVWorldSpec(u_size=10, v_size=10, x_min=0.0, x_max=1.0, x_size=3.0, y_min=0.0, y_size=1.0)
    with pytest.raises(TypeError):
        BEVWorldSpec(x_size=10)  # the reference's own __main__ smoke block fails the same way
    b.set_keep(x_max=None, x_size=10.0)
    assert b.x_max == 11.0


def test_frozen_class():
    class P(FrozenClass):
        def __init__(self):
            self.a = 1
            self._freeze()
    p = P()
    p.a = 2
    with pytest.raises(TypeError):
        p.b = 3
    b = BEVWorldSpec(u_size=4, v_size=4, x_min=0, x_size=1, y_min=0, y_size=1)
    with pytest.raises(TypeError):
        b.not_an_attribute = 1


def test_calib_modes_and_rt():
    p = KAT["presets"]["KoPER_1"]
    K, T = A(p["K"]), A(p["T"])
    c = Calib(K=K, T=T, u_size=656, v_size=494)
    assert c.mode == "from_KRt" and np.allclose(c.R, T[:3, :3])
    c2 = Calib(K=K, R=T[:3, :3], t=T[:3, 3], u_size=656, v_size=494)  # reference raises here (App. C)
    assert close(c2.gen_H_world_img(), c.gen_H_world_img(), 1e-6)
    with pytest.raises(AssertionError):
        c.gen_H_world_img(mode="bogus")


def test_brno_bspec_and_compose():
    k = KAT["homo_from_vps"]
    c = Calib(vp1=A(k["vp1"]), vp2=A(k["vp2"]), height=10, u_size=1920, v_size=1080)
    e = KAT["brno_5_1"]
    spec = {kk: v for kk, v in e["bspec"].items() if v is not None and kk not in ("x_max", "y_max")}
    b = BEVWorldSpec(**spec)
    assert (b.u_size, b.v_size) == (288, 448)
    assert close(homo.compose_H_bev_img(c, b), e["H_bev_img"], 1e-9)
    assert e["bspec"] == e["bspec_yaml"]


def test_cfg4_cameras_reproducible():
    for cam in util.load_json("cfg4_cams.json"):
        c = Calib(vp1=A(cam["vp1"]), vp2=A(cam["vp2"]), height=cam["height"], u_size=1920, v_size=1080)
        spec = {kk: v for kk, v in cam["bspec"].items() if v is not None and kk not in ("x_max", "y_max")}
        b = BEVWorldSpec(**spec)
        assert close(homo.compose_H_bev_img(c, b), cam["H_bev_img"], 1e-9), cam["k"]


def test_invert_homography_bit_equal_to_oracle():
    from oracle import warp_oracle as wo
    rng = np.random.default_rng(3)
    for _ in range(100):
        H = rng.normal(size=(3, 3)) * 10
        assert util.bits_equal(homo.invert_homography(H), wo.invert3x3(H))
