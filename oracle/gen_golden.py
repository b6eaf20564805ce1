"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/.

Runs ONLY in the build container: it imports the unmodified reference from /root/reference
(with the two harness shims of SURVEY.md App. D -- ``np.float = float`` and a ``yaml.load``
that defaults to SafeLoader -- no reference file is edited or copied) and the cv2 build the
reference resolves to (opencv-python-headless 4.13.0.92), and stores their outputs as small
fixtures so that the GPU box, which has no /root/reference, can still check against them.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz, *.json

Fixtures written:
  homo_kat.json   host-geometry known answers (homo.py, calib.py, bev.py, presets)   [SURVEY App. B]
  rbox_kat.npz    inputs + reference float64 (bev/rbox.py) and float32 (bev/rbox_torch.py) outputs
  warp_small.npz  small cv2.warpPerspective cases (seeded inputs, stored outputs), all dtypes/flags
  warp_hash.json  sha256 of cv2 outputs for seeded full-size cases (inputs regenerated from seed)
  cfg4_cams.json  the 8 BrnoCompSpeed-shaped cameras of BASELINE configs[3]: H_bev_img + BEV size
  iou_kat.npz     rotated-box pairs with cv2.rotatedRectangleIntersection areas (float32 inside
                  OpenCV) next to the float64 IoU oracle's values, incl. degenerate pairs
  resize_kat.json sha256 of cv2.resize (uint8, INTER_LINEAR) outputs on seeded frames + the
                  small-frame chain of vis_homo.py:73-78,90-91 (Calib.scale -> H_bev_img_small)
"""
import hashlib
import json
import os
import sys

import numpy as np
import yaml

np.float = float  # shim 1 (numpy >= 1.24 removed the alias; bev/bev.py:71 uses it)
_yl = yaml.load
yaml.load = lambda f, Loader=yaml.SafeLoader: _yl(f, Loader=Loader)  # shim 2 (homo_constr.py:227)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
import torch  # noqa: E402
import bev.rbox as ref_rbox  # noqa: E402
import bev.rbox_torch as ref_rbox_torch  # noqa: E402
import bev.homo as ref_homo  # noqa: E402
from bev.calib import Calib  # noqa: E402
from bev.bev import BEVWorldSpec  # noqa: E402
from bev.constructor.homo_constr import preset_calib, preset_bspec, load_bspec  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.synth import seeded_frame, compo_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
OUT = os.path.normpath(OUT)


def L(a):
    return np.asarray(a, dtype=np.float64).tolist()


def h_canon(scale=1):
    """SURVEY.md 8d synthetic homography (1080p -> 1024^2; scale=2 for 4K -> 2048^2)."""
    src = np.array([[700, 420], [1220, 420], [1900, 1060], [20, 1060]], np.float64) * scale
    dst = np.array([[200, 0], [824, 0], [824, 1024], [200, 1024]], np.float64) * scale
    return ref_homo.homo_from_pts(src, dst)


def cfg4_camera(k):
    """SURVEY.md 8d cfg 4 camera k."""
    calib = Calib(vp1=np.array([1100.0 + 40 * k, -320.0 - 15 * k]),
                  vp2=np.array([-4200.0 + 100 * k, 80.0 + 10 * k]),
                  height=9 + 0.5 * k, u_size=1920, v_size=1080)
    ids = [4.1, 4.2, 4.3, 5.1, 5.2, 5.3, 6.1, 6.2]
    bspec = preset_bspec("BrnoCompSpeed", ids[k], calib)
    H_world_img = calib.gen_H_world_img()
    H_world_bev = bspec.gen_H_world_bev()
    H_bev_img = np.linalg.inv(H_world_bev).dot(H_world_img)  # vis_homo.py:61-63
    return calib, bspec, H_bev_img, ids[k]


def bspec_dict(b):
    keys = ["u_size", "v_size", "u_axis", "v_axis", "x_size", "y_size", "x_min", "x_max", "y_min",
            "y_max", "u_min", "u_max", "v_min", "v_max"]
    return {k: (getattr(b, k) if not isinstance(getattr(b, k), (np.floating, np.integer))
                else float(getattr(b, k))) for k in keys}


def gen_homo_kat():
    kat = {}
    kat["h_canon"] = L(h_canon())
    kat["h_canon_4k"] = L(h_canon(2))
    vp1, vp2 = np.array([1200.0, -300.0]), np.array([-4000.0, 100.0])
    H = ref_homo.homo_from_vps(vp1, vp2, 10, 1920, 1080)
    pp = np.array([(1920 - 1) * 0.5, (1080 - 1) * 0.5])
    kat["homo_from_vps"] = {"vp1": L(vp1), "vp2": L(vp2), "height": 10, "u_size": 1920,
                            "v_size": 1080, "H_img_world": L(H),
                            "focal": ref_homo.get_focal(vp1, vp2, pp),
                            "vps_back": [L(v) for v in ref_homo.get_vps_from_homo(H)]}
    K, focal, R, t = ref_homo.get_KRt_from_homo(H, pp)
    kat["get_KRt_from_homo"] = {"K": L(K), "focal": focal, "R": L(R), "t": L(t)}
    Kx = np.array([[900.0, 0, 960], [0, 910.0, 540], [0, 0, 1]])
    Rt = np.eye(4)
    Rt[:3, :3] = cv2.Rodrigues(np.array([1.9, 0.2, -0.1]))[0]
    Rt[:3, 3] = [0.5, 1.0, 12.0]
    kat["homo_from_KRt"] = {"K": L(Kx), "Rt": L(Rt),
                            "H_Rt_homo": L(ref_homo.homo_from_KRt(Kx, Rt_homo=Rt)),
                            "H_R_t": L(ref_homo.homo_from_KRt(Kx, R=Rt[:3, :3], t=Rt[:3, 3]))}
    calib = Calib(vp1=vp1.copy(), vp2=vp2.copy(), height=10, u_size=1920, v_size=1080)
    kat["calib_vps"] = {"H_world_img": L(calib.gen_H_world_img()),
                        "center_in_world": L(calib.gen_center_in_world())}
    for tag, kw in (("scale_f", dict(align_corners=False, new_u=852, new_v=480)),
                    ("scale_t", dict(align_corners=True, new_u=852, new_v=480))):
        c2 = calib.scale(**kw)
        kat["calib_vps_" + tag] = {"H_world_img": L(c2.gen_H_world_img()), "u_size": c2.u_size,
                                   "v_size": c2.v_size}
    c2 = calib.pad(10, 20, 30, 40)
    kat["calib_vps_pad"] = {"H_world_img": L(c2.gen_H_world_img()), "u_size": c2.u_size,
                            "v_size": c2.v_size}
    c2 = calib.flip(lr=True, tb=True)
    kat["calib_vps_flip"] = {"H_world_img": L(c2.gen_H_world_img())}

    presets = {}
    for name, sub in (("KoPER", 1), ("KoPER", 4), ("lturn", None), ("roundabout", None)):
        c = preset_calib(name, sub)
        b = preset_bspec(name, sub, c)
        Hwi = c.gen_H_world_img()
        Hwb = b.gen_H_world_bev()
        Hbi = np.linalg.inv(Hwb).dot(Hwi)
        rb = np.array([[5.0, -3.0, 1.8, 4.5, 0.3]])
        ent = {"mode": c.mode, "u_size": c.u_size, "v_size": c.v_size, "bspec": bspec_dict(b),
               "H_world_img": L(Hwi), "H_world_bev": L(Hwb), "H_bev_img": L(Hbi),
               "rbox_world": L(rb),
               "rbox_bev": L(ref_rbox.rbox_world_bev(rb, np.linalg.inv(Hwb), "world"))}
        if c.mode == "from_KRt":
            ent["K"] = L(c.K)
            ent["T"] = L(c.T)
            for tag, c2 in (("scale_f", c.scale(False, new_u=328, new_v=247)),
                            ("pad", c.pad(3, 5, 7, 9)), ("flip", c.flip(lr=True))):
                ent["H_world_img_" + tag] = L(c2.gen_H_world_img())
        else:
            ent["pts_world"] = L(c.pts_world)
            ent["pts_image"] = L(c.pts_image)
            for tag, c2 in (("scale_f", c.scale(False, new_u=426, new_v=240)),
                            ("pad", c.pad(3, 5, 7, 9)), ("flip", c.flip(tb=True))):
                ent["H_world_img_" + tag] = L(c2.gen_H_world_img())
        for tag, b2 in (("scale_f", b.scale(False, new_u=b.u_size // 2, new_v=b.v_size // 2)),
                        ("scale_t", b.scale(True, new_u=b.u_size // 2, new_v=b.v_size // 2)),
                        ("pad", b.pad(4, 8, 12, 16)), ("flip", b.flip(lr=True, tb=True))):
            ent["H_world_bev_" + tag] = L(b2.gen_H_world_bev())
            ent["bspec_" + tag] = bspec_dict(b2)
        presets["%s_%s" % (name, sub)] = ent
    kat["presets"] = presets

    b51 = preset_bspec("BrnoCompSpeed", 5.1, calib)
    b51y = load_bspec("BrnoCompSpeed", 5.1, calib)
    kat["brno_5_1"] = {"bspec": bspec_dict(b51), "bspec_yaml": bspec_dict(b51y),
                       "H_bev_img": L(np.linalg.inv(b51.gen_H_world_bev()).dot(calib.gen_H_world_img()))}
    # every axis convention of BEVWorldSpec.gen_bev_corners_in_world (bev.py:81-105)
    axes = {}
    for ua, va in (("x", "y"), ("x", "-y"), ("-x", "-y"), ("-x", "y"), ("y", "x"), ("y", "-x"),
                   ("-y", "-x"), ("-y", "x")):
        b = BEVWorldSpec(u_size=320, v_size=200, u_axis=ua, v_axis=va, x_min=-3.0, x_size=40.0,
                         y_min=2.0, y_size=25.0)
        axes["%s,%s" % (ua, va)] = {"corners": L(b.gen_bev_corners_in_world()),
                                   "H_world_bev": L(b.gen_H_world_bev())}
    kat["axes"] = axes
    with open(os.path.join(OUT, "homo_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)


def gen_cfg4():
    cams = []
    for k in range(8):
        calib, bspec, H, sid = cfg4_camera(k)
        cams.append({"k": k, "id": sid, "vp1": L(calib.vp1), "vp2": L(calib.vp2),
                     "height": calib.height, "bspec": bspec_dict(bspec), "H_bev_img": L(H),
                     "H_world_img": L(calib.gen_H_world_img()),
                     "H_world_bev": L(bspec.gen_H_world_bev())})
    with open(os.path.join(OUT, "cfg4_cams.json"), "w") as f:
        json.dump(cams, f, indent=1)


def gen_rbox_kat():
    rng = np.random.default_rng(20261018)
    n = 256
    box = np.stack([rng.uniform(0, 1024, n), rng.uniform(0, 1024, n), rng.uniform(4, 40, n),
                    rng.uniform(8, 120, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    box[0] = [100, 200, 20, 50, 0.3]  # SURVEY App. B vector
    box[1, 4] = 0.0
    box[2, 4] = np.float32(np.pi)
    box[3, 4] = np.float32(-np.pi / 2)
    Hc = h_canon()
    Hc_inv = np.linalg.inv(Hc)
    # a similarity with reflection (KoPER cam 1 H_world_bev) and one with rotation
    c1 = preset_calib("KoPER", 1)
    Hwb = preset_bspec("KoPER", 1, c1).gen_H_world_bev()
    Hwb_r = preset_bspec("roundabout", None, preset_calib("roundabout")).gen_H_world_bev()
    out = {"box": box, "H_canon": Hc, "H_canon_inv": Hc_inv, "H_sim_a": Hwb, "H_sim_b": Hwb_r}
    b64 = box.astype(np.float64)
    bt = torch.from_numpy(box)
    for mode in ("bev", "world"):
        c = ref_rbox.xywhr2xyxy(b64, mode)
        out["xywhr2xyxy_" + mode] = c
        out["xywhr2xyxy_t32_" + mode] = ref_rbox_torch.xywhr2xyxy(bt, mode).numpy()
        c32 = c.astype(np.float32)
        out["xy8_in_" + mode] = c32
        out["xy82xywhr_" + mode] = ref_rbox.xy82xywhr(c32.astype(np.float64), mode)
        out["xywhr2xyvec_" + mode] = ref_rbox.xywhr2xyvec(b64, mode)
        out["xywhr2xyvec_t32_" + mode] = ref_rbox_torch.xywhr2xyvec(bt, mode).numpy()
        out["yaw2v_" + mode] = ref_rbox.yaw2v(b64[:, 4], mode)
        out["yaw2mat_" + mode] = ref_rbox.yaw2mat(b64[:, 4], mode)
        out["v2yaw_" + mode] = ref_rbox.v2yaw(b64[:, :2] - 512.0, mode)
        for tag, Hs in (("a", Hwb), ("b", Hwb_r)):
            Hs_use = Hs if mode == "bev" else np.linalg.inv(Hs)
            out["rbox_world_bev_%s_%s" % (tag, mode)] = ref_rbox.rbox_world_bev(b64, Hs_use, mode)
            out["rbox_world_bev_t32_%s_%s" % (tag, mode)] = ref_rbox_torch.rbox_world_bev(
                bt, torch.from_numpy(Hs_use.astype(np.float32)), mode).numpy()
        # cfg-3 chain: BEV boxes -> image corners (vis_rbox, rbox_vis.py:39-54) and back
        img = cv2.perspectiveTransform(c.reshape(-1, 1, 2), Hc_inv).reshape(-1, 8)
        img_np = ref_rbox.pts_world_bev(c.reshape(-1, 2), Hc_inv).reshape(-1, 8)
        out["img_corners_" + mode] = img_np
        out["img_corners_cv2_" + mode] = img
        img32 = img_np.astype(np.float32)
        out["img_corners_in_" + mode] = img32
        back = ref_rbox.pts_world_bev(img32.astype(np.float64).reshape(-1, 2), Hc).reshape(-1, 8)
        out["back_xywhr_" + mode] = ref_rbox.xy82xywhr(back, mode)
    out["xy82xyvec"] = ref_rbox.xy82xyvec(out["xy8_in_bev"].astype(np.float64))
    pts = rng.uniform(0, 1024, (n, 2)).astype(np.float32)
    out["pts"] = pts
    out["pts_proj"] = ref_rbox.pts_world_bev(pts.astype(np.float64), Hc_inv)
    pts3 = np.concatenate([pts, rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)], 1)
    out["pts3"] = pts3
    out["pts3_proj"] = ref_rbox.pts_world_bev(pts3.astype(np.float64), Hc_inv)
    out["rbox_world_img"] = ref_rbox.rbox_world_img(b64, Hc_inv)
    np.savez_compressed(os.path.join(OUT, "rbox_kat.npz"), **out)


def gen_warp():
    rng = np.random.default_rng(7)
    small = {}
    Hs = {
        "persp": cv2.findHomography(np.array([[5, 3], [57, 6], [62, 44], [2, 40]], np.float64),
                                    np.array([[0, 0], [49, 0], [49, 36], [0, 36]], np.float64))[0],
        "tie2": np.diag([2.0, 2.0, 1.0]),
        "tiehalf": np.array([[0.5, 0, 0.25], [0, 0.5, 0.25], [0, 0, 1.0]]),
        "horizon": np.array([[1.0, 0.2, -3.0], [0.1, 1.1, -2.0], [0.0, 0.03, -0.5]]),
        "shift_out": np.array([[1.0, 0, 40.0], [0, 1.0, -30.0], [0, 0, 1.0]]),
    }
    idx = 0
    for hname, H in Hs.items():
        for dtype, ch in (("uint8", 3), ("uint8", 1), ("uint8", 4), ("float32", 3),
                          ("float32", 1)):
            idx_img = idx
            img = seeded_frame(100 + idx_img, 48, 64, ch, dtype)
            if ch == 1:
                img = img[:, :, 0]
            for flags in (0, 1, 0 | 16, 1 | 16):
                if dtype == "float32" and ch in (1, 4) and (flags & 1) == 0:
                    # float32 C1/C4 + INTER_NEAREST is the one combination this cv2 build hands
                    # to IPP (ippiWarpPerspectiveNearest), whose rounding of exact-tie and border
                    # coordinates differs from cv2's own code path (half-away-like instead of
                    # half-even; independent of cv2.ipp.setUseIPP).  The oracle follows cv2's own
                    # path for every dtype, so these cases are not fixtures ("Oracle").
                    continue
                for bv in (0, (9, 77, 200, 5)):
                    if bv != 0 and (flags & 16 or hname not in ("persp", "shift_out")):
                        continue
                    out = cv2.warpPerspective(img, H, (50, 37), flags=flags, borderValue=bv)
                    key = "c%03d" % idx
                    small[key + "_seed"] = np.array([100 + idx_img, 48, 64, ch])
                    small[key + "_dtype"] = np.array(dtype)
                    small[key + "_H"] = H
                    small[key + "_flags"] = np.array(flags)
                    small[key + "_bv"] = np.array(bv if bv != 0 else (0, 0, 0, 0), np.float64)
                    small[key + "_dst"] = out
                    idx += 1
    small["n"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "warp_small.npz"), **small)

    gen_warp_hashes()


def gen_warp_hashes():
    """sha256 of cv2's full-size outputs on seeded frames (tests/golden/warp_hash.json).  float16
    cases: cv2 has no float16 warp, the contract is float16(cv2(float32(src))), so
    the seeded float16 frame is upcast, warped by cv2 in float32 and rounded once to float16."""
    hashes = []
    Hc, Hc4 = h_canon(), h_canon(2)
    full = [
        ("cfg1_1080p_to_1024_lin", 1234, (1080, 1920, 3), "uint8", Hc, (1024, 1024), 1),
        ("cfg1_1080p_to_1024_nn", 1234, (1080, 1920, 3), "uint8", Hc, (1024, 1024), 0),
        ("cfg1_inv_1024_to_1080p_lin", 1235, (1024, 1024, 3), "uint8", np.linalg.inv(Hc),
         (1920, 1080), 1),
        ("cfg1_inv_wim_1024_to_1080p_lin", 1235, (1024, 1024, 3), "uint8", Hc, (1920, 1080), 17),
        ("cfg5_4k_to_2048_lin", 1236, (2160, 3840, 3), "uint8", Hc4, (2048, 2048), 1),
        ("cfg5_4k_to_2048_nn", 1236, (2160, 3840, 3), "uint8", Hc4, (2048, 2048), 0),
        ("cfg5_inv_2048_to_4k_lin", 1237, (2048, 2048, 3), "uint8", np.linalg.inv(Hc4),
         (3840, 2160), 1),
        ("cfg5_4k_to_2048_lin_f32", 1238, (2160, 3840, 3), "float32", Hc4, (2048, 2048), 1),
        ("c1_1080p_to_1024_lin", 1239, (1080, 1920, 1), "uint8", Hc, (1024, 1024), 1),
        ("c4_1080p_to_1024_lin", 1240, (1080, 1920, 4), "uint8", Hc, (1024, 1024), 1),
        # BASELINE configs[4] in both directions and both dtypes, bilinear and nearest
        ("cfg5_inv_2048_to_4k_nn", 1237, (2048, 2048, 3), "uint8", np.linalg.inv(Hc4), (3840, 2160), 0),
        ("cfg5_4k_to_2048_lin_f16", 1241, (2160, 3840, 3), "float16", Hc4, (2048, 2048), 1),
        ("cfg5_4k_to_2048_nn_f16", 1241, (2160, 3840, 3), "float16", Hc4, (2048, 2048), 0),
        ("cfg5_inv_2048_to_4k_lin_f16", 1242, (2048, 2048, 3), "float16", np.linalg.inv(Hc4),
         (3840, 2160), 1),
        ("cfg5_inv_2048_to_4k_nn_f16", 1242, (2048, 2048, 3), "float16", np.linalg.inv(Hc4),
         (3840, 2160), 0),
    ]
    for k in range(8):
        _, bspec, H, sid = cfg4_camera(k)
        full.append(("cfg4_cam%d_lin" % k, 1234 + k, (1080, 1920, 3), "uint8", H,
                     (int(bspec.u_size), int(bspec.v_size)), 1))
    for name, seed, shape, dtype, H, dsize, flags in full:
        img = seeded_frame(seed, *shape, dtype)
        if shape[2] == 1:
            img = img[:, :, 0]
        if dtype == "float16":
            out = cv2.warpPerspective(img.astype(np.float32), H, dsize, flags=flags).astype(np.float16)
        else:
            out = cv2.warpPerspective(img, H, dsize, flags=flags)
        hashes.append({"name": name, "seed": seed, "shape": list(shape), "dtype": dtype,
                       "H": L(H), "dsize": list(dsize), "flags": flags,
                       "sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest(),
                       "sum": float(out.astype(np.float64).sum())})
    with open(os.path.join(OUT, "warp_hash.json"), "w") as f:
        json.dump({"cv2_version": cv2.__version__, "cases": hashes}, f, indent=1)


def compo_camera():
    """The fixed geometry of the compositing fixture (a KoPER-like camera over a 160x120 BEV)."""
    K = np.array([[300.0, 0, 120], [0, 300.0, 80], [0, 0, 1]])
    RT = np.eye(4)
    RT[:3, :3] = cv2.Rodrigues(np.array([2.2, 0.05, -0.03]))[0]
    RT[:3, 3] = [0.3, 1.5, 14.0]
    H_world2bev = np.array([[8.0, 0, 80], [0, -8.0, 100], [0, 0, 1]])
    Kfix = np.array([[310.0, 0, 118], [0, 305.0, 82], [0, 0, 1]])
    RTfix = np.eye(4)
    RTfix[:3, :3] = cv2.Rodrigues(np.array([2.15, 0.02, 0.04]))[0]
    RTfix[:3, 3] = [-0.2, 1.2, 15.0]
    H_img2world_fix = np.linalg.inv(ref_homo.homo_from_KRt(Kfix, Rt_homo=RTfix))
    return K, RT, H_world2bev, H_img2world_fix


def gen_compo():
    """Outputs of the reference's bev/tool/compo.py (composite_reg_img / composite_bev_img)."""
    import bev.tool.compo as ref_compo
    out = {}
    bg, fg, mask = compo_inputs(4242, 161, 241)  # odd sizes: pixel count not a multiple of 4
    out["reg"] = ref_compo.composite_reg_img(bg, fg, mask)
    out["reg_bw"] = ref_compo.composite_reg_img(bg, fg, mask, bw_mode=True)
    K, RT, H_world2bev, H_img2world_fix = compo_camera()
    bg, fg, mask = compo_inputs(4343, 160, 240)
    for tag, bw in (("bev", False), ("bev_bw", True)):
        c, Hcam = ref_compo.composite_bev_img(bg, fg, mask, H_world2bev, H_img2world_fix, K, RT,
                                              160, 120, bw_mode=bw)
        out[tag] = c
        out[tag + "_Hcam"] = Hcam
    for k, v in (("K", K), ("RT", RT), ("H_world2bev", H_world2bev), ("H_img2world_fix", H_img2world_fix)):
        out[k] = v
    np.savez_compressed(os.path.join(OUT, "compo_kat.npz"), **out)


def gen_rbox7():
    """Outputs of the reference's 7-dof box functions (bev/rbox.py:228-314)."""
    rng = np.random.default_rng(77)
    n = 300
    zt = np.stack([rng.uniform(-20, 20, n), rng.uniform(5, 60, n), rng.uniform(1.5, 2.5, n),
                   rng.uniform(3, 12, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(0, 0.5, n),
                   rng.uniform(1.2, 3.5, n)], 1)
    K, RT, H_world2bev, _ = compo_camera()
    out = {"zt": zt, "K": K, "Rt": RT, "H": H_world2bev}
    out["zt2tt"] = ref_rbox.rbox_zt2tt_world(zt.copy(), K, RT)
    out["tt_bev"] = ref_rbox.rboxtt_world_bev(out["zt2tt"].copy(), H_world2bev, "world")
    out["tt_back"] = ref_rbox.rboxtt_world_bev(out["tt_bev"].copy(), np.linalg.inv(H_world2bev), "bev")
    out["zt_bev"] = ref_rbox.rboxzt_world_bev(zt.copy(), H_world2bev, K, RT, "world")
    np.savez_compressed(os.path.join(OUT, "rbox7_kat.npz"), **out)


RESIZE_CASES = [  # (src_h, src_w, channels, dst_w, dst_h)
    (1080, 1920, 3, 852, 480),   # vis_homo.py:90 with the default --calib-new-u/v
    (1080, 1920, 3, 960, 540),   # exact 2x (cv2 switches to its area path: same bytes)
    (1080, 1920, 3, 1920, 1080), # identity
    (720, 1280, 3, 852, 480),
    (480, 640, 3, 1000, 700),    # upscale
    (100, 100, 1, 200, 200),
    (37, 53, 3, 20, 11),
    (64, 48, 4, 33, 77),
    (9, 7, 3, 31, 29),
    (2, 2, 3, 5, 5),
    (1, 1, 3, 4, 3),
    (5, 1, 1, 3, 7),
    (1, 9, 3, 4, 1),
    (300, 200, 2, 100, 50),
    (17, 1000, 3, 1000, 17),
    (50, 50, 3, 49, 51),
    (211, 173, 3, 97, 131),
    (131, 97, 4, 211, 173),
]


def gen_resize():
    """cv2.resize fixtures; the numpy oracle must reproduce every one before it is stored."""
    from oracle import resize_oracle
    cases = []
    for i, (h, w, c, dw, dh) in enumerate(RESIZE_CASES):
        seed = 9000 + i
        src = seeded_frame(seed, h, w, c, "uint8")
        ref = cv2.resize(src if c > 1 else src[:, :, 0], (dw, dh))
        got = resize_oracle.resize(src if c > 1 else src[:, :, 0], (dw, dh))
        assert np.array_equal(ref, got), ("oracle != cv2", h, w, c, dw, dh)
        cases.append({"seed": seed, "shape": [h, w, c], "dsize": [dw, dh],
                      "sha256": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()})
    # the small-frame chain of vis_homo.py:73-78,90-91 on cfg-4 camera 0
    calib, bspec, _, _ = cfg4_camera(0)
    new_u, new_v = 852, 480
    calib_small = calib.scale(align_corners=False, new_u=new_u, new_v=new_v)
    H_world_bev = bspec.gen_H_world_bev()
    H_small = np.linalg.inv(H_world_bev).dot(calib_small.gen_H_world_img())
    img = seeded_frame(9100, 1080, 1920, 3, "uint8")
    small = cv2.resize(img, (new_u, new_v))
    bev_small = cv2.warpPerspective(small, H_small, (bspec.u_size, bspec.v_size))
    chain = {"seed": 9100, "new_uv": [new_u, new_v], "H_bev_img_small": L(H_small),
             "bev_size": [int(bspec.u_size), int(bspec.v_size)],
             "sha256_small": hashlib.sha256(small.tobytes()).hexdigest(),
             "sha256_bev_small": hashlib.sha256(bev_small.tobytes()).hexdigest()}
    with open(os.path.join(OUT, "resize_kat.json"), "w") as f:
        json.dump({"cv2": cv2.__version__, "cases": cases, "small_frame_chain": chain}, f, indent=1)


def gen_iou():
    """Rotated-box IoU fixtures: the float64 oracle is checked against OpenCV's own rotated
    rectangle intersection (third party, float32) before its values are stored."""
    from oracle import iou_oracle
    rng = np.random.default_rng(77)
    n = 400
    b1 = np.stack([rng.uniform(0, 60, n), rng.uniform(0, 60, n), rng.uniform(1.5, 20, n),
                   rng.uniform(1.5, 20, n), rng.uniform(-np.pi, np.pi, n)], 1)
    b2 = b1.copy()
    b2[:, :2] += rng.normal(0, 5, (n, 2))
    b2[:, 2:4] = rng.uniform(1.5, 20, (n, 2))
    b2[:, 4] = rng.uniform(-np.pi, np.pi, n)
    b2[::10] = b1[::10]                      # identical boxes
    b2[1::10] = b1[1::10]
    b2[1::10, 0] += 1.0                      # shifted copies: coincident edge lines
    b2[2::10, :2] = b1[2::10, :2] + 100      # far apart
    b1 = b1.astype(np.float32).astype(np.float64)   # exactly representable in float32
    b2 = b2.astype(np.float32).astype(np.float64)
    inter_cv = np.zeros(n)
    inter_or = np.zeros(n)
    for i in range(n):
        r1 = ((b1[i, 0], b1[i, 1]), (b1[i, 2], b1[i, 3]), float(np.degrees(b1[i, 4])))
        r2 = ((b2[i, 0], b2[i, 1]), (b2[i, 2], b2[i, 3]), float(np.degrees(b2[i, 4])))
        ret, pts = cv2.rotatedRectangleIntersection(r1, r2)
        if ret != 0 and pts is not None and len(pts) >= 3:
            inter_cv[i] = cv2.contourArea(cv2.convexHull(pts))
        inter_or[i] = iou_oracle.intersection_area(b1[i], b2[i])
    scale = np.maximum(b1[:, 2] * b1[:, 3], b2[:, 2] * b2[:, 3])
    assert np.all(np.abs(inter_cv - inter_or) <= 5e-4 * scale), float(np.max(np.abs(inter_cv - inter_or) / scale))
    iou_pairs = np.array([iou_oracle.box2d_iou(b1[i:i + 1], b2[i:i + 1])[0, 0] for i in range(n)])
    dets = np.concatenate([b1[:24], rng.uniform(0, 1, (24, 1))], 1)   # (24, 6): with a score column
    trks = b2[:17]
    np.savez_compressed(os.path.join(OUT, "iou_kat.npz"), b1=b1, b2=b2, inter_cv=inter_cv,
                        iou_pairs=iou_pairs, dets=dets, trks=trks,
                        iou_matrix=iou_oracle.box2d_iou(dets, trks),
                        iou_tracker=iou_oracle.iou_batch_rbox(dets, trks))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "warp_hash":
        gen_warp_hashes()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "compo":
        gen_compo()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rbox7":
        gen_rbox7()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "iou":
        gen_iou()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "resize":
        gen_resize()
        sys.exit(0)
    gen_homo_kat()
    gen_cfg4()
    gen_rbox_kat()
    gen_warp()
    gen_compo()
    gen_rbox7()
    gen_resize()
    gen_iou()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
