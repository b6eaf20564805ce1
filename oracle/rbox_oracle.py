"""oracle/rbox_oracle.py -- TEST INFRASTRUCTURE ONLY.

float64 numpy restatement of the reference's rotated-box / point projection math
(/root/reference/bev/rbox.py).  SURVEY.md 8c makes the *float64* numpy path, evaluated on the
float32 inputs upcast, the oracle for the CUDA projection kernels (the reference's own float32
torch twin is up to 2.3e-4 relative off after a perspective H).

Pinned against the reference itself (imported from /root/reference in the build container) by
oracle/gen_golden.py -> tests/golden/rbox_kat.npz, checked in tests/test_oracle_rbox.py.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  bev_b200/ never does.
"""
import numpy as np

_MODES = ("bev", "world")


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def v2yaw(v, mode):
    """rbox.py:20-27 -- bev: atan2(u, v) (yaw 0 = +v axis); world: atan2(y, x)."""
    assert mode in _MODES
    v = _f64(v)
    return np.arctan2(v[:, 0], v[:, 1]) if mode == "bev" else np.arctan2(v[:, 1], v[:, 0])


def yaw2v(yaw, mode):
    """rbox.py:29-36 -- unit heading vector of a yaw angle."""
    assert mode in _MODES
    yaw = _f64(yaw)
    s, c = np.sin(yaw), np.cos(yaw)
    return np.stack((s, c), axis=1) if mode == "bev" else np.stack((c, s), axis=1)


def yaw2mat(yaw, mode):
    """rbox.py:38-48 -- per-box 2x2 rotation, sign layout differs per mode."""
    assert mode in _MODES
    yaw = _f64(yaw).reshape(-1)
    s, c = np.sin(yaw), np.cos(yaw)
    m = np.empty((yaw.shape[0], 2, 2))
    m[:, 0, 0] = c
    m[:, 1, 1] = c
    if mode == "bev":
        m[:, 0, 1] = s
        m[:, 1, 0] = -s
    else:
        m[:, 0, 1] = -s
        m[:, 1, 0] = s
    return m


def xywhr2xyxy(box, mode):
    """rbox.py:65-112 (external_aa=False; the True branch is dead in the reference).

    Corner order tl, bl, br, tr of the un-rotated template; bev: w along u, h along v;
    world: h along x, w along y.
    """
    assert mode in _MODES
    box = _f64(box)
    hw, hh = box[:, 2] / 2, box[:, 3] / 2
    if mode == "bev":
        tx = np.stack((-hw, -hw, hw, hw), axis=1)
        ty = np.stack((-hh, hh, hh, -hh), axis=1)
    else:
        tx = np.stack((-hh, hh, hh, -hh), axis=1)
        ty = np.stack((-hw, -hw, hw, hw), axis=1)
    R = yaw2mat(box[:, 4], mode)
    cx = R[:, 0, 0, None] * tx + R[:, 0, 1, None] * ty + box[:, 0, None]
    cy = R[:, 1, 0, None] * tx + R[:, 1, 1, None] * ty + box[:, 1, None]
    out = np.empty((box.shape[0], 8))
    out[:, 0::2] = cx
    out[:, 1::2] = cy
    return out


def xy82xywhr(xy8, mode):
    """rbox.py:50-63 -- w = |tr-tl|, h = |bl-tl|, centre = (bl+tr)/2, yaw = v2yaw(tl-bl)."""
    assert mode in _MODES
    xy8 = _f64(xy8)
    tl, bl, tr = xy8[:, 0:2], xy8[:, 2:4], xy8[:, 6:8]
    w = np.sqrt(((tr - tl) ** 2).sum(1))
    h = np.sqrt(((bl - tl) ** 2).sum(1))
    c = 0.5 * (bl + tr)
    r = v2yaw(tl - bl, mode)
    return np.stack((c[:, 0], c[:, 1], w, h, r), axis=1)


def xywhr2xyvec(box, mode):
    """rbox.py:114-125 -- heading segment [x, y, x + h*dx, y + h*dy]."""
    assert mode in _MODES
    box = _f64(box)
    d = yaw2v(box[:, 4], mode) * box[:, 3:4]
    return np.stack((box[:, 0], box[:, 1], box[:, 0] + d[:, 0], box[:, 1] + d[:, 1]), axis=1)


def xy82xyvec(xy8):
    """rbox.py:127-134 -- start = (tl+br)/2, direction = bl - tl."""
    xy8 = _f64(xy8)
    d = xy8[:, 2:4] - xy8[:, 0:2]
    cx = 0.5 * (xy8[:, 0] + xy8[:, 4])
    cy = 0.5 * (xy8[:, 1] + xy8[:, 5])
    return np.stack((cx, cy, cx + d[:, 0], cy + d[:, 1]), axis=1)


def pts_world_bev(pts, H):
    """rbox.py:136-151 -- homogeneous projection with divide; (N,2)->(N,2), (N,3)->(N,3)."""
    pts = _f64(pts)
    if pts.ndim == 1:
        pts = pts[None]
    H = _f64(H)
    homo = pts.shape[1] == 3
    if not homo:
        assert pts.shape[1] == 2
        pts = np.concatenate((pts, np.ones((pts.shape[0], 1))), axis=1)
    q = pts @ H.T
    q = q / q[:, 2:3]
    return q if homo else q[:, :2]


def dist_world_bev(d, H):
    """rbox.py:153-160 -- isotropic scale of a similarity (column norms, asserted equal)."""
    H = _f64(H)
    s0 = np.sqrt(H[0, 0] ** 2 + H[1, 0] ** 2)
    s1 = np.sqrt(H[0, 1] ** 2 + H[1, 1] ** 2)
    assert abs(s0 - s1) < 1e-5
    return s0 * _f64(d)


def angle_world_bev(yaw, H, src):
    """rbox.py:162-171 -- yaw -> unit vector -> linear part of H -> yaw in the other frame."""
    assert src in _MODES
    tgt = "world" if src == "bev" else "bev"
    v = yaw2v(_f64(yaw).reshape(-1), src)
    H = _f64(H)
    t = v @ H[:2, :2].T
    return v2yaw(t, tgt)


def rbox_world_bev(box, H, src):
    """rbox.py:173-219 -- similarity transform of (N,5) xywh-yaw boxes."""
    assert src in _MODES
    H = _f64(H)
    H = H / H[2, 2]
    assert abs(H[2, 0]) + abs(H[2, 1]) < 1e-5
    box = _f64(box)
    if len(box) == 0:
        return box
    r = angle_world_bev(box[:, 4], H, src)
    xy = pts_world_bev(box[:, :2], H)
    wh = dist_world_bev(box[:, 2:4], H)
    return np.concatenate((xy, wh, r[:, None]), axis=1)


def rbox_world_img(box, H_img_world):
    """rbox.py:221-226 -- box centres through a full homography."""
    return pts_world_bev(_f64(box)[:, :2], H_img_world)


# --- composite chains of BASELINE.json configs[2] (rbox_vis.py:38-55 and its inverse) ----------

def xywhr_to_img_corners(box, H, mode):
    """xywhr2xyxy -> perspective projection of the 4 corners (vis_rbox chain, rbox_vis.py:39-54)."""
    c = xywhr2xyxy(box, mode).reshape(-1, 2)
    return pts_world_bev(c, H).reshape(-1, 8)


def img_corners_to_xywhr(xy8, H, mode):
    """The way back: project 4 corners with H, then xy82xywhr."""
    c = pts_world_bev(_f64(xy8).reshape(-1, 2), H).reshape(-1, 8)
    return xy82xywhr(c, mode)


# --- 7-dof boxes (ground box + height tail), rbox.py:228-314 -------------------------------------

def homo_from_KRt(K, Rt):
    """bev/homo.py:6-26 (Rt_homo form): K [r1 r2 t]."""
    return _f64(K)[:, :3].dot(_f64(Rt)[:3][:, [0, 1, 3]])


def rbox_zt2tt_world(rboxzt, K, Rt):
    """rbox.py:228-256."""
    rboxzt, K, Rt = _f64(rboxzt), _f64(K), _f64(Rt)
    H_world_cam = np.linalg.inv(homo_from_KRt(K, Rt))

    def ground(xyz):
        cam = Rt[:3, :3].dot(xyz) + Rt[:3, [3]]
        uvd = K.dot(cam)
        uv1 = uvd / np.clip(uvd[2], a_min=1e-2, a_max=None)
        xy1 = H_world_cam.dot(uv1)
        return xy1 / xy1[2]

    low = rboxzt[:, [0, 1, 5]].T
    high = low.copy()
    high[2] = high[2] + rboxzt[:, 6]
    xy_low, xy_high = ground(low), ground(high)
    return np.concatenate([xy_low[:2].T, rboxzt[:, 2:5], (xy_high[:2] - xy_low[:2]).T], axis=1)


def rboxtt_world_bev(box, H, src):
    """rbox.py:258-288."""
    assert src in _MODES
    box, H = _f64(box), _f64(H)
    if len(box) == 0:
        return box
    H = H / H[2, 2]
    assert abs(H[2, 0]) + abs(H[2, 1]) < 1e-5
    assert box.shape[1] == 7
    one = np.ones((box.shape[0], 1))
    start = H.dot(np.concatenate([box[:, :2], one], axis=1).T)
    end = H.dot(np.concatenate([box[:, :2] + box[:, 5:], one], axis=1).T)
    dudv = (end - start)[:2].T
    return np.concatenate([rbox_world_bev(box[:, :5], H, src), dudv], axis=1)


def rboxzt_world_bev(box, H, K, Rt, src):
    """rbox.py:291-314 (world -> bev only, like the reference)."""
    assert src in _MODES
    if src != "world":
        raise NotImplementedError("rboxzt_world_bev only supports converting from world to bev")
    return rboxtt_world_bev(rbox_zt2tt_world(box, K, Rt), H, src)
