"""Calib: camera calibration -> image/world homography (host side).

Mirrors /root/reference/bev/calib.py:7-267: the same keyword attributes and three ways to define
the ground-plane homography -- ``from_KRt`` (K + 4x4 T), ``from_vps`` (two vanishing points, camera
height, image size) and ``from_pts`` (image/world point pairs) -- plus scale / pad / flip that
return a new Calib.  Output is always a float64 3x3 ``H_world_img``.
"""
import numpy as np

from .frozen_class import FrozenClass
from .homo import homo_from_KRt, homo_from_pts, homo_from_vps

_MODES = ("from_KRt", "from_pts", "from_vps")


def _T_from_Rt(R, t):
    top = np.concatenate((R, np.reshape(t, (3, 1))), axis=1)
    return np.concatenate((top, np.array([[0, 0, 0, 1]])), axis=0).astype(np.float32)


def _map_coord(c, ratio, align_corners):
    """Pixel coordinate under a resize: plain scaling, or the half-pixel rule (calib.py:165-166)."""
    return c * ratio if align_corners else (c + 0.5) * ratio - 0.5


class Calib(FrozenClass):
    def __init__(self, **kwargs):
        # intrinsics
        self.K = None
        self.fx = self.fy = self.cx = self.cy = 0
        self.dist_coeff = None
        # extrinsics (T is the 4x4 world->camera transform)
        self.R = None
        self.t = None
        self.T = None
        # point correspondences / cached homographies
        self.pts_world = None
        self.pts_image = None
        self.H_world_img = None
        self.H_img_world = None
        # vanishing-point description
        self.vp1 = None
        self.vp2 = None
        self.pp = None
        self.height = None
        self.u_size = None
        self.v_size = None

        self.mode = None
        self._freeze()
        self.__dict__.update(kwargs)

        if self.pts_image is not None and self.pts_world is not None:
            self.mode = "from_pts"
        elif self.vp1 is not None and self.vp2 is not None:
            self.mode = "from_vps"
        else:
            self.mode = "from_KRt"

        self.update()
        self.check_validity()

    def update(self):
        """Complete K / R / t / T from whichever of them were given (reference calib.py:64-90)."""
        absent = [self.K is None, self.R is None, self.t is None, self.T is None]
        if all(absent):
            pass
        elif any(absent):
            assert self.K is not None or all(v is not None for v in (self.fx, self.fy, self.cx, self.cy))
            if self.K is None:
                self.K = np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]],
                                  dtype=np.float32)
            if self.T is not None and self.R is None and self.t is None:
                self.R = self.T[:3, :3]
                self.t = self.T[:3, 3]
            elif self.T is None and self.R is not None and self.t is not None:
                # the reference's expression for this case is malformed and raises
                # (calib.py:76); building T here is the evident intent
                self.T = _T_from_Rt(self.R, self.t)
            else:
                raise ValueError("R,t,T not valid", self.R, self.t, self.T)
        else:
            assert np.allclose(self.T, _T_from_Rt(self.R, self.t)), \
                "{} {} {}".format(self.R, self.t, self.T)

        if self.mode == "from_vps" and self.pp is None:
            self.pp = np.zeros_like(self.vp1)
            self.pp[0] = (self.u_size - 1) * 0.5
            self.pp[1] = (self.v_size - 1) * 0.5

    def check_validity(self):
        absent = [self.K is None, self.R is None, self.t is None, self.T is None]
        assert all(absent) or not any(absent)
        if not any(absent):
            assert np.allclose(self.T, _T_from_Rt(self.R, self.t)), \
                "{} {} {}".format(self.R, self.t, self.T)
        if self.mode == "from_pts":
            assert self.pts_image is not None and self.pts_world is not None
        elif self.mode == "from_vps":
            assert all(v is not None for v in (self.vp1, self.vp2, self.pp, self.height,
                                               self.u_size, self.v_size))

    def gen_H_world_img(self, mode=None):
        """float64 3x3 H with world ~ H * image-pixel (reference calib.py:109-127)."""
        self.check_validity()
        mode = self.mode if mode is None else mode
        assert mode in _MODES, mode
        if mode == "from_pts":
            assert self.pts_image is not None and self.pts_world is not None
            return homo_from_pts(self.pts_image, self.pts_world[:, :2])
        if mode == "from_vps":
            H_img_world = homo_from_vps(self.vp1, self.vp2, self.height, self.u_size, self.v_size,
                                        self.pp)
        else:
            assert self.R is not None and self.t is not None
            H_img_world = homo_from_KRt(self.K, Rt_homo=self.T)
        return np.linalg.inv(H_img_world)

    def gen_center_in_world(self):
        """World (x, y, 1) of the image centre pixel ((u-1)/2, (v-1)/2) (calib.py:129-140)."""
        H = self.gen_H_world_img()
        p = H.dot(np.array([(self.u_size - 1) / 2, (self.v_size - 1) / 2, 1.0]))
        return (p / p[2]).reshape(-1)

    def scale(self, align_corners, new_u=None, new_v=None, scale_ratio_u=None, scale_ratio_v=None):
        """Calib of the resized image (reference calib.py:142-198); see BEVWorldSpec.scale for the
        meaning of ``align_corners`` and the size/ratio arguments."""
        if scale_ratio_u is None and scale_ratio_v is None:
            assert new_u is not None and new_v is not None
            if align_corners:
                scale_ratio_u = (new_u - 1) / (self.u_size - 1)
                scale_ratio_v = (new_v - 1) / (self.v_size - 1)
            else:
                scale_ratio_u = new_u / self.u_size
                scale_ratio_v = new_v / self.v_size
        elif align_corners:
            new_u = scale_ratio_u * (self.u_size - 1) + 1
            new_v = scale_ratio_v * (self.v_size - 1) + 1
        else:
            new_u = scale_ratio_u * self.u_size
            new_v = scale_ratio_v * self.v_size

        def mu(c):
            return _map_coord(c, scale_ratio_u, align_corners)

        def mv(c):
            return _map_coord(c, scale_ratio_v, align_corners)

        if self.mode == "from_KRt":
            K = self.K.copy()
            K[0, 0] *= scale_ratio_u
            K[1, 1] *= scale_ratio_v
            K[0, 2] = mu(K[0, 2])
            K[1, 2] = mv(K[1, 2])
            return Calib(K=K, T=self.T.copy(), u_size=new_u, v_size=new_v)
        if self.mode == "from_vps":
            vp1, vp2, pp = self.vp1.copy(), self.vp2.copy(), self.pp.copy()
            for p in (vp1, vp2, pp):
                p[0] = mu(p[0])
                p[1] = mv(p[1])
            return Calib(vp1=vp1, vp2=vp2, height=self.height, u_size=new_u, v_size=new_v, pp=pp)
        pts_image = self.pts_image.copy()
        pts_image[:, 0] = mu(pts_image[:, 0])
        pts_image[:, 1] = mv(pts_image[:, 1])
        return Calib(pts_image=pts_image, pts_world=self.pts_world.copy(), u_size=new_u, v_size=new_v)

    def pad(self, pad_left, pad_top, pad_right, pad_bottom):
        """Calib of the padded image; pads may be negative (reference calib.py:200-229)."""
        new_u = self.u_size + pad_left + pad_right
        new_v = self.v_size + pad_top + pad_bottom
        if self.mode == "from_KRt":
            K = self.K.copy()
            K[0, 2] += pad_left
            K[1, 2] += pad_top
            return Calib(K=K, T=self.T.copy(), u_size=new_u, v_size=new_v)
        if self.mode == "from_vps":
            vp1, vp2, pp = self.vp1.copy(), self.vp2.copy(), self.pp.copy()
            for p in (vp1, vp2, pp):
                p[0] += pad_left
                p[1] += pad_top
            return Calib(vp1=vp1, vp2=vp2, height=self.height, u_size=new_u, v_size=new_v, pp=pp)
        pts_image = self.pts_image.copy()
        pts_image[:, 0] += pad_left
        pts_image[:, 1] += pad_top
        return Calib(pts_image=pts_image, pts_world=self.pts_world.copy(), u_size=new_u, v_size=new_v)

    def flip(self, lr=False, tb=False):
        """Calib of the mirrored image (reference calib.py:231-267).  For ``from_KRt`` the flip is
        folded into K (negative focal entry), as in the reference."""
        if self.mode == "from_KRt":
            K = self.K.copy()
            if lr:
                K[0, 2] = self.u_size - 1 - K[0, 2]
                K[0, 0] = -K[0, 0]
            if tb:
                K[1, 2] = self.v_size - 1 - K[1, 2]
                K[1, 1] = -K[1, 1]
            return Calib(K=K, T=self.T.copy(), u_size=self.u_size, v_size=self.v_size)
        if self.mode == "from_vps":
            vp1, vp2, pp = self.vp1.copy(), self.vp2.copy(), self.pp.copy()
            for p in (vp1, vp2, pp):
                if lr:
                    p[0] = self.u_size - 1 - p[0]
                if tb:
                    p[1] = self.v_size - 1 - p[1]
            return Calib(vp1=vp1, vp2=vp2, height=self.height, u_size=self.u_size,
                         v_size=self.v_size, pp=pp)
        pts_image = self.pts_image.copy()
        if lr:
            pts_image[:, 0] = self.u_size - 1 - pts_image[:, 0]
        if tb:
            pts_image[:, 1] = self.v_size - 1 - pts_image[:, 1]
        return Calib(pts_image=pts_image, pts_world=self.pts_world.copy(), u_size=self.u_size,
                     v_size=self.v_size)
