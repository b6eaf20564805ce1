"""BEVWorldSpec: BEV raster <-> world rectangle correspondence (host side).

Mirrors /root/reference/bev/bev.py:6-175 -- same constructor keywords, attributes, methods and
assertions -- so existing configuration code keeps working.  It defines the output size
``(u_size, v_size)`` and the coordinate conventions the warp kernel honours:

  * BEV pixel coordinates are corner-of-corner-pixel: the raster corners (0,0),(0,v),(u,v),(u,0)
    (not u-1 / v-1) map to the world rectangle corners (bev.py:70-77);
  * ``u_axis`` / ``v_axis`` in {"x","y","-x","-y"} say which world axis each raster axis follows.
"""
import numpy as np

from .frozen_class import FrozenClass
from .homo import homo_from_pts

# raster corner order tl, bl, br, tr -> index into the world corners
# [(x_min,y_min), (x_min,y_max), (x_max,y_max), (x_max,y_min)]   (reference bev.py:85-100)
_CORNER_PERM = {
    ("x", "y"): (0, 1, 2, 3),
    ("x", "-y"): (1, 0, 3, 2),
    ("-x", "-y"): (2, 3, 0, 1),
    ("-x", "y"): (3, 2, 1, 0),
    ("y", "x"): (0, 3, 2, 1),
    ("y", "-x"): (3, 0, 1, 2),
    ("-y", "-x"): (2, 1, 0, 3),
    ("-y", "x"): (1, 2, 3, 0),
}
_AXES = ("x", "y", "-x", "-y")


def _close_interval(lo, hi, size):
    """Fill the one missing value of (lo, hi, size); all three given must already agree."""
    missing = [v is None for v in (lo, hi, size)]
    if any(missing):
        assert sum(missing) == 1, np.array([lo, hi, size])
        if lo is None:
            lo = hi - size
        if hi is None:
            hi = lo + size
        if size is None:
            size = hi - lo
    else:
        assert size == hi - lo
    return lo, hi, size


class BEVWorldSpec(FrozenClass):
    def __init__(self, u_size, v_size, **kwargs):
        self.u_size = u_size
        self.v_size = v_size
        self.u_axis = "-x"
        self.v_axis = "y"
        for name in ("x_size", "y_size", "x_min", "x_max", "y_min", "y_max",
                     # raster coordinates of the world rectangle after scale()/pad()
                     "u_min", "u_max", "v_min", "v_max"):
            setattr(self, name, None)
        self._freeze()
        self.__dict__.update(kwargs)
        self.update()
        self.check_validity()

    def set_keep(self, **kwargs):
        """Overwrite attributes; set the dependent one to None in the same call so update() refills it."""
        self.__dict__.update(kwargs)
        self.update()

    def update(self):
        self.x_min, self.x_max, self.x_size = _close_interval(self.x_min, self.x_max, self.x_size)
        self.y_min, self.y_max, self.y_size = _close_interval(self.y_min, self.y_max, self.y_size)

    def check_validity(self):
        for trio in ((self.x_min, self.x_max, self.x_size), (self.y_min, self.y_max, self.y_size)):
            assert all(v is not None for v in trio)
            assert np.isclose(trio[2], trio[1] - trio[0])
        assert self.u_axis in _AXES
        assert self.v_axis in _AXES
        assert ("x" in self.u_axis and "y" in self.v_axis) or ("y" in self.u_axis and "x" in self.v_axis)

    def _raster_rect(self):
        if self.u_min is None:
            return 0, self.u_size, 0, self.v_size
        return self.u_min, self.u_max, self.v_min, self.v_max

    def gen_H_world_bev(self):
        """3x3 float64 H with world ~ H * bev (4-point homography, reference bev.py:67-79)."""
        self.check_validity()
        u0, u1, v0, v1 = self._raster_rect()
        pts_bev = np.array([[u0, v0], [u0, v1], [u1, v1], [u1, v0]], dtype=float)
        return homo_from_pts(pts_bev, self.gen_bev_corners_in_world())

    def gen_bev_corners_in_world(self):
        """World coordinates of the raster's top-left, bottom-left, bottom-right, top-right."""
        rect = np.array([[self.x_min, self.y_min], [self.x_min, self.y_max],
                         [self.x_max, self.y_max], [self.x_max, self.y_min]], dtype=float)
        key = (self.u_axis, self.v_axis)
        if key not in _CORNER_PERM:
            raise ValueError("illegal u_axis and v_axis combo", self.u_axis, self.v_axis)
        return rect[list(_CORNER_PERM[key])]

    def _derive(self, u_size, v_size, rect, u_axis=None, v_axis=None):
        u0, u1, v0, v1 = rect
        return BEVWorldSpec(u_size=u_size, v_size=v_size,
                            u_axis=self.u_axis if u_axis is None else u_axis,
                            v_axis=self.v_axis if v_axis is None else v_axis,
                            x_size=self.x_size, y_size=self.y_size, x_min=self.x_min, y_min=self.y_min,
                            u_min=u0, v_min=v0, u_max=u1, v_max=v1)

    def scale(self, align_corners, new_u=None, new_v=None, scale_ratio_u=None, scale_ratio_v=None):
        """New spec for a resized raster (reference bev.py:107-140).

        ``align_corners=True`` aligns corner pixel centres (ratio (new-1)/(old-1)); otherwise pixel
        corners are aligned (ratio new/old) and coordinates follow the half-pixel rule
        ``(c + 0.5) * s - 0.5``.  Give either the new size or the ratios.
        """
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

        u0, u1, v0, v1 = self._raster_rect()
        if align_corners:
            rect = (u0 * scale_ratio_u, u1 * scale_ratio_u, v0 * scale_ratio_v, v1 * scale_ratio_v)
        else:
            rect = ((u0 + 0.5) * scale_ratio_u - 0.5, (u1 + 0.5) * scale_ratio_u - 0.5,
                    (v0 + 0.5) * scale_ratio_v - 0.5, (v1 + 0.5) * scale_ratio_v - 0.5)
        return self._derive(new_u, new_v, rect)

    def pad(self, pad_left, pad_top, pad_right, pad_bottom):
        """New spec for a padded (or, with negative values, cropped) raster (bev.py:142-160)."""
        u0, u1, v0, v1 = self._raster_rect()
        rect = (u0 + pad_left, u1 + pad_left, v0 + pad_top, v1 + pad_top)
        return self._derive(self.u_size + pad_left + pad_right, self.v_size + pad_top + pad_bottom,
                            rect)

    def flip(self, lr=False, tb=False):
        """New spec for a mirrored raster: flips the sign of the axis name (bev.py:162-175)."""
        def neg(a):
            return a[1] if "-" in a else "-" + a
        u_axis = neg(self.u_axis) if lr else self.u_axis
        v_axis = neg(self.v_axis) if tb else self.v_axis
        return BEVWorldSpec(u_size=self.u_size, v_size=self.v_size, x_size=self.x_size,
                            y_size=self.y_size, x_min=self.x_min, y_min=self.y_min,
                            u_min=self.u_min, v_min=self.v_min, u_max=self.u_max, v_max=self.v_max,
                            u_axis=u_axis, v_axis=v_axis)
