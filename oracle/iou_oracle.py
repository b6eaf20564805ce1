"""oracle/iou_oracle.py -- TEST INFRASTRUCTURE ONLY.

float64 restatement of the rotated-box IoU matrix behind the reference tracker's association
step (/root/reference/bev/tracker/rbox_tracker.py:87-92, used at :383-405):

    iou_batch_rbox(bb_test, bb_gt) = d3d.box.box2d_iou(bb_test[:, :5] + [0,0,0,0,pi/2],
                                                       bb_gt[:, :5]  + [0,0,0,0,pi/2], method="rbox")

``d3d`` (cmpute/d3d) is a third-party dependency that is NOT in /root/reference and not pinned by
it (setup.py lists no requirements); it is not installed in this image either.  Its published
operation is the plain geometric one: boxes [x, y, w, h, r] are rectangles centred at (x, y) with
side w along (cos r, sin r) and side h along (-sin r, cos r); IoU = area(A n B) / (area A + area B
- area(A n B)).  That box convention is the common one (it is also cv2.RotatedRect's) but could not
be checked against d3d's source here: **the convention is unpinned**.  The geometry is pinned:
oracle/gen_golden.py checks this module against cv2.rotatedRectangleIntersection + contourArea
(float32 inside OpenCV, hence a 2e-4 tolerance) before writing tests/golden/iou_kat.npz.

The algorithm here (Sutherland-Hodgman clipping of one quad by the other, shoelace area) is
deliberately not the one the CUDA kernel uses (clipped-edge boundary integrals).
Never imported by bev_b200/.
"""
import numpy as np


def corners(box):
    """CCW corners (4, 2) of [x, y, w, h, r]: w along (cos r, sin r), h along (-sin r, cos r)."""
    x, y, w, h, r = (float(v) for v in box[:5])
    c, s = np.cos(r), np.sin(r)
    u = np.array([c, s]) * (w / 2)
    v = np.array([-s, c]) * (h / 2)
    ctr = np.array([x, y])
    return np.array([ctr - u - v, ctr + u - v, ctr + u + v, ctr - u + v])


def _clip(poly, a, b):
    """Keep the part of convex polygon ``poly`` on the left of the directed line a -> b."""
    out = []
    n = len(poly)
    d = b - a
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = d[0] * (p[1] - a[1]) - d[1] * (p[0] - a[0])
        sq = d[0] * (q[1] - a[1]) - d[1] * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return out


def intersection_area(b1, b2):
    p = list(corners(b1))
    q = corners(b2)
    for i in range(4):
        if len(p) < 3:
            return 0.0
        p = _clip(p, q[i], q[(i + 1) % 4])
    if len(p) < 3:
        return 0.0
    p = np.array(p)
    x, y = p[:, 0], p[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def box2d_iou(boxes1, boxes2):
    """(N, >=5), (M, >=5) -> (N, M) float64 IoU matrix (d3d.box.box2d_iou, method="rbox")."""
    b1 = np.asarray(boxes1, np.float64)
    b2 = np.asarray(boxes2, np.float64)
    out = np.zeros((b1.shape[0], b2.shape[0]))
    for i in range(b1.shape[0]):
        for j in range(b2.shape[0]):
            inter = intersection_area(b1[i], b2[j])
            union = abs(b1[i, 2] * b1[i, 3]) + abs(b2[j, 2] * b2[j, 3]) - inter
            out[i, j] = inter / union if union > 0 else 0.0
    return out


def iou_batch_rbox(bb_test, bb_gt):
    """rbox_tracker.py:87-92."""
    a = np.asarray(bb_test, np.float64)[:, :5].copy()
    b = np.asarray(bb_gt, np.float64)[:, :5].copy()
    a[:, 4] += np.pi / 2
    b[:, 4] += np.pi / 2
    return box2d_iou(a, b)
