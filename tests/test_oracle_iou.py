"""CPU: the rotated-box IoU oracle (oracle/iou_oracle.py) against the fixtures of
tests/golden/iou_kat.npz -- OpenCV's rotated-rectangle intersection areas computed in the build
container -- and against closed-form cases."""
import numpy as np

from oracle import iou_oracle as io
from tests import util

K = util.load_npz("iou_kat.npz")


def test_intersection_areas_match_opencv():
    b1, b2 = K["b1"], K["b2"]
    scale = np.maximum(b1[:, 2] * b1[:, 3], b2[:, 2] * b2[:, 3])
    mine = np.array([io.intersection_area(b1[i], b2[i]) for i in range(len(b1))])
    assert np.all(np.abs(mine - K["inter_cv"]) <= 5e-4 * scale)  # OpenCV works in float32


def test_stored_values_reproduce():
    b1, b2 = K["b1"], K["b2"]
    pairs = np.array([io.box2d_iou(b1[i:i + 1], b2[i:i + 1])[0, 0] for i in range(len(b1))])
    assert np.allclose(pairs, K["iou_pairs"], rtol=0, atol=1e-12)
    assert np.allclose(io.box2d_iou(K["dets"], K["trks"]), K["iou_matrix"], rtol=0, atol=1e-12)
    assert np.allclose(io.iou_batch_rbox(K["dets"], K["trks"]), K["iou_tracker"], rtol=0, atol=1e-12)


def test_closed_forms():
    a = [[0, 0, 4, 2, 0]]
    got = io.box2d_iou(a, [[0, 0, 4, 2, 0], [1, 0, 4, 2, 0], [10, 0, 4, 2, 0], [0, 0, 2, 4, np.pi / 2],
                           [0, 0, 4, 2, np.pi], [0, 0, 1, 1, 0.3]])[0]
    assert np.allclose(got, [1.0, 0.6, 0.0, 1.0, 1.0, 1.0 / 8.0], atol=1e-12)
    # a square turned by 45 degrees inside a concentric square of the same size: an octagon
    s = 2.0
    octagon = s * s * (2 * np.sqrt(2) - 2)
    assert np.isclose(io.box2d_iou([[0, 0, s, s, 0]], [[0, 0, s, s, np.pi / 4]])[0, 0],
                      octagon / (2 * s * s - octagon), atol=1e-12)
    # the tracker's wrapper turns both boxes by 90 degrees about their own centres
    d, t = [[0, 0, 4, 2, 0.0]], [[3, 0, 4, 2, 0.0]]
    assert np.isclose(io.box2d_iou(d, t)[0, 0], 2.0 / 14.0)
    assert np.isclose(io.iou_batch_rbox(d, t)[0, 0], 0.0)
