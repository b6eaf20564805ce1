"""GPU parity: rotated-box IoU matrix (CUDA, through the C ABI) vs the float64 oracle
(oracle/iou_oracle.py, itself pinned against OpenCV's rotated-rectangle intersection in
tests/golden/iou_kat.npz).  Tolerance: 1e-5 absolute on the IoU for float32 boxes (the boxes of
the fixtures are exactly representable in float32), 1e-9 for float64."""
import numpy as np
import pytest
import torch

from bev_b200 import rbox_torch
from oracle import iou_oracle as io
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = util.load_npz("iou_kat.npz")


def cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-9)])
def test_golden_pairs_and_matrix(dtype, tol):
    b1, b2 = K["b1"], K["b2"]
    m = rbox_torch.box2d_iou(cu(b1, dtype), cu(b2, dtype)).cpu().numpy()
    assert m.shape == (len(b1), len(b2))
    assert np.max(np.abs(np.diag(m) - K["iou_pairs"])) <= tol
    out = rbox_torch.box2d_iou(cu(K["dets"], dtype), cu(K["trks"], dtype)).cpu().numpy()
    assert np.max(np.abs(out - K["iou_matrix"])) <= tol
    trk = rbox_torch.iou_batch_rbox(cu(K["dets"], dtype), cu(K["trks"], dtype)).cpu().numpy()
    assert np.max(np.abs(trk - K["iou_tracker"])) <= tol


def test_degenerate_pairs():
    """Identical boxes, coincident edge lines, touching boxes, containment, zero-area boxes."""
    a = np.array([[0, 0, 4, 2, 0.0], [5, 5, 3, 7, 0.7], [0, 0, 4, 2, 0.0], [0, 0, 4, 2, 0.0],
                  [0, 0, 4, 2, 0.0], [0, 0, 6, 6, 0.2], [1, 1, 0, 3, 0.3], [2, 2, 2, 2, np.pi / 4]])
    b = np.array([[0, 0, 4, 2, 0.0], [5, 5, 3, 7, 0.7], [1, 0, 4, 2, 0.0], [4, 0, 4, 2, 0.0],
                  [0, 0, 2, 4, np.pi / 2], [0, 0, 1, 1, 1.0], [1, 1, 2, 3, 0.3], [2, 2, 2, 2, 0.0]])
    ref = np.array([io.box2d_iou(a[i:i + 1], b[i:i + 1])[0, 0] for i in range(len(a))])
    out = rbox_torch.box2d_iou(cu(a, torch.float64), cu(b, torch.float64)).cpu().numpy()
    assert np.max(np.abs(np.diag(out) - ref)) <= 1e-9
    assert np.allclose(np.diag(out)[:5], [1.0, 1.0, 0.6, 0.0, 1.0], atol=1e-9)


def test_large_random_matrix_properties():
    """4096 x 3000 boxes: symmetric in its arguments, in [0, 1], diagonal of a self-matrix is 1,
    and a random sample of entries matches the oracle."""
    rng = np.random.default_rng(5)
    n, m = 4096, 3000
    def boxes(k):
        return np.stack([rng.uniform(0, 200, k), rng.uniform(0, 200, k), rng.uniform(1, 30, k),
                         rng.uniform(1, 30, k), rng.uniform(-4, 4, k)], 1).astype(np.float32)
    a, b = boxes(n), boxes(m)
    ta, tb = cu(a), cu(b)
    ab = rbox_torch.box2d_iou(ta, tb)
    ba = rbox_torch.box2d_iou(tb, ta)
    assert ab.shape == (n, m)
    assert float((ab - ba.T).abs().max()) <= 2e-6
    assert float(ab.min()) >= 0.0 and float(ab.max()) <= 1.0 + 1e-6
    aa = rbox_torch.box2d_iou(ta, ta)
    assert float((aa.diagonal() - 1).abs().max()) <= 1e-6
    host = ab.cpu().numpy()
    idx = np.argwhere(host > 0.02)
    pick = idx[rng.choice(len(idx), 300, replace=False)]
    for i, j in pick:
        ref = io.box2d_iou(a[i:i + 1].astype(np.float64), b[j:j + 1].astype(np.float64))[0, 0]
        assert abs(host[i, j] - ref) <= 1e-5, (i, j)


def test_shapes_and_errors():
    t = torch.zeros(3, 6, device=DEV)
    assert rbox_torch.iou_batch_rbox(t[:0], t).shape == (0, 3)
    assert rbox_torch.iou_batch_rbox(t, t[:0]).shape == (3, 0)
    with pytest.raises(ValueError):
        rbox_torch.box2d_iou(t[:, :4], t)
    with pytest.raises(TypeError):
        rbox_torch.box2d_iou(t, t.double())
    with pytest.raises(RuntimeError):
        rbox_torch.box2d_iou(t.cpu(), t.cpu())
    with pytest.raises(AssertionError):
        rbox_torch.box2d_iou(t, t, method="box")
