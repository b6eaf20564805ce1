"""torch.ops.bev_cuda (the PyTorch C++ extension in front of libbev_b200.so) against the oracle and
against the ctypes route: same kernels, identical bytes."""
import numpy as np
import pytest
import torch

from bev_b200 import homo, rbox_torch, torch_ops
from oracle import rbox_oracle
from oracle import warp_oracle as wo
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _ops():
    torch_ops.load()


def test_warp_operator_matches_oracle_and_ctypes_route():
    S = np.diag([0.25, 0.25, 1.0])
    H = S @ util.h_canon() @ np.linalg.inv(S)  # 480x270 -> 256x256
    frames = np.stack([util.seeded_frame(900 + i, 270, 480, 3, "uint8") for i in range(9)])
    t = torch.from_numpy(frames).to(DEV)
    Ht = torch.from_numpy(H)
    for flags in (1, 0):
        out = torch.ops.bev_cuda.warp_perspective(t, Ht, 256, 256, flags, 0, 0.0, None)
        assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (9, 256, 256, 3)
        assert torch.equal(out, homo.warp_perspective(t, H, (256, 256), flags=flags))
        for i in (0, 8):
            assert util.bits_equal(out[i].cpu().numpy(), wo.warp_perspective(frames[i], H, (256, 256), flags))
    # matrix table + per-frame index, WARP_INVERSE_MAP, a single (H, W, C) frame
    Hs = torch.from_numpy(np.stack([np.linalg.inv(H), np.linalg.inv(H) @ np.diag([1.0, 0.9, 1.0])]))
    idx = torch.tensor([i % 2 for i in range(9)], dtype=torch.int32)
    out = torch.ops.bev_cuda.warp_perspective(t, Hs, 256, 256, 17, 0, 0.0, idx)
    for i in (1, 4):
        ref = wo.warp_perspective(frames[i], Hs[int(idx[i])].numpy(), (256, 256), flags=17)
        assert util.bits_equal(out[i].cpu().numpy(), ref)
    one = torch.ops.bev_cuda.warp_perspective(t[3], Ht, 256, 256)
    assert tuple(one.shape) == (256, 256, 3) and torch.equal(one, homo.warp_perspective(t[3], H, (256, 256)))


def test_projection_operators():
    rng = np.random.default_rng(3)
    n = 10007
    box = np.stack([rng.uniform(0, 1024, n), rng.uniform(0, 1024, n), rng.uniform(4, 40, n),
                    rng.uniform(8, 120, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    H_back = util.h_canon()
    H_fwd = np.linalg.inv(H_back)
    tb = torch.from_numpy(box).to(DEV)
    img = torch.ops.bev_cuda.rbox_corners_project(tb, torch.from_numpy(H_fwd), 0)
    ref = rbox_oracle.xywhr_to_img_corners(box, H_fwd, "bev")
    assert util.rel_err(img.cpu().numpy(), ref) <= 1e-5
    assert torch.equal(img, rbox_torch.xywhr_to_img_corners(tb, H_fwd, "bev"))
    back = torch.ops.bev_cuda.corners_to_rbox(img, torch.from_numpy(H_back), 0)
    assert torch.equal(back, rbox_torch.img_corners_to_xywhr(img, H_back, "bev"))
    plain = torch.ops.bev_cuda.rbox_corners_project(tb, None, 1)
    assert torch.equal(plain, rbox_torch.xywhr2xyxy(tb, "world"))
    pts = torch.from_numpy(rng.uniform(0, 1000, (999, 2)).astype(np.float32)).to(DEV)
    assert torch.equal(torch.ops.bev_cuda.project_points(pts, torch.from_numpy(H_fwd)),
                       rbox_torch.pts_world_bev(pts, H_fwd))
    Hsim = np.array([[0.0, -0.125, 45.0], [-0.125, 0.0, 27.0], [0.0, 0.0, 1.0]])
    sim = torch.ops.bev_cuda.rbox_similarity(tb, torch.from_numpy(Hsim), 1)
    assert torch.equal(sim, rbox_torch.rbox_world_bev(tb, Hsim, "world"))


def test_errors_like_the_reference_and_no_cpu_fallback():
    tb = torch.zeros((4, 5), dtype=torch.float32, device=DEV)
    Hpersp = torch.from_numpy(util.h_canon())
    with pytest.raises(RuntimeError, match="AssertionError"):  # non-affine H where rbox_torch.py:140 asserts
        torch.ops.bev_cuda.rbox_similarity(tb, Hpersp, 1)
    with pytest.raises((RuntimeError, NotImplementedError)):   # CPU tensors: no kernel registered
        torch.ops.bev_cuda.rbox_corners_project(tb.cpu(), None, 0)
    with pytest.raises(RuntimeError, match="CPU tensor"):      # homographies stay on the host
        torch.ops.bev_cuda.project_points(tb[:, :2].contiguous(), Hpersp.to(DEV))
