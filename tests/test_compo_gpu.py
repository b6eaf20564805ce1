"""GPU parity: bev_b200.compo (CUDA, through the C ABI) vs the compositing oracle and the outputs
of the reference's bev/tool/compo.py (tests/golden/compo_kat.npz).  Bar: bit-exact."""
import numpy as np
import pytest
import torch

from bev_b200 import compo
from oracle import compo_oracle as co
from oracle.synth import compo_inputs
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = util.load_npz("compo_kat.npz")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_blend_golden():
    bg, fg, mask = compo_inputs(4242, 161, 241)  # 38801 pixels: exercises the tail pixels
    assert np.array_equal(compo.composite_reg_img(cu(bg), cu(fg), cu(mask)).cpu().numpy(), K["reg"])
    out = compo.composite_reg_img(cu(bg), cu(fg), cu(mask), bw_mode=True).cpu().numpy()
    assert np.array_equal(out, K["reg_bw"])


def test_blend_every_value_pair():
    """All 256 x 256 x 256 (bg, fg, mask) byte triples against the float64 numpy expression."""
    v = np.arange(256, dtype=np.uint8)
    bg, fg = np.meshgrid(v, v, indexing="ij")
    for m in range(256):
        b3 = np.repeat(bg[..., None], 3, 2)
        f3 = np.repeat(fg[..., None], 3, 2)
        m3 = np.full_like(b3, m)
        out = compo.composite_reg_img(cu(b3), cu(f3), cu(m3)).cpu().numpy()
        assert np.array_equal(out, co.composite_reg_img(b3, f3, m3)), m


def test_bev_composite_golden_and_batch():
    bg, fg, mask = compo_inputs(4343, 160, 240)
    args = (K["H_world2bev"], K["H_img2world_fix"], K["K"], K["RT"], 160, 120)
    for tag, bw in (("bev", False), ("bev_bw", True)):
        c, Hcam = compo.composite_bev_img(cu(bg), cu(fg), cu(mask), *args, bw_mode=bw)
        assert np.array_equal(c.cpu().numpy(), K[tag]), tag
        assert np.allclose(Hcam, K[tag + "_Hcam"], rtol=0, atol=1e-12)
    # a batch of composites in one call == the single-frame results
    n = 5
    sets = [compo_inputs(5000 + 3 * i, 160, 240) for i in range(n)]
    B, F, M = (np.stack([s[j] for s in sets]) for j in range(3))
    c, _ = compo.composite_bev_img(cu(B), cu(F), cu(M), *args)
    for i in range(n):
        ref, _ = co.composite_bev_img(sets[i][0], sets[i][1], sets[i][2], *args)
        assert np.array_equal(c[i].cpu().numpy(), ref), i


def test_argument_errors():
    t = torch.zeros(4, 4, 3, dtype=torch.uint8, device=DEV)
    with pytest.raises(TypeError):
        compo.composite_reg_img("bg.png", t, t)
    with pytest.raises(ValueError):
        compo.composite_reg_img(t, t[:2], t)
    with pytest.raises(TypeError):
        compo.composite_reg_img(t.float(), t, t)


# ---- fused compositor (bevk_composite_bev_u8c3) vs three warps + blend vs the oracle ------------
def _quad_h(rng, src_wh, dst_wh, jitter):
    """Homography taking a jittered quad of the source onto the jittered dst rectangle."""
    import cv2
    sw, sh = src_wh
    dw, dh = dst_wh
    s = np.float32([[0, 0], [sw, 0], [sw, sh], [0, sh]]) + rng.uniform(-jitter, jitter, (4, 2)).astype(np.float32) * [sw, sh]
    d = np.float32([[0, 0], [dw, 0], [dw, dh], [0, dh]]) + rng.uniform(-jitter, jitter, (4, 2)).astype(np.float32) * [dw, dh]
    return cv2.getPerspectiveTransform(s.astype(np.float32), d.astype(np.float32)).astype(np.float64)


def _oracle_batch(B, F, M, Hb, Hf, dsize):
    from oracle import warp_oracle
    out = []
    for i in range(F.shape[0]):
        b = B[i if B.shape[0] > 1 else 0]
        hb = Hb[i if Hb.shape[0] > 1 else 0]
        hf = Hf[i if Hf.shape[0] > 1 else 0]
        out.append(co.composite_reg_img(warp_oracle.warp_perspective(b, hb, dsize),
                                        warp_oracle.warp_perspective(F[i], hf, dsize),
                                        warp_oracle.warp_perspective(M[i], hf, dsize)))
    return np.stack(out)


@pytest.mark.parametrize("n,n_bg,n_mats,bg_hw,fg_hw,dsize", [
    (3, 3, 1, (60, 88), (60, 88), (64, 50)),      # shared cameras, a background per frame
    (5, 1, 1, (72, 100), (40, 64), (96, 33)),     # one background, renders of another size
    (4, 1, 4, (50, 64), (64, 48), (36, 70)),      # a rendering camera per frame
    (2, 2, 2, (33, 12), (9, 8), (8, 9)),          # tiny frames: every window touches a border
    (40, 1, 40, (24, 32), (24, 32), (32, 16)),    # more camera pairs than one launch carries
])
def test_fused_bev_composite(n, n_bg, n_mats, bg_hw, fg_hw, dsize):
    rng = np.random.default_rng(n * 1000 + n_mats)
    B = rng.integers(0, 256, (n_bg,) + bg_hw + (3,), dtype=np.uint8)
    F = rng.integers(0, 256, (n,) + fg_hw + (3,), dtype=np.uint8)
    M = np.repeat(rng.integers(0, 256, (n,) + fg_hw + (1,), dtype=np.uint8), 3, axis=3)
    M[:, : fg_hw[0] // 3] = 255
    Hb = np.stack([_quad_h(rng, bg_hw[::-1], dsize, 0.3) for _ in range(n_mats)])
    Hf = np.stack([_quad_h(rng, fg_hw[::-1], dsize, 0.3) for _ in range(n_mats)])
    ref = _oracle_batch(B, F, M, Hb, Hf, dsize)
    fused = compo.composite_bev_batch(cu(B), cu(F), cu(M), Hb, Hf, dsize, fused=True).cpu().numpy()
    assert np.array_equal(fused, ref)
    unfused = compo.composite_bev_batch(cu(B), cu(F), cu(M), Hb, Hf, dsize, fused=False).cpu().numpy()
    assert np.array_equal(unfused, ref)


def test_fused_bev_composite_shape_rules():
    """Widths that are not multiples of 4 take the three-warp route automatically; forcing the
    fused kernel on them is an argument error, never a silent fallback."""
    rng = np.random.default_rng(77)
    B = rng.integers(0, 256, (1, 30, 41, 3), dtype=np.uint8)
    F = rng.integers(0, 256, (2, 30, 41, 3), dtype=np.uint8)
    M = rng.integers(0, 256, (2, 30, 41, 3), dtype=np.uint8)
    H = _quad_h(rng, (41, 30), (37, 22), 0.2)
    ref = _oracle_batch(B, F, M, H[None], H[None], (37, 22))
    out = compo.composite_bev_batch(cu(B), cu(F), cu(M), H, H, (37, 22)).cpu().numpy()
    assert np.array_equal(out, ref)
    from bev_b200._native import NativeError
    with pytest.raises(NativeError, match="multiples of 4"):
        compo.composite_bev_batch(cu(B), cu(F), cu(M), H, H, (37, 22), fused=True)


def test_fused_bev_composite_full_size():
    """1080p renders over a 1080p background into a 1024^2 BEV: fused == three warps + blend."""
    g = torch.Generator(device=DEV).manual_seed(5)
    n = 6
    B = torch.randint(0, 256, (1, 1080, 1920, 3), dtype=torch.uint8, device=DEV, generator=g)
    F = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device=DEV, generator=g)
    M = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device=DEV, generator=g)
    Hb = util.h_canon()
    rng = np.random.default_rng(9)
    Hf = np.stack([_quad_h(rng, (1920, 1080), (1024, 1024), 0.15) for _ in range(n)])
    Hbn = np.repeat(np.asarray(Hb, np.float64)[None], n, 0)
    a = compo.composite_bev_batch(B, F, M, Hbn, Hf, (1024, 1024), fused=True)
    b = compo.composite_bev_batch(B, F, M, Hbn, Hf, (1024, 1024), fused=False)
    assert torch.equal(a, b)
    a1 = compo.composite_bev_batch(B, F, M, Hb, Hf[0], (1024, 1024), fused=True)
    b1 = compo.composite_bev_batch(B, F, M, Hb, Hf[0], (1024, 1024), fused=False)
    assert torch.equal(a1, b1)
    assert torch.equal(a1[0], a[0])


@pytest.mark.parametrize("dsize", [(37, 22), (1023, 1023), (5, 3)])
def test_single_frame_odd_bev_and_unaligned_views(dsize):
    """One image into a BEV whose byte size is not a multiple of 4: the unfused route blends the
    two halves of ONE concatenated warp output, so the mask half starts on an odd byte offset --
    it must still work (the reference handles any size).  Same for user views at odd offsets."""
    rng = np.random.default_rng(78)
    B = rng.integers(0, 256, (1, 30, 41, 3), dtype=np.uint8)
    F = rng.integers(0, 256, (1, 30, 41, 3), dtype=np.uint8)
    M = rng.integers(0, 256, (1, 30, 41, 3), dtype=np.uint8)
    H = _quad_h(rng, (41, 30), dsize, 0.2)
    ref = _oracle_batch(B, F, M, H[None], H[None], dsize)
    out = compo.composite_bev_batch(cu(B), cu(F), cu(M), H, H, dsize).cpu().numpy()
    assert np.array_equal(out, ref)
    # composite_reg_img on views that start 1, 2, 3 bytes into a buffer
    flat = torch.from_numpy(rng.integers(0, 256, 3 * 30 * 41 * 3 + 16, dtype=np.uint8)).to(DEV)
    for off in (1, 2, 3):
        n = 30 * 41 * 3
        bg, fg, mk = (flat[off + i * n: off + (i + 1) * n].view(30, 41, 3) for i in range(3))
        ref = co.composite_reg_img(bg.cpu().numpy(), fg.cpu().numpy(), mk.cpu().numpy())
        assert np.array_equal(compo.composite_reg_img(bg, fg, mk).cpu().numpy(), ref)


def test_out_arguments_are_validated():
    from bev_b200 import _native
    a = torch.zeros((4, 8, 8, 3), dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        _native.composite_u8c3(a, a, a, out=torch.zeros((3, 8, 8, 3), dtype=torch.uint8, device=DEV))
    with pytest.raises(ValueError):
        _native.composite_u8c3(a, a, a, out=torch.zeros((4, 8, 8, 3), dtype=torch.float32, device=DEV))
    with pytest.raises(ValueError):
        _native.composite_bev_u8c3(a[:1], a, a, np.eye(3), np.eye(3), (8, 8),
                                   out=torch.zeros((2, 8, 8, 3), dtype=torch.uint8, device=DEV))
    with pytest.raises(ValueError):  # resize: dst with too few frames would be written out of bounds
        _native.resize(a, (4, 4), dst=torch.zeros((2, 4, 4, 3), dtype=torch.uint8, device=DEV))
