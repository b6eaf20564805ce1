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
    for m in range(0, 256, 5):
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
