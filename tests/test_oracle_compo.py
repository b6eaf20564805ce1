"""CPU: the compositing oracle (oracle/compo_oracle.py) against outputs of the reference's own
bev/tool/compo.py (tests/golden/compo_kat.npz, written by oracle/gen_golden.py in the build
container with cv2 4.13)."""
import numpy as np

from oracle import compo_oracle as co
from oracle.synth import compo_inputs
from tests import util

K = util.load_npz("compo_kat.npz")


def test_blend_matches_the_reference():
    bg, fg, mask = compo_inputs(4242, 161, 241)
    assert np.array_equal(co.composite_reg_img(bg, fg, mask), K["reg"])
    assert np.array_equal(co.composite_reg_img(bg, fg, mask, bw_mode=True), K["reg_bw"])


def test_bev_composite_matches_the_reference():
    bg, fg, mask = compo_inputs(4343, 160, 240)
    for tag, bw in (("bev", False), ("bev_bw", True)):
        c, Hcam = co.composite_bev_img(bg, fg, mask, K["H_world2bev"], K["H_img2world_fix"], K["K"],
                                       K["RT"], 160, 120, bw_mode=bw)
        assert np.array_equal(c, K[tag]), tag
        assert np.allclose(Hcam, K[tag + "_Hcam"], rtol=0, atol=1e-12)


def test_blend_edge_values():
    """mask 0 / 255 select bg / fg exactly; exact .5 sums round half to even like np.round."""
    bg = np.array([[[0, 255, 7]]], np.uint8)
    fg = np.array([[[255, 0, 8]]], np.uint8)
    assert np.array_equal(co.composite_reg_img(bg, fg, np.zeros_like(bg)), bg)
    assert np.array_equal(co.composite_reg_img(bg, fg, np.full_like(bg, 255)), fg)


def test_integer_blend_is_the_float64_blend_for_every_byte_triple():
    """The kernels blend in integers (compo.cu); this is the proof obligation: the integer form
    equals the reference's float64 numpy expression for all 256^3 (bg, fg, mask) values."""
    v = np.arange(256, dtype=np.uint8)
    bg, fg = np.meshgrid(v, v, indexing="ij")
    for k in range(256):
        m = np.full_like(bg, k)
        assert np.array_equal(co.blend_integer(bg, fg, m), co.composite_reg_img(bg, fg, m)), k
