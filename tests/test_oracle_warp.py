"""CPU: pins oracle/warp_oracle.c against the reference's arithmetic.

The reference ships no tests for this path, so the pins are (a) outputs of
cv2.warpPerspective 4.13.0 -- the implementation the reference calls at vis_homo.py:89 -- stored
by oracle/gen_golden.py, and (b) the live cv2 build when it is importable.
"""
import numpy as np
import pytest

from oracle import warp_oracle as wo
from tests import util


def test_small_golden_bit_exact():
    n = 0
    for case in util.small_cases():
        out = wo.warp_perspective(case["src"], case["H"], (50, 37), flags=case["flags"],
                                  borderValue=case["bv"])
        assert util.bits_equal(out, case["dst"]), case["name"]
        n += 1
    assert n >= 100


@pytest.mark.parametrize("case", util.hash_cases(), ids=lambda c: c["name"])
def test_full_size_hashes(case):
    src = util.hash_case_input(case)
    out = wo.warp_perspective(src, np.array(case["H"]), case["dsize"], flags=case["flags"])
    assert util.sha256(out) == case["sha256"]


def test_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = util.seeded_frame(77, 270, 480, 3, "uint8")
    for i in range(8):
        s = np.array([[0, 0], [479, 0], [479, 269], [0, 269]], np.float64) + rng.normal(size=(4, 2)) * 60
        d = np.array([[0, 0], [199, 0], [199, 149], [0, 149]], np.float64) + rng.normal(size=(4, 2)) * 20
        H, _ = cv2.findHomography(s, d)
        for flags in (0, 1, 17):
            a = cv2.warpPerspective(img, H, (200, 150), flags=flags)
            b = wo.warp_perspective(img, H, (200, 150), flags=flags)
            assert util.bits_equal(a, b), (i, flags)
    f = util.seeded_frame(78, 270, 480, 3, "float32")
    a = cv2.warpPerspective(f, H, (200, 150), flags=1)
    assert util.bits_equal(a, wo.warp_perspective(f, H, (200, 150), flags=1))


def test_invert_matches_cv2_formula():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(200):
        A = rng.normal(size=(3, 3)) * rng.choice([1e-3, 1.0, 1e3])
        assert util.bits_equal(cv2.invert(A)[1], wo.invert3x3(A))
    assert not wo.invert3x3(np.zeros((3, 3))).any()  # singular -> zero matrix, as cv2


def test_touched_pixels_survey_constants():
    # canonical 1080p -> 1024^2 map touches 971 287 px (bilinear), 557 111 (nearest)
    H = util.h_canon()
    t, r0, r1 = wo.touched_pixels((1920, 1080), (1024, 1024), H, 1)
    assert t == 971287 and (r0, r1) == (420, 1058)
    assert wo.algo_bytes((1920, 1080), (1024, 1024), H, 3, 1, 1) == 6059589
    assert wo.touched_pixels((1920, 1080), (1024, 1024), H, 0)[0] == 557111


def test_fp16_goes_through_fp32():
    H = util.h_canon()
    src = util.seeded_frame(3, 108, 192, 3, "float16")
    Hs = np.diag([0.1, 0.1, 1.0]) @ H @ np.diag([10.0, 10.0, 1.0])
    a = wo.warp_perspective(src, Hs, (100, 100), 1)
    b = wo.warp_perspective(src.astype(np.float32), Hs, (100, 100), 1).astype(np.float16)
    assert a.dtype == np.float16 and util.bits_equal(a, b)


def test_edge_shapes():
    H = np.array([[1.0, 0, 0.5], [0, 1.0, 0.5], [0, 0, 1.0]])
    src = util.seeded_frame(1, 1, 1, 3, "uint8")
    out = wo.warp_perspective(src, H, (3, 2), 1)
    assert out.shape == (2, 3, 3)
    out = wo.warp_perspective(util.seeded_frame(2, 5, 7, 1, "uint8")[:, :, 0], H, (1, 1), 0)
    assert out.shape == (1, 1)
