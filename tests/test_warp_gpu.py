"""GPU parity: bev_b200.homo.warp_perspective (CUDA, through the C ABI) vs the oracle
(oracle/warp_oracle.c, pinned to cv2 4.13) and the committed cv2 fixtures.

Bars (BASELINE.json north_star / SURVEY.md 8c): nearest bit-exact; bilinear uint8 contract <= 1 LSB,
asserted bit-exact here; float32 bit-exact; float16 == float16(oracle(float32(src)))."""
import numpy as np
import pytest
import torch

from bev_b200 import _native, homo
from oracle import warp_oracle as wo
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# "fast" forces the staged TMA kernel wherever the shape qualifies (uint8 x 3, zero border), also
# for the tiny batches that "auto" hands to the direct-gather kernel; other shapes run as "auto".
PATHS = ["generic", "auto", "fast"]
_path = {"name": "auto"}


def gpu_warp(src_np, H, dsize, flags=1, bv=0, **kw):
    t = torch.from_numpy(np.ascontiguousarray(src_np)).to(DEV)
    try:
        out = homo.warp_perspective(t, H, dsize, flags=flags, borderValue=bv, **kw)
    except _native.NativeError as e:
        if _path["name"] != "fast" or "does not qualify" not in str(e):
            raise
        _native.set_warp_path("auto")
        try:
            out = homo.warp_perspective(t, H, dsize, flags=flags, borderValue=bv, **kw)
        finally:
            _native.set_warp_path("fast")
    assert out.device.type == "cuda" and out.dtype == t.dtype
    return out.cpu().numpy()


@pytest.fixture(params=PATHS)
def path(request):
    _native.set_warp_path(request.param)
    _path["name"] = request.param
    yield request.param
    _native.set_warp_path("auto")
    _path["name"] = "auto"


def test_native_library_is_the_one_running():
    sm, major, minor = _native.device_info()
    assert major == 10 and sm >= 100


def test_small_golden_fixtures(path):
    for case in util.small_cases():
        out = gpu_warp(case["src"], case["H"], (50, 37), case["flags"], case["bv"])
        assert util.bits_equal(out, case["dst"]), (path, case["name"])


@pytest.mark.parametrize("case", util.hash_cases(), ids=lambda c: c["name"])
def test_full_size_vs_cv2_hash(case, path):
    src = util.hash_case_input(case)
    out = gpu_warp(src, np.array(case["H"]), case["dsize"], case["flags"])
    assert util.sha256(out) == case["sha256"], path


@pytest.mark.parametrize("dtype", ["uint8", "float16", "float32"])
@pytest.mark.parametrize("ch", [1, 2, 3, 4])
@pytest.mark.parametrize("flags", [0, 1, 16, 17])
def test_dtype_channel_matrix_vs_oracle(dtype, ch, flags, path):
    rng = np.random.default_rng(ch * 31 + flags)
    src = util.seeded_frame(50 + ch, 135, 240, ch, dtype)
    s = np.array([[0, 0], [239, 0], [239, 134], [0, 134]], np.float64) + rng.normal(size=(4, 2)) * 25
    d = np.array([[0, 0], [159, 0], [159, 99], [0, 99]], np.float64) + rng.normal(size=(4, 2)) * 10
    H = homo.homo_from_pts(s, d)
    if flags & 16:
        H = np.linalg.inv(H)
    ref = wo.warp_perspective(src, H, (160, 100), flags=flags, borderValue=(3, 50, 100, 250))
    out = gpu_warp(src, H, (160, 100), flags, (3, 50, 100, 250))
    assert util.bits_equal(out, ref)


def test_batch_one_matrix_matches_per_frame(path):
    H = util.h_canon()
    S = np.diag([0.25, 0.25, 1.0])
    Hs = S @ H @ np.linalg.inv(S)  # canonical geometry at quarter size: 480x270 -> 256x256
    frames = np.stack([util.seeded_frame(900 + i, 270, 480, 3, "uint8") for i in range(37)])
    out = gpu_warp(frames, Hs, (256, 256), 1)
    assert out.shape == (37, 256, 256, 3)
    for i in (0, 1, 17, 36):
        assert util.bits_equal(out[i], wo.warp_perspective(frames[i], Hs, (256, 256), 1)), i


def test_many_matrices_one_call(path):
    cams = util.load_json("cfg4_cams.json")
    S = np.diag([0.25, 0.25, 1.0])
    n = 24
    frames = np.stack([util.seeded_frame(700 + i, 270, 480, 3, "uint8") for i in range(n)])
    # all cameras rendered to one BEV size so they can share a batch
    Hs = np.stack([np.diag([0.5, 0.5, 1.0]) @ np.array(c["H_bev_img"]) @ np.linalg.inv(S) for c in cams])
    idx = np.array([(i * 5) % 8 for i in range(n)], np.int32)  # interleaved, not contiguous
    out = gpu_warp(frames, Hs, (160, 320), 1, mat_index=idx)
    for i in range(n):
        assert util.bits_equal(out[i], wo.warp_perspective(frames[i], Hs[idx[i]], (160, 320), 1)), i
    # one matrix per frame, no index
    out2 = gpu_warp(frames[:8], Hs, (160, 320), 0)
    for i in range(8):
        assert util.bits_equal(out2[i], wo.warp_perspective(frames[i], Hs[i], (160, 320), 0)), i


def test_split_launch_on_camera_homographies(path):
    """Real camera geometry at a large BEV: 5-15 % of the staged kernel's tiles (the nearest BEV
    rows) have source boxes no tile shape can stage; the launch is split -- staged kernel for the
    tiles that stage, direct-gather kernel over the tiles it marked.  Every pixel must come out
    exactly once and exact."""
    cams = util.load_json("cfg4_cams.json")
    frames = np.stack([util.seeded_frame(900 + i, 1080, 1920, 3, "uint8") for i in range(5)])
    for k, flags in ((0, 1), (6, 1), (3, 0)):
        c = cams[k]
        u, v = c["bspec"]["u_size"], c["bspec"]["v_size"]
        H = np.diag([1024.0 / u, 1024.0 / v, 1.0]) @ np.array(c["H_bev_img"])
        out = gpu_warp(frames, H, (1024, 1024), flags)
        for i in (0, 4):
            assert util.bits_equal(out[i], wo.warp_perspective(frames[i], H, (1024, 1024), flags)), (k, i)
    # float16 frames take the same split (staged float16 policy + the generic float16 kernel)
    f16 = np.stack([util.seeded_frame(950 + i, 1080, 1920, 3, "float16") for i in range(4)])
    c = cams[1]
    H = np.diag([1024.0 / c["bspec"]["u_size"], 1024.0 / c["bspec"]["v_size"], 1.0]) @ np.array(c["H_bev_img"])
    out = gpu_warp(f16, H, (1024, 1024), 1)
    assert util.bits_equal(out[3], wo.warp_perspective(f16[3], H, (1024, 1024), 1))


def test_round_trip_img_bev_img(path):
    """Size-independent property at BASELINE size: warping a smooth image to BEV and back with the
    inverse map reproduces it inside the BEV footprint (up to interpolation blur)."""
    H = util.h_canon()
    yy, xx = np.mgrid[0:1080, 0:1920]
    img = ((xx * 0.05 + yy * 0.11) % 256).astype(np.uint8)
    img = np.stack([img, 255 - img, (img // 2)], -1)
    bev = gpu_warp(img, H, (1024, 1024), 1)
    back = gpu_warp(bev, H, (1920, 1080), 1 | 16)
    ref_bev = wo.warp_perspective(img, H, (1024, 1024), 1)
    assert util.bits_equal(bev, ref_bev)
    assert util.bits_equal(back, wo.warp_perspective(ref_bev, H, (1920, 1080), 17))


def test_linearity_in_the_source(path):
    """Bilinear uint8 is linear up to rounding: warp(a) + warp(b) is within 1 LSB of warp(a + b)."""
    H = util.h_canon()
    a = util.seeded_frame(1, 1080, 1920, 3, "uint8") // 2
    b = util.seeded_frame(2, 1080, 1920, 3, "uint8") // 2
    wa = gpu_warp(a, H, (1024, 1024), 1).astype(np.int32)
    wb = gpu_warp(b, H, (1024, 1024), 1).astype(np.int32)
    wab = gpu_warp(a + b, H, (1024, 1024), 1).astype(np.int32)
    assert np.abs(wa + wb - wab).max() <= 1


def test_identity_and_shift_are_copies(path):
    src = util.seeded_frame(4, 300, 500, 3, "uint8")
    for flags in (0, 1):
        out = gpu_warp(src, np.eye(3), (500, 300), flags)
        assert util.bits_equal(out, src)
    T = np.array([[1.0, 0, 7], [0, 1.0, -3], [0, 0, 1.0]])
    out = gpu_warp(src, T, (500, 300), 0)
    assert util.bits_equal(out[:297, 7:], src[3:, :493])
    assert not out[297:].any() and not out[:, :7].any()  # BORDER_CONSTANT 0


def test_edge_shapes_and_empty(path):
    H = np.array([[1.0, 0, 0.5], [0, 1.0, 0.5], [0, 0, 1.0]])
    src = util.seeded_frame(1, 1, 1, 3, "uint8")
    assert util.bits_equal(gpu_warp(src, H, (3, 2), 1), wo.warp_perspective(src, H, (3, 2), 1))
    g = util.seeded_frame(2, 5, 7, 1, "uint8")[:, :, 0]
    out = gpu_warp(g, H, (1, 1), 0)
    assert out.shape == (1, 1) and util.bits_equal(out, wo.warp_perspective(g, H, (1, 1), 0))
    empty = torch.zeros((0, 16, 16, 3), dtype=torch.uint8, device=DEV)
    assert homo.warp_perspective(empty, np.eye(3), (8, 8)).shape == (0, 8, 8, 3)
    # degenerate matrix: cv2 maps everything to the zero matrix inverse -> source (0,0) everywhere
    src = util.seeded_frame(9, 20, 30, 3, "uint8")
    Z = np.zeros((3, 3))
    assert util.bits_equal(gpu_warp(src, Z, (16, 8), 1), wo.warp_perspective(src, Z, (16, 8), 1))
    # horizon crossing inside the output (w changes sign)
    Hh = np.array([[1.0, 0.2, -3.0], [0.1, 1.1, -2.0], [0.0, 0.03, -0.5]])
    src = util.seeded_frame(10, 96, 128, 3, "uint8")
    assert util.bits_equal(gpu_warp(src, Hh, (128, 96), 1), wo.warp_perspective(src, Hh, (128, 96), 1))


def test_argument_validation():
    t = torch.zeros(8, 8, 3, dtype=torch.uint8, device=DEV)
    with pytest.raises(TypeError):
        homo.warp_perspective(t.to(torch.int32), np.eye(3), (4, 4))
    with pytest.raises(_native.NativeError):
        homo.warp_perspective(t, np.eye(3), (4, 4), flags=2)       # INTER_CUBIC
    with pytest.raises(_native.NativeError):
        homo.warp_perspective(t, np.eye(3), (4, 4), borderMode=1)  # BORDER_REPLICATE
    with pytest.raises(ValueError):
        homo.warp_perspective(t, np.eye(4), (4, 4))
    dst = torch.empty(4, 4, 3, dtype=torch.uint8, device=DEV)
    assert homo.warp_perspective(t, np.eye(3), (4, 4), dst=dst) is dst


def test_host_buffer_entry_point():
    H = util.h_canon()
    S = np.diag([0.25, 0.25, 1.0])
    Hs = S @ H @ np.linalg.inv(S)
    frames = np.stack([util.seeded_frame(300 + i, 270, 480, 3, "uint8") for i in range(19)])
    out = _native.warp_perspective_host(frames, Hs, (256, 256), flags=1)
    for i in (0, 9, 18):
        assert util.bits_equal(out[i], wo.warp_perspective(frames[i], Hs, (256, 256), 1)), i
    pinned = torch.from_numpy(frames).pin_memory()
    dst = torch.empty((19, 256, 256, 3), dtype=torch.uint8).pin_memory()
    _native.warp_perspective_host(pinned, Hs, (256, 256), dst=dst, flags=0)
    assert util.bits_equal(dst.numpy()[5], wo.warp_perspective(frames[5], Hs, (256, 256), 0))
    # horizon-crossing map uploads the whole frame
    Hh = np.array([[1.0, 0.2, -3.0], [0.1, 1.1, -2.0], [0.0, 0.003, -0.5]])
    out = _native.warp_perspective_host(frames[:3], Hh, (200, 100), flags=1)
    assert util.bits_equal(out[2], wo.warp_perspective(frames[2], Hh, (200, 100), 1))


def test_touched_pixels_matches_oracle():
    H = util.h_canon()
    assert _native.warp_touched_pixels((1920, 1080), (1024, 1024), H, 1) == (971287, 420, 1058)
    assert _native.warp_touched_pixels((1920, 1080), (1024, 1024), H, 0) == \
        wo.touched_pixels((1920, 1080), (1024, 1024), H, 0)


def test_calibration_object_wrappers():
    from bev_b200 import BEVWorldSpec, Calib
    cam = util.load_json("cfg4_cams.json")[3]
    c = Calib(vp1=np.array(cam["vp1"]), vp2=np.array(cam["vp2"]), height=cam["height"],
              u_size=1920, v_size=1080)
    spec = {k: v for k, v in cam["bspec"].items() if v is not None and k not in ("x_max", "y_max")}
    b = BEVWorldSpec(**spec)
    frames = torch.from_numpy(np.stack([util.seeded_frame(1234 + 3, 1080, 1920, 3, "uint8")] * 2)).to(DEV)
    bev = homo.warp_img_to_bev(frames, c, b)
    assert tuple(bev.shape) == (2, b.v_size, b.u_size, 3)
    case = [h for h in util.hash_cases() if h["name"] == "cfg4_cam3_lin"][0]
    assert util.sha256(bev[0].cpu().numpy()) == case["sha256"]
    img = homo.warp_bev_to_img(bev, c, b)
    assert tuple(img.shape) == (2, 1080, 1920, 3)


@pytest.mark.parametrize("dtype", ["uint8", "float16"])
@pytest.mark.parametrize("flags", [0, 1])
def test_staged_kernel_box_shapes(flags, dtype):
    """Shapes that push the staged kernel (both pixel formats) through every staging mode: one
    tensor box, several boxes per frame (tall source boxes), boxes too wide / large for the ring
    (direct global loads inside the same kernel), partial tiles, and frame counts that do not fill
    the ring stages."""
    _native.set_warp_path("fast")
    try:
        rng = np.random.default_rng(5)
        frames = np.stack([util.seeded_frame(300 + i, 540, 960, 3, dtype) for i in range(7)])
        quad = np.array([[0, 0], [959, 0], [959, 539], [0, 539]], np.float64)
        cases = [
            ((256, 256), 0.0),    # ~3.7x / 2.1x minification: wide boxes, ~17 rows
            ((40, 36), 0.0),      # extreme minification: boxes exceed the menu -> global loads
            ((130, 520), 30.0),   # tall thin output, vertical magnification, partial tiles
            ((1000, 12), 10.0),   # wide flat output: strong vertical minification -> many rows
        ]
        for (dw, dh), jitter in cases:
            d = np.array([[0, 0], [dw - 1, 0], [dw - 1, dh - 1], [0, dh - 1]], np.float64)
            H = homo.homo_from_pts(quad + rng.normal(size=(4, 2)) * jitter, d)
            dw4 = (dw + 3) // 4 * 4
            out = gpu_warp(frames, H, (dw4, dh), flags)
            for i in (0, 3, 6):
                ref = wo.warp_perspective(frames[i], H, (dw4, dh), flags)
                assert util.bits_equal(out[i], ref), ((dw, dh), flags, i)
    finally:
        _native.set_warp_path("auto")


@pytest.mark.parametrize("dtype", ["uint8", "float16"])
def test_staged_kernel_long_runs_through_shallow_rings(dtype):
    """Many frames through the ring shapes that hold few of them: two-slot rings of single frames
    (boxes of a third of the ring: the slot just released is refilled while the other one is
    consumed) and frames of several tensor boxes, 29 frames each (a stage count that fills no ring
    evenly), every frame against the oracle."""
    _native.set_warp_path("fast")
    try:
        rng = np.random.default_rng(11)
        frames = np.stack([util.seeded_frame(900 + i, 540, 960, 3, dtype) for i in range(29)])
        quad = np.array([[0, 0], [959, 0], [959, 539], [0, 539]], np.float64)
        for (dw, dh), jitter in [((256, 256), 0.0), ((512, 120), 6.0), ((1000, 12), 10.0)]:
            d = np.array([[0, 0], [dw - 1, 0], [dw - 1, dh - 1], [0, dh - 1]], np.float64)
            H = homo.homo_from_pts(quad + rng.normal(size=(4, 2)) * jitter, d)
            out = gpu_warp(frames, H, (dw, dh), 1)
            for i in range(len(frames)):
                ref = wo.warp_perspective(frames[i], H, (dw, dh), 1)
                assert util.bits_equal(out[i], ref), ((dw, dh), i)
    finally:
        _native.set_warp_path("auto")


@pytest.mark.parametrize("flags", [1, 0])
def test_cfg2_full_batch(flags):
    """BASELINE configs[1] at full size: 256 x 1080p -> 1024^2 in one call.  Three frames are
    checked bit for bit against the oracle; the rest through a size-independent property -- the
    batch contains every frame twice (i and i + 128), so both halves of the output must agree,
    whatever chunk, tile or ring slot a frame went through."""
    H = util.h_canon()
    g = torch.Generator(device=DEV).manual_seed(7)
    half = torch.randint(0, 256, (128, 1080, 1920, 3), dtype=torch.uint8, device=DEV, generator=g)
    frames = torch.cat([half, half], 0)
    del half
    out = homo.warp_perspective(frames, H, (1024, 1024), flags=flags)
    assert tuple(out.shape) == (256, 1024, 1024, 3)
    assert torch.equal(out[:128], out[128:])
    for i in (0, 77, 255):
        ref = wo.warp_perspective(frames[i].cpu().numpy(), H, (1024, 1024), flags)
        assert util.bits_equal(out[i].cpu().numpy(), ref), i


@pytest.mark.parametrize("flags", [1, 0, 17])
def test_float16_staged_batch(flags):
    """float16 x 3 through the staged kernel (the fp16 leg of BASELINE configs[4], at 1/4 size):
    == float16(oracle(float32(src))) bit for bit, batch of 21 frames, two homographies."""
    _native.set_warp_path("fast")
    try:
        S = np.diag([0.5, 0.5, 1.0])
        H = S @ util.h_canon() @ np.linalg.inv(S)  # 960x540 -> 512x512
        H2 = np.array([[0.9, 0.08, 12.0], [-0.04, 1.05, 6.0], [2e-5, 1e-4, 1.0]])
        frames = np.stack([util.seeded_frame(800 + i, 540, 960, 3, "float16") for i in range(21)])
        idx = np.array([i % 2 for i in range(21)], np.int32)
        Hs = np.stack([H, H2])
        if flags & 16:
            Hs = np.stack([np.linalg.inv(H), np.linalg.inv(H2)])
        out = gpu_warp(frames, Hs, (512, 512), flags, mat_index=idx)
        assert out.dtype == np.float16
        for i in (0, 1, 10, 19, 20):
            ref = wo.warp_perspective(frames[i], Hs[idx[i]], (512, 512), flags=flags)
            assert util.bits_equal(out[i], ref), (flags, i)
    finally:
        _native.set_warp_path("auto")


def test_fuzz_paths_agree_on_random_homographies():
    """Differential fuzz: 60 random quad-to-quad homographies (strong perspective, partly outside
    the source, both interpolations, uint8 and float16, odd batch sizes) -- the automatic route
    (staged kernel, split launches, direct-gather) and the forced staged kernel must give the bytes
    of the generic one-thread-per-pixel kernels, and a sample of frames must match the oracle."""
    import cv2
    rng = np.random.default_rng(20261018)
    sizes = [(368, 640), (540, 960), (720, 1280), (1080, 1920)]
    try:
        for case in range(60):
            sh, sw = sizes[int(rng.integers(len(sizes)))]
            dw, dh = int(rng.integers(16, 200)) * 4, int(rng.integers(40, 700))
            n = int(rng.integers(4, 10))
            jit = float(rng.uniform(0.05, 0.45))
            s = np.float32([[0, 0], [sw, 0], [sw, sh], [0, sh]]) + (rng.uniform(-jit, jit, (4, 2)) * [sw, sh]).astype(np.float32)
            d = np.float32([[0, 0], [dw, 0], [dw, dh], [0, dh]]) + (rng.uniform(-jit, jit, (4, 2)) * [dw, dh]).astype(np.float32)
            H = cv2.getPerspectiveTransform(s, d).astype(np.float64)
            flags = int(rng.integers(0, 2)) | (16 if rng.random() < 0.3 else 0)
            if flags & 16:
                H = np.linalg.inv(H)
            dtype = "float16" if case % 5 == 4 else "uint8"
            frames = np.stack([util.seeded_frame(3000 + 10 * case + i, sh, sw, 3, dtype) for i in range(n)])
            t = torch.from_numpy(frames).to(DEV)
            outs = {}
            for pth in ("generic", "auto", "fast"):
                _native.set_warp_path(pth)
                try:
                    outs[pth] = homo.warp_perspective(t, H, (dw, dh), flags=flags).cpu().numpy()
                except RuntimeError as e:  # shape does not qualify for the forced staged path
                    assert pth == "fast" and "does not qualify" in str(e)
            for pth in outs:
                assert util.bits_equal(outs[pth], outs["generic"]), (case, pth, sh, sw, dw, dh, flags, dtype)
            if case % 6 == 0:
                ref = wo.warp_perspective(frames[n - 1], H, (dw, dh), flags=flags)
                assert util.bits_equal(outs["generic"][n - 1], ref), (case, "oracle")
    finally:
        _native.set_warp_path("auto")


@pytest.mark.parametrize("dtype", ["uint8", "float16"])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("flags", [1, 0])
def test_cfg5_full_batch(flags, inverse, dtype):
    """BASELINE configs[4] at full size: 64 frames, 4K -> 2048^2 and 2048^2 -> 4K, uint8 and float16,
    bilinear and nearest, one call each.  The batch holds every frame twice (i and i + 32): both
    halves of the output must agree whatever chunk, tile, ring slot or kernel path a frame took;
    three frames are compared bit for bit with the oracle (float16: float16(oracle(float32)))."""
    H = util.h_canon(2)
    ssize, dsize = (3840, 2160), (2048, 2048)
    if inverse:
        H, ssize, dsize = np.linalg.inv(H), dsize, ssize
    g = torch.Generator(device=DEV).manual_seed(11 + flags)
    half = torch.randint(0, 256, (32, ssize[1], ssize[0], 3), dtype=torch.uint8, device=DEV, generator=g)
    if dtype == "float16":
        half = (half.to(torch.float32) / 255.0).to(torch.float16)
    frames = torch.cat([half, half], 0)
    del half
    out = homo.warp_perspective(frames, H, dsize, flags=flags)
    assert tuple(out.shape) == (64, dsize[1], dsize[0], 3) and out.dtype == frames.dtype
    assert torch.equal(out[:32].view(torch.uint8), out[32:].view(torch.uint8))
    for i in (0, 21, 63):
        ref = wo.warp_perspective(frames[i].cpu().numpy(), H, dsize, flags)
        assert util.bits_equal(out[i].cpu().numpy(), ref), (flags, inverse, dtype, i)


def test_cfg4_shaped_call_eight_matrices_one_launch(path):
    """BASELINE configs[3] in one call: 1080p frames of eight camera streams interleaved in one
    batch, a table of eight homographies and a per-frame camera index (one launch serves several
    cameras).  The eight cameras of tests/golden/cfg4_cams.json stretched to one
    common 1024 x 1024 BEV so that they fit a single output tensor; every frame against the oracle."""
    import json
    import os
    cams = json.load(open(os.path.join(util.GOLDEN, "cfg4_cams.json")))
    Hs = []
    for c in cams:
        w, h = int(c["bspec"]["u_size"]), int(c["bspec"]["v_size"])
        S = np.diag([1024.0 / w, 1024.0 / h, 1.0])
        Hs.append(S @ np.array(c["H_bev_img"], np.float64))
    Hs = np.stack(Hs)
    n = 24
    idx = np.array([(5 * i) % 8 for i in range(n)], np.int32)  # cameras interleaved, 3 frames each
    frames = np.stack([util.seeded_frame(4000 + i, 1080, 1920, 3, "uint8") for i in range(n)])
    out = gpu_warp(frames, Hs, (1024, 1024), 1, mat_index=idx)
    for i in range(n):
        ref = wo.warp_perspective(frames[i], Hs[idx[i]], (1024, 1024), flags=1)
        assert util.bits_equal(out[i], ref), (path, i, int(idx[i]))


def test_concurrent_streams_and_threads():
    """Launches from several CUDA streams and two host threads at once: every launch owns its
    stream-ordered scratch (item counter, split flags, set-up records), so results must stay
    bit-identical to a serial run.  Camera homographies at 1024^2 give split launches (staged +
    direct-gather kernel), H_canon a plain staged one."""
    import json
    import os
    import threading
    cams = json.load(open(os.path.join(util.GOLDEN, "cfg4_cams.json")))
    c = cams[2]
    w, h = int(c["bspec"]["u_size"]), int(c["bspec"]["v_size"])
    H_cam = np.diag([1024.0 / w, 1024.0 / h, 1.0]) @ np.array(c["H_bev_img"], np.float64)
    jobs = [(util.h_canon(), 1, "auto"), (H_cam, 1, "auto"), (util.h_canon(), 0, "fast"), (H_cam, 1, "generic")]
    g = torch.Generator(device=DEV).manual_seed(3)
    frames = torch.randint(0, 256, (12, 1080, 1920, 3), dtype=torch.uint8, device=DEV, generator=g)
    serial = [homo.warp_perspective(frames, H, (1024, 1024), flags=f) for H, f, _ in jobs]
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(tid):
        try:
            streams = [torch.cuda.Stream(device=DEV) for _ in range(4)]
            outs = []
            for rep in range(6):
                for j, (H, f, pth) in enumerate(jobs):  # the kernel family is a per-call argument
                    with torch.cuda.stream(streams[(j + rep) % 4]):
                        outs.append((j, homo.warp_perspective(frames, H, (1024, 1024), flags=f, path=pth)))
            for st in streams:
                st.synchronize()
            results[tid] = outs
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for tid in results:
        for j, o in results[tid]:
            assert torch.equal(o, serial[j]), (tid, j)


@pytest.mark.parametrize("dtype,ch", [("uint8", 1), ("uint8", 4), ("float32", 3), ("uint8", 3), ("float16", 3)])
@pytest.mark.parametrize("flags", [1, 0, 17, 16])
def test_staged_kernel_pixel_formats(dtype, ch, flags):
    """Every pixel-format policy of the staged (TMA) kernel, forced: uint8 x 1 / x 3 / x 4, float16 x 3,
    float32 x 3, bilinear and nearest, forward and WARP_INVERSE_MAP, over maps that magnify, minify and
    leave the source (zero border): bit-identical to the oracle (float32 negative values included)."""
    rng = np.random.default_rng(17 * ch + flags)
    n = 6
    frames = np.stack([util.seeded_frame(5000 + 7 * ch + i, 270, 480, ch, dtype) for i in range(n)])
    if dtype == "float32":
        frames = frames * 2.0 - 1.0  # signed values: sums of products in cv2's order, signs of zero
        frames = frames.astype(np.float32)
    quad = np.array([[0, 0], [479, 0], [479, 269], [0, 269]], np.float64)
    cases = [((256, 256), 30.0), ((64, 48), 5.0), ((800, 600), 40.0), ((160, 400), 90.0)]
    for (dw, dh), jitter in cases:
        d = np.array([[0, 0], [dw - 1, 0], [dw - 1, dh - 1], [0, dh - 1]], np.float64)
        H = homo.homo_from_pts(quad + rng.normal(size=(4, 2)) * jitter, d)
        if flags & 16:
            H = np.linalg.inv(H)
        t = torch.from_numpy(frames).to(DEV)
        out = homo.warp_perspective(t, H, (dw, dh), flags=flags, path="fast").cpu().numpy()
        gen = homo.warp_perspective(t, H, (dw, dh), flags=flags, path="generic").cpu().numpy()
        assert util.bits_equal(out, gen), (dtype, ch, flags, dw, dh, "staged vs direct-gather kernels")
        for i in (0, n - 1):
            ref = wo.warp_perspective(frames[i] if ch > 1 else frames[i][:, :, 0], H, (dw, dh), flags=flags)
            got = out[i] if ch > 1 else out[i][:, :, 0]
            assert util.bits_equal(got, ref), (dtype, ch, flags, dw, dh, i)


def test_pair_path_orientations():
    """The shared-window pair path of the staged uint8 x 3 bilinear kernel has three row patterns
    (both pixels of a vertical pair on the same source rows, the lower dst pixel one source row
    further down, or -- maps that flip the image vertically -- one row further up) and per-pixel
    column offsets to either side.  Magnifying maps in all eight orientations (flips, transposes,
    rotations, with shear so that columns drift both ways) must stay bit-exact."""
    frames = np.stack([util.seeded_frame(6000 + i, 270, 480, 3, "uint8") for i in range(5)])
    t = torch.from_numpy(frames).to(DEV)
    W, Hh = 480.0, 270.0
    base = [
        np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1.0]]),                 # upright
        np.array([[1, 0, 0], [0, -1, Hh - 1], [0, 0, 1.0]]),           # vertical flip   -> row step -1
        np.array([[-1, 0, W - 1], [0, 1, 0], [0, 0, 1.0]]),            # horizontal flip -> column step -1
        np.array([[-1, 0, W - 1], [0, -1, Hh - 1], [0, 0, 1.0]]),      # rotation by 180 degrees
        np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1.0]]),                 # transpose
        np.array([[0, -1, Hh - 1], [1, 0, 0], [0, 0, 1.0]]),           # rotation by 90 degrees
    ]
    for bi, B in enumerate(base):
        for zoom, shear, persp in ((3.7, 0.35, 1.0), (1.6, -0.5, 1.0), (6.0, 0.0, 1.0), (4.3, 0.0, 0.0),
                                   (2.2, 0.0, 0.0)):
            # persp = 0, shear = 0: every dst row maps to ONE source row, so whole warps share a row
            # pattern (the warp-uniform variants of the pair path); the others mix patterns in a warp
            Z = np.array([[zoom * 0.9, shear, 3.0], [shear * 0.4, zoom, -5.0],
                          [1e-5 * persp, 2e-5 * persp, 1.0]])
            H = Z @ B
            dsize = (512, 384)
            out = homo.warp_perspective(t, H, dsize, flags=1, path="fast").cpu().numpy()
            for i in (0, 4):
                ref = wo.warp_perspective(frames[i], H, dsize, flags=1)
                assert util.bits_equal(out[i], ref), (bi, zoom, shear, persp, i)
