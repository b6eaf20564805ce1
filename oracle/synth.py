"""oracle/synth.py -- TEST INFRASTRUCTURE ONLY: seeded synthetic frames shared by
oracle/gen_golden.py (which hashes cv2's output on them) and tests/ (which regenerate them)."""
import numpy as np


def seeded_frame(seed, h, w, c, dtype):
    """i.i.d. uniform noise frame (h, w, c); float dtypes are uint8 noise / 255."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    if str(dtype) != "uint8":
        a = (a.astype(np.float32) / 255.0).astype(dtype)
    return a


def compo_inputs(seed, h, w):
    """Seeded background / foreground / soft mask (3-channel, like cv2.imread gives) of one size:
    the inputs of the compositing fixtures (tests/golden/compo_kat.npz)."""
    rng = np.random.default_rng(seed)
    bg = seeded_frame(seed, h, w, 3, "uint8")
    fg = seeded_frame(seed + 1, h, w, 3, "uint8")
    m = rng.integers(0, 256, (h, w, 1)).astype(np.uint8)
    m[: h // 4] = 0
    m[-(h // 4):] = 255
    return bg, fg, np.repeat(m, 3, axis=2)
