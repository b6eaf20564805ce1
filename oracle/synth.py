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
