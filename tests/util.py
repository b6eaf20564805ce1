"""Shared helpers for the test-suite (tests may import oracle/; the product never does)."""
import hashlib
import json
import os

import numpy as np

from oracle.synth import seeded_frame  # noqa: F401  (re-exported)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_cases():
    """Yield the cv2 fixtures of warp_small.npz as dicts (src regenerated from its seed)."""
    d = load_npz("warp_small.npz")
    for i in range(int(d["n"])):
        k = "c%03d" % i
        seed, h, w, ch = [int(v) for v in d[k + "_seed"]]
        dtype = str(d[k + "_dtype"])
        src = seeded_frame(seed, h, w, ch, dtype)
        if ch == 1:
            src = src[:, :, 0]
        yield {"name": k, "src": src, "H": d[k + "_H"], "flags": int(d[k + "_flags"]),
               "bv": d[k + "_bv"], "dst": d[k + "_dst"]}


def hash_cases():
    return load_json("warp_hash.json")["cases"]


def hash_case_input(case):
    h, w, c = case["shape"]
    src = seeded_frame(case["seed"], h, w, c, case["dtype"])
    return src[:, :, 0] if c == 1 else src


def h_canon(scale=1):
    kat = load_json("homo_kat.json")
    return np.array(kat["h_canon" if scale == 1 else "h_canon_4k"], dtype=np.float64)


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes()


def rel_err(out, ref):
    """SURVEY.md 8c projection metric: |out - ref| / max(|ref|, 1), element-wise max."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0))) if ref.size else 0.0


def yaw_err(out, ref):
    d = np.asarray(out, dtype=np.float64) - np.asarray(ref, dtype=np.float64)
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return float(np.max(np.abs(d))) if d.size else 0.0
