"""oracle/ref_cpu.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's own CPU path for the warp, timed by bench.py (``cpu_baseline`` and
``--impl reference``): the reference implements the warp as a Python loop that calls
``cv2.warpPerspective`` once per frame with default flags (/root/reference/vis_homo.py:85-89), so
that loop over in-memory frames -- cv2 using every host thread it wants -- IS the reference arm
(kind "reference").  If cv2 cannot be imported the C oracle port is timed instead, threaded over
frames from Python (kind "port").  The projection baseline is the reference's numpy float64 chain
(bev/rbox.py:65,136,50), restated in oracle/rbox_oracle.py.

Never imported by bev_b200/.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def make_warp_runner(threads=None):
    """Returns (run(frames, H, dsize, flags) -> list of outputs, kind, cores_used)."""
    threads = threads or host_cores()
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        cv2.setNumThreads(threads)
        used = cv2.getNumThreads()

        def run(frames, H, dsize, flags):
            out = []
            for f in frames:  # the structure of vis_homo.py:85-89
                out.append(cv2.warpPerspective(f, H, dsize, flags=flags))
            return out
        return run, "reference", used

    from oracle import warp_oracle
    pool = ThreadPoolExecutor(threads)

    def run(frames, H, dsize, flags):
        return list(pool.map(lambda f: warp_oracle.warp_perspective(f, H, dsize, flags), frames))
    return run, "port", threads


def time_warp(frames, H, dsize, flags, min_seconds=10.0, max_seconds=30.0, threads=None):
    """Time the reference loop over `frames` repeatedly for about `min_seconds`.
    Returns dict(value=Mpix/s of output pixels, ms_per_frame, kind, cores, sample)."""
    run, kind, used = make_warp_runner(threads)
    run(frames[:2], H, dsize, flags)  # warm the thread pool / page in
    n_frames, t_total, best = 0, 0.0, None
    while t_total < min_seconds:
        t0 = time.perf_counter()
        run(frames, H, dsize, flags)
        dt = time.perf_counter() - t0
        t_total += dt
        n_frames += len(frames)
        best = dt if best is None else min(best, dt)
        if t_total > max_seconds:
            break
    px = dsize[0] * dsize[1]
    return {"value": n_frames * px / t_total / 1e6, "best_pass_value": len(frames) * px / best / 1e6,
            "ms_per_frame": 1e3 * t_total / n_frames, "kind": kind, "cores": used,
            "sample": "%d passes over %d frames (%.1f s)" % (n_frames // len(frames), len(frames), t_total),
            "cpu": cpu_model()}


def time_rbox_chain(box, H_fwd, H_back, mode="bev", min_seconds=5.0):
    """numpy float64 chain xywhr -> img corners -> back to xywhr (1 thread, as the reference)."""
    from oracle import rbox_oracle as ro
    n, t_total = 0, 0.0
    while t_total < min_seconds:
        t0 = time.perf_counter()
        img = ro.xywhr_to_img_corners(box, H_fwd, mode)
        ro.img_corners_to_xywhr(img.astype(np.float32), H_back, mode)
        t_total += time.perf_counter() - t0
        n += 2 * len(box)
    return {"value": n / t_total / 1e6, "unit": "Mproj/s", "kind": "port", "cores": 1,
            "sample": "%d boxes fwd+back, %.1f s" % (len(box), t_total), "cpu": cpu_model()}
