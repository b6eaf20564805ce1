#!/usr/bin/env python
"""tools/kbench.py -- kernel-only timing of the warp workloads (development aid, not the bench).

    python tools/kbench.py [--steps K] [--path auto|generic|fast] [workload ...]

Prints: workload, ms per launch (CUDA events), Mpix/s, fraction of the measured HBM peak
(algorithmic bytes of SURVEY.md 8d)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bev_b200 import _native, homo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--check", action="store_true", help="compare frame 0 and the last frame with the oracle")
    ap.add_argument("workloads", nargs="*")
    args = ap.parse_args()
    wls = args.workloads or ["cfg2_1080p_to_bev1024_u8c3_bilinear_x256", "cfg2_nearest",
                             "cfg5_4k_to_bev2048_u8c3_x64", "cfg5_inv_bev2048_to_4k_u8c3_x64"]
    _native.set_warp_path(args.path)
    peak, _ = bench.measured_peak()
    dev = torch.device("cuda", 0)
    for wl in wls:
        n, ssize, dsize, ch, dtype, flags, hscale, inverse = bench.WORKLOADS[wl]
        H = bench.h_canon(hscale)
        if inverse:
            H = np.linalg.inv(H)
        tdtype = {"uint8": torch.uint8, "float16": torch.float16, "float32": torch.float32}[dtype]
        es = {"uint8": 1, "float16": 2, "float32": 4}[dtype]
        g = torch.Generator(device=dev).manual_seed(1234)
        frames = torch.randint(0, 256, (n, ssize[1], ssize[0], ch), dtype=torch.uint8, device=dev, generator=g)
        if tdtype != torch.uint8:
            frames = (frames.to(torch.float32) / 255.0).to(tdtype)
        out = torch.empty((n, dsize[1], dsize[0], ch), dtype=tdtype, device=dev)
        T, _, _ = _native.warp_touched_pixels(ssize, dsize, H, flags)
        algo = (T + dsize[0] * dsize[1]) * ch * es * n
        for _ in range(3):
            homo.warp_perspective(frames, H, dsize, dst=out, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            homo.warp_perspective(frames, H, dsize, dst=out, flags=flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        msg = "%s %.4f ms %.0f Mpix/s frac %.4f" % (wl, ms, n * dsize[0] * dsize[1] / ms / 1e3,
                                                   algo / (ms * 1e-3) / 1e9 / peak)
        if args.check:
            from oracle import warp_oracle as wo
            bad = 0
            for i in (0, n - 1):
                src = frames[i].cpu().numpy()
                if tdtype == torch.float16:
                    ref = wo.warp_perspective(src, H, dsize, flags=flags)
                else:
                    ref = wo.warp_perspective(src, H, dsize, flags=flags)
                bad += int(np.count_nonzero(out[i].cpu().numpy() != ref))
            msg += " mismatches %d" % bad
        print(msg, flush=True)
        del frames, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
