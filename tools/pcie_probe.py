#!/usr/bin/env python
"""tools/pcie_probe.py -- what the host link can do for the cfg-2 host-buffer step (development aid).

Copies the same byte counts as one `bevk_warp_perspective_host` step of cfg 2 (945 MB of source row
bands up, 805 MB of BEVs down) with plain pinned-memory copies on two streams: up alone, down
alone, and both directions at once.  The last figure is the ceiling of the `e2e` bench number."""
import torch

dev = torch.device("cuda", 0)
UP, DOWN = 945_000_000, 805_306_368
h_up = torch.empty(UP, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
d_up = torch.empty(UP, dtype=torch.uint8, device=dev)
d_dn = torch.empty(DOWN, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
print("H2D alone  %.2f ms  %.1f GB/s" % (t_up, UP / t_up / 1e6))
print("D2H alone  %.2f ms  %.1f GB/s" % (t_dn, DOWN / t_dn / 1e6))
print("both       %.2f ms  %.1f GB/s total -> ceiling %.1f Gpix/s for 256 x 1024^2 outputs"
      % (t_both, (UP + DOWN) / t_both / 1e6, 256 * 1024 * 1024 / t_both / 1e6))
