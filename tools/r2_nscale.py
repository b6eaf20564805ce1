"""Development aid: cfg-2 warp time against the number of frames in the batch (8 ... 256): the marginal cost of
a frame and the part of a launch that does not scale with frames."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bev_b200 import homo, _native
dev = torch.device("cuda", 0)
H = bench.h_canon(1)
g = torch.Generator(device=dev).manual_seed(1234)
frames = torch.randint(0, 256, (256, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
out = torch.empty((256, 1024, 1024, 3), dtype=torch.uint8, device=dev)
for n in (8, 16, 32, 64, 128, 256, 8, 16, 32, 64, 128, 256):
    f, o = frames[:n], out[:n]
    for _ in range(3): homo.warp_perspective(f, H, (1024, 1024), dst=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): homo.warp_perspective(f, H, (1024, 1024), dst=o)
    e1.record(); torch.cuda.synchronize()
    print("n=%d %.4f ms  (%.4f us/frame)" % (n, e0.elapsed_time(e1) / 30, 1e3 * e0.elapsed_time(e1) / 30 / n), flush=True)
