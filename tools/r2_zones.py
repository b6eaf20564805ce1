"""Development aid: time the staged warp kernel on horizontal bands of the cfg-2 BEV (same
homography, translated) to see which zone of the map costs what."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bev_b200 import homo, _native
if os.environ.get("ZLIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["ZLIB"])

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1234)
frames = torch.randint(0, 256, (256, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
H = bench.h_canon(1)
bands = [(768, 1024), (0, 1024)] if os.environ.get("ZB") else [(0, 256), (256, 512), (512, 768), (768, 1024), (0, 512), (512, 1024), (0, 1024)]
for y0, y1 in bands:
    T = np.array([[1, 0, 0], [0, 1, -y0], [0, 0, 1]], np.float64)
    Hb = T @ H
    dsize = (1024, y1 - y0)
    out = torch.empty((256, dsize[1], dsize[0], 3), dtype=torch.uint8, device=dev)
    for _ in range(3):
        homo.warp_perspective(frames, Hb, dsize, dst=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        homo.warp_perspective(frames, Hb, dsize, dst=out)
    e1.record()
    torch.cuda.synchronize()
    print("rows %4d-%4d: %.4f ms" % (y0, y1, e0.elapsed_time(e1) / 20), flush=True)
