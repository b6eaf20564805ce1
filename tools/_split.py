import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bev_b200 import _native, homo
dev = torch.device("cuda", 0)
cams = json.load(open("tests/golden/cfg4_cams.json"))
g = torch.Generator(device=dev).manual_seed(1234)
n = 256
frames = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
def t(H, dsize, path):
    _native.set_warp_path(path)
    out = torch.empty((n, dsize[1], dsize[0], 3), dtype=torch.uint8, device=dev)
    try:
        for _ in range(2): homo.warp_perspective(frames, H, dsize, dst=out)
    except Exception as e:
        return float("nan")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): homo.warp_perspective(frames, H, dsize, dst=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
for c in cams:
    u, v = int(c["bspec"]["u_size"]), int(c["bspec"]["v_size"])
    H = np.diag([1024.0 / u, 1024.0 / v, 1.0]) @ np.array(c["H_bev_img"])
    res = []
    for cut in (896, 960):
        top = t(H, (1024, cut), "fast")
        Tm = np.array([[1, 0, 0], [0, 1, -cut], [0, 0, 1.0]])
        bot = t(Tm @ H, (1024, 1024 - cut), "generic")
        res.append("cut %d: staged top %.3f + direct bottom %.3f = %.3f" % (cut, top, bot, top + bot))
    full = t(H, (1024, 1024), "generic")
    print("cam %s: direct all %.3f | %s" % (c["id"], full, " | ".join(res)))
