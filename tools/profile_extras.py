#!/usr/bin/env python
"""tools/profile_extras.py -- one launch of each secondary kernel at a representative size, for
`ncu -k regex:...` captures (profiles/README.md).  Not a benchmark."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bev_b200 import compo, homo, rbox_torch  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n = 32
B = torch.randint(0, 256, (1, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
F = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
M = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device=dev, generator=g)
H = bench.h_canon(1)
for _ in range(2):
    out = compo.composite_bev_batch(B, F, M, H, H, (1024, 1024), fused=True)      # composite_bev_kernel
    blend = compo.composite_reg_img(F, M, B.expand(n, -1, -1, -1).contiguous())  # composite_kernel
    small = homo.resize(F, (852, 480))                                            # resize_u8c3_kernel
    u = torch.rand((2, 4096, 5), device=dev, generator=g)
    lo = torch.tensor([0.0, 0.0, 1.0, 1.0, -3.2], device=dev)
    hi = torch.tensor([300.0, 300.0, 30.0, 30.0, 3.2], device=dev)
    iou = rbox_torch.iou_batch_rbox((lo + u[0] * (hi - lo)).contiguous(), (lo + u[1] * (hi - lo)).contiguous())
torch.cuda.synchronize()
print("ok", tuple(out.shape), tuple(blend.shape), tuple(small.shape), tuple(iou.shape))
