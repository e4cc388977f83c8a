"""Generates tests/golden/color_edges.npz with the REAL third-party code paths of the reference's inference
edges (cv2.cvtColor, torchvision v2.ToImage/ToDtype) in this container.  Run from the repo root:
    python oracle/make_golden_color.py"""
import os

import cv2
import numpy as np
import torch
from torchvision.transforms import v2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(2026)
B, H, W = 2, 37, 53
bgr = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
bgr[0, :16, :16] = np.stack(list(np.meshgrid(np.arange(0, 256, 16), np.arange(0, 256, 16), indexing="ij")) + [np.zeros((16, 16))], -1)
bgr[1, 0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255]]
tf = v2.Compose([v2.ToImage(), v2.ToDtype(dtype=torch.float32, scale=True)])          # a015:56-60
vis = torch.stack([tf(cv2.cvtColor(src=bgr[i], code=cv2.COLOR_BGR2YCrCb)) for i in range(B)])   # a015:86-93
y, crcb = vis[:, 0:1].numpy(), vis[:, 1:3].numpy()                                      # a017:68
fus_y = (rng.random((B, 1, H, W), dtype=np.float32) * 1.4 - 0.2).astype(np.float32)     # outside [0,1] too: clamp
rgb = []
for i in range(B):                                                                      # a017:83-88
    f = torch.clamp_(torch.from_numpy(fus_y[i:i + 1].copy()), min=0, max=1)
    ycc = torch.concat([f, torch.from_numpy(crcb[i:i + 1])], dim=1).squeeze(0).permute(1, 2, 0).numpy()
    rgb.append(torch.from_numpy(cv2.cvtColor(src=ycc, code=cv2.COLOR_YCrCb2RGB)).permute(2, 0, 1).numpy())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "color_edges.npz"), bgr=bgr, y=y, crcb=crcb, fus_y=fus_y,
                    rgb=np.stack(rgb), cv2_version=np.array(cv2.__version__))
print("wrote color_edges.npz", bgr.shape, y.shape, crcb.shape, np.stack(rgb).shape)
