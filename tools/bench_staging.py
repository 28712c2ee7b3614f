"""Throughput of the input-staging kernel (crop + cv2-exact bilinear resize), next to cv2.resize on the host cores."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hse_facerec_tf_b200 as hfr  # noqa: E402
import cv2  # noqa: E402

rs = np.random.RandomState(0)
frames = torch.from_numpy(rs.randint(0, 256, (8, 1080, 1920, 3)).astype(np.uint8)).cuda()
n = 4096
boxes = []
for _ in range(n):
    x1, y1 = rs.randint(0, 1500), rs.randint(0, 700)
    s = rs.randint(60, 360)
    boxes.append([x1, y1, x1 + s, y1 + s])
fidx = rs.randint(0, 8, n)
for _ in range(3):
    out = hfr.crop_resize(frames, boxes, 224, frame_index=fidx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = hfr.crop_resize(frames, boxes, 224, frame_index=fidx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fr = frames.cpu().numpy()
t0 = time.perf_counter()
for b, f in list(zip(boxes, fidx))[:512]:
    cv2.resize(fr[f][b[1]:b[3], b[0]:b[2]], (224, 224))
cpu = 512 / (time.perf_counter() - t0)
out_bytes = n * 224 * 224 * 3
print(json.dumps(dict(metric="face crops/sec (crop + INTER_LINEAR resize to 224x224)", value=round(n / (ms * 1e-3)), unit="crops/s",
                      ms_per_batch=round(ms, 4), batch=n, out_gbs=round(out_bytes / (ms * 1e-3) / 1e9, 1),
                      cpu_baseline=dict(value=round(cpu), unit="crops/s", kind="reference", cores=1,
                                        sample="cv2.resize on 512 of the crops, one thread (the reference's per-face call)"))))

# ---- the file path's resize: scipy.misc.imresize(..., 'bilinear') = Pillow BILINEAR (facerec_test.py:84,93)
from PIL import Image  # noqa: E402

m = 512
imgs = torch.from_numpy(rs.randint(0, 256, (m, 250, 250, 3)).astype(np.uint8)).cuda()     # LFW-sized inputs, resident
for _ in range(3):
    out = hfr.resize_pil(imgs, 192)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    out = hfr.resize_pil(imgs, 192)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
hi = imgs[:256].cpu().numpy()
t0 = time.perf_counter()
for a in hi:
    np.asarray(Image.fromarray(a).resize((192, 192), resample=Image.BILINEAR))
cpu = len(hi) / (time.perf_counter() - t0)
print(json.dumps(dict(metric="images/sec (Pillow-exact bilinear resize 250x250 -> 192x192)", value=round(m / (ms * 1e-3)),
                      unit="images/s", ms_per_batch=round(ms, 4), batch=m,
                      cpu_baseline=dict(value=round(cpu), unit="images/s", kind="reference", cores=1,
                                        sample="PIL.Image.resize on 256 of the images, one thread"))))
