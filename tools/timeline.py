#!/usr/bin/env python
"""Prints the device timeline (CUDA events) of one hot-path step: which launch groups ran when, on the two streams of
mavd_farneback.  Usage: python tools/timeline.py [pairs]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mav_detection_b200 import engine, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = 1920, 1080
seq = synth.make_sequence(W, H, B + 1, seq=0)
eng = engine.Engine(W, H, engine.SAMPLE_PARAMS, max_pairs=B)
frames = torch.from_numpy(seq.frames).cuda()
seg = torch.from_numpy(seq.segmentation[1:].copy()).cuda()
rs = np.random.RandomState(1)
samples = torch.from_numpy(np.stack([np.concatenate([rs.randint(0, H, 2000), rs.randint(0, W, 2000)]) for _ in range(B)])
                           .astype(np.int32)).cuda()
imu = engine.make_imu(B, derotate=True)
fixed = torch.empty((B, H, W), dtype=torch.uint8, device='cuda')
for _ in range(3):
    eng.process(frames, imu, samples, seg=seg, fixed_out=fixed)
torch.cuda.synchronize()
eng.profile_enable(True)
eng.process(frames, imu, samples, seg=seg, fixed_out=fixed)
torch.cuda.synchronize()
tl = eng.profile_timeline()
end = max(t[2] for t in tl)
print('step %.3f ms, %d launch groups' % (end, len(tl)))
for name, t0, t1 in tl:
    print('%-15s %8.3f -> %8.3f  (%.3f ms)' % (name, t0, t1, t1 - t0))
