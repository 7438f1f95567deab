"""Synthetic AirSim-like sequences — input generation only, no hot-path compute.

A smooth random texture is viewed under a 1 %/frame zoom about a fixed focus of
expansion, with one small independently moving textured blob (the "MAV").  The
generator runs on the host; frames are handed to the CUDA path as uint8 arrays.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np


@dataclass
class SyntheticSequence:
    frames: np.ndarray        # (F, H, W) uint8 gray
    segmentation: np.ndarray  # (F, H, W) uint8, 255 on the blob
    sky_mask: np.ndarray      # (H, W) bool, all False
    foe: Tuple[float, float]  # ground-truth FoE in pixels (x, y)
    omega: np.ndarray         # (F, 3) float64 IMU angle deltas per frame [rad]
    dt: float                 # seconds between frames


def _bicubic_upsample8(lo: np.ndarray) -> np.ndarray:
    import cv2  # host-side input synthesis only
    return cv2.resize(lo, None, fx=8, fy=8, interpolation=cv2.INTER_CUBIC)


def make_sequence(width: int, height: int, n_frames: int, seq: int = 0,
                  expansion: float = 0.01, with_rotation: bool = False) -> SyntheticSequence:
    import cv2
    seed = 1000 + seq
    rng = np.random.default_rng(seed)
    lo = rng.random((height // 8 + 8, width // 8 + 8), dtype=np.float32)
    canvas = _bicubic_upsample8(lo)[:height + 32, :width + 32]
    blob_lo = rng.random((3 + 2, 5 + 2), dtype=np.float32)
    blob_tex = _bicubic_upsample8(blob_lo)[:24, :40]
    foe = (0.4 * width, 0.45 * height)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float32)
    frames = np.empty((n_frames, height, width), np.uint8)
    seg = np.zeros((n_frames, height, width), np.uint8)
    bx0, by0 = 0.6 * width, 0.55 * height
    for t in range(n_frames):
        z = (1.0 + expansion) ** t
        mx = (foe[0] + (xs - foe[0]) / z + 16).astype(np.float32)
        my = (foe[1] + (ys - foe[1]) / z + 16).astype(np.float32)
        img = cv2.remap(canvas, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        bx = int(round(bx0 + 3 * t)) % max(width - 40, 1)
        by = int(round(by0 - 2 * t)) % max(height - 24, 1)
        img[by:by + 24, bx:bx + 40] = 0.5 * img[by:by + 24, bx:bx + 40] + 0.5 * blob_tex
        seg[t, by:by + 24, bx:bx + 40] = 255
        frames[t] = np.clip(img * 255.0, 0, 255).astype(np.uint8)
    omega = np.zeros((n_frames, 3), np.float64)
    if with_rotation:
        omega[:] = np.array([0.002, -0.001, 0.0005])
    return SyntheticSequence(frames, seg, np.zeros((height, width), bool), foe, omega, 1.0 / 30.0)


def make_pair(width: int, height: int, seq: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    s = make_sequence(width, height, 2, seq)
    return s.frames[0], s.frames[1]
