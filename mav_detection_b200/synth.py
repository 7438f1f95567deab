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
                  expansion: float = 0.01, with_rotation: bool = False, motion: str = 'zoom') -> SyntheticSequence:
    """motion 'zoom': expansion about the FoE (radial flow, sparse detection masks).  motion 'translate': the camera
    slides sideways (6 px per frame, uniform flow): the flow lines are parallel, the FoE estimate finds no consensus
    (or a far-away one) and most pixels end up above the angle thresholds — the dense-mask stress case."""
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
        if motion == 'translate':
            ph = t % 8
            sh = 6.0 * (ph if ph <= 4 else 8 - ph)       # triangle wave 0..24 px: stays inside the canvas margin
            mx = (xs + 4.0 + sh).astype(np.float32)
            my = (ys + 16.0).astype(np.float32)
        else:
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


class SyntheticDataset:
    """The subset of the reference's Dataset interface that Processor.run_detection touches
    (/root/reference/src/datasets/dataset.py:152-175,205-230,266-350, sim_data.py:56-81), backed by a
    SyntheticSequence held in memory.  Stands in for SimData in tests, the bench and the examples."""

    def __init__(self, seq: SyntheticSequence, sequence: str = 'synthetic', flows: np.ndarray = None,
                 results_path: str = None, bgr: bool = False, gt_flows: np.ndarray = None) -> None:
        self.seq = seq
        self.sequence = sequence
        self.N = int(seq.frames.shape[0])
        self.capture_size = (int(seq.frames.shape[2]), int(seq.frames.shape[1]))
        self.resolution = np.array(self.capture_size)
        self.results_path = results_path
        self.flows = flows            # optional precomputed (N-1, H, W, 2) float32: the get_flow_uv seam
        self.gt_flows = gt_flows      # optional (N-1, H, W, 2) float32 ground-truth flow: the get_gt_of seam
        self.bgr = bgr
        self._next = 0

    def get_frame(self) -> np.ndarray:
        f = self.seq.frames[min(self._next, self.N - 1)]
        self._next += 1
        return np.repeat(f[..., None], 3, axis=2) if self.bgr else f

    def get_flow_uv(self, i: int) -> np.ndarray:
        if self.flows is None:
            raise ValueError('Could not load flow field.')
        return self.flows[i]

    def get_gt_of(self, i: int):
        return None if self.gt_flows is None else self.gt_flows[i]

    def get_sky_segmentation(self, i: int) -> np.ndarray:
        return self.seq.sky_mask

    def get_segmentation(self, i: int) -> np.ndarray:
        s = self.seq.segmentation[i]
        return np.repeat(s[..., None], 3, axis=2)

    def get_depth(self, i: int):
        return None

    def get_gt_foe(self, i: int):
        return self.seq.foe

    def get_time(self, i: int) -> float:
        return i * self.seq.dt

    def get_delta_time(self, i: int) -> float:
        return self.seq.dt

    def get_angular_difference(self, first: int, second: int) -> np.ndarray:
        return self.seq.omega[second].copy()

    def release(self) -> None:
        pass
