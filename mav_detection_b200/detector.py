"""Detector with the reference's API for the hot path (/root/reference/src/detector.py).

Kept: the Algorithm enum (:15-21, 1-tuple values as in the reference), the constructor's state and its
draws from the global legacy NumPy generator (:33-36, so the random stream of get_FOE_dense is the
reference's), derotate (:70-117, on the device through mavd_derotate) and is_homography_based
(:430-433).  The homography / sliding-window branch (:119-428) is unreachable with the default
Algorithm.ESSENTIAL and is out of scope."""
from __future__ import annotations

from enum import Enum
from typing import Any, Optional

import numpy as np

from .lucas_kanade import LucasKanade


class Detector:
    class Algorithm(Enum):
        NONE = 0,
        FOE = 1,
        AFFINE = 2,
        HOMOGRAPHY = 3,
        FUNDAMENTAL = 4,
        ESSENTIAL = 5,

    def __init__(self, dataset: Any, algorithm: 'Detector.Algorithm' = None, use_sparse_of: bool = False,
                 engine: Any = None) -> None:
        self.dataset = dataset
        self.algorithm = Detector.Algorithm.ESSENTIAL if algorithm is None else algorithm
        self.use_sparse_of = use_sparse_of
        flow_width, flow_height = self.dataset.capture_size[0], self.dataset.capture_size[1]
        self.sample_size = 1000
        self.border_offset = 20
        self.sample_y = np.random.randint(self.border_offset, flow_height - self.border_offset, self.sample_size)
        self.sample_x = np.random.randint(self.border_offset, flow_width - self.border_offset, self.sample_size)
        self.coords = np.column_stack((self.sample_x, self.sample_y))
        self.history_length = 20
        self.history_index = 0
        self.confidence: int = 0
        self.use_optimization = False
        self.prev_frame = np.zeros((flow_height, flow_width, 3), dtype=np.uint8)
        self.lucas_kanade = LucasKanade(self.prev_frame)
        self.fov = 90  # degrees
        self.focal_length = 1 / np.tan(np.deg2rad(self.fov) / 2)
        self._engine = engine

    def _eng(self):
        if self._engine is None:
            from . import engine
            self._engine = engine.shared_engine(self.dataset.capture_size[0], self.dataset.capture_size[1])
        return self._engine

    def imu_for(self, previous_frame_index: int, current_frame_index: int):
        """(ang, dt, derotate) — the IMU inputs derotate() reads from the dataset (:83-88)."""
        if current_frame_index < 1:
            return np.zeros(3), 1.0, False
        dt = self.dataset.get_delta_time(current_frame_index)
        ang = np.asarray(self.dataset.get_angular_difference(previous_frame_index, current_frame_index), np.float64)
        return ang, float(dt), True

    def derotate(self, previous_frame_index: int, current_frame_index: int, flow_uv: np.ndarray) -> np.ndarray:
        """Derotate the flow field according to IMU data — detector.py:70-117.  float32 (H, W, 2) in,
        float64 out; frame index < 1 passes the input through untouched."""
        if current_frame_index < 1:
            return flow_uv
        import torch
        from . import engine
        ang, dt, _ = self.imu_for(previous_frame_index, current_frame_index)
        eng = self._eng()
        flow = np.ascontiguousarray(flow_uv)
        if flow.dtype not in (np.float32, np.float64):
            flow = flow.astype(np.float64)
        out = eng.derotate(torch.from_numpy(flow[None]).to(eng.device), engine.make_imu(1, ang[None], dt, derotate=True))
        return out[0].cpu().numpy()

    def is_homography_based(self) -> bool:
        return self.algorithm in [Detector.Algorithm.HOMOGRAPHY]
