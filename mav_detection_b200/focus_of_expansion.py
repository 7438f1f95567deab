"""FocusOfExpansion with the reference's API (/root/reference/src/focus_of_expansion.py:13-184) on the
CUDA path.  get_FOE_dense draws its 2 x 2000 sample indices from the process-global legacy NumPy
generator exactly as the reference does (:69-71: rows first, then columns) and hands them to the device
kernel; ransac and get_phi run on the device as well.  The sparse (Lucas-Kanade) estimator and the
drawing helpers are out of scope."""
from __future__ import annotations

from typing import Any, Tuple

import numpy as np

from ._lib import N_SAMPLE_PAIRS


class FocusOfExpansion:
    def __init__(self, lucas_kanade: Any, engine: Any = None) -> None:
        self.lucas_kanade = lucas_kanade
        self.time = 0
        self.roll_back = 20
        self.num_features = 0
        self.enable_plots = False
        self.max_flow = 0.0  # maximum phi in the image (degrees)
        self.radial_threshold = np.cos(np.deg2rad(15))
        self.magnitude_threshold = 2.5
        self.ransac_threshold = 30.0  # pixels
        n = self.lucas_kanade.total_num_corners
        # the reference's constructor draws (focus_of_expansion.py:24,26) — kept for the random stream
        self.color = np.random.randint(0, 255, (n, 3))
        self.random_lines = np.random.randint(0, n, n)
        self.flow_height, self.flow_width = self.lucas_kanade.old_frame.shape[0], self.lucas_kanade.old_frame.shape[1]
        self._engine = engine

    def _eng(self):
        if self._engine is None:
            from . import engine
            self._engine = engine.shared_engine(self.flow_width, self.flow_height)
        eng = self._engine
        eng.detect_params.magnitude_threshold = float(self.magnitude_threshold)
        eng.detect_params.ransac_threshold = float(self.ransac_threshold)
        return eng

    @staticmethod
    def draw_sample_indices(height: int, width: int) -> np.ndarray:
        """The two np.random.randint draws of focus_of_expansion.py:69-71 as one int32 [ry(2000) | rx(2000)] row."""
        ry = np.random.randint(0, height, N_SAMPLE_PAIRS * 2)
        rx = np.random.randint(0, width, N_SAMPLE_PAIRS * 2)
        return np.concatenate([ry, rx]).astype(np.int32)

    def ransac(self, estimates: np.ndarray) -> Tuple[float, float]:
        """focus_of_expansion.py:32-54."""
        import torch
        eng = self._eng()
        est = np.ascontiguousarray(estimates, dtype=np.float64).reshape(-1, 2)
        out = eng.ransac(torch.from_numpy(est).to(eng.device)).cpu().numpy()
        if out[0] == 0.0 and out[1] == 0.0:
            return (0.0, 0.0)
        return (out[0], out[1])

    def get_FOE_dense(self, flow_uv: np.ndarray) -> Tuple[float, float]:
        """focus_of_expansion.py:56-86 — (x, y) in pixels; (0.0, 0.0) means no consensus."""
        import torch
        eng = self._eng()
        flow = np.ascontiguousarray(flow_uv)
        if flow.dtype not in (np.float32, np.float64):
            flow = flow.astype(np.float64)
        samples = self.draw_sample_indices(flow.shape[0], flow.shape[1])
        foe, _ = eng.foe_dense(torch.from_numpy(flow[None]).to(eng.device), torch.from_numpy(samples[None]).to(eng.device))
        out = foe[0].cpu().numpy()
        if out[0] == 0.0 and out[1] == 0.0:
            return (0.0, 0.0)
        return (out[0], out[1])

    def get_phi(self, derotated_flow_uv: np.ndarray, FoE: Tuple[float, float]) -> np.ndarray:
        """focus_of_expansion.py:150-184 — phi per pixel in degrees, in the flow's dtype; sets max_flow."""
        if FoE[0] is np.nan:
            return np.zeros(0)
        import torch
        eng = self._eng()
        flow = np.ascontiguousarray(derotated_flow_uv)
        if flow.dtype not in (np.float32, np.float64):
            flow = flow.astype(np.float64)
        foe = torch.tensor([[float(FoE[0]), float(FoE[1])]], dtype=torch.float64)
        phi, mx = eng.get_phi(torch.from_numpy(flow[None]).to(eng.device), foe)
        out = phi[0].cpu().numpy()
        self.max_flow = out.dtype.type(mx[0].item())
        return out
