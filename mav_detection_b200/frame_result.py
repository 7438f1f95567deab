"""FrameResult — the per-frame output contract of Processor.run_detection.

Mirrors /root/reference/src/frame_result.py:4-17: a plain attribute bag whose twelve attributes are the JSON
keys Validator.load_results reads back (/root/reference/src/validator.py:141-152)."""
from __future__ import annotations

from typing import Tuple

KEYS = ('time', 'tpr', 'fpr', 'tpr_fixed', 'fpr_fixed', 'sky_tpr', 'sky_fpr', 'drone_size_pixels',
        'drone_flow_pixels', 'foe_dense', 'foe_gt', 'center_phi')


class FrameResult:
    def __init__(self) -> None:
        self.time = 0.0
        self.tpr = 0.0
        self.fpr = 0.0
        self.tpr_fixed = 0.0
        self.fpr_fixed = 0.0
        self.sky_tpr = 0.0
        self.sky_fpr = 0.0
        self.drone_size_pixels = 0.0
        self.drone_flow_pixels: Tuple[float, float] = (0.0, 0.0)
        self.foe_dense: Tuple[float, float] = (0.0, 0.0)
        self.foe_gt: Tuple[float, float] = (0.0, 0.0)
        self.center_phi = 0.0
