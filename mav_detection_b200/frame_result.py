"""FrameResult — the per-frame output contract of Processor.run_detection.

Same attribute set as /root/reference/src/frame_result.py:4-17: a plain attribute bag whose twelve attributes are the
JSON keys Validator.load_results reads back (/root/reference/src/validator.py:141-152).  The attribute table below is
the single source of truth for the writer (processor.py), the tests and the JSON emitters."""
from __future__ import annotations

from typing import Any, Dict, Tuple

# attribute -> value before the first frame has been processed (scalars 0.0, points (0.0, 0.0))
DEFAULTS: Dict[str, Any] = {
    'time': 0.0,
    'tpr': 0.0, 'fpr': 0.0,                  # dynamic-threshold mask vs ground truth
    'tpr_fixed': 0.0, 'fpr_fixed': 0.0,      # fixed 15-degree mask vs ground truth
    'sky_tpr': 0.0, 'sky_fpr': 0.0,          # sky segmentation vs depth
    'drone_size_pixels': 0.0,
    'drone_flow_pixels': (0.0, 0.0),
    'foe_dense': (0.0, 0.0),
    'foe_gt': (0.0, 0.0),
    'center_phi': 0.0,
}
KEYS: Tuple[str, ...] = tuple(DEFAULTS)


class FrameResult:
    """Attribute bag; `vars(result)` is exactly what utils.get_json serialises."""

    def __init__(self, **values: Any) -> None:
        unknown = set(values) - set(DEFAULTS)
        if unknown:
            raise TypeError('unknown FrameResult fields: %s' % ', '.join(sorted(unknown)))
        for name, default in DEFAULTS.items():
            setattr(self, name, values.get(name, default))

    def __repr__(self) -> str:
        return 'FrameResult(%s)' % ', '.join('%s=%r' % (k, getattr(self, k)) for k in KEYS)
