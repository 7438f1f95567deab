"""Host-side helpers with the reference's names (/root/reference/src/utils.py): Rectangle (:13-104, the
part the hot loop uses), get_json (:350-361) and the .flo codec (:204-257).  No numerics live here —
line_intersection (:183-197) runs on the device inside mavd_foe."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np

TAG_FLOAT = 202021.25


class Rectangle:
    def __init__(self, topleft: Tuple[float, float], size: Tuple[float, float]) -> None:
        self.topleft = topleft
        self.size = size

    @classmethod
    def from_points(cls, topleft: Tuple[float, float], bottomright: Tuple[float, float]) -> 'Rectangle':
        # size excludes the last pixel, as in the reference (utils.py:26-30)
        return Rectangle(topleft, (bottomright[0] - topleft[0], bottomright[1] - topleft[1]))

    @classmethod
    def from_center(cls, center: Tuple[float, float], size: Tuple[float, float]) -> 'Rectangle':
        return Rectangle((center[0] - size[0] / 2, center[1] - size[1] / 2), size)

    def get_topleft(self) -> Tuple[float, float]:
        return (self.topleft[0], self.topleft[1])

    def get_bottomright(self) -> Tuple[float, float]:
        return (self.topleft[0] + self.size[0], self.topleft[1] + self.size[1])

    def get_center(self) -> Tuple[float, float]:
        return (self.topleft[0] + self.size[0] / 2, self.topleft[1] + self.size[1] / 2)

    def get_left(self) -> float:
        return self.topleft[0]

    def get_right(self) -> float:
        return self.topleft[0] + self.size[0]

    def get_top(self) -> float:
        return self.topleft[1]

    def get_bottom(self) -> float:
        return self.topleft[1] + self.size[1]

    def get_area(self) -> float:
        return max(1.0, self.size[0] * self.size[1])

    def to_yolo(self, img_size: np.ndarray, obj_id: int = 0) -> str:
        img_size = np.asarray(img_size).astype(np.float64)
        center = np.array(self.get_center()) / img_size
        size = np.array(self.size) / img_size
        return f'{obj_id} {center[0]} {center[1]} {size[0]} {size[1]}\n'

    @classmethod
    def calculate_iou(cls, r1: 'Rectangle', r2: 'Rectangle') -> float:
        left, right = max(r1.get_left(), r2.get_left()), min(r1.get_right(), r2.get_right())
        top, bottom = max(r1.get_top(), r2.get_top()), min(r1.get_bottom(), r2.get_bottom())
        aoo = (right - left) * (bottom - top)
        return aoo / (r1.get_area() + r2.get_area() - aoo)


def get_json(obj: Any) -> Dict[str, Any]:
    """utils.py:350-361: objects through __dict__, everything else JSON cannot encode through str()."""
    return json.loads(json.dumps(obj, default=lambda o: getattr(o, '__dict__', str(o))))


def create_if_not_exists(path: str) -> None:
    if not os.path.exists(path):
        os.makedirs(path)


def read_flow(filename: str) -> np.ndarray:
    """.flo reader (utils.py:204-223): float32 tag 202021.25, int32 w, int32 h, h*w*2 float32 (u, v)."""
    with open(filename, 'rb') as f:
        tag = np.fromfile(f, np.float32, count=1)[0]
        if tag != np.float32(TAG_FLOAT):
            raise AssertionError('Flow number %r incorrect. Invalid .flo file' % tag)
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
    return np.resize(data, (h, w, 2))


def write_flow(filename: str, uv: np.ndarray, v: Optional[np.ndarray] = None) -> None:
    """.flo writer (utils.py:226-257)."""
    if v is None:
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u = uv
    assert u.shape == v.shape
    h, w = u.shape
    with open(filename, 'wb') as f:
        f.write(np.array([TAG_FLOAT], np.float32).tobytes())
        np.array(w).astype(np.int32).tofile(f)
        np.array(h).astype(np.int32).tofile(f)
        np.stack([u, v], -1).astype(np.float32).tofile(f)
