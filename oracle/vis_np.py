"""NumPy restatement of the visualisation Farneback.process() returns (/root/reference/src/farneback.py:83-99).

TEST INFRASTRUCTURE (see oracle/__init__.py).  `process_visualisation_cv2` is the reference's own statement
sequence run with cv2 (the library the reference calls); `process_visualisation` restates the arithmetic cv2
4.13.0 performs (cartToPolar -> fastAtan2 polynomial with FMAs, float32 NORM_MINMAX, truncating HSV2BGR) so
that the CUDA kernel can be checked without cv2's internals; tests/test_oracle_vis.py pins the restatement
bit-exactly against the cv2 version.

cv2 quirk [probe, cv2 4.13.0]: the 8-bit HSV2BGR converts whole 32-pixel blocks of a row with a SIMD body that
TRUNCATES the scaled float, and the remaining (width mod 32) pixels of each row with a scalar tail that ROUNDS.
The restatement (and the CUDA kernel) use the body arithmetic for every pixel, so they equal cv2 exactly for
widths that are multiples of 32 (1920, 3840, 640, 752 = 23.5 blocks is NOT) and differ by at most 1 grey level
in those tail columns otherwise.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def process_visualisation_cv2(flow: np.ndarray, shape3) -> tuple:
    """farneback.py:83-99 verbatim (img only supplies the shape/dtype).  Returns (bgr, invalid_frame)."""
    import cv2
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    hsv = np.zeros(shape3, np.uint8)
    with np.errstate(invalid='ignore'):
        hsv[:, :, 0] = ang * 180 / np.pi / 2
        hsv[:, :, 1] = 255
        hsv[:, :, 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX) * 2.0
    invalid_frame = np.sum(hsv[:, :, 2]) < 1
    mask = hsv[:, :, 2] < 1
    hsv[mask, 0] = 127
    hsv[mask, 2] = 255
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR), bool(invalid_frame)


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def cart_to_polar(x: np.ndarray, y: np.ndarray):
    """cv2.cartToPolar(x, y) (radians): magnitude sqrt(fma(x, x, y*y)), angle = fastAtan2 polynomial."""
    mag = np.sqrt(_fma(x, x, (y * y).astype(F))).astype(F)
    p1, p3, p5, p7 = (F(0.9997878412794807) * F(180 / np.pi), F(-0.3258083974640975) * F(180 / np.pi),
                      F(0.1555786518463281) * F(180 / np.pi), F(-0.04432655554792128) * F(180 / np.pi))
    ax, ay = np.abs(x), np.abs(y)
    c = (np.minimum(ax, ay) / (np.maximum(ax, ay) + F(2.220446049250313e-16))).astype(F)
    c2 = (c * c).astype(F)
    full = lambda v: np.full_like(c, v)
    a = (_fma(_fma(_fma(full(p7), c2, full(p5)), c2, full(p3)), c2, full(p1)) * c).astype(F)
    a = np.where(ax < ay, F(90) - a, a)
    a = np.where(x < 0, F(180) - a, a)
    a = np.where(y < 0, F(360) - a, a)
    return mag, (a.astype(F) * F(np.pi / 180)).astype(F)


def hsv2bgr_s255(h8: np.ndarray, v8: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(COLOR_HSV2BGR) on uint8 with S == 255 (exhaustively equal to cv2 4.13.0)."""
    v = v8.astype(F) * F(1 / 255.)
    h = h8.astype(F) * F(6 / 180.)
    for _ in range(3):
        h = np.where(h >= 6, h - F(6), h)
    sec = np.floor(h).astype(np.int32)
    h = (h - sec).astype(F)
    tab = np.stack([v, v * F(0), v * (F(1) - h), v * (F(1) - (F(1) - h))], -1).astype(F)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    idx = sd[np.clip(sec, 0, 5)]
    out = np.stack([np.take_along_axis(tab, idx[..., k:k + 1], -1)[..., 0] for k in range(3)], -1)
    return np.clip(np.trunc(out * F(255)), 0, 255).astype(np.uint8)


def process_visualisation(flow: np.ndarray) -> tuple:
    """Restatement of farneback.py:83-99.  Returns (bgr uint8 (H, W, 3), invalid_frame)."""
    mag, ang = cart_to_polar(flow[..., 0].astype(F), flow[..., 1].astype(F))
    h8 = (((ang * F(180)).astype(F) / F(np.pi)).astype(F) / F(2)).astype(F).astype(np.int64).astype(np.uint8)
    mn, mx = float(mag.min()), float(mag.max())
    span = mx - mn
    scale = 255.0 * (1.0 / span if span > 2.220446049250313e-16 else 0.0)
    nrm = ((mag * F(scale)).astype(F) + F(0.0 - mn * scale)).astype(F)
    v8 = (nrm * F(2)).astype(F).astype(np.int64).astype(np.uint8)          # truncate, wrap modulo 256
    invalid = int(v8.astype(np.int64).sum()) < 1
    zero = v8 < 1
    h8 = np.where(zero, np.uint8(127), h8)
    v8 = np.where(zero, np.uint8(255), v8)
    return hsv2bgr_s255(h8, v8), invalid
