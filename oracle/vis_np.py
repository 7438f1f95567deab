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


# ------------------------------------------------------------------------------------------------
# Image payloads of Processor.run_detection (/root/reference/src/processor.py:324,364-376,385-392)
# ------------------------------------------------------------------------------------------------
# OpenCV's COLORMAP_JET table, BGR order (cv2.applyColorMap(arange(256), COLORMAP_JET), cv2 4.13.0); pinned against the
# live cv2 in tests/test_flo_and_vis.py
JET_LUT = np.array([
    128, 0, 0, 132, 0, 0, 136, 0, 0, 140, 0, 0, 144, 0, 0, 148, 0, 0, 152, 0, 0, 156, 0, 0,
    160, 0, 0, 164, 0, 0, 168, 0, 0, 172, 0, 0, 176, 0, 0, 180, 0, 0, 184, 0, 0, 188, 0, 0,
    192, 0, 0, 196, 0, 0, 200, 0, 0, 204, 0, 0, 208, 0, 0, 212, 0, 0, 216, 0, 0, 220, 0, 0,
    224, 0, 0, 228, 0, 0, 232, 0, 0, 236, 0, 0, 240, 0, 0, 244, 0, 0, 248, 0, 0, 252, 0, 0,
    255, 0, 0, 255, 4, 0, 255, 8, 0, 255, 12, 0, 255, 16, 0, 255, 20, 0, 255, 24, 0, 255, 28, 0,
    255, 32, 0, 255, 36, 0, 255, 40, 0, 255, 44, 0, 255, 48, 0, 255, 52, 0, 255, 56, 0, 255, 60, 0,
    255, 64, 0, 255, 68, 0, 255, 72, 0, 255, 76, 0, 255, 80, 0, 255, 84, 0, 255, 88, 0, 255, 92, 0,
    255, 96, 0, 255, 100, 0, 255, 104, 0, 255, 108, 0, 255, 112, 0, 255, 116, 0, 255, 120, 0, 255, 124, 0,
    255, 128, 0, 255, 132, 0, 255, 136, 0, 255, 140, 0, 255, 144, 0, 255, 148, 0, 255, 152, 0, 255, 156, 0,
    255, 160, 0, 255, 164, 0, 255, 168, 0, 255, 172, 0, 255, 176, 0, 255, 180, 0, 255, 184, 0, 255, 188, 0,
    255, 192, 0, 255, 196, 0, 255, 200, 0, 255, 204, 0, 255, 208, 0, 255, 212, 0, 255, 216, 0, 255, 220, 0,
    255, 224, 0, 255, 228, 0, 255, 232, 0, 255, 236, 0, 255, 240, 0, 255, 244, 0, 255, 248, 0, 255, 252, 0,
    254, 255, 2, 250, 255, 6, 246, 255, 10, 242, 255, 14, 238, 255, 18, 234, 255, 22, 230, 255, 26, 226, 255, 30,
    222, 255, 34, 218, 255, 38, 214, 255, 42, 210, 255, 46, 206, 255, 50, 202, 255, 54, 198, 255, 58, 194, 255, 62,
    190, 255, 66, 186, 255, 70, 182, 255, 74, 178, 255, 78, 174, 255, 82, 170, 255, 86, 166, 255, 90, 162, 255, 94,
    158, 255, 98, 154, 255, 102, 150, 255, 106, 146, 255, 110, 142, 255, 114, 138, 255, 118, 134, 255, 122, 130, 255, 126,
    126, 255, 130, 122, 255, 134, 118, 255, 138, 114, 255, 142, 110, 255, 146, 106, 255, 150, 102, 255, 154, 98, 255, 158,
    94, 255, 162, 90, 255, 166, 86, 255, 170, 82, 255, 174, 78, 255, 178, 74, 255, 182, 70, 255, 186, 66, 255, 190,
    62, 255, 194, 58, 255, 198, 54, 255, 202, 50, 255, 206, 46, 255, 210, 42, 255, 214, 38, 255, 218, 34, 255, 222,
    30, 255, 226, 26, 255, 230, 22, 255, 234, 18, 255, 238, 14, 255, 242, 10, 255, 246, 6, 255, 250, 1, 255, 254,
    0, 252, 255, 0, 248, 255, 0, 244, 255, 0, 240, 255, 0, 236, 255, 0, 232, 255, 0, 228, 255, 0, 224, 255,
    0, 220, 255, 0, 216, 255, 0, 212, 255, 0, 208, 255, 0, 204, 255, 0, 200, 255, 0, 196, 255, 0, 192, 255,
    0, 188, 255, 0, 184, 255, 0, 180, 255, 0, 176, 255, 0, 172, 255, 0, 168, 255, 0, 164, 255, 0, 160, 255,
    0, 156, 255, 0, 152, 255, 0, 148, 255, 0, 144, 255, 0, 140, 255, 0, 136, 255, 0, 132, 255, 0, 128, 255,
    0, 124, 255, 0, 120, 255, 0, 116, 255, 0, 112, 255, 0, 108, 255, 0, 104, 255, 0, 100, 255, 0, 96, 255,
    0, 92, 255, 0, 88, 255, 0, 84, 255, 0, 80, 255, 0, 76, 255, 0, 72, 255, 0, 68, 255, 0, 64, 255,
    0, 60, 255, 0, 56, 255, 0, 52, 255, 0, 48, 255, 0, 44, 255, 0, 40, 255, 0, 36, 255, 0, 32, 255,
    0, 28, 255, 0, 24, 255, 0, 20, 255, 0, 16, 255, 0, 12, 255, 0, 8, 255, 0, 4, 255, 0, 0, 255,
    0, 0, 252, 0, 0, 248, 0, 0, 244, 0, 0, 240, 0, 0, 236, 0, 0, 232, 0, 0, 228, 0, 0, 224,
    0, 0, 220, 0, 0, 216, 0, 0, 212, 0, 0, 208, 0, 0, 204, 0, 0, 200, 0, 0, 196, 0, 0, 192,
    0, 0, 188, 0, 0, 184, 0, 0, 180, 0, 0, 176, 0, 0, 172, 0, 0, 168, 0, 0, 164, 0, 0, 160,
    0, 0, 156, 0, 0, 152, 0, 0, 148, 0, 0, 144, 0, 0, 140, 0, 0, 136, 0, 0, 132, 0, 0, 128,
], np.uint8).reshape(256, 3)


def to_rgb(img: np.ndarray, max_value=None) -> np.ndarray:
    """im_helpers.to_rgb (im_helpers.py:162-173) = GRAY2RGB(to_int(img, uint8, normalize=True, max_value))
    (im_helpers.py:176-201), evaluated in img's own dtype like NumPy does."""
    if max_value is None:
        max_value = np.max(img)
    elif max_value <= 0.0:
        max_value = 1.0
    with np.errstate(invalid='ignore', divide='ignore'):
        v = np.around(np.abs(img) * 255 / max_value)
        v = np.where(np.isnan(v), 0, v).astype(np.int64).astype(np.uint8)
    return np.repeat(v[..., None], 3, axis=-1)


def apply_colormap_jet(rgb: np.ndarray, max_value=None) -> np.ndarray:
    """im_helpers.apply_colormap on a 3-channel uint8 image (im_helpers.py:112-135): cv2.applyColorMap(COLORMAP_JET)
    reduces it with BGR2GRAY, (B*3735 + G*19235 + R*9798 + 16384) >> 15, then looks the gray value up in the JET table.
    With max_value the reference writes max_value into pixel [0, 0], maps, and restores that pixel from `old_value` —
    which is a VIEW of the pixel it has just overwritten (im_helpers.py:130-133), so pixel [0, 0] comes out as the
    colour of uint8(max_value); every other pixel is unaffected (the lookup does not normalise)."""
    b, g, r = (rgb[..., k].astype(np.uint32) for k in range(3))
    gray = (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15
    out = JET_LUT[gray]
    if max_value is not None:
        out[0, 0] = JET_LUT[int(np.array(max_value).astype(np.uint8))]
    return out


def mask_overlay(frame: np.ndarray, fixed: np.ndarray):
    """processor.py:364 and :385-392: (to_rgb(255 * estimate_fixed) restated, mask_vis).  cv2.addWeighted on uint8:
    saturate(round_half_even(float32(a) * 0.2 + float32(b) * 0.8))."""
    fixed = np.asarray(fixed).astype(bool)
    f3 = frame if frame.ndim == 3 else np.repeat(frame[..., None], 3, axis=-1)
    painted = f3.copy()
    painted[fixed] = (150, 0, 150)
    t = f3.astype(F) * F(0.2) + painted.astype(F) * F(0.8)
    vis = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    mask_rgb = np.repeat((fixed.astype(np.uint8) * 255)[..., None], 3, axis=-1)
    return vis, mask_rgb
