"""The three im_helpers functions on the hot path, with the reference's signatures
(/root/reference/src/im_helpers.py:55-84, 150-159, 244-252), evaluated on the device through the C ABI.
Inside Processor.run_detection the same quantities come out of the fused residual kernel; these
stand-alone versions exist for callers that use the helpers on their own.  to_rgb / apply_colormap (the phi image of
processor.py:324,376) and mask_overlay (processor.py:385-392) are the visualisation payloads of run_detection, also
on the device; get_flow_vis needs the third-party flow_vis package, whose source is not part of the reference, and
is not rebuilt."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _lib, utils
from ._lib import check


def _cuda():
    import torch
    if not torch.cuda.is_available():
        raise _lib.MavdError('no CUDA device: the mav-detection hot path has no CPU fallback')
    return torch


def _stream(torch) -> int:
    return torch.cuda.current_stream().cuda_stream


def get_magnitude(img: np.ndarray) -> np.ndarray:
    """np.linalg.norm(img, axis=-1) for an (..., 2) float32/float64 vector field — im_helpers.py:150-159."""
    torch = _cuda()
    a = np.ascontiguousarray(img)
    if a.shape[-1] != 2:
        raise ValueError('get_magnitude expects a (..., 2) vector field')
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    d = torch.from_numpy(a).cuda()
    out = torch.empty(a.shape[:-1], dtype=d.dtype, device=d.device)
    check(_lib.load().mavd_magnitude(d.data_ptr(), 1 if a.dtype == np.float64 else 0, out.numel(), out.data_ptr(),
                                     _stream(torch)))
    return out.cpu().numpy()


def get_simple_bounding_box(img: np.ndarray) -> utils.Rectangle:
    """Bounding box around all pixels brighter than 0.1 * max — im_helpers.py:55-84.  Coordinates are -1 when
    nothing exceeds the threshold; the size excludes the last pixel (Rectangle.from_points)."""
    torch = _cuda()
    a = np.ascontiguousarray(img)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise ValueError('get_simple_bounding_box expects a uint8 (H, W) or (H, W, C) image')
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    d = torch.from_numpy(a).cuda()
    out = torch.empty((5,), dtype=torch.int32, device=d.device)
    check(_lib.load().mavd_simple_bbox(d.data_ptr(), w, h, c, out.data_ptr(), _stream(torch)))
    x0, y0, x1, y1, _ = (int(v) for v in out.cpu().numpy())
    return utils.Rectangle.from_points((x0, y0), (x1, y1))


def calculate_tpr_fpr(gt_img: np.ndarray, img: np.ndarray) -> Tuple[float, float]:
    """true/false positive rates of `img` (e.g. 255 * mask) against a uint8 ground truth — im_helpers.py:244-252.
    Empty classes give nan / inf exactly as the NumPy division in the reference does."""
    torch = _cuda()
    gt = np.ascontiguousarray(gt_img)
    if gt.dtype != np.uint8:
        raise ValueError('calculate_tpr_fpr expects a uint8 ground-truth image')
    im = np.ascontiguousarray(np.broadcast_to(np.asarray(img), gt.shape)).astype(np.int64)
    dg, di = torch.from_numpy(gt).cuda(), torch.from_numpy(im).cuda()
    out = torch.empty((4,), dtype=torch.int64, device=dg.device)
    check(_lib.load().mavd_tpr_fpr_counts(dg.data_ptr(), di.data_ptr(), gt.size, out.data_ptr(), _stream(torch)))
    pos, neg, tp, fp = (int(v) for v in out.cpu().numpy())
    with np.errstate(divide='ignore', invalid='ignore'):
        return (float(np.float64(tp) / np.float64(pos)), float(np.float64(fp) / np.float64(neg)))


def _phi_images(img: np.ndarray, max_value, want_gray: bool, want_jet: bool):
    torch = _cuda()
    a = np.ascontiguousarray(img)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)                # integer input: the reference's to_int works in float64 then
    if max_value is None:
        max_value = float(np.max(a))            # im_helpers.py:188-189
    d = torch.from_numpy(a).cuda()
    gray = torch.empty(a.shape + (3,), dtype=torch.uint8, device=d.device) if want_gray else None
    jet = torch.empty(a.shape + (3,), dtype=torch.uint8, device=d.device) if want_jet else None
    check(_lib.load().mavd_phi_colormap(d.data_ptr(), 1 if a.dtype == np.float64 else 0, a.size, float(max_value),
                                        gray.data_ptr() if want_gray else None, jet.data_ptr() if want_jet else None,
                                        _stream(torch)))
    return (gray.cpu().numpy() if want_gray else None), (jet.cpu().numpy() if want_jet else None)


def to_rgb(img: np.ndarray, max_value: float = None) -> np.ndarray:
    """Grayscale float image -> (H, W, 3) uint8 — im_helpers.py:162-173 with to_int(normalize=True):
    uint8(around(|img| * 255 / max_value)) on three equal channels."""
    return _phi_images(img, max_value, True, False)[0]


def apply_colormap(img: np.ndarray, max_value: float = None) -> np.ndarray:
    """cv2.applyColorMap(img, COLORMAP_JET) as im_helpers.py:112-135 applies it.  A float image is normalised like
    to_rgb first; a uint8 (H, W, 3) image with equal channels (to_rgb's output, the case at processor.py:325,376) is
    looked up directly.  With max_value the reference writes it into pixel [0, 0] before the lookup and restores the
    pixel from a VIEW of the value it has just overwritten (im_helpers.py:130-133): pixel [0, 0] of the result is the
    colour of uint8(max_value), all other pixels are unaffected.  Reproduced as is."""
    a = np.asarray(img)
    if a.dtype in (np.float32, np.float64):
        out = _phi_images(a, max_value, False, True)[1]
    else:
        if a.dtype != np.uint8:
            raise ValueError('apply_colormap expects a float image or a uint8 image')
        if a.ndim == 3:
            if not (np.array_equal(a[..., 0], a[..., 1]) and np.array_equal(a[..., 0], a[..., 2])):
                raise ValueError('apply_colormap: only gray (equal-channel) uint8 images are supported')
            a = a[..., 0]
        # uint8 v -> v * 255 / 255 = v: the same kernel does the lookup
        out = _phi_images(a.astype(np.float32), 255.0, False, True)[1]
    if max_value is not None:
        v = np.zeros((1, 1), np.float32) + np.float32(int(np.array(max_value).astype(np.uint8)))
        out[0, 0] = _phi_images(v, 255.0, False, True)[1][0, 0]
    return out


def mask_overlay(frame: np.ndarray, estimate_fixed: np.ndarray, want_mask_rgb: bool = False):
    """The detection overlay of processor.py:385-392: (150, 0, 150) where estimate_fixed is set, blended
    0.2 * frame + 0.8 * painted frame with cv2.addWeighted's rounding.  frame: (H, W, 3) BGR or (H, W) gray uint8.
    With want_mask_rgb also returns im_helpers.to_rgb(255 * estimate_fixed) (processor.py:364)."""
    torch = _cuda()
    f = np.ascontiguousarray(frame)
    m = np.ascontiguousarray(np.asarray(estimate_fixed).astype(np.uint8))
    if f.dtype != np.uint8 or f.ndim not in (2, 3) or (f.ndim == 3 and f.shape[2] != 3) or f.shape[:2] != m.shape:
        raise ValueError('mask_overlay expects a uint8 (H, W[, 3]) frame and an (H, W) mask')
    df, dm = torch.from_numpy(f).cuda(), torch.from_numpy(m).cuda()
    out = torch.empty(m.shape + (3,), dtype=torch.uint8, device=df.device)
    rgb = torch.empty_like(out) if want_mask_rgb else None
    check(_lib.load().mavd_mask_overlay(df.data_ptr(), 3 if f.ndim == 3 else 1, dm.data_ptr(), m.size, out.data_ptr(),
                                        rgb.data_ptr() if want_mask_rgb else None, _stream(torch)))
    return (out.cpu().numpy(), rgb.cpu().numpy()) if want_mask_rgb else out.cpu().numpy()
