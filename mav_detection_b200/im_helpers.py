"""The three im_helpers functions on the hot path, with the reference's signatures
(/root/reference/src/im_helpers.py:55-84, 150-159, 244-252), evaluated on the device through the C ABI.
Inside Processor.run_detection the same quantities come out of the fused residual kernel; these
stand-alone versions exist for callers that use the helpers on their own.  The visualisation helpers of
the reference module (to_rgb, apply_colormap, get_flow_vis, ...) are out of scope."""
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
