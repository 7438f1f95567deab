"""Thin host wrapper over the C ABI: owns one mavd handle and passes torch CUDA tensors' data pointers
and the current CUDA stream to libmavd.  PyTorch is plumbing here (device memory, streams,
torch.distributed) — all compute runs in the hand-written sm_100a kernels of csrc/."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Config, DetectParams, FarnebackParams, FrameRecord, FrameStats, Imu, check

# cv2.calcOpticalFlowFarneback arguments used by the reference (src/farneback.py:78-80)
REFERENCE_PARAMS = dict(pyr_scale=0.4, levels=1, winsize=12, iterations=10, poly_n=8, poly_sigma=1.2, flags=0)
# OpenCV's documented sample values (BASELINE config 2)
SAMPLE_PARAMS = dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)

RECORD_DTYPE = np.dtype(FrameRecord)
STATS_DTYPE = np.dtype(FrameStats)


def make_imu(n: int, ang: Optional[np.ndarray] = None, dt=None, derotate=None) -> C.Array:
    """Build the mavd_imu array for n frames.  ang (n,3), dt (n,) or scalar, derotate (n,) bool."""
    arr = (Imu * n)()
    for i in range(n):
        a = (0.0, 0.0, 0.0) if ang is None else np.asarray(ang, np.float64).reshape(-1, 3)[i]
        arr[i].ang[0], arr[i].ang[1], arr[i].ang[2] = float(a[0]), float(a[1]), float(a[2])
        arr[i].dt = float(dt if np.isscalar(dt) else (1.0 if dt is None else dt[i]))
        d = True if derotate is None else (derotate if isinstance(derotate, (bool, int)) else derotate[i])
        arr[i].derotate = 1 if d else 0
    return arr


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One handle per (device, W, H, Farneback parameters) — SURVEY.md §8b threading/state."""

    def __init__(self, width: int, height: int, params: Optional[Dict] = None, max_pairs: int = 1,
                 device: int = 0) -> None:
        if not torch.cuda.is_available():
            raise _lib.MavdError('no CUDA device: the mav-detection hot path has no CPU fallback')
        self.lib = _lib.load()
        p = dict(REFERENCE_PARAMS)
        if params:
            p.update(params)
        self.params = p
        self.width, self.height, self.max_pairs = int(width), int(height), int(max_pairs)
        self.device = torch.device('cuda', device)
        cfg = Config(device, self.width, self.height, self.max_pairs,
                     FarnebackParams(float(p['pyr_scale']), int(p['levels']), int(p['winsize']),
                                     int(p['iterations']), int(p['poly_n']), float(p['poly_sigma']),
                                     int(p['flags'])))
        h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            check(self.lib.mavd_create(C.byref(cfg), C.byref(h)))
        self._h = h
        n = C.c_int32()
        ws, hs = (C.c_int32 * 16)(), (C.c_int32 * 16)()
        check(self.lib.mavd_level_info(self._h, C.byref(n), ws, hs))
        self.levels: List[Tuple[int, int]] = [(ws[i], hs[i]) for i in range(n.value)]
        self.detect_params = DetectParams()
        self.lib.mavd_default_detect_params(C.byref(self.detect_params))

    # -- lifecycle --------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.mavd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        out = C.c_size_t()
        check(self.lib.mavd_workspace_bytes(self._h, C.byref(out)))
        return out.value

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check_frames(self, frames: torch.Tensor, n_pairs: int, pair_stride: int) -> None:
        need = n_pairs + 1 if pair_stride == 1 else 2 * n_pairs
        if frames.dtype != torch.uint8 or not frames.is_cuda or not frames.is_contiguous():
            raise ValueError('frames must be a contiguous CUDA uint8 tensor (F, H, W)')
        if frames.dim() != 3 or frames.shape[1] != self.height or frames.shape[2] != self.width:
            raise ValueError('frames must have shape (F, %d, %d), got %s' % (self.height, self.width,
                                                                             tuple(frames.shape)))
        if frames.shape[0] < need:
            raise ValueError('%d pairs with pair_stride %d need %d frames, got %d'
                             % (n_pairs, pair_stride, need, frames.shape[0]))

    # -- stage 0 -----------------------------------------------------------------------------
    def bgr2gray(self, bgr: torch.Tensor) -> torch.Tensor:
        """cv2.cvtColor(img, COLOR_BGR2GRAY) — src/farneback.py:74.  (..., 3) uint8 -> (...) uint8."""
        if bgr.dtype != torch.uint8 or not bgr.is_cuda or bgr.shape[-1] != 3:
            raise ValueError('bgr must be a CUDA uint8 tensor (..., 3)')
        bgr = bgr.contiguous()
        out = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=bgr.device)
        check(self.lib.mavd_bgr2gray(bgr.data_ptr(), out.data_ptr(), out.numel(), self._stream()))
        return out

    # -- stage 1 -----------------------------------------------------------------------------
    def farneback(self, frames: torch.Tensor, n_pairs: Optional[int] = None, pair_stride: int = 1,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Dense flow for every pair — cv2.calcOpticalFlowFarneback (src/farneback.py:76-80)."""
        if n_pairs is None:
            n_pairs = frames.shape[0] - 1 if pair_stride == 1 else frames.shape[0] // 2
        self._check_frames(frames, n_pairs, pair_stride)
        if out is None:
            out = torch.empty((n_pairs, self.height, self.width, 2), dtype=torch.float32, device=self.device)
        check(self.lib.mavd_farneback(self._h, frames.data_ptr(), n_pairs, pair_stride, out.data_ptr(),
                                      self._stream()))
        return out

    def tap(self, kind: str, level: int, index: int = 0) -> torch.Tensor:
        kinds = {'img': (0, 1), 'R': (1, 5), 'M': (2, 5), 'flow': (3, 2)}
        k, ch = kinds[kind]
        w, h = self.levels[level]
        shape = (h, w) if ch == 1 else (h, w, ch)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        check(self.lib.mavd_farneback_tap(self._h, k, level, index, out.data_ptr(), self._stream()))
        return out

    # -- stage 1.5 ----------------------------------------------------------------------------
    def derotate(self, flow: torch.Tensor, imu) -> torch.Tensor:
        n = flow.shape[0]
        out = torch.empty(flow.shape, dtype=torch.float64, device=self.device)
        check(self.lib.mavd_derotate(self._h, flow.data_ptr(), n, imu, out.data_ptr(), self._stream()))
        return out

    # -- stage 2 -----------------------------------------------------------------------------
    def foe(self, flow: torch.Tensor, imu, samples: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n = flow.shape[0]
        if samples.dtype != torch.int32 or tuple(samples.shape) != (n, _lib.SAMPLES_PER_FRAME):
            raise ValueError('samples must be int32 (n, %d): [ry(2000) | rx(2000)]' % _lib.SAMPLES_PER_FRAME)
        foe = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        check(self.lib.mavd_foe(self._h, flow.data_ptr(), n, imu, C.byref(self.detect_params),
                                samples.data_ptr(), foe.data_ptr(), cnt.data_ptr(), self._stream()))
        return foe, cnt

    # -- stage 3 -----------------------------------------------------------------------------
    def residual_masks(self, flow: torch.Tensor, imu, foe: torch.Tensor, sky: Optional[torch.Tensor] = None,
                       seg: Optional[torch.Tensor] = None, want_phi: bool = True, want_stats: bool = True):
        n = flow.shape[0]
        npx = self.width * self.height
        phi = torch.zeros((n, self.height, self.width), dtype=torch.float64, device=self.device) if want_phi else None
        total = torch.empty((n, self.height, self.width), dtype=torch.uint8, device=self.device)
        fixed = torch.empty_like(total)
        stats = torch.zeros((n, STATS_DTYPE.itemsize), dtype=torch.uint8, device=self.device) if want_stats else None

        def stride(t):
            if t is None:
                return 0
            if t.dtype not in (torch.uint8, torch.bool) or not t.is_contiguous():
                raise ValueError('sky/seg must be contiguous uint8/bool tensors')
            return 0 if t.dim() == 2 else npx
        check(self.lib.mavd_residual_masks(self._h, flow.data_ptr(), n, imu, C.byref(self.detect_params),
                                           foe.data_ptr(), _ptr(sky), stride(sky), _ptr(seg), stride(seg),
                                           _ptr(phi), total.data_ptr(), fixed.data_ptr(), _ptr(stats),
                                           self._stream()))
        return phi, total, fixed, stats

    # -- stage 4 -----------------------------------------------------------------------------
    def ccl(self, mask: torch.Tensor, max_boxes: int = _lib.MAX_BOXES):
        n = mask.shape[0]
        labels = torch.empty((n, self.height, self.width), dtype=torch.int32, device=self.device)
        boxes = torch.zeros((n, max_boxes, 5), dtype=torch.int32, device=self.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        check(self.lib.mavd_ccl(self._h, mask.data_ptr(), n, labels.data_ptr(), boxes.data_ptr(), max_boxes,
                                cnt.data_ptr(), self._stream()))
        return labels, boxes, cnt

    # -- whole path ----------------------------------------------------------------------------
    def process(self, frames: torch.Tensor, imu, samples: torch.Tensor, n_pairs: Optional[int] = None,
                pair_stride: int = 1, sky: Optional[torch.Tensor] = None, seg: Optional[torch.Tensor] = None,
                flow_out: Optional[torch.Tensor] = None, total_out: Optional[torch.Tensor] = None,
                fixed_out: Optional[torch.Tensor] = None, records: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Farneback -> derotate -> FoE -> phi/masks -> components, device buffers in and out.
        Returns the (n_pairs, sizeof(mavd_frame_record)) uint8 CUDA tensor of records."""
        if n_pairs is None:
            n_pairs = frames.shape[0] - 1 if pair_stride == 1 else frames.shape[0] // 2
        self._check_frames(frames, n_pairs, pair_stride)
        npx = self.width * self.height
        if records is None:
            records = torch.empty((n_pairs, RECORD_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        sky_stride = 0 if (sky is None or sky.dim() == 2) else npx
        seg_stride = 0 if (seg is None or seg.dim() == 2) else npx
        check(self.lib.mavd_process(self._h, frames.data_ptr(), n_pairs, pair_stride, imu,
                                    C.byref(self.detect_params), samples.data_ptr(), _ptr(sky), sky_stride,
                                    _ptr(seg), seg_stride, _ptr(flow_out), _ptr(total_out), _ptr(fixed_out),
                                    records.data_ptr(), self._stream()))
        return records

    def process_host(self, frames: np.ndarray, imu, samples: np.ndarray, n_pairs: Optional[int] = None,
                     pair_stride: int = 1, sky: Optional[np.ndarray] = None, seg: Optional[np.ndarray] = None,
                     flow_out: Optional[np.ndarray] = None, fixed_out: Optional[np.ndarray] = None,
                     records: Optional[np.ndarray] = None) -> np.ndarray:
        """The end-to-end call: HOST buffers in, HOST records (and optional masks / flow) out."""
        if n_pairs is None:
            n_pairs = frames.shape[0] - 1 if pair_stride == 1 else frames.shape[0] // 2
        npx = self.width * self.height
        if records is None:
            records = np.empty((n_pairs,), dtype=RECORD_DTYPE)
        for name, a in (('frames', frames), ('samples', samples), ('sky', sky), ('seg', seg),
                        ('flow_out', flow_out), ('fixed_out', fixed_out)):
            if a is not None and not a.flags['C_CONTIGUOUS']:
                raise ValueError('%s must be C-contiguous' % name)
        if frames.dtype != np.uint8 or samples.dtype != np.int32:
            raise ValueError('frames must be uint8 and samples int32')

        def hp(a):
            return None if a is None else a.ctypes.data
        sky_stride = 0 if (sky is None or sky.ndim == 2) else npx
        seg_stride = 0 if (seg is None or seg.ndim == 2) else npx
        with torch.cuda.device(self.device):
            check(self.lib.mavd_process_host(self._h, hp(frames), n_pairs, pair_stride, imu,
                                             C.byref(self.detect_params), hp(samples), hp(sky), sky_stride,
                                             hp(seg), seg_stride, hp(flow_out), hp(fixed_out),
                                             records.ctypes.data, self._stream()))
        return records

    @staticmethod
    def records_to_numpy(records: torch.Tensor) -> np.ndarray:
        return records.cpu().numpy().view(RECORD_DTYPE).reshape(-1)

    @staticmethod
    def stats_to_numpy(stats: torch.Tensor) -> np.ndarray:
        return stats.cpu().numpy().view(STATS_DTYPE).reshape(-1)

    def force_generic_iteration(self, on: bool = True) -> None:
        """Tests only: use the non-TMA iteration kernel even where the TMA kernel applies."""
        check(self.lib.mavd_debug_force_generic_iteration(self._h, 1 if on else 0))

    def profile_enable(self, on: bool = True) -> None:
        check(self.lib.mavd_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """{kernel class: (summed device ms, timed launch groups)} since profile_enable()."""
        prof = _lib.Profile()
        check(self.lib.mavd_profile_read(self._h, C.byref(prof)))
        return {n: (prof.ms[i], int(prof.launches[i])) for i, n in enumerate(_lib.PROF_NAMES)}

    def launch_count(self) -> int:
        return int(self.lib.mavd_launch_count())
