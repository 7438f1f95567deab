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


def _hp(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


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
        self._inflight: Dict[int, tuple] = {}

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
        """Detector.derotate for a batch — detector.py:70-117.  float32 or float64 (n, H, W, 2) in, float64 out."""
        n = flow.shape[0]
        kind = self._flow_kind(flow)
        out = torch.empty(flow.shape, dtype=torch.float64, device=self.device)
        fn = self.lib.mavd_derotate_f64 if kind else self.lib.mavd_derotate
        check(fn(self._h, flow.data_ptr(), n, imu, out.data_ptr(), self._stream()))
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

    # -- reference-literal seams (FocusOfExpansion methods on a flow array of either dtype) -----
    def _flow_kind(self, flow: torch.Tensor) -> int:
        if flow.dtype not in (torch.float32, torch.float64) or not flow.is_cuda or not flow.is_contiguous():
            raise ValueError('flow must be a contiguous CUDA float32/float64 tensor (n, H, W, 2)')
        if tuple(flow.shape[1:]) != (self.height, self.width, 2):
            raise ValueError('flow must have shape (n, %d, %d, 2), got %s' % (self.height, self.width, tuple(flow.shape)))
        return 1 if flow.dtype == torch.float64 else 0

    def foe_dense(self, flow: torch.Tensor, samples: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """FocusOfExpansion.get_FOE_dense(flow_uv) on the flow as given (no IMU) — focus_of_expansion.py:56-86."""
        n = flow.shape[0]
        if samples.dtype != torch.int32 or tuple(samples.shape) != (n, _lib.SAMPLES_PER_FRAME):
            raise ValueError('samples must be int32 (n, %d): [ry(2000) | rx(2000)]' % _lib.SAMPLES_PER_FRAME)
        foe = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        check(self.lib.mavd_foe_dense(self._h, flow.data_ptr(), self._flow_kind(flow), n, C.byref(self.detect_params),
                                      samples.data_ptr(), foe.data_ptr(), cnt.data_ptr(), self._stream()))
        return foe, cnt

    def ransac(self, estimates: torch.Tensor, threshold: Optional[float] = None) -> torch.Tensor:
        """FocusOfExpansion.ransac(estimates) — focus_of_expansion.py:32-54.  (K, 2) float64 -> (2,) float64."""
        if estimates.dtype != torch.float64 or estimates.dim() != 2 or estimates.shape[1] != 2:
            raise ValueError('estimates must be a (K, 2) float64 tensor')
        estimates = estimates.contiguous()
        out = torch.empty((2,), dtype=torch.float64, device=self.device)
        thr = self.detect_params.ransac_threshold if threshold is None else float(threshold)
        check(self.lib.mavd_ransac(self._h, estimates.data_ptr() if estimates.shape[0] else None, estimates.shape[0],
                                   thr, out.data_ptr(), self._stream()))
        return out

    def get_phi(self, flow: torch.Tensor, foe: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """FocusOfExpansion.get_phi(flow, FoE) — focus_of_expansion.py:150-184.  Returns (phi in the flow's
        dtype, per-frame max phi = the max_flow side effect)."""
        n = flow.shape[0]
        kind = self._flow_kind(flow)
        phi = torch.empty((n, self.height, self.width), dtype=flow.dtype, device=self.device)
        mx = torch.empty((n,), dtype=torch.float64, device=self.device)
        foe = foe.to(device=self.device, dtype=torch.float64).contiguous()
        if kind == 0:
            # the C ABI lays float32 phi at the front of each frame's 8-byte-per-pixel slot
            buf = torch.empty((n, self.height, self.width), dtype=torch.float64, device=self.device)
            check(self.lib.mavd_get_phi(self._h, flow.data_ptr(), 0, n, foe.data_ptr(), buf.data_ptr(), mx.data_ptr(),
                                        self._stream()))
            npx = self.height * self.width
            phi = buf.view(torch.float32).view(n, 2 * npx)[:, :npx].reshape(n, self.height, self.width).clone()
        else:
            check(self.lib.mavd_get_phi(self._h, flow.data_ptr(), 1, n, foe.data_ptr(), phi.data_ptr(), mx.data_ptr(),
                                        self._stream()))
        return phi, mx

    # -- stage 4 -----------------------------------------------------------------------------
    def ccl(self, mask: torch.Tensor, max_boxes: int = _lib.MAX_BOXES, want_labels: bool = True):
        n = mask.shape[0]
        labels = torch.empty((n, self.height, self.width), dtype=torch.int32, device=self.device) if want_labels else None
        boxes = torch.zeros((n, max_boxes, 5), dtype=torch.int32, device=self.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        check(self.lib.mavd_ccl(self._h, mask.data_ptr(), n, _ptr(labels), boxes.data_ptr(), max_boxes,
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

    def detect(self, flow: torch.Tensor, imu, samples: torch.Tensor, sky: Optional[torch.Tensor] = None,
               seg: Optional[torch.Tensor] = None, total_out: Optional[torch.Tensor] = None,
               fixed_out: Optional[torch.Tensor] = None, records: Optional[torch.Tensor] = None) -> torch.Tensor:
        """derotate -> FoE -> phi/masks -> components from a given float32 flow (the Dataset.get_flow_uv seam)."""
        n = flow.shape[0]
        if flow.dtype != torch.float32 or not flow.is_cuda or not flow.is_contiguous() or \
                tuple(flow.shape[1:]) != (self.height, self.width, 2):
            raise ValueError('flow must be a contiguous CUDA float32 tensor (n, %d, %d, 2)' % (self.height, self.width))
        npx = self.width * self.height
        if records is None:
            records = torch.empty((n, RECORD_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        sky_stride = 0 if (sky is None or sky.dim() == 2) else npx
        seg_stride = 0 if (seg is None or seg.dim() == 2) else npx
        check(self.lib.mavd_detect(self._h, flow.data_ptr(), n, imu, C.byref(self.detect_params), samples.data_ptr(),
                                   _ptr(sky), sky_stride, _ptr(seg), seg_stride, _ptr(total_out), _ptr(fixed_out),
                                   records.data_ptr(), self._stream()))
        return records

    def _host_args(self, frames, samples, n_pairs, pair_stride, sky, seg, flow_out, fixed_out, records):
        if n_pairs is None:
            n_pairs = frames.shape[0] - 1 if pair_stride == 1 else frames.shape[0] // 2
        npx = self.width * self.height
        if records is None:
            records = np.empty((n_pairs,), dtype=RECORD_DTYPE)
        for name, a in (('frames', frames), ('samples', samples), ('sky', sky), ('seg', seg),
                        ('flow_out', flow_out), ('fixed_out', fixed_out)):
            if a is not None and not a.flags['C_CONTIGUOUS']:
                raise ValueError('%s must be C-contiguous' % name)
        if samples.dtype != np.int32:
            raise ValueError('samples must be int32')
        need = n_pairs + 1 if pair_stride == 1 else 2 * n_pairs
        if frames is not None and (frames.dtype != np.uint8 or frames.shape[0] < need or
                                   tuple(frames.shape[1:]) not in ((self.height, self.width),
                                                                   (self.height, self.width, 3))):
            raise ValueError('frames must be uint8 (>=%d, %d, %d) gray or (>=%d, %d, %d, 3) BGR'
                             % (need, self.height, self.width, need, self.height, self.width))
        sky_stride = 0 if (sky is None or sky.ndim == 2) else npx
        seg_stride = 0 if (seg is None or seg.ndim == 2) else npx
        return n_pairs, records, sky_stride, seg_stride

    def process_host(self, frames: np.ndarray, imu, samples: np.ndarray, n_pairs: Optional[int] = None,
                     pair_stride: int = 1, sky: Optional[np.ndarray] = None, seg: Optional[np.ndarray] = None,
                     flow_out: Optional[np.ndarray] = None, fixed_out: Optional[np.ndarray] = None,
                     records: Optional[np.ndarray] = None) -> np.ndarray:
        """The end-to-end call: HOST buffers in, HOST records (and optional masks / flow) out.  Frames may be gray
        (F, H, W) or BGR (F, H, W, 3)."""
        if frames.ndim == 4:
            self.wait_host(0)
            records = self.submit_host(0, frames, imu, samples, n_pairs, pair_stride, sky, seg, flow_out, fixed_out,
                                       records)
            self.wait_host(0)
            return records
        n_pairs, records, sky_stride, seg_stride = self._host_args(frames, samples, n_pairs, pair_stride, sky, seg,
                                                                   flow_out, fixed_out, records)
        with torch.cuda.device(self.device):
            check(self.lib.mavd_process_host(self._h, _hp(frames), n_pairs, pair_stride, imu,
                                             C.byref(self.detect_params), _hp(samples), _hp(sky), sky_stride,
                                             _hp(seg), seg_stride, _hp(flow_out), _hp(fixed_out),
                                             records.ctypes.data, self._stream()))
        return records

    def submit_host(self, slot: int, frames: np.ndarray, imu, samples: np.ndarray, n_pairs: Optional[int] = None,
                    pair_stride: int = 1, sky: Optional[np.ndarray] = None, seg: Optional[np.ndarray] = None,
                    flow_out: Optional[np.ndarray] = None, fixed_out: Optional[np.ndarray] = None,
                    records: Optional[np.ndarray] = None) -> np.ndarray:
        """Asynchronous process_host: returns at once; the outputs are valid after wait_host(slot).  Keeping
        up to _lib.HOST_SLOTS batches in flight overlaps the host<->device copies with the compute."""
        n_pairs, records, sky_stride, seg_stride = self._host_args(frames, samples, n_pairs, pair_stride, sky, seg,
                                                                   flow_out, fixed_out, records)
        submit = self.lib.mavd_submit_host_bgr if frames.ndim == 4 else self.lib.mavd_submit_host
        with torch.cuda.device(self.device):
            check(submit(self._h, slot, _hp(frames), n_pairs, pair_stride, imu,
                         C.byref(self.detect_params), _hp(samples), _hp(sky), sky_stride,
                         _hp(seg), seg_stride, _hp(flow_out), _hp(fixed_out),
                         records.ctypes.data, self._stream()))
        # keep the host buffers alive until the wait
        self._inflight[slot] = (frames, imu, samples, sky, seg, flow_out, fixed_out, records)
        return records

    def wait_host(self, slot: int) -> None:
        check(self.lib.mavd_wait_host(self._h, slot))
        self._inflight.pop(slot, None)

    def detect_host(self, flow: np.ndarray, imu, samples: np.ndarray, sky: Optional[np.ndarray] = None,
                    seg: Optional[np.ndarray] = None, fixed_out: Optional[np.ndarray] = None,
                    records: Optional[np.ndarray] = None) -> np.ndarray:
        """Detection from a HOST float32 flow (n, H, W, 2): what Processor.run_detection does with
        Dataset.get_flow_uv (processor.py:305-362)."""
        if flow.dtype != np.float32 or tuple(flow.shape[1:]) != (self.height, self.width, 2) or \
                not flow.flags['C_CONTIGUOUS']:
            raise ValueError('flow must be C-contiguous float32 (n, %d, %d, 2)' % (self.height, self.width))
        n, records, sky_stride, seg_stride = self._host_args(None, samples, flow.shape[0], 1, sky, seg, None, fixed_out,
                                                             records)
        with torch.cuda.device(self.device):
            check(self.lib.mavd_detect_host(self._h, flow.ctypes.data, n, imu, C.byref(self.detect_params),
                                            _hp(samples), _hp(sky), sky_stride, _hp(seg), seg_stride, _hp(fixed_out),
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

    def force_exact_residual(self, on: bool = True) -> None:
        """Tests only: evaluate every pixel of the residual stage in float64 (no float32 pre-decision)."""
        check(self.lib.mavd_debug_force_exact_residual(self._h, 1 if on else 0))

    def profile_enable(self, on: bool = True) -> None:
        check(self.lib.mavd_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """{kernel class: (summed device ms, timed launch groups)} since profile_enable()."""
        prof = _lib.Profile()
        check(self.lib.mavd_profile_read(self._h, C.byref(prof)))
        return {n: (prof.ms[i], int(prof.launches[i])) for i, n in enumerate(_lib.PROF_NAMES)}

    def profile_timeline(self, max_records: int = 4096):
        """[(kernel class, start ms, end ms)] of every timed launch group since profile_enable(), in launch order."""
        buf = (C.c_double * (3 * max_records))()
        n = C.c_int32()
        check(self.lib.mavd_profile_timeline(self._h, buf, max_records, C.byref(n)))
        names = _lib.PROF_NAMES
        return [(names[int(buf[3 * i])] if int(buf[3 * i]) < len(names) else str(int(buf[3 * i])), buf[3 * i + 1],
                 buf[3 * i + 2]) for i in range(n.value)]

    def launch_count(self) -> int:
        return int(self.lib.mavd_launch_count())


_SHARED: Dict[tuple, 'Engine'] = {}


def shared_engine(width: int, height: int, params: Optional[Dict] = None, max_pairs: int = 1,
                  device: Optional[int] = None) -> 'Engine':
    """One cached Engine per (device, W, H, Farneback parameters, max_pairs) for the reference-named classes
    (Farneback, Detector, FocusOfExpansion, Processor), which all work on the same frame geometry."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    p = dict(REFERENCE_PARAMS)
    if params:
        p.update(params)
    key = (device, int(width), int(height), int(max_pairs), tuple(sorted(p.items())))
    eng = _SHARED.get(key)
    if eng is None or not eng._h.value:
        eng = Engine(width, height, p, max_pairs=max_pairs, device=device)
        _SHARED[key] = eng
    return eng
