"""GPU: the reference-named Python classes (the drop-in surface of SURVEY.md §8b) against the golden vectors
produced by the reference's own modules (tests/golden/make_golden.py) and against the oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class FakeDataset:
    """The two IMU accessors Detector.derotate reads (same stand-in make_golden.py used with the reference)."""

    def __init__(self, w, h, ang, dt):
        self.capture_size = (w, h)
        self._ang = np.asarray(ang, np.float64)
        self._dt = dt

    def get_delta_time(self, i):
        return self._dt

    def get_angular_difference(self, a, b):
        return self._ang.copy()


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_detector_and_foe_classes_reproduce_the_reference(golden_dir, ci):
    """Same call sequence, same seeds as make_golden.py ran on the reference classes."""
    from mav_detection_b200 import im_helpers
    from mav_detection_b200.detector import Detector
    from mav_detection_b200.focus_of_expansion import FocusOfExpansion
    from mav_detection_b200.lucas_kanade import LucasKanade
    g = np.load(os.path.join(golden_dir, 'detect_%d.npz' % ci))
    flow, fi = g['flow'], int(g['frame_index'])
    h, w = flow.shape[:2]
    ds = FakeDataset(w, h, g['ang'], float(g['dt']))
    np.random.seed(7)
    det = Detector(ds)
    assert det.algorithm == Detector.Algorithm.ESSENTIAL and not det.is_homography_based()
    assert Detector.Algorithm.ESSENTIAL.value == (5,)                 # 1-tuple values, as in the reference
    foe_obj = FocusOfExpansion(LucasKanade(np.zeros((h, w, 3), np.uint8)))
    fd = det.derotate(fi - 1, fi, flow)
    assert fd.dtype == g['flow_derot'].dtype and np.array_equal(fd, g['flow_derot'])
    np.random.seed(int(g['seed']))
    foe = foe_obj.get_FOE_dense(fd)
    assert foe == (g['foe'][0], g['foe'][1])
    # the call consumed exactly the reference's draws: the next draw of the global stream agrees
    nxt = np.random.randint(0, 1 << 30)
    np.random.seed(int(g['seed']))
    np.random.randint(0, h, 2000)
    np.random.randint(0, w, 2000)
    assert nxt == np.random.randint(0, 1 << 30)
    phi = foe_obj.get_phi(fd, foe)
    assert phi.dtype == g['phi'].dtype and phi.shape == g['phi'].shape
    if phi.dtype == np.float64:
        assert np.abs(phi - g['phi']).max() < 1e-9
    else:
        from oracle import detect_np as dn
        assert np.array_equal(phi, dn.get_phi(fd, foe, cr_arccos_f32=True))       # see tests/test_gpu_detect.py
        ulp = np.abs(phi.view(np.int32).astype(np.int64) - g['phi'].view(np.int32).astype(np.int64))
        assert ulp.max() <= 8
    assert abs(float(foe_obj.max_flow) - float(g['phi'].max())) < 1e-4
    mag = im_helpers.get_magnitude(fd)
    assert mag.dtype == fd.dtype and np.array_equal(mag, np.linalg.norm(fd, axis=-1))
    box = im_helpers.get_simple_bounding_box(g['seg'])
    x0, y0, x1, y1 = (int(v) for v in g['bbox'])
    assert tuple(box.topleft) == (x0, y0) and tuple(box.size) == (x1 - x0, y1 - y0)
    assert im_helpers.calculate_tpr_fpr(g['seg'], 255 * g['total_mask']) == (g['rates'][0], g['rates'][1])
    assert im_helpers.calculate_tpr_fpr(g['seg'], 255 * g['estimate_fixed']) == (g['rates'][2], g['rates'][3])


def test_im_helpers_edge_cases():
    from mav_detection_b200 import im_helpers
    from oracle import detect_np as dn
    empty = np.zeros((40, 60), np.uint8)
    box = im_helpers.get_simple_bounding_box(empty)
    assert tuple(box.topleft) == (-1, -1) and tuple(box.size) == (0, 0)
    rng = np.random.default_rng(3)
    img = (rng.random((37, 53, 3)) * 255).astype(np.uint8)
    img[:5] = 0
    img[:, :7] = 0
    box = im_helpers.get_simple_bounding_box(img)
    x0, y0, x1, y1 = dn.simple_bounding_box(img)
    assert tuple(box.topleft) == (x0, y0) and tuple(box.size) == (x1 - x0, y1 - y0)
    tpr, fpr = im_helpers.calculate_tpr_fpr(empty, 255 * np.ones((40, 60), np.int64))
    assert np.isnan(tpr) and fpr == 1.0                                # no positives: 0/0, as the reference's NumPy division
    f32 = rng.normal(0, 3, (33, 45, 2)).astype(np.float32)
    assert np.array_equal(im_helpers.get_magnitude(f32), np.linalg.norm(f32, axis=-1))


def test_ransac_method_edge_cases():
    from mav_detection_b200.focus_of_expansion import FocusOfExpansion
    from mav_detection_b200.lucas_kanade import LucasKanade
    from oracle import detect_np as dn
    foe_obj = FocusOfExpansion(LucasKanade(np.zeros((48, 64, 3), np.uint8)))
    assert foe_obj.ransac(np.zeros((0, 2))) == (0.0, 0.0)
    assert foe_obj.ransac(np.array([[5.0, 6.0]])) == (0.0, 0.0)        # a single estimate has no support
    E = np.array([[10.0, 10.0], [500.0, 2.0], [12.0, 9.0], [11.0, 30.0]])
    assert foe_obj.ransac(E) == dn.ransac(E)


@pytest.mark.parametrize('flow_source', ['dataset', 'farneback', 'farneback-bgr'])
def test_processor_run_detection_matches_oracle_chain(tmp_path, flow_source):
    """Processor.run_detection over a synthetic sequence: FrameResult per frame == the oracle chained over the same
    frames with the same global random stream; JSON files carry the keys Validator.load_results reads."""
    import logging
    cv2 = pytest.importorskip('cv2')
    from mav_detection_b200 import engine, synth
    from mav_detection_b200.frame_result import KEYS
    from mav_detection_b200.processor import Processor
    from mav_detection_b200.run_config import RunConfig
    from oracle import detect_np as dn
    W, H, F = 320, 240, 8
    params = dict(engine.SAMPLE_PARAMS)
    seq = synth.make_sequence(W, H, F, seq=2, with_rotation=True)
    p = params
    cvflow = np.stack([cv2.calcOpticalFlowFarneback(seq.frames[i], seq.frames[i + 1], None, p['pyr_scale'], p['levels'],
                                                    p['winsize'], p['iterations'], p['poly_n'], p['poly_sigma'], p['flags'])
                       for i in range(F - 1)])
    bgr = flow_source.endswith('-bgr')       # frames delivered as (H, W, 3) BGR, converted to gray on the device
    flow_source = flow_source.split('-')[0]
    ds = synth.SyntheticDataset(seq, flows=cvflow, results_path=str(tmp_path / 'results'), bgr=bgr)
    RunConfig.register_dataset(RunConfig.DatasetType.SIMULATION, lambda logger, sequence: ds)
    cfg = RunConfig(logging.getLogger('test'), 'simulation', 'synthetic', False, False, False, True, False, False,
                    'FLOW_FOE_CLUSTERING')
    np.random.seed(99)
    proc = Processor(cfg, flow_source=flow_source, batch_frames=3, farneback_params=params)
    res = proc.run_detection()
    proc.release()
    assert sorted(res) == list(range(F - 1)) and res is proc.detection_results and cfg.results[0] is res[0]

    # the oracle with the same global random stream: constructor draws first (detector.py:33-36,
    # lucas_kanade.py:32, focus_of_expansion.py:24,26), then 2 x 2000 indices per frame in frame order
    np.random.seed(99)
    np.random.randint(20, H - 20, 1000); np.random.randint(20, W - 20, 1000)
    n = 2000 + 2000 // 3
    np.random.randint(0, 255, (n, 3)); np.random.randint(0, 255, (n, 3)); np.random.randint(0, n, n)
    for i in range(F - 1):
        ry, rx = dn.draw_sample_indices(H, W)
        if flow_source == 'dataset':
            flow = cvflow[i]
        else:
            eng = proc.engine
            import torch
            flow = eng.farneback(torch.from_numpy(seq.frames[i:i + 2]).to(eng.device))[0].cpu().numpy()
            assert np.linalg.norm(flow - cvflow[i], axis=-1).mean() < 1e-3
        # frame 0 stays float32 in the reference: the oracle evaluates its arccos correctly rounded, as the CUDA path
        # does (NumPy's float32 arccos is up to 2 ulp off and CPU-dependent, see tests/test_gpu_detect.py)
        fd, foe, phi, total, fixed = dn.frame_pipeline(i, flow, seq.omega[i], seq.dt, seq.sky_mask, ry, rx,
                                                       cr_arccos_f32=True)
        fr = res[i]
        assert fr.foe_dense == foe, (i, fr.foe_dense, foe)
        seg = seq.segmentation[i]
        tpr, fpr = dn.tpr_fpr(seg, total)
        tprf, fprf = dn.tpr_fpr(seg, fixed)
        assert (fr.tpr, fr.fpr, fr.tpr_fixed, fr.fpr_fixed) == (tpr, fpr, tprf, fprf), i
        assert fr.drone_size_pixels == int((seg > 127).sum())
        assert fr.time == i * seq.dt and fr.foe_gt == seq.foe
        x0, y0, x1, y1 = dn.simple_bounding_box(seg)
        cx, cy = x0 + (x1 - x0) / 2, y0 + (y1 - y0) / 2
        assert fr.center_phi == np.rad2deg(np.arctan2(cy - seq.foe[1], cx - seq.foe[0]))
        avg = fd[seg > 127].astype(np.float64).mean(axis=0)
        assert np.allclose(fr.drone_flow_pixels, avg, rtol=1e-9, atol=1e-12)
        with open(tmp_path / 'results' / ('image_%05d.json' % i)) as f:
            js = json.load(f)
        assert sorted(js) == sorted(KEYS)
        assert js['foe_dense'] == [foe[0], foe[1]]


class FakeCapture:
    """cv2.VideoCapture stand-in: read() -> (ok, BGR frame)."""

    def __init__(self, frames_bgr):
        self.frames, self.i = frames_bgr, 0

    def read(self):
        f = self.frames[min(self.i, len(self.frames) - 1)]
        self.i += 1
        return True, f.copy()


def test_farneback_class_process():
    """Farneback(capture, output).process() (farneback.py:72-107): gray conversion, flow with the reference's
    parameters, visualisation bytes."""
    cv2 = pytest.importorskip('cv2')
    from mav_detection_b200 import synth
    from mav_detection_b200.farneback import PROCESS_PARAMS, Farneback
    from oracle import vis_np
    W, H = 320, 240
    seq = synth.make_sequence(W, H, 3, seq=6)
    rng = np.random.default_rng(8)
    tint = rng.integers(0, 40, (3, H, W, 3)).astype(np.int16)
    bgr = np.clip(seq.frames[..., None].astype(np.int16) + tint - 20, 0, 255).astype(np.uint8)
    fb = Farneback(FakeCapture(list(bgr)), None)
    assert np.array_equal(fb.prevgray, cv2.cvtColor(bgr[0], cv2.COLOR_BGR2GRAY))
    p = PROCESS_PARAMS
    for t in (1, 2):
        out = fb.process()
        g0, g1 = cv2.cvtColor(bgr[t - 1], cv2.COLOR_BGR2GRAY), cv2.cvtColor(bgr[t], cv2.COLOR_BGR2GRAY)
        assert np.array_equal(fb.prevgray, g1)
        ref_flow = cv2.calcOpticalFlowFarneback(g0, g1, None, p['pyr_scale'], p['levels'], p['winsize'], p['iterations'],
                                                p['poly_n'], p['poly_sigma'], p['flags'])
        flow = fb.flow.cpu().numpy()
        assert np.linalg.norm(flow - ref_flow, axis=-1).mean() < 1e-3
        assert out.shape == (H, W, 3) and out.dtype == np.uint8
        # bytes: exactly the restated cv2 arithmetic applied to OUR flow ...
        mine, invalid = vis_np.process_visualisation(flow)
        assert not invalid and np.array_equal(out, mine), int((out != mine).any(-1).sum())
        # ... and the reference's cv2 statements on cv2's flow agree except where a 1e-6 px flow difference
        # crosses a truncation boundary
        ref_img, _ = vis_np.process_visualisation_cv2(ref_flow, (H, W, 3))
        assert (out == ref_img).all(-1).mean() > 0.99
    # invalid_frame branch (farneback.py:97-98): an all-zero value channel returns the previous result.  It needs an
    # exactly uniform flow magnitude, which Farneback never produces (even identical frames give non-zero flow at
    # the right/bottom border, in cv2 as well), so the kernel's counter is checked directly on synthetic fields.
    import torch
    from mav_detection_b200 import _lib
    lib = _lib.load()
    for name, field, expect_invalid in (('zero', np.zeros((H, W, 2), np.float32), True),
                                        ('const', np.full((H, W, 2), 1.5, np.float32), True),
                                        ('ramp', flow, False)):
        d = torch.from_numpy(np.ascontiguousarray(field)).cuda()
        out_d = torch.empty((H, W, 3), dtype=torch.uint8, device='cuda')
        scratch = torch.zeros((3,), dtype=torch.int32, device='cuda')
        _lib.check(lib.mavd_flow_vis(d.data_ptr(), H * W, out_d.data_ptr(), scratch.data_ptr(), None))
        torch.cuda.synchronize()
        ref_img, ref_invalid = vis_np.process_visualisation(field)
        assert ref_invalid == expect_invalid, name
        assert (int(scratch[2].item()) == 0) == expect_invalid, name
        assert np.array_equal(out_d.cpu().numpy(), ref_img), name


def test_detector_derotate_float64_input():
    """Detector.derotate on a flow that is already float64 (detector.py:117: flow - derotation, no promotion)."""
    from mav_detection_b200.detector import Detector
    from oracle import detect_np as dn
    w, h = 96, 72
    ang, dt = np.array([0.003, -0.002, 0.001]), 0.04
    rng = np.random.default_rng(2)
    flow64 = rng.normal(0, 3, (h, w, 2))
    det = Detector(FakeDataset(w, h, ang, dt))
    out = det.derotate(4, 5, flow64)
    assert out.dtype == np.float64 and np.array_equal(out, dn.derotate(5, flow64, ang, dt))
    assert det.derotate(-1, 0, flow64) is flow64                 # frame index < 1: untouched, same object
