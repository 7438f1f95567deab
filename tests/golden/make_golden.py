"""Regenerates tests/golden/*.npz from the REFERENCE ITSELF, run in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py

* farneback_*.npz  : cv2.calcOpticalFlowFarneback (the call at
                     /root/reference/src/farneback.py:76-80) on small synthetic pairs.
* detect_*.npz     : the reference's own Detector.derotate, FocusOfExpansion.get_FOE_dense /
                     get_phi, im_helpers.get_simple_bounding_box / calculate_tpr_fpr imported from
                     /root/reference/src (stub modules for matplotlib/imutils/flow_vis/airsim), plus
                     the mask expressions of processor.py:333-341 evaluated verbatim on those outputs.
The script also asserts that oracle/ reproduces every stored vector, so a stale oracle cannot be
committed together with fresh goldens.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for m in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'imutils', 'flow_vis', 'airsim',
              'airsim.types']:
        sys.modules[m] = types.ModuleType(m)
    sys.modules['matplotlib.pyplot'].rcParams = {}
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['airsim.types'].Vector3r = object
    np.lib.angle = np.angle
    sys.path.insert(0, '/root/reference/src')
    import detector
    import focus_of_expansion
    import im_helpers
    import utils
    return detector, focus_of_expansion, im_helpers, utils


class FakeDataset:
    def __init__(self, w, h, ang, dt):
        self.capture_size = (w, h)
        self._ang = np.asarray(ang, np.float64)
        self._dt = dt

    def get_delta_time(self, i):
        return self._dt

    def get_angular_difference(self, a, b):
        return self._ang.copy()


def radial_flow(w, h, foe, rate, rng, noise=0.05):
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    f = np.stack([(xs - foe[0]) * rate, (ys - foe[1]) * rate], -1)
    f += rng.normal(0, noise, f.shape)
    return f.astype(np.float32)


def main():
    import cv2
    from mav_detection_b200 import synth
    from oracle import detect_np as dn
    from oracle import farneback_np as fb

    # ---------------- Farneback ----------------
    cases = {
        'ref':   (0.4, 1, 12, 10, 8, 1.2, 0),     # farneback.py:78-80
        'c2':    (0.5, 5, 15, 3, 5, 1.2, 0),
        'gauss': (0.5, 3, 15, 3, 5, 1.1, 256),
        'odd':   (0.8, 4, 9, 2, 7, 1.5, 0),
        'lvl0':  (0.5, 0, 15, 3, 5, 1.2, 0),
    }
    for i, (name, p) in enumerate(cases.items()):
        w, h = (160, 112) if name != 'odd' else (150, 101)
        a, b = synth.make_pair(w, h, seq=i)
        flow = cv2.calcOpticalFlowFarneback(a, b, None, *p)
        mine = fb.calc_optical_flow_farneback(a, b, None, *p)
        epe = np.linalg.norm(flow - mine, axis=-1)
        assert epe.mean() < 1e-5 and epe.max() < 1e-4, (name, epe.mean(), epe.max())
        np.savez_compressed(os.path.join(HERE, 'farneback_%s.npz' % name), prev=a, next=b,
                            params=np.array(p, np.float64), flow=flow)
        print('farneback', name, 'oracle-vs-cv2 mean EPE %.2e max %.2e' % (epe.mean(), epe.max()))

    # ---------------- FoE / phi / masks ----------------
    detector, foe_mod, im_helpers, utils = import_reference()
    import lucas_kanade
    for ci, (w, h, ang, dt, frame_index) in enumerate([
            (160, 120, (0.0, 0.0, 0.0), 1 / 30, 3),
            (160, 120, (0.002, -0.001, 0.0005), 1 / 30, 3),
            (200, 96, (0.001, 0.003, -0.002), 0.05, 0),      # frame 0: float32 passthrough
            (96, 130, (-0.004, 0.001, 0.002), 0.04, 7)]):
        rng = np.random.default_rng(100 + ci)
        foe_true = (0.4 * w, 0.45 * h)
        flow = radial_flow(w, h, foe_true, 0.06, rng)
        flow[40:52, 100 % (w - 20):100 % (w - 20) + 16] += np.float32(4.0)   # the "MAV"
        flow[5:10, 5:10] = 0                                                  # zero-flow patch (8/mag = inf)
        sky = np.zeros((h, w), bool)
        sky[:15, :] = True
        seg = np.zeros((h, w), np.uint8)
        seg[40:52, 100 % (w - 20):100 % (w - 20) + 16] = 255

        ds = FakeDataset(w, h, ang, dt)
        np.random.seed(7)
        det = detector.Detector(ds)
        foe_obj = foe_mod.FocusOfExpansion(lucas_kanade.LucasKanade(np.zeros((h, w, 3), np.uint8)))
        fd = det.derotate(frame_index - 1, frame_index, flow)
        seed = 1234 + ci
        np.random.seed(seed)
        state = np.random.get_state()
        foe = foe_obj.get_FOE_dense(fd)
        np.random.set_state(state)
        ry, rx = dn.draw_sample_indices(h, w)
        phi = foe_obj.get_phi(fd, foe)
        flow_mag = im_helpers.get_magnitude(fd)
        # processor.py:333-341, evaluated verbatim
        angle_threshold_max = phi > (0.25 + (0.5 + 8 / flow_mag))
        angle_threshold_min = phi < (0.25 - (0.5 + 8 / flow_mag))
        angle_threshold = np.logical_or(angle_threshold_min, angle_threshold_max)
        total_mask = (flow_mag > 0.5) * ~sky * angle_threshold
        estimate_fixed = phi * (flow_mag > 1.0) * ~sky > 15
        bbox = im_helpers.get_simple_bounding_box(seg)
        tpr_fixed, fpr_fixed = im_helpers.calculate_tpr_fpr(seg, 255 * estimate_fixed)
        tpr, fpr = im_helpers.calculate_tpr_fpr(seg, 255 * total_mask)

        # the oracle must reproduce all of it
        ofd, ofoe, ophi, otot, ofix = dn.frame_pipeline(frame_index, flow, np.array(ang), dt, sky, ry, rx)
        assert ofd.dtype == fd.dtype and np.array_equal(ofd, fd)
        assert ofoe == tuple(float(v) for v in foe), (ofoe, foe)
        assert np.array_equal(ophi, phi) and ophi.dtype == phi.dtype
        assert np.array_equal(otot, total_mask) and np.array_equal(ofix, estimate_fixed)
        x0, y0, x1, y1 = dn.simple_bounding_box(seg)
        assert (x0, y0) == tuple(bbox.topleft) and (x1 - x0, y1 - y0) == tuple(bbox.size)
        assert dn.tpr_fpr(seg, total_mask) == (tpr, fpr)
        assert dn.tpr_fpr(seg, estimate_fixed) == (tpr_fixed, fpr_fixed)

        np.savez_compressed(
            os.path.join(HERE, 'detect_%d.npz' % ci), flow=flow, ang=np.array(ang, np.float64),
            dt=np.float64(dt), frame_index=np.int64(frame_index), sky=sky, seg=seg, seed=np.int64(seed),
            ry=ry.astype(np.int32), rx=rx.astype(np.int32), flow_derot=fd, foe=np.array(foe, np.float64),
            phi=phi, total_mask=total_mask, estimate_fixed=estimate_fixed,
            bbox=np.array([x0, y0, x1, y1], np.int32),
            rates=np.array([tpr, fpr, tpr_fixed, fpr_fixed], np.float64))
        print('detect', ci, 'foe', foe, 'dtype', fd.dtype, 'mask px', int(total_mask.sum()),
              int(estimate_fixed.sum()))


if __name__ == '__main__':
    main()
