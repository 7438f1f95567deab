"""GPU: derotation, FoE, residual angle, masks, metrics and components against the oracle / golden
vectors.  Bars (BASELINE.json): FoE within 0.5 px under identical seeds (we require exact equality of
the discrete selection), masks bit-exact when fed the reference flow, labels bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load(golden_dir, ci):
    return np.load(os.path.join(golden_dir, 'detect_%d.npz' % ci))


def _engine(w, h, n=1):
    from mav_detection_b200 import engine
    return engine.Engine(w, h, engine.SAMPLE_PARAMS, max_pairs=n)


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_golden_frame(golden_dir, ci):
    import torch
    from mav_detection_b200 import engine
    g = _load(golden_dir, ci)
    h, w = g['flow'].shape[:2]
    fi = int(g['frame_index'])
    eng = _engine(w, h)
    flow = torch.from_numpy(g['flow'][None]).cuda()
    imu = engine.make_imu(1, g['ang'][None], float(g['dt']), derotate=[fi >= 1])
    samples = torch.from_numpy(np.concatenate([g['ry'], g['rx']])[None].astype(np.int32)).cuda()
    # derotation: bit-exact float64
    if fi >= 1:
        fd = eng.derotate(flow, imu)[0].cpu().numpy()
        assert np.array_equal(fd, g['flow_derot'])
    # FoE: the same discrete choice -> identical doubles
    foe, cnt = eng.foe(flow, imu, samples)
    foe_h = foe.cpu().numpy()[0]
    assert np.array_equal(foe_h, g['foe']), (foe_h, g['foe'])
    # phi and masks, fed the reference's FoE
    phi, total, fixed, stats = eng.residual_masks(flow, imu, foe, sky=torch.from_numpy(g['sky']).cuda(),
                                                  seg=torch.from_numpy(g['seg']).cuda())
    total = total[0].cpu().numpy().astype(bool)
    fixed = fixed[0].cpu().numpy().astype(bool)
    if fi >= 1:
        p = phi[0].cpu().numpy()
        assert np.abs(p - g['phi']).max() < 1e-9            # acos within a few ulp of glibc
        assert np.array_equal(total, g['total_mask']) and np.array_equal(fixed, g['estimate_fixed'])
    else:
        p = phi.view(torch.float32)[0].reshape(-1)[:h * w].reshape(h, w).cpu().numpy()
        assert np.abs(p - g['phi']).max() < 1e-4            # float32 path: NumPy's own f32 arccos
        assert (total != g['total_mask']).sum() <= 2 and (fixed != g['estimate_fixed']).sum() <= 2
    st = eng.stats_to_numpy(stats)[0]
    assert st['n_total'] == total.sum() and st['n_fixed'] == fixed.sum()
    assert abs(st['max_phi'] - g['phi'].max()) < 1e-4
    assert tuple(st['seg_bbox']) == tuple(g['bbox'])
    seg = g['seg']
    assert st['positives'] == (seg > 127).sum() and st['negatives'] == ((255 - seg) > 127).sum()
    if fi >= 1:
        r = g['rates']
        assert st['tp_total'] / st['positives'] == r[0] and st['fp_total'] / st['negatives'] == r[1]
        assert st['tp_fixed'] / st['positives'] == r[2] and st['fp_fixed'] / st['negatives'] == r[3]
        ref_sum = g['flow_derot'][seg > 127].sum(axis=0)
        assert np.allclose(st['seg_flow_sum'], ref_sum, rtol=1e-9)
    eng.close()


def test_foe_many_seeds_and_degenerate_flows():
    import torch
    from mav_detection_b200 import engine
    from oracle import detect_np as dn
    h, w = 120, 160
    rng = np.random.default_rng(9)
    ys, xs = np.mgrid[0:h, 0:w]
    flows, samples, expect = [], [], []
    for i in range(8):
        f = np.stack([(xs - 70) * 0.07, (ys - 50) * 0.07], -1) + rng.normal(0, 0.1 * (i % 3), (h, w, 2))
        if i == 5:
            f[:] = 0.0                      # nothing above the magnitude threshold -> (0, 0)
        if i == 6:
            f[:] = (3.0, 0.0)               # parallel lines: div == 0 everywhere -> (0, 0)
        f = f.astype(np.float32)
        np.random.seed(100 + i)
        ry, rx = dn.draw_sample_indices(h, w)
        flows.append(f)
        samples.append(np.concatenate([ry, rx]).astype(np.int32))
        expect.append(dn.foe_dense(f, ry, rx))          # float32 path (derotate off)
    eng = _engine(w, h, 8)
    imu = engine.make_imu(8, derotate=False)
    foe, cnt = eng.foe(torch.from_numpy(np.stack(flows)).cuda(), imu, torch.from_numpy(np.stack(samples)).cuda())
    got = foe.cpu().numpy()
    assert np.array_equal(got, np.array(expect))
    assert tuple(got[5]) == (0.0, 0.0) and tuple(got[6]) == (0.0, 0.0)
    eng.close()


def test_masks_bit_exact_on_larger_derotated_field():
    """640x480 noisy radial field with rotation: many pixels sit near both thresholds."""
    import torch
    from mav_detection_b200 import engine
    from oracle import detect_np as dn
    h, w = 480, 640
    rng = np.random.default_rng(4)
    ys, xs = np.mgrid[0:h, 0:w]
    flow = (np.stack([(xs - 250) * 0.02, (ys - 210) * 0.02], -1) + rng.normal(0, 0.3, (h, w, 2))).astype(np.float32)
    flow[100:110, 100:120] = 0
    ang, dt = np.array([0.002, -0.001, 0.0005]), 1 / 30
    np.random.seed(3)
    ry, rx = dn.draw_sample_indices(h, w)
    sky = np.zeros((h, w), bool)
    sky[:40] = True
    fd, foe_ref, phi_ref, total_ref, fixed_ref = dn.frame_pipeline(5, flow, ang, dt, sky, ry, rx)
    eng = _engine(w, h)
    imu = engine.make_imu(1, ang[None], dt, derotate=True)
    fl = torch.from_numpy(flow[None]).cuda()
    foe, _ = eng.foe(fl, imu, torch.from_numpy(np.concatenate([ry, rx])[None].astype(np.int32)).cuda())
    assert tuple(foe.cpu().numpy()[0]) == foe_ref
    phi, total, fixed, _ = eng.residual_masks(fl, imu, foe, sky=torch.from_numpy(sky).cuda())
    assert np.array_equal(total[0].cpu().numpy().astype(bool), total_ref)
    assert np.array_equal(fixed[0].cpu().numpy().astype(bool), fixed_ref)
    assert total_ref.sum() > 1000 and fixed_ref.sum() > 1000
    eng.close()


@pytest.mark.parametrize('shape,density', [((48, 64), 0.5), ((120, 160), 0.3), ((257, 333), 0.55), ((480, 640), 0.02),
                                           ((96, 128), 1.0), ((96, 128), 0.0)])
def test_ccl_labels_bit_exact(shape, density):
    import torch
    from oracle import ccl_np
    rng = np.random.default_rng(11)
    h, w = shape
    masks = (rng.random((3, h, w)) < density).astype(np.uint8)
    masks[1, 10:30, 5:25] = 1          # a solid blob
    masks[2, ::2, :] = 0               # stripes
    eng = _engine(w, h, 3)
    labels, boxes, cnt = eng.ccl(torch.from_numpy(masks).cuda())
    labels, boxes, cnt = labels.cpu().numpy(), boxes.cpu().numpy(), cnt.cpu().numpy()
    for i in range(3):
        ref, stats = ccl_np.label(masks[i])
        assert cnt[i] == ref.max()
        assert np.array_equal(labels[i], ref)
        k = min(32, stats.shape[0])
        assert np.array_equal(boxes[i, :k], stats[:k])
        assert not boxes[i, k:].any()
    eng.close()


def test_bgr2gray_matches_opencv_fixed_point():
    import torch
    rng = np.random.default_rng(2)
    bgr = rng.integers(0, 256, (3, 37, 53, 3), dtype=np.uint8)
    eng = _engine(64, 48)
    got = eng.bgr2gray(torch.from_numpy(bgr).cuda()).cpu().numpy()
    b, g, r = (bgr[..., i].astype(np.uint32) for i in range(3))
    ref = ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
    assert np.array_equal(got, ref)
    try:
        import cv2
        assert np.array_equal(got[0], cv2.cvtColor(bgr[0], cv2.COLOR_BGR2GRAY))
    except ImportError:
        pass
    eng.close()
