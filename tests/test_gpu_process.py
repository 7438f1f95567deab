"""GPU: the whole path (device-buffer and host-buffer calls) against the oracle chained end to end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_chain(seq, params, samples, first_index=0):
    import cv2
    from oracle import ccl_np
    from oracle import detect_np as dn
    out = []
    for p in range(seq.frames.shape[0] - 1):
        flow = cv2.calcOpticalFlowFarneback(seq.frames[p], seq.frames[p + 1], None, params['pyr_scale'],
                                            params['levels'], params['winsize'], params['iterations'],
                                            params['poly_n'], params['poly_sigma'], params['flags'])
        out.append((flow,) + dn.frame_pipeline(first_index + p, flow, seq.omega[p + 1], seq.dt, seq.sky_mask,
                                               samples[p, :2000], samples[p, 2000:]))
    return out


@pytest.mark.parametrize('host', [False, True])
def test_whole_path_matches_chained_oracle(host):
    pytest.importorskip('cv2')
    import torch
    from mav_detection_b200 import engine, synth
    from oracle import ccl_np
    from oracle import detect_np as dn
    W, H, F = 640, 480, 5
    seq = synth.make_sequence(W, H, F, seq=4, with_rotation=True)
    params = dict(engine.SAMPLE_PARAMS)
    np.random.seed(21)
    samples = np.stack([np.concatenate(dn.draw_sample_indices(H, W)) for _ in range(F - 1)]).astype(np.int32)
    imu = engine.make_imu(F - 1, seq.omega[1:], seq.dt, derotate=[i >= 1 for i in range(F - 1)])
    eng = engine.Engine(W, H, params, max_pairs=F - 1)
    if host:
        flow = np.empty((F - 1, H, W, 2), np.float32)
        fixed = np.empty((F - 1, H, W), np.uint8)
        rec = eng.process_host(seq.frames, imu, samples, seg=seq.segmentation[1:].copy(), flow_out=flow,
                               fixed_out=fixed)
    else:
        flow_t = torch.empty((F - 1, H, W, 2), dtype=torch.float32, device='cuda')
        fixed_t = torch.empty((F - 1, H, W), dtype=torch.uint8, device='cuda')
        rec = eng.process(torch.from_numpy(seq.frames).cuda(), imu, torch.from_numpy(samples).cuda(),
                          seg=torch.from_numpy(seq.segmentation[1:].copy()).cuda(), flow_out=flow_t,
                          fixed_out=fixed_t)
        rec = eng.records_to_numpy(rec)
        flow, fixed = flow_t.cpu().numpy(), fixed_t.cpu().numpy()
    ref = _oracle_chain(seq, params, samples)
    for p in range(F - 1):
        rflow, fd, foe, phi, total, fix = ref[p]
        epe = np.linalg.norm(flow[p] - rflow, axis=-1).mean()
        assert epe < 1e-3, epe
        # FoE from OUR flow vs the oracle fed OUR flow: exact; vs the cv2-flow chain: within 0.5 px
        own = dn.frame_pipeline(p, flow[p], seq.omega[p + 1], seq.dt, seq.sky_mask, samples[p, :2000], samples[p, 2000:],
                                cr_arccos_f32=True)      # frame 0 is float32: see tests/test_gpu_detect.py
        assert tuple(rec[p]['foe']) == own[1]
        assert abs(rec[p]['foe'][0] - foe[0]) < 0.5 and abs(rec[p]['foe'][1] - foe[1]) < 0.5
        assert np.array_equal(fixed[p].astype(bool), own[4]), p
        assert rec[p]['stats']['n_total'] == own[3].sum(), p
        lab, stats = ccl_np.label(fixed[p])
        assert rec[p]['n_labels'] == lab.max()
        k = min(32, stats.shape[0])
        assert np.array_equal(rec[p]['boxes'][:k], stats[:k])
        assert tuple(rec[p]['stats']['seg_bbox']) == dn.simple_bounding_box(seq.segmentation[p + 1])
    eng.close()


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_empty_batches_are_no_ops():
    """n = 0 through every batched entry point: status OK, nothing written, no launch."""
    import torch
    from mav_detection_b200 import engine
    W, H = 64, 48
    eng = engine.Engine(W, H, engine.SAMPLE_PARAMS, max_pairs=2)
    before = eng.launch_count()
    one = torch.zeros((1, H, W), dtype=torch.uint8, device='cuda')
    flow = eng.farneback(one)                                   # one frame = zero pairs
    assert tuple(flow.shape) == (0, H, W, 2)
    imu = engine.make_imu(1)
    samples = torch.zeros((0, 4000), dtype=torch.int32, device='cuda')
    rec = eng.process(one, imu, samples, n_pairs=0)
    assert tuple(rec.shape) == (0, engine.RECORD_DTYPE.itemsize)
    rec = eng.detect(torch.zeros((0, H, W, 2), dtype=torch.float32, device='cuda'), imu, samples)
    assert rec.shape[0] == 0
    assert eng.launch_count() == before
    with pytest.raises(ValueError):
        eng.farneback(torch.zeros((4, H, W), dtype=torch.uint8, device='cuda'))      # 3 pairs > max_pairs 2
    eng.close()
