"""CPU: the .flo codec against files written / read by the reference itself (tests/golden/make_golden_vis_flo.py), and
the visualisation oracle (oracle/vis_np.py) against the reference's im_helpers outputs and live cv2."""
import os

import numpy as np
import pytest


def test_flo_codec_matches_the_reference_files(golden_dir, tmp_path):
    """utils.read_flow reads the file the reference's write_flow produced; utils.write_flow reproduces it byte for
    byte (/root/reference/src/utils.py:204-257)."""
    from mav_detection_b200 import utils
    g = np.load(os.path.join(golden_dir, 'flo_ref.npz'))
    ref_file = os.path.join(golden_dir, 'flo_ref.flo')
    got = utils.read_flow(ref_file)
    assert got.dtype == np.float32 and got.shape == g['flow'].shape
    assert np.array_equal(got, g['read_back']) and np.array_equal(got, g['flow'])
    mine = str(tmp_path / 'mine.flo')
    utils.write_flow(mine, g['flow'])
    assert open(mine, 'rb').read() == open(ref_file, 'rb').read()
    utils.write_flow(mine, g['flow'][..., 0], g['flow'][..., 1])          # the (u, v) form
    assert open(mine, 'rb').read() == open(ref_file, 'rb').read()
    bad = str(tmp_path / 'bad.flo')
    with open(bad, 'wb') as f:
        f.write(np.array([1.0], np.float32).tobytes() + b'\0' * 8)
    with pytest.raises(AssertionError):
        utils.read_flow(bad)


def test_flo_cache_host_batches_and_provider(golden_dir, tmp_path):
    """FloCache writes the reference's path convention and byte layout; CachedFlowDataset serves get_flow_uv."""
    from mav_detection_b200 import flow_cache, utils
    g = np.load(os.path.join(golden_dir, 'flo_ref.npz'))
    rng = np.random.default_rng(3)
    flows = rng.normal(0, 2, (5,) + g['flow'].shape).astype(np.float32)
    flows[2] = g['flow']
    cache = flow_cache.FloCache(str(tmp_path / 'output' / 'inference' / 'run.epoch-0-flow-field'))
    cache.put_batch(10, flows[:3])
    cache.put_batch(13, flows[3:])
    assert sorted(os.listdir(cache.directory)) == ['%06d.flo' % i for i in range(10, 15)]
    assert open(flow_cache.flo_path(cache.directory, 12), 'rb').read() == \
        open(os.path.join(golden_dir, 'flo_ref.flo'), 'rb').read()
    for k in range(5):
        assert np.array_equal(cache.get_flow_uv(10 + k), flows[k])
        assert np.array_equal(utils.read_flow(flow_cache.flo_path(cache.directory, 10 + k)), flows[k])

    class Inner:
        N = 7

        def get_flow_uv(self, i):
            raise AssertionError('the cache must answer')
    ds = flow_cache.CachedFlowDataset(Inner(), cache)
    assert ds.N == 7 and np.array_equal(ds.get_flow_uv(11), flows[1])


def test_vis_oracle_matches_the_reference_outputs(golden_dir):
    """oracle/vis_np.py: phi image (to_rgb + JET) and mask overlay, against what the reference's im_helpers / the
    statements of processor.py:385-392 produced."""
    from oracle import vis_np
    g = np.load(os.path.join(golden_dir, 'vis_ref.npz'))
    assert np.array_equal(vis_np.JET_LUT, g['jet_lut'])
    for name, phi in (('f64', g['phi64']), ('f32', g['phi32'])):
        rgb = vis_np.to_rgb(phi, 180.0)
        assert np.array_equal(rgb, g['phi_rgb_' + name]), name
        jet = vis_np.apply_colormap_jet(rgb)
        assert np.array_equal(jet, g['phi_jet_' + name]), name
        assert np.array_equal(vis_np.apply_colormap_jet(rgb, 180.0), g['phi_jet_max_' + name]), name
    vis, mask_rgb = vis_np.mask_overlay(g['frame'], g['fixed'])
    assert np.array_equal(vis, g['mask_vis']) and np.array_equal(mask_rgb, g['result_img'])
    assert np.array_equal(vis_np.mask_overlay(g['frame'], np.zeros_like(g['fixed']))[1], g['empty_img'])


def test_vis_oracle_against_live_cv2():
    """The JET table and the addWeighted rounding for every byte value, against the cv2 on this machine."""
    cv2 = pytest.importorskip('cv2')
    from oracle import vis_np
    v = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(vis_np.apply_colormap_jet(np.repeat(v[..., None], 3, 2)),
                          cv2.applyColorMap(np.repeat(v[..., None], 3, 2), cv2.COLORMAP_JET))
    frame = np.stack([v, v[::-1], v.T], -1).copy()
    for fixed in (np.zeros((16, 16), bool), np.ones((16, 16), bool)):
        painted = frame.copy()
        painted[fixed] = (150, 0, 150)
        ref = cv2.addWeighted(frame, 0.2, painted, 0.8, 0.0)
        assert np.array_equal(vis_np.mask_overlay(frame, fixed)[0], ref)
