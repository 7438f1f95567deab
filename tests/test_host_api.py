"""CPU: host-side pieces of the drop-in surface that need no device — RunConfig, FrameResult, utils
(Rectangle, get_json, .flo codec), enum shapes, the RNG side effects of the constructors."""
import json
import logging

import numpy as np
import pytest


def test_run_config_api():
    from mav_detection_b200.run_config import RunConfig
    cfg = RunConfig(logging.getLogger('t'), 'midgard', 'seq', False, False, False, True, False, False, 'FLOW_RADIAL')
    assert cfg.mode == RunConfig.Mode.FLOW_RADIAL and cfg.uses_nn_for_detection()
    assert RunConfig.Mode.FLOW_UV.value == (1,) and str(RunConfig.Mode.FLOW_UV) == 'FLOW_UV'      # run_config.py:14-22
    assert str(RunConfig.DatasetType.MIDGARD) == 'MIDGARD'
    assert cfg.get_dataset_type('simulation') == RunConfig.DatasetType.SIMULATION
    assert str(cfg) == 'midgard/seq/FLOW_RADIAL'
    with pytest.raises(ValueError, match='is not a valid mode type'):
        cfg.get_mode('nope')
    with pytest.raises(ValueError, match='is not a valid dataset type'):
        cfg.get_dataset_type('nope')
    cfg2 = RunConfig(logging.getLogger('t'), 'simulation', 's', False, False, False, True, False, False,
                     'FLOW_FOE_CLUSTERING')
    assert not cfg2.uses_nn_for_detection()
    RunConfig.dataset_factories.pop(RunConfig.DatasetType.VIS_DRONE, None)
    cfg3 = RunConfig(logging.getLogger('t'), 'vis_drone', 's', False, False, False, True, False, False, 'FLOW_UV')
    with pytest.raises(ValueError, match='Invalid dataset type'):
        cfg3.get_dataset()


def test_frame_result_keys_are_what_the_validator_reads():
    """validator.py:141-152 copies exactly these keys out of results/image_%05d.json."""
    from mav_detection_b200 import utils
    from mav_detection_b200.frame_result import KEYS, FrameResult
    fr = FrameResult()
    assert sorted(vars(fr)) == sorted(KEYS) and len(KEYS) == 12
    fr.foe_dense = (np.float64(1.5), np.float64(2.5))
    fr.drone_size_pixels = np.int64(7)           # NumPy integers go through default= -> str, as in utils.py:350-361
    js = json.loads(json.dumps(utils.get_json(fr), indent=4, sort_keys=True))
    assert js['foe_dense'] == [1.5, 2.5] and js['drone_size_pixels'] == '7' and sorted(js) == sorted(KEYS)


def test_rectangle_semantics():
    from mav_detection_b200.utils import Rectangle
    r = Rectangle.from_points((10, 20), (30, 50))
    assert r.size == (20, 30) and r.get_center() == (20.0, 35.0)           # size excludes the last pixel (utils.py:26-30)
    assert r.get_bottomright() == (30, 50) and r.get_area() == 600
    assert Rectangle.from_points((-1, -1), (-1, -1)).get_center() == (-1.0, -1.0)
    assert Rectangle.from_center((5, 5), (4, 2)).get_topleft() == (3.0, 4.0)
    a, b = Rectangle((0, 0), (10, 10)), Rectangle((5, 5), (10, 10))
    assert abs(Rectangle.calculate_iou(a, b) - 25 / 175) < 1e-12


def test_flo_codec_round_trip(tmp_path):
    from mav_detection_b200 import utils
    rng = np.random.default_rng(1)
    flow = rng.normal(0, 3, (37, 53, 2)).astype(np.float32)
    path = str(tmp_path / 'a.flo')
    utils.write_flow(path, flow)
    raw = open(path, 'rb').read()
    assert np.frombuffer(raw[:4], np.float32)[0] == np.float32(202021.25)            # utils.py:204-223 layout
    assert tuple(np.frombuffer(raw[4:12], np.int32)) == (53, 37) and len(raw) == 12 + flow.nbytes
    assert np.array_equal(utils.read_flow(path), flow)
    utils.write_flow(path, flow[..., 0], flow[..., 1])
    assert np.array_equal(utils.read_flow(path), flow)
    open(path, 'wb').write(b'\x00' * 64)
    with pytest.raises(AssertionError):
        utils.read_flow(path)


def test_constructor_draws_keep_the_global_random_stream():
    """Detector / LucasKanade / FocusOfExpansion constructors consume the legacy global generator exactly as the
    reference's do (detector.py:33-36, lucas_kanade.py:32, focus_of_expansion.py:24,26)."""
    from mav_detection_b200.detector import Detector
    from mav_detection_b200.focus_of_expansion import FocusOfExpansion

    class DS:
        capture_size = (64, 48)
    np.random.seed(3)
    det = Detector(DS())
    foe = FocusOfExpansion(det.lucas_kanade)
    after = np.random.randint(0, 1 << 30)
    np.random.seed(3)
    sy = np.random.randint(20, 48 - 20, 1000)
    sx = np.random.randint(20, 64 - 20, 1000)
    n = 2000 + 2000 // 3
    np.random.randint(0, 255, (n, 3))
    np.random.randint(0, 255, (n, 3))
    np.random.randint(0, n, n)
    assert after == np.random.randint(0, 1 << 30)
    assert np.array_equal(det.sample_y, sy) and np.array_equal(det.sample_x, sx)
    assert foe.magnitude_threshold == 2.5 and foe.ransac_threshold == 30.0 and (foe.flow_height, foe.flow_width) == (48, 64)
    idx = FocusOfExpansion.draw_sample_indices(48, 64)
    assert idx.dtype == np.int32 and idx.shape == (4000,) and idx[:2000].max() < 48 and idx[2000:].max() < 64
