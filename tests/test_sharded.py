"""Host logic of the multi-GPU path (mav_detection_b200/sharded.py) on CPU: world_size-2 gloo run of
run_sharded with a stand-in batch function, compared with the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_batch(frames, first_pair, samples):
    """Deterministic stand-in for Engine.process_host: a record per pair that depends on the pair's two
    frames, its global index and its sample indices (so a wrong shard, order or RNG slice is caught)."""
    from mav_detection_b200.engine import RECORD_DTYPE
    n = frames.shape[0] - 1
    rec = np.zeros((n,), RECORD_DTYPE)
    for k in range(n):
        rec[k]['foe'][0] = float(frames[k].astype(np.int64).sum() * 3 + frames[k + 1].astype(np.int64).sum())
        rec[k]['foe'][1] = float(first_pair + k)
        rec[k]['n_intersections'] = int(samples[k].astype(np.int64).sum() % 1000003)
        rec[k]['n_labels'] = int(samples[k, 0])
    return rec


def _worker(rank, world, port, n_frames, out_dir):
    import torch.distributed as dist
    from mav_detection_b200 import sharded
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 255, (n_frames, 12, 16), dtype=np.uint8)
    rec = sharded.run_sharded(frames, _fake_batch, seed=11, batch_pairs=3)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), rec.view(np.uint8))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [12, 8, 2])
def test_two_rank_gloo_equals_single_process(tmp_path, n_frames):
    import torch.multiprocessing as mp
    from mav_detection_b200 import sharded
    from mav_detection_b200.engine import RECORD_DTYPE
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 255, (n_frames, 12, 16), dtype=np.uint8)
    single = sharded.run_sharded(frames, _fake_batch, seed=11, batch_pairs=3)       # no process group: world 1
    assert single.shape[0] == n_frames - 1
    mp.spawn(_worker, args=(2, _free_port(), n_frames, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), 'rank%d.npy' % r)).view(RECORD_DTYPE).reshape(-1)
        assert got.shape == single.shape
        assert got.tobytes() == single.tobytes(), 'rank %d gathered records differ from the single-process run' % r


def test_shard_ranges_cover_everything_once():
    from mav_detection_b200 import sharded
    for n in (0, 1, 7, 16, 999):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharded.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_sample_stream_is_the_legacy_global_stream():
    """draw_all_samples(seed) == what np.random.seed(seed) + FocusOfExpansion.get_FOE_dense would draw, frame by frame
    (/root/reference/src/focus_of_expansion.py:69-71: rows first, then columns)."""
    from mav_detection_b200 import sharded
    h, w, n = 48, 64, 5
    got = sharded.draw_all_samples(n, h, w, seed=123)
    np.random.seed(123)
    for i in range(n):
        ry = np.random.randint(0, h, 2000)
        rx = np.random.randint(0, w, 2000)
        assert np.array_equal(got[i, :2000], ry) and np.array_equal(got[i, 2000:], rx)
