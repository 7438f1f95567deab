#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): a synthetic sequence's frame pairs are
sharded over the ranks (mav_detection_b200/sharded.py), each rank runs the whole hot path on its share, the
per-pair records are all_gathered, and rank 0 compares them byte for byte with a single-GPU run of the full
sequence.  Prints one line: SHARDED_OK / SHARDED_MISMATCH."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mav_detection_b200 import engine, sharded, synth  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, F, B = 640, 480, 41, 8
    params = dict(engine.SAMPLE_PARAMS)
    seq = synth.make_sequence(W, H, F, seq=3, with_rotation=True)
    eng = engine.Engine(W, H, params, max_pairs=B, device=local)

    def batch(frames, first_pair, samples):
        n = frames.shape[0] - 1
        imu = engine.make_imu(n, seq.omega[first_pair:first_pair + n], seq.dt,
                              derotate=[(first_pair + i) >= 1 for i in range(n)])
        seg = np.ascontiguousarray(seq.segmentation[first_pair:first_pair + n])
        return eng.process_host(np.ascontiguousarray(frames), imu, np.ascontiguousarray(samples), seg=seg).copy()

    got = sharded.run_sharded(seq.frames, batch, seed=77, batch_pairs=B)
    ok = True
    if rank == 0:
        samples = sharded.draw_all_samples(F - 1, H, W, seed=77)
        ref = np.concatenate([batch(seq.frames[b:min(b + B, F - 1) + 1], b, samples[b:min(b + B, F - 1)])
                              for b in range(0, F - 1, B)])
        # every field bit-equal, except the double-precision flow sums, which are accumulated with atomics
        # (summation order varies from run to run): those to 1e-12 relative
        ok = got.shape == ref.shape
        bad = []
        if ok:
            for name in ('foe', 'n_intersections', 'n_labels', 'boxes'):
                if not np.array_equal(got[name], ref[name]):
                    bad.append(name)
            for name in ref['stats'].dtype.names:
                a, b = got['stats'][name], ref['stats'][name]
                same = np.allclose(a, b, rtol=1e-12, atol=1e-9) if name == 'seg_flow_sum' else np.array_equal(a, b)
                if not same:
                    bad.append('stats.' + name)
            ok = not bad
        if bad:
            print('mismatching fields:', bad)
        print('SHARDED_OK' if ok else 'SHARDED_MISMATCH', 'world', world, 'pairs', got.shape[0],
              'foe[0]', got[0]['foe'].tolist(), 'labels', got['n_labels'][:6].tolist())
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
