#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): mav_detection_b200.sharded.parity_check —
a synthetic sequence's frame pairs are sharded over the ranks, each rank runs the whole hot path on its share, the
per-pair records are all_gathered, and rank 0 compares them field by field with a single-GPU run of the full
sequence.  bench.py runs the same check before timing whenever it is launched on more than one GPU.
Prints one line: SHARDED_OK / SHARDED_MISMATCH."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mav_detection_b200 import sharded  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok, msg = sharded.parity_check(local, n_frames=41)
    if dist.get_rank() == 0:
        print('SHARDED_OK' if ok else 'SHARDED_MISMATCH', msg)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
