"""Frame pairs of a sequence sharded over the GPUs of one box.

A pair (t, t+1) depends only on its two frames, so rank r of G takes a contiguous range of pairs and
runs the whole hot path on it with no data-path collective.  Two things cross ranks:
  * the FoE sample indices: the reference draws them from ONE process-global legacy NumPy stream in
    frame order (focus_of_expansion.py:69-71), so every rank replays the same seeded stream and slices
    out its own frames' rows — results are identical to a single-process run;
  * the per-pair records (FoE point, detection boxes, scalar metrics; ~0.8 KB each), gathered with one
    all_gather per run (NCCL over NVLink on GPUs, gloo in the CPU tests).
Frame data never crosses GPUs."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import N_SAMPLE_PAIRS


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank: ceil(n/world) items per rank, the tail ranks may be short or empty."""
    per = -(-n_items // world) if world > 0 else n_items
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def draw_all_samples(n_frames: int, height: int, width: int, seed: Optional[int] = None,
                     rng: Optional[np.random.RandomState] = None) -> np.ndarray:
    """(n_frames, 4000) int32 [ry(2000) | rx(2000)] drawn in frame order from a legacy RandomState —
    the stream np.random.randint produces after np.random.seed(seed)."""
    rs = rng if rng is not None else np.random.RandomState(seed)
    out = np.empty((n_frames, 4 * N_SAMPLE_PAIRS), np.int32)
    for i in range(n_frames):
        out[i, :2 * N_SAMPLE_PAIRS] = rs.randint(0, height, 2 * N_SAMPLE_PAIRS)
        out[i, 2 * N_SAMPLE_PAIRS:] = rs.randint(0, width, 2 * N_SAMPLE_PAIRS)
    return out


def gather_records(local: np.ndarray, counts: Sequence[int], group=None) -> np.ndarray:
    """all_gather of per-rank record arrays of (possibly) different lengths, returned in pair order.
    `local` is a 1-D structured array; every rank must pass the same `counts` (pairs per rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and counts[rank] == local.shape[0]
    item = local.dtype.itemsize
    width = max(max(counts), 1) * item
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    buf = torch.zeros((width,), dtype=torch.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[:raw.shape[0]] = torch.from_numpy(raw.copy())
    buf = buf.to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = [outs[r].cpu().numpy()[:counts[r] * item].view(local.dtype) for r in range(world)]
    return np.concatenate(parts) if parts else local[:0]


def run_sharded(frames: np.ndarray, process_batch: Callable[[np.ndarray, int, np.ndarray], np.ndarray],
                seed: int, batch_pairs: int = 16, group=None) -> np.ndarray:
    """Run `process_batch(frames[lo:hi+1], first_pair_index, samples[lo:hi]) -> records` over this rank's
    contiguous share of the sequence's pairs, batch by batch, and return ALL ranks' records in pair order.
    `frames` is the whole (F, H, W) sequence (each rank only touches its own slice plus one frame)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_pairs = frames.shape[0] - 1
    height, width = frames.shape[1], frames.shape[2]
    samples = draw_all_samples(n_pairs, height, width, seed)       # every rank replays the same stream
    lo, hi = shard_range(n_pairs, world, rank)
    out: List[np.ndarray] = []
    for b0 in range(lo, hi, batch_pairs):
        b1 = min(b0 + batch_pairs, hi)
        out.append(process_batch(frames[b0:b1 + 1], b0, samples[b0:b1]))
    if out:
        local = np.concatenate(out)
    else:
        from .engine import RECORD_DTYPE
        local = np.empty((0,), RECORD_DTYPE)
    if world == 1:
        return local
    counts = [shard_range(n_pairs, world, r)[1] - shard_range(n_pairs, world, r)[0] for r in range(world)]
    return gather_records(local, counts, group)


def compare_records(got: np.ndarray, ref: np.ndarray) -> List[str]:
    """Names of the record fields that differ.  Every field must be bit-equal except the float64 flow sums, which
    are accumulated with atomics (summation order varies from run to run): those to 1e-12 relative."""
    if got.shape != ref.shape:
        return ['shape %s != %s' % (got.shape, ref.shape)]
    bad = []
    for name in ('foe', 'n_intersections', 'n_labels', 'boxes'):
        if not np.array_equal(got[name], ref[name]):
            bad.append(name)
    for name in ref['stats'].dtype.names:
        a, b = got['stats'][name], ref['stats'][name]
        same = np.allclose(a, b, rtol=1e-12, atol=1e-9) if name in ('seg_flow_sum', 'gt_flow_sum') else np.array_equal(a, b)
        if not same:
            bad.append('stats.' + name)
    return bad


def parity_check(device: int, width: int = 640, height: int = 480, n_frames: int = 33, batch_pairs: int = 8,
                 seed: int = 77, group=None) -> Tuple[bool, str]:
    """Multi-GPU parity, run by every rank of an initialised process group: a synthetic sequence's pairs are sharded
    over the ranks (run_sharded), the records gathered, and rank 0 compares them field by field with its own
    single-GPU run of the whole sequence.  Returns (ok, message) on every rank (the verdict is broadcast)."""
    import torch
    import torch.distributed as dist
    from . import engine, synth
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    params = dict(engine.SAMPLE_PARAMS)
    seq = synth.make_sequence(width, height, n_frames, seq=3, with_rotation=True)
    eng = engine.Engine(width, height, params, max_pairs=batch_pairs, device=device)

    def batch(frames, first_pair, samples):
        n = frames.shape[0] - 1
        imu = engine.make_imu(n, seq.omega[first_pair + 1:first_pair + 1 + n], seq.dt,
                              derotate=[(first_pair + i) >= 1 for i in range(n)])
        seg = np.ascontiguousarray(seq.segmentation[first_pair + 1:first_pair + 1 + n])
        return eng.process_host(np.ascontiguousarray(frames), imu, np.ascontiguousarray(samples), seg=seg).copy()

    got = run_sharded(seq.frames, batch, seed=seed, batch_pairs=batch_pairs, group=group)
    ok, msg = True, ''
    if rank == 0:
        n_pairs = n_frames - 1
        samples = draw_all_samples(n_pairs, height, width, seed=seed)
        ref = np.concatenate([batch(seq.frames[b:min(b + batch_pairs, n_pairs) + 1], b,
                                    samples[b:min(b + batch_pairs, n_pairs)]) for b in range(0, n_pairs, batch_pairs)])
        bad = compare_records(got, ref)
        ok = not bad
        msg = ('%d pairs of %dx%d sharded over %d ranks == single-GPU run, field by field' % (n_pairs, width, height, world)
               if ok else 'sharded records differ from the single-GPU run in: ' + ', '.join(bad))
    eng.close()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32,
                            device=torch.device('cuda', device) if dist.get_backend(group) == 'nccl' else 'cpu')
        dist.broadcast(flag, src=0, group=group)
        ok = bool(flag.item())
    return ok, msg
