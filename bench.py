#!/usr/bin/env python
"""Benchmark of the mav-detection hot path (Farneback flow -> derotate -> FoE -> phi/masks -> components).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1|c3|c4]

One "step" = one pass of the whole hot path over one batch of frame pairs taken from a synthetic
sequence that is resident in HBM (device arm) / in pinned host memory (e2e).  Prints ONE JSON line.
Metric: BASELINE.json -> 1080p frame-pairs/sec (flow+FoE+mask), plus % of HBM roofline of the fused
Farneback iteration kernel.  `--impl reference` times the reference's own CPU path (cv2's Farneback +
the NumPy restatement of the reference's FoE/phi/mask code in oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

JSON_OUT = sys.stdout     # main() swaps this for a private handle on the real stdout

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, farneback params, pairs per step, label)
    'c2': (1920, 1080, dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0), 64,
           'C2: synthetic AirSim-like 1920x1080 sequence, Farneback (0.5,5,15,3,5,1.2,0) 6 pyramid images'),
    'c2ref': (1920, 1080, dict(pyr_scale=0.4, levels=1, winsize=12, iterations=10, poly_n=8, poly_sigma=1.2, flags=0), 32,
              "C2': 1920x1080 with the reference's own parameters (0.4,1,12,10,8,1.2,0)"),
    'c1': (640, 480, dict(pyr_scale=0.4, levels=1, winsize=12, iterations=10, poly_n=8, poly_sigma=1.2, flags=0), 1,
           'C1: one 640x480 pair, reference parameters'),
    'c3': (640, 480, dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0), 64,
           'C3: 640x480, 64 pairs per launch'),
    'c4': (3840, 2160, dict(pyr_scale=0.5, levels=7, winsize=15, iterations=10, poly_n=5, poly_sigma=1.2, flags=0), 8,
           'C4: 3840x2160, 7 pyramid images, winsize 15, 10 iterations'),
}
N_BATCHES = 4          # distinct resident batches cycled through (working set per step >> L2 anyway)


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def build_workload(name: str, rank: int = 0, pairs: int = 0):
    from mav_detection_b200 import synth
    W, H, params, B, label = WORKLOADS[name]
    B = pairs or B
    n_frames = N_BATCHES * B + 1
    seq = synth.make_sequence(W, H, n_frames, seq=rank, with_rotation=False)
    rs = np.random.RandomState(1000 + rank)   # legacy generator = the stream np.random.randint draws from
    samples = np.empty((N_BATCHES * B, 4000), np.int32)
    for i in range(N_BATCHES * B):
        samples[i, :2000] = rs.randint(0, H, 2000)
        samples[i, 2000:] = rs.randint(0, W, 2000)
    return dict(name=name, W=W, H=H, params=params, B=B, label=label, seq=seq, samples=samples)


def algorithmic_bytes_iter(W, H, B):
    """Fused iteration, finest level: M 20 + R0 20 + R1 20 -> flow 8 + M' 20 = 88 B/px."""
    return 88.0 * W * H * B


def cpu_one_pair(args):
    """The reference's CPU path for one pair: cv2 Farneback + restated derotate/FoE/phi/masks."""
    import cv2
    from oracle import detect_np as dn
    prev, nxt, params, omega, dt, ry, rx, idx = args
    t0 = time.perf_counter()
    flow = cv2.calcOpticalFlowFarneback(prev, nxt, None, params['pyr_scale'], params['levels'], params['winsize'],
                                        params['iterations'], params['poly_n'], params['poly_sigma'], params['flags'])
    t1 = time.perf_counter()
    sky = np.zeros(prev.shape, bool)
    dn.frame_pipeline(idx, flow, omega, dt, sky, ry, rx)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def _cpu_init():
    import cv2
    cv2.setNumThreads(1)   # one pair per core: OpenCV's Farneback core is single-threaded anyway


class CpuReference:
    """The reference's CPU path on a pool of host processes (one frame pair per process at a time)."""

    def __init__(self, wl, workers: int):
        import multiprocessing as mp
        self.wl, self.workers = wl, workers
        self.pool = mp.get_context('spawn').Pool(workers, initializer=_cpu_init)

    def jobs(self, n_pairs: int):
        seq, wl = self.wl['seq'], self.wl
        out = []
        for i in range(n_pairs):
            p = i % (seq.frames.shape[0] - 1)
            out.append((seq.frames[p], seq.frames[p + 1], wl['params'], seq.omega[p + 1], seq.dt,
                        wl['samples'][p, :2000], wl['samples'][p, 2000:], 1 + p))
        return out

    def step(self, n_pairs: int):
        """Returns (pairs/s, mean flow ms, mean post ms, seconds)."""
        jobs = self.jobs(n_pairs)
        t0 = time.perf_counter()
        res = self.pool.map(cpu_one_pair, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        return (n_pairs / dt, 1e3 * float(np.mean([r[0] for r in res])), 1e3 * float(np.mean([r[1] for r in res])), dt)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_workers() -> int:
    return max(1, min(os.cpu_count() or 1, 32))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import cv2
    wl = build_workload(args.workload, 0, args.pairs)
    workers = cpu_workers()
    per_step = workers
    ref = CpuReference(wl, workers)
    vals = []
    for s in range(args.warmup + args.steps):
        r = ref.step(per_step)
        if s >= args.warmup:
            vals.append(r)
    ref.close()
    v = float(np.sum([per_step for _ in vals]) / np.sum([x[3] for x in vals]))
    line = {
        'impl': 'reference',
        'metric': '1080p frame-pairs/sec (flow+FoE+mask)' if wl['W'] == 1920 else 'frame-pairs/sec (flow+FoE+mask)',
        'value': v, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean([x[3] for x in vals])), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['label'], 'pairs_per_step': per_step},
        'cpu_baseline': {'value': v, 'unit': 'pairs/s', 'cores': workers, 'kind': 'port',
                         'sample': '%d pairs/step on %d processes (cv2.calcOpticalFlowFarneback %s, the library the '
                                   'reference calls, + oracle/detect_np restatement of derotate/FoE/phi/masks); per '
                                   'pair: flow %.0f ms, post %.0f ms; host has %d logical CPUs'
                                   % (per_step, workers, cv2.__version__, float(np.mean([x[1] for x in vals])),
                                      float(np.mean([x[2] for x in vals])), os.cpu_count() or 0)},
        'e2e': {'value': v, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mav_detection_b200 import engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    wl = build_workload(args.workload, rank, args.pairs)
    W, H, B, params = wl['W'], wl['H'], wl['B'], wl['params']
    seq = wl['seq']
    eng = engine.Engine(W, H, params, max_pairs=B, device=local)
    dev = torch.device('cuda', local)

    # resident inputs (device arm) and pinned host inputs (e2e arm)
    frames_d = torch.from_numpy(seq.frames).to(dev)
    seg_d = torch.from_numpy(seq.segmentation).to(dev)
    samples_d = torch.from_numpy(wl['samples']).to(dev)
    fixed_d = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    records_d = torch.empty((B, engine.RECORD_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(records_d) for _ in range(world)] if world > 1 else None
    imus = []
    for b in range(N_BATCHES):
        first = b * B
        imus.append(engine.make_imu(B, seq.omega[first + 1:first + 1 + B], seq.dt,
                                    derotate=[(first + i) >= 1 for i in range(B)]))
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    def step(s):
        b = s % N_BATCHES
        f0 = b * B
        eng.process(frames_d[f0:f0 + B + 1], imus[b], samples_d[f0:f0 + B], seg=seg_d[f0 + 1:f0 + B + 1],
                    fixed_out=fixed_d, records=records_d)
        if world > 1:
            # only detection boxes / FoE points cross NVLink: one small all_gather per batch, side stream
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                dist.all_gather(gathered, records_d)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for s in range(args.warmup):
        step(s)
    sync_all()
    eng.profile_enable(True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(side)
    e1.record()
    sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    prof = eng.profile_read()
    eng.profile_enable(False)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI host call, copies inside the timed region ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    frames_h = pinned(seq.frames)
    seg_h = pinned(seq.segmentation)
    samples_h = pinned(wl['samples'])
    from mav_detection_b200._lib import HOST_SLOTS
    fixed_h = [pinned(np.zeros((B, H, W), np.uint8)) for _ in range(HOST_SLOTS)]
    rec_h = [pinned(np.zeros((B,), engine.RECORD_DTYPE).view(np.uint8)).view(engine.RECORD_DTYPE)
             for _ in range(HOST_SLOTS)]

    def step_host(s):
        # the public host call, pipelined: batch s's host->device copies overlap batch s-1's kernels and
        # batch s-2's device->host copies (mavd_submit_host / mavd_wait_host, HOST_SLOTS staging sets)
        b = s % N_BATCHES
        f0 = b * B
        slot = s % HOST_SLOTS
        eng.wait_host(slot)
        eng.submit_host(slot, frames_h[f0:f0 + B + 1], imus[b], samples_h[f0:f0 + B], seg=seg_h[f0 + 1:f0 + B + 1],
                        fixed_out=fixed_h[slot], records=rec_h[slot])

    def drain_host():
        for slot in range(HOST_SLOTS):
            eng.wait_host(slot)
    for s in range(min(args.warmup, 3)):
        step_host(s)
    drain_host()
    sync_all()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(args.warmup + s)
    drain_host()                       # every step's records and masks have landed in host memory
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s
    npx = W * H
    h2d = (B + 1) * npx + B * npx + B * 4000 * 4
    d2h = B * engine.RECORD_DTYPE.itemsize + B * npx

    if rank == 0:
        peak, peak_src = hbm_peak()
        it_ms, it_n = prof['iter_full']
        roof = None
        if it_n > 0:
            dur_s = it_ms * 1e-3 / it_n
            achieved = algorithmic_bytes_iter(W, H, B) / dur_s / 1e9
            traffic = None
            try:
                with open(os.path.join(ROOT, 'profiles', 'iter_kernel_traffic.json')) as f:
                    per_pair = json.load(f).get(args.workload + '_per_pair')
                traffic = per_pair * B if per_pair else None      # bytes per launch, like `achieved`
            except Exception:
                pass
            roof = {'bound': 'hbm', 'kernel': 'iter_box_tma_kernel<m, not-last> (fused Farneback iteration), finest level', 'achieved': achieved,
                    'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': achieved / peak,
                    'traffic': traffic, 'us_per_launch': dur_s * 1e6, 'us_per_pair': dur_s * 1e6 / B,
                    'algorithmic_bytes_per_launch': algorithmic_bytes_iter(W, H, B), 'launches_timed': it_n}
        total_ms = sum(v[0] for v in prof.values())
        shares = {k: round(v[0] / total_ms, 4) for k, v in prof.items() if v[1] > 0} if total_ms > 0 else {}
        cpu = None
        if world == 1 and not args.no_cpu:
            workers = cpu_workers()
            ref = CpuReference(wl, workers)
            ref.step(workers)                                  # warm-up: page in cv2, first touch
            v, flow_ms, post_ms, dt = ref.step(2 * workers if workers <= 8 else workers)
            ref.close()
            import cv2
            cpu = {'value': v, 'unit': 'pairs/s', 'cores': workers, 'kind': 'port',
                   'sample': '%d processes, one pair of this workload at a time (cv2 %s Farneback + oracle/detect_np); '
                             'per pair: flow %.0f ms, FoE/phi/masks %.0f ms; host has %d logical CPUs'
                             % (workers, cv2.__version__, flow_ms, post_ms, os.cpu_count() or 0)}
        line = {
            'metric': '1080p frame-pairs/sec (flow+FoE+mask)' if W == 1920 else 'frame-pairs/sec (flow+FoE+mask)',
            'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['label'], 'pairs_per_step': B, 'frames_resident': int(seq.frames.shape[0]),
                       'l2_policy': 'inputs larger than L2: each step streams a %.1f GB working set'
                                    % (eng.workspace_bytes / 1e9),
                       'sharding': 'each rank runs its own sequence; records all_gathered over NCCL per step'
                                   if world > 1 else 'single GPU'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'call': 'mavd_submit_host/mavd_wait_host (C ABI, pinned host buffers, %d batches in flight)' % HOST_SLOTS},
            'gpu_launches': int(launches),
            'roofline': roof,
            'kernel_time_shares': shares,
            'kernel_ms_per_step': {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1] > 0},
            'cpu_baseline': cpu,
        }
        print(json.dumps(line), file=JSON_OUT, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--pairs', type=int, default=0, help='frame pairs per step (default: the workload\'s)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    # stdout carries the JSON line and nothing else: keep a private handle on the real stdout for it and point file
    # descriptor 1 at stderr, so that whatever a library prints (NCCL's version banner, worker processes) lands there
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
