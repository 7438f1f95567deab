#!/usr/bin/env python
"""Condense ncu CSV exports into the small tables kept under profiles/.

  launches : ncu --metrics gpu__time_duration.sum --csv log  -> per-kernel totals and share of one step
  raw      : ncu -i rep --page raw --csv                     -> one row of key counters per profiled launch
  source   : ncu -i rep --page source --csv                  -> stall samples / instructions per code phase

Usage: python tools/ncu_summary.py launches <csv> [steps_in_capture]
       python tools/ncu_summary.py raw <csv>
       python tools/ncu_summary.py source <csv>
"""
from __future__ import annotations

import csv
import re
import sys
from collections import OrderedDict, defaultdict


def short(name: str) -> str:
    name = re.sub(r'\(.*$', '', name.replace('mavd::', '').replace('void ', ''))
    return name.strip()


def launches(path: str, steps: float = 1.0) -> str:
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) >= 15 and r[0].isdigit()]
    tot = defaultdict(float)
    cnt = defaultdict(int)
    grid = {}
    for r in rows:
        if r[12] != 'gpu__time_duration.sum':
            continue
        k = short(r[4])
        v = float(r[14]) / (1e3 if r[13] == 'ns' else 1.0)
        tot[k] += v
        cnt[k] += 1
        grid.setdefault(k, r[8])
    total = sum(tot.values())
    out = ['| kernel | launches | total us | us/launch | share | first grid |', '|---|---:|---:|---:|---:|---|']
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append('| %s | %d | %.1f | %.1f | %.1f%% | %s |' % (k, cnt[k], v, v / cnt[k], 100 * v / total, grid[k]))
    out.append('| **all** | %d | %.1f | | 100%% | (%.1f us per step over %g steps) |'
               % (sum(cnt.values()), total, total / steps, steps))
    return '\n'.join(out)


RAW_KEYS = OrderedDict([
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram rd'),
    ('dram__bytes_write.sum', 'dram wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu wavefronts %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem'),
    ('launch__shared_mem_per_block_static', 'static smem'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('smsp__inst_executed.sum', 'warp inst'),
    ('sm__inst_executed_pipe_fp64.sum', 'fp64 inst'),
])


def raw(path: str) -> str:
    rows = list(csv.reader(open(path, errors='replace')))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in RAW_KEYS if k in idx]
    out = ['| kernel | grid | ' + ' | '.join(RAW_KEYS[k] for k in cols) + ' |', '|---|---|' + '---:|' * len(cols)]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        cells = []
        for k in cols:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v)
                v = ('%.0f' % f) if abs(f) >= 100 or f == int(f) else ('%.2f' % f)
            except ValueError:
                pass
            cells.append((v + ' ' + u).strip())
        out.append('| %s | %s | %s |' % (short(r[idx['Kernel Name']]), r[idx['Grid Size']], ' | '.join(cells)))
    return '\n'.join(out)


def source(path: str) -> str:
    rows = list(csv.reader(open(path, errors='replace')))
    name = rows[0][1] if len(rows[0]) > 1 else ''
    hdr = rows[1]
    i_s, i_a, i_i = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    body = [r for r in rows[2:] if len(r) > i_a]
    total_s = sum(int(r[i_a]) for r in body) or 1
    total_i = sum(int(r[i_i]) for r in body) or 1
    out = ['kernel: ' + short(name), '', '| phase ends at (SASS line, marker) | stall samples | share | warp instructions | share |',
           '|---|---:|---:|---:|---:|']
    s = i = 0
    for n, r in enumerate(body):
        s += int(r[i_a])
        i += int(r[i_i])
        src = r[i_s].strip()
        if any(m in src for m in ('BAR.SYNC', 'UTMALDG', 'SYNCS.PHASECHK', 'EXIT')):
            out.append('| %d `%s` | %d | %.1f%% | %d | %.1f%% |' % (n, src[:48], s, 100 * s / total_s, i, 100 * i / total_i))
            s = i = 0
    ops = defaultdict(int)
    for r in body:
        t = r[i_s].strip().split()
        if not t:
            continue
        op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
        ops[op] += int(r[i_i])
    out += ['', 'instruction mix: ' + ', '.join('%s %.1f%%' % (k, 100 * v / total_i)
                                               for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14])]
    return '\n'.join(out)


if __name__ == '__main__':
    mode, path = sys.argv[1], sys.argv[2]
    if mode == 'launches':
        print(launches(path, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0))
    elif mode == 'raw':
        print(raw(path))
    else:
        print(source(path))
