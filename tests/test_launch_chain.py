"""Static check of the programmatic-dependent-launch invariant (mav_detection_b200/csrc/common.cuh): every kernel that is
launched through launch_chained() must execute griddepcontrol.wait (pdl_entry) as its FIRST statement — before any early
exit, so that no grid of the chain can complete ahead of its predecessor, and before any global access."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mav_detection_b200', 'csrc')


def _sources():
    return '\n'.join(open(os.path.join(CSRC, f)).read() for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh')))


def _chained_kernels(src):
    names = set()
    for m in re.finditer(r'launch_chained\(', src):
        # second argument: a kernel name, a template-id, or `cond ? kernel_a : kernel_b`
        j, depth, start = m.end(), 0, None
        args, cur = [], []
        while True:
            c = src[j]
            if c in '(<[':
                depth += 1
            elif c in ')>]':
                if depth == 0:
                    args.append(''.join(cur))
                    break
                depth -= 1
            if c == ',' and depth == 0:
                args.append(''.join(cur))
                cur = []
                if len(args) == 2:
                    break
            else:
                cur.append(c)
            j += 1
        if len(args) >= 2:
            names.update(re.findall(r'([A-Za-z_][A-Za-z_0-9]*_kernel)\b', args[1]))
    return names


def test_every_chained_kernel_waits_first():
    src = _sources()
    names = _chained_kernels(src)
    assert len(names) >= 20, sorted(names)          # the whole per-batch chain goes through launch_chained
    for n in sorted(names):
        m = re.search(r'__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?' + n + r'\s*\(', src)
        assert m, 'no definition of %s' % n
        j, depth = m.end(), 1
        while depth:
            depth += (src[j] == '(') - (src[j] == ')')
            j += 1
        body = src[src.index('{', j) + 1:].lstrip()
        assert body.startswith('pdl_entry();'), '%s does not start with pdl_entry()' % n


def test_chain_is_broken_after_event_waits():
    """A programmatic edge may only join two chained kernels: after an event wait (or a memset / copy, next test) on a
    stream the next launch must be a plain one (pdl_break)."""
    for f in ('farneback.cu', 'api.cu', 'detect.cu'):
        lines = open(os.path.join(CSRC, f)).read().split('\n')
        for i, line in enumerate(lines):
            if 'cudaStreamWaitEvent(' in line and 'MAVD_CUDA' in line and 'h->s_main' not in line and 'user' not in line:
                window = '\n'.join(lines[i:i + 3])
                if 's_in' in line or 's_out' in line or 'S.ev' in window or 'ev_in' in window or 'ev_done' in window:
                    continue                        # copy-stream plumbing of the host calls: no chained kernels there
                assert 'pdl_break' in window, '%s:%d: event wait without pdl_break' % (f, i + 1)


def test_chain_is_broken_after_memsets():
    for f in ('farneback.cu', 'api.cu', 'detect.cu'):
        lines = open(os.path.join(CSRC, f)).read().split('\n')
        for i, line in enumerate(lines):
            if 'cudaMemsetAsync(' not in line and 'cudaMemcpyAsync(' not in line:
                continue
            for k in range(i + 1, min(i + 8, len(lines))):
                if '<<<' in lines[k] or 'return' in lines[k]:
                    break
                if 'launch_chained(' in lines[k]:
                    assert any('pdl_break' in l for l in lines[i:k]), '%s:%d: memset / copy, then a chained launch' % (f, i + 1)
                    break
