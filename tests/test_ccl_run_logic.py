"""CPU: the contact enumeration of the run-based labelling kernels (csrc/detect.cu: ccl_init_kernel / ccl_merge_kernel),
restated with Python integers as 128-bit unit masks and a plain union-find, against the labelling oracle.  It checks the
LOGIC the CUDA passes implement — the left continuation init links without an atomic, the contacts with the row above
merge unions, and that together they connect exactly the 8-connected components — on widths that are / are not multiples of the
128-pixel unit, including units that straddle rows and rows shorter than a unit."""
import numpy as np
import pytest

UNIT = 128


def _bit(v, t):
    return (v >> t) & 1 if t >= 0 else 0


def _first_set_at_or_above(v, t):
    v >>= t
    assert v
    return t + (v & -v).bit_length() - 1


def _last_set_at_or_below(v, t):
    v &= (1 << (t + 1)) - 1
    assert v
    return v.bit_length() - 1


class Unit:
    def __init__(self, flat, ub, w):
        npx = flat.size

        def px(i):
            return int(flat[i] != 0) if 0 <= i < npx else 0
        self.ub, self.w = ub, w
        self.M = sum(px(ub + t) << t for t in range(UNIT))
        self.RS = sum(int((ub + t) % w == 0) << t for t in range(UNIT))
        self.x0 = ub % w
        full = (1 << UNIT) - 1
        prev = (self.M << 1) & full
        nxt = self.M >> 1
        rs_next = (self.RS >> 1) | (1 << (UNIT - 1))
        self.S = self.M & (~prev | self.RS) & full
        self.E = self.M & (~nxt | rs_next) & full
        self.U = sum(px(ub - w + t) << t for t in range(UNIT))
        self.prevpix, self.um1, self.u128 = px(ub - 1), px(ub - w - 1), px(ub - w + UNIT)
        self.Uprev = ((self.U << 1) | self.um1) & full
        self.US = self.U & (~self.Uprev | self.RS) & full


def _find(par, i):
    while par[i] != i:
        par[i] = par[par[i]]
        i = par[i]
    return i


def _union(par, a, b):
    a, b = _find(par, a), _find(par, b)
    if a != b:
        par[max(a, b)] = min(a, b)


def emulate(mask):
    """Returns (labels by first raster appearance, number of unions merge performed)."""
    h, w = mask.shape
    flat = mask.reshape(-1)
    npx = flat.size
    par = np.arange(npx)
    units = [Unit(flat, ub, w) for ub in range(0, npx, UNIT)]
    units = [u for u in units if u.M]
    for u in units:                                              # ccl_init_kernel
        left_cont = bool(_bit(u.M, 0) and not _bit(u.RS, 0) and u.prevpix)
        for t in range(UNIT):
            if _bit(u.M, t):
                par[u.ub + t] = u.ub - 1 if (t == 0 and left_cont) else u.ub + _last_set_at_or_below(u.S, t)
    unions = 0
    for u in units:                                              # ccl_merge_kernel
        if not (u.U or u.um1 or u.u128):
            continue
        left_cont = bool(u.prevpix and not _bit(u.RS, 0) and _bit(u.M, 0))
        for t in range(UNIT):
            g = u.ub + t
            mt, rs = _bit(u.M, t), _bit(u.RS, t)
            if mt and _bit(u.S, t) and not rs and not (t == 0 and left_cont) and _bit(u.Uprev, t):
                _union(par, g, g - w - 1)                                           # (B)
                unions += 1
            if _bit(u.US, t):                                                        # (A)
                if mt:
                    _union(par, g, g - w)
                    unions += 1
                elif t >= 1 and not rs and _bit(u.M, t - 1):
                    _union(par, g - 1, g - w)
                    unions += 1
            if t == UNIT - 1 and mt and u.u128 and not _bit(u.U, UNIT - 1) and (u.x0 + UNIT) % w != 0:
                _union(par, g, g - w + 1)
                unions += 1
    labels = np.zeros(npx, np.int32)
    order = {}
    for i in range(npx):
        if flat[i]:
            r = _find(par, i)
            labels[i] = order.setdefault(r, len(order) + 1)
    return labels.reshape(h, w), unions


@pytest.mark.parametrize('size', [(128, 9), (256, 7), (64, 11), (100, 13), (203, 9), (37, 21), (129, 6), (8, 40)])
def test_run_contacts_connect_exactly_the_components(size):
    from oracle import ccl_np
    w, h = size
    rng = np.random.default_rng(w * 31 + h)
    ys, xs = np.mgrid[0:h, 0:w]
    masks = [np.ones((h, w), np.uint8), ((xs + ys) % 2).astype(np.uint8), (xs % 3 == 0).astype(np.uint8),
             ((ys % 2 == 0) * (xs % 7 != 3)).astype(np.uint8), ((xs // 5 + ys // 3) % 2).astype(np.uint8)]
    masks += [(rng.random((h, w)) < p).astype(np.uint8) for p in (0.05, 0.3, 0.5, 0.7, 0.93)]
    for i, m in enumerate(masks):
        got, unions = emulate(m)
        ref, _ = ccl_np.label(m)
        assert np.array_equal(got, ref), (size, i)
        if i == 0:
            assert unions == h - 1      # a fully set frame: init links the rows, merge makes one union per row
