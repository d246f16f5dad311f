"""CPU model of k_gather_pass0_sparse (csrc/rs_kernels.cu): the first visits of pass 0 get their patches by SEARCH (the
K-1 smallest offsets, in table order, among the earlier target points, their wrapped aliases and the context pixels)
instead of by scanning the sorted offsets table.  Checked here without a GPU:
  * the 64-bit key (x^2 + y^2, rank in reverse row-major order) orders offsets exactly as the sorted table does
    (lib/engine.c:465-497 through the host restatement rs_host_sorted_offsets, itself equal to the oracle's table);
  * selecting by key among the valued pixels == scanning the table (lib/synthesize.h:189-241), with and without tiling.
The kernel itself is compared bit for bit through whole images in the GPU parity suite."""
import ctypes as C

import numpy as np
import pytest

from resynthesizer_b200 import api


def table(tw, th, cw, ch):
    L = api.lib()
    ow, oh = min(tw, cw), min(th, ch)
    n = (2 * ow - 1) * (2 * oh - 1)
    a = np.zeros((n, 2), np.int32)
    assert L.rs_host_sorted_offsets(tw, th, cw, ch, a.ctypes.data, n) == n
    return a, ow, oh


def key(dx, dy, ow, oh):
    return ((dx * dx + dy * dy) << 32) | ((oh - 1 - dy) * (2 * ow - 1) + (ow - 1 - dx))


@pytest.mark.parametrize("dims", [(7, 5, 7, 5), (16, 16, 9, 12), (33, 20, 40, 17), (64, 64, 64, 64)])
def test_key_order_is_table_order(built_lib, dims):
    a, ow, oh = table(*dims)
    k = np.array([key(int(x), int(y), ow, oh) for x, y in a], dtype=object)
    assert all(k[i] < k[i + 1] for i in range(len(k) - 1))
    assert len(set(map(tuple, a))) == len(a)


def scan(a, valued, px, py, tw, th, htile, vtile, K1):
    """The reference's rule: offsets in table order (entry 0 = the point itself), wrap or clip, keep valued pixels."""
    out = []
    for dx, dy in a[1:]:
        x, y = px + int(dx), py + int(dy)
        if x < 0:
            if not htile: continue
            x += tw
        elif x >= tw:
            if not htile: continue
            x -= tw
        if y < 0:
            if not vtile: continue
            y += th
        elif y >= th:
            if not vtile: continue
            y -= th
        if valued[y, x]:
            out.append((int(dx), int(dy), y * tw + x))
            if len(out) == K1: break
    return out


def search(earlier, ctx, px, py, tw, th, ow, oh, htile, vtile, K1):
    """What the kernel does: candidates = earlier points (+ aliases when tiling) and context pixels, K1 smallest keys."""
    cand = []
    for qx, qy in earlier:
        dx1, dy1 = qx - px, qy - py
        for ay in range(2 if vtile else 1):
            for ax in range(2 if htile else 1):
                if (ax and dx1 == 0) or (ay and dy1 == 0): continue
                dx = (dx1 - tw if dx1 > 0 else dx1 + tw) if ax else dx1
                dy = (dy1 - th if dy1 > 0 else dy1 + th) if ay else dy1
                if abs(dx) < ow and abs(dy) < oh:
                    cand.append((key(dx, dy, ow, oh), dx, dy, qy * tw + qx))
    if not (htile or vtile):
        ys, xs = np.nonzero(ctx)
        for x, y in zip(xs.tolist(), ys.tolist()):
            dx, dy = x - px, y - py
            if abs(dx) < ow and abs(dy) < oh:
                cand.append((key(dx, dy, ow, oh), dx, dy, y * tw + x))
    cand.sort()
    return [(dx, dy, q) for _, dx, dy, q in cand[:K1]]


@pytest.mark.parametrize("tiles", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("dims", [(24, 20, 24, 20), (31, 17, 12, 40)])
def test_search_equals_scan(built_lib, dims, tiles):
    tw, th, cw, ch = dims
    htile, vtile = tiles
    a, ow, oh = table(*dims)
    rng = np.random.default_rng(7 + tw + 2 * htile + vtile)
    for trial in range(12):
        ctx = np.zeros((th, tw), bool)
        if not (htile or vtile) and trial % 2:
            ctx[:, : 3 + trial // 2] = True      # context on one side, as beside a hole
            ctx[rng.integers(0, th, 5), rng.integers(0, tw, 5)] = True
        free = np.argwhere(~ctx)
        order = free[rng.permutation(len(free))][: 40]
        v = int(rng.integers(0, len(order)))
        py, px = (int(t) for t in order[v])
        earlier = [(int(x), int(y)) for y, x in order[:v]]
        valued = ctx.copy()
        for x, y in earlier:
            valued[y, x] = True
        for K1 in (1, 8, 29):
            want = scan(a, valued, px, py, tw, th, htile, vtile, K1)
            got = search(earlier, ctx, px, py, tw, th, ow, oh, htile, vtile, K1)
            assert got == want, (dims, tiles, trial, K1)
