"""not-gpu: host preparation of libresynthesizer_b200.so (no CUDA call) must produce the reference's arrays:
metric tables, sorted neighbour offsets, target visit order for all 9 modes, pass schedule, format indices --
compared with the oracle array by array."""
import ctypes as C

import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api


@pytest.mark.parametrize("sens,mw", [(0.117, 0.5), (0.05, 0.0), (0.3, 0.25012680), (1.0, 1.0), (0.117, 0.39778528)])
def test_metric_tables(built_oracle, built_lib, sens, mw):
    L = api.lib(); port = R.load_port()
    c1 = np.zeros(512, np.uint16); m1 = np.zeros(512, np.uint32)
    c2 = np.zeros(512, np.uint16); m2 = np.zeros(512, np.uint32)
    L.rs_host_metric_tables(sens, mw, c1.ctypes.data, m1.ctypes.data)
    port.port_luts(sens, mw, c2.ctypes.data, m2.ctypes.data)
    assert (c1 == c2).all() and (m1 == m2).all()
    assert c1[0] == 65535 and c1[256] == 0
    # both functions are even: the device indexes them by |difference|
    assert (c1[257:] == c1[255:0:-1]).all() and (m1[257:] == m1[255:0:-1]).all()


@pytest.mark.parametrize("dims", [(3, 3, 3, 3), (5, 7, 9, 4), (64, 48, 32, 32), (130, 97, 200, 120)])
def test_sorted_offsets(built_oracle, built_lib, dims):
    L = api.lib(); port = R.load_port()
    tw, th, cw, ch = dims
    w, h = min(tw, cw), min(th, ch)
    n = (2 * w - 1) * (2 * h - 1)
    a = np.zeros((n, 2), np.int32); b = np.zeros((n, 2), np.int32)
    assert L.rs_host_sorted_offsets(tw, th, cw, ch, a.ctypes.data, n) == n
    assert port.port_offsets(tw, th, cw, ch, b.ctypes.data, n) == n
    assert (a == b).all()
    assert tuple(a[0]) == (0, 0) and [tuple(x) for x in a[1:5]] == [(0, 1), (1, 0), (-1, 0), (0, -1)]  # SURVEY A-3
    d = (a.astype(np.int64) ** 2).sum(axis=1)
    assert (np.diff(d) >= 0).all()


def _points(w, h, hole=None, seed=0):
    ys, xs = np.mgrid[0:h, 0:w]
    sel = np.ones((h, w), bool)
    if hole == "ring":
        sel = ((xs - w // 2) ** 2 + (ys - h // 2) ** 2 < (min(w, h) // 2) ** 2) & ((xs - w // 2) ** 2 + (ys - h // 2) ** 2 > 9)
    elif hole == "random":
        sel = np.random.RandomState(seed).rand(h, w) < 0.4
    return np.stack([xs[sel], ys[sel]], axis=1).astype(np.int32).copy()


@pytest.mark.parametrize("mode", list(range(9)))
@pytest.mark.parametrize("shape", [(1, 1, None), (2, 1, None), (17, 13, None), (64, 40, "ring"), (90, 70, "random")])
def test_target_order(built_oracle, built_lib, mode, shape):
    L = api.lib(); port = R.load_port()
    pts = _points(*shape)
    a, b = pts.copy(), pts.copy()
    assert L.rs_host_order_targets(mode, a.ctypes.data, len(a), 1198472) == 0
    assert port.port_order(mode, b.ctypes.data, len(b), 1198472) == 0
    assert (a == b).all()
    assert sorted(map(tuple, a)) == sorted(map(tuple, pts))


@pytest.mark.parametrize("mode", [0, 2, 3, 5])
def test_target_order_large(built_oracle, built_lib, mode):
    """Above 2^18 points the draws come from a producer thread and the brushfire keys from several: same order."""
    L = api.lib(); port = R.load_port()
    pts = _points(760, 620, "ring")
    assert len(pts) > (1 << 18)
    a, b = pts.copy(), pts.copy()
    assert L.rs_host_order_targets(mode, a.ctypes.data, len(a), 1198472) == 0
    assert port.port_order(mode, b.ctypes.data, len(b), 1198472) == 0
    assert (a == b).all()


def test_target_order_bad_mode(built_lib):
    L = api.lib()
    a = _points(4, 4)
    assert L.rs_host_order_targets(9, a.ctypes.data, len(a), 1) == abi.IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE


@pytest.mark.parametrize("n", [1, 2, 3, 4096, 65536, 1048576, 4194304])
def test_pass_schedule(built_lib, n):
    L = api.lib()
    ends = (C.c_uint32 * 6)()
    total = L.rs_host_pass_schedule(n, ends)
    want, m = [n], n
    for _ in range(5):
        want.append(m); m = m * 3 // 4
    assert list(ends) == want and total == sum(want)
    if n == 4096:
        assert total == 16592


def test_format_indices(built_oracle, built_lib):
    L = api.lib(); port = R.load_port()
    for args in [(3, 0, 0, 0, 0), (3, 0, 1, 1, 0), (1, 0, 0, 0, 0), (1, 0, 1, 1, 0), (3, 3, 0, 0, 1), (3, 1, 1, 0, 1),
                 (1, 1, 0, 1, 1), (3, 3, 1, 1, 1)]:
        a = abi.TFormatIndices(); b = abi.TFormatIndices()
        L.prepareImageFormatIndices(C.byref(a), *args)
        port.prepareImageFormatIndices(C.byref(b), *args)
        for f, _t in abi.TFormatIndices._fields_:
            if f == "alpha_bip" and not (args[2] or args[3]):
                continue
            assert getattr(a, f) == getattr(b, f), (args, f)
    for fmt, bpp in ((abi.T_RGB, 4), (abi.T_RGBA, 5), (abi.T_Gray, 2), (abi.T_GrayA, 3)):
        a = abi.TFormatIndices()
        assert L.prepareImageFormatIndicesFromFormatType(C.byref(a), fmt) == 0 and a.total_bpp == bpp
    assert L.prepareImageFormatIndicesFromFormatType(C.byref(a), 666) == abi.IMAGE_SYNTH_ERROR_INVALID_IMAGE_FORMAT
    assert [L.countPixelelsPerPixelForFormat(f) for f in (0, 1, 2, 3, 9)] == [3, 4, 1, 2, 0]


@pytest.mark.parametrize("n,count", [(1, 10), (1000, 5000), (1 << 20, 300000), (3000000, 200000), (2863311531, 150000)])
def test_raw_stream_draws_equal_direct_draws(built_lib, n, count):
    """The engine starts producing raw PRNG words before it knows how many target points there are and reduces them to
    the range afterwards: the draws must be those of g_rand_int_range(0, n) in sequence, rejections included."""
    L = api.lib()
    L.rs_host_draws.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
    a = np.zeros(count, np.uint32); b = np.zeros(count, np.uint32)
    L.rs_host_draws(1198472, n, count, a.ctypes.data, 0)
    L.rs_host_draws(1198472, n, count, b.ctypes.data, 1)
    assert (a == b).all() and int(a.max()) < n
