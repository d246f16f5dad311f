"""Device-side cache of visit orders: same key -> hit and identical output; different selection, seed or mode -> miss."""
import numpy as np
import pytest

from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _job(seed_img, hole, mode=1):
    img = G(96, 80, 3, seed_img)
    m = centered_mask(96, 80, hole, hole)
    p = abi.make_params(0, 0, mode, 0.5, 0.117, 16, 60)
    tp = np.ascontiguousarray(np.concatenate([m[:, :, None], img], axis=2))
    cp = np.ascontiguousarray(np.concatenate([(255 - m)[:, :, None], img], axis=2))
    return p, tp, cp


def _run(p, tp, cp):
    fi = api.format_indices(3, 0, False, False, False)
    t = tp.copy()
    assert api.engine(p, fi, t, cp) == 0
    return t, api.last_stats()


def test_hit_is_identical_and_misses_are_misses():
    api.order_cache(False)          # drop whatever earlier tests left
    api.order_cache(True)
    try:
        api.set_seed(1198472)
        p, tp, cp = _job(7, 24)
        a, sa = _run(p, tp, cp)
        b, sb = _run(p, tp, cp)
        assert sa["order_cache_hit"] == 0 and sb["order_cache_hit"] == 1
        assert (a == b).all()
        # another image with the SAME selection: the order is shared, the result is that image's own
        p2, tp2, cp2 = _job(8, 24)
        c, sc = _run(p2, tp2, cp2)
        assert sc["order_cache_hit"] == 1
        api.order_cache(False)
        c_ref, s_ref = _run(p2, tp2, cp2)
        assert s_ref["order_cache_hit"] == 0 and (c == c_ref).all()
        api.order_cache(True)
        # different selection / context type / seed: all misses
        _run(p, tp, cp)
        assert _run(*_job(7, 26))[1]["order_cache_hit"] == 0
        assert _run(*_job(7, 24, mode=2))[1]["order_cache_hit"] == 0
        api.set_seed(5)
        assert _run(p, tp, cp)[1]["order_cache_hit"] == 0
        assert _run(p, tp, cp)[1]["order_cache_hit"] == 1
    finally:
        api.set_seed(1198472)
        api.order_cache(True)


def test_gather_rate_and_timeline():
    assert api.gather_rate(1 << 18, 4, 1) > 1e9
    api.keep_result(True)
    try:
        p, tp, cp = _job(9, 32)
        _run(p, tp, cp)
        st = api.last_stats()
        t = api.last_timeline(0)
        assert len(t) >= 1 and t[0] == 0
        assert st["passes_run"] >= 1 and st["ms_pass"][0] > 0 and st["kernel_launches"] > 0
        assert st["synth_launches_run"] >= st["passes_run"]
    finally:
        api.keep_result(False)


def test_concurrent_jobs_share_a_fresh_entry():
    """A batch whose jobs all have the same selection: the first job to order its points publishes the entry while the
    others are already looking for it (the entry's upload is on another stream).  Every result must equal the
    single-job result."""
    api.order_cache(False)
    api.order_cache(True)
    try:
        api.set_seed(1198472)
        fi = api.format_indices(3, 0, False, False, False)
        singles = []
        for k in range(4):
            p, tp, cp = _job(20 + k, 28)
            api.order_cache(False)
            singles.append(_run(p, tp, cp)[0])
        for rep in range(3):
            api.order_cache(False)      # drop the entries: every repetition starts cold
            api.order_cache(True)
            jobs = []
            for i in range(24):
                p, tp, cp = _job(20 + i % 4, 28)
                jobs.append((p, fi, tp.copy(), cp))
            assert not any(api.engine_batch(jobs, 8))
            for i, jb in enumerate(jobs):
                assert (jb[2] == singles[i % 4]).all(), "job %d of repetition %d differs" % (i, rep)
    finally:
        api.order_cache(True)


@pytest.mark.parametrize("mode", [2, 3, 4, 5, 6, 7, 8])
def test_device_sort_equals_host_sort(mode):
    """Orderings 2-8: sorting the (key, point) pairs on the device gives the same visit order, hence the same image,
    sources and counters as the host's radix sort."""
    api.order_cache(False)
    api.keep_result(True)
    try:
        p, tp, cp = _job(31, 36, mode=mode)
        api.set_device_sort_min(1 << 30)
        a, sa = _run(p, tp, cp)
        ta, srca = api.last_result()
        api.set_device_sort_min(1)
        b, sb = _run(p, tp, cp)
        tb, srcb = api.last_result()
        assert (ta == tb).all() and (srca == srcb).all() and (a == b).all()
        assert sa["evals"] == sb["evals"] and sa["betters"] == sb["betters"]
    finally:
        api.set_device_sort_min(1 << 16)
        api.keep_result(False)
        api.order_cache(True)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("hole", [3, 40])
def test_device_shuffle_equals_host_shuffle(mode, hole):
    """Orderings 0/1: the chain of swaps resolved on the device gives the visit order the host's swap loop gives."""
    api.order_cache(False)
    api.keep_result(True)
    try:
        p, tp, cp = _job(33, hole, mode=mode)
        api.set_device_shuffle_min(1 << 30)
        a, sa = _run(p, tp, cp)
        ta, srca = api.last_result()
        api.set_device_shuffle_min(1)
        b, sb = _run(p, tp, cp)
        tb, srcb = api.last_result()
        assert len(ta) == len(tb) and (ta == tb).all() and (srca == srcb).all() and (a == b).all()
        assert sa["evals"] == sb["evals"] and sa["betters"] == sb["betters"]
    finally:
        api.set_device_shuffle_min(1 << 15)
        api.keep_result(False)
        api.order_cache(True)


@pytest.mark.parametrize("seed", [1198472, 0, 1, 0xFFFFFFFF, 20241017])
def test_device_prng_stream_is_glib_mt19937(seed):
    """k_mt19937_raw: the raw 32-bit words of g_rand_new_with_seed(seed) / g_rand_int (MT19937, init_genrand seeding -- the
    stream the reference draws its visit order from, lib/orderTarget.h:38-53), made on the device, equal numpy's legacy
    MT19937 word for word; lengths around the kernel's rounds of 623 words, the state size, and the 2^18 words of a CTA."""
    import ctypes as C
    L = api.lib()
    L.rs_cuda_mt19937_raw.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    api.set_device(0)
    want = np.random.RandomState(seed).randint(0, 2 ** 32, 4400000, dtype=np.uint32)
    # up to 2^18 words: one CTA; above: CTA q starts q * 2^18 words into the stream by jump-ahead (1.2 M words: 5 CTAs,
    # 4.4 M, the order of a 4 Mi-point job: 17)
    for n in (1, 226, 227, 228, 453, 454, 455, 622, 623, 624, 625, 1246, 1247, 100000, 262144, 262145, 524288 + 5, 1200000, 4400000):
        got = np.zeros(n + 8, np.uint32)
        got[n:] = 0xDEADBEEF
        assert L.rs_cuda_mt19937_raw(seed, n, got.ctypes.data) == 0
        assert (got[:n] == want[:n]).all(), (seed, n)
        assert (got[n:] == 0xDEADBEEF).all()


@pytest.mark.parametrize("prng", ["device", "host_raw", "host_draws"])
def test_device_shuffle_large(prng, monkeypatch):
    """200 k target points: the device-resolved shuffle (PRNG stream made on the device; raw words of the host's producer
    reduced on the device; draws reduced on the host) orders them exactly as the host's swap loop does."""
    if prng == "host_raw":
        monkeypatch.setenv("RS_HOST_PRNG", "1")
    if prng == "host_draws":
        monkeypatch.setenv("RS_NO_RAW_STREAM", "1")
    api.order_cache(False)
    api.keep_result(True)
    try:
        img = G(640, 512, 3, 77)
        m = np.zeros((512, 640), np.uint8); m[40:440, 60:560] = 255
        p = abi.make_params(0, 0, 1, 0.5, 0.117, 9, 20)
        tp = np.ascontiguousarray(np.concatenate([m[:, :, None], img], axis=2))
        cp = np.ascontiguousarray(np.concatenate([(255 - m)[:, :, None], img], axis=2))
        api.set_device_shuffle_min(1 << 30)
        a, sa = _run(p, tp, cp)
        ta, srca = api.last_result()
        api.set_device_shuffle_min(1 << 15)
        b, sb = _run(p, tp, cp)
        tb, srcb = api.last_result()
        assert len(ta) == 200000 and (ta == tb).all() and (srca == srcb).all() and (a == b).all()
    finally:
        api.set_device_shuffle_min(1 << 15)
        api.keep_result(False)
        api.order_cache(True)


@pytest.mark.parametrize("mode", list(range(9)))
def test_device_resolved_order_equals_oracle_order(mode):
    """The visit order as the DEVICE paths resolve it (modes 0/1: chain of swaps traced on the device from raw PRNG
    words; modes 2-8: pair sort on the device) against the oracle's orderTargetPoints restatement (port_order,
    lib/orderTarget.h:268-343), directly, on 200 000 points."""
    from oracle import refdriver as R
    import subprocess, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "port"])
    port = R.load_port(R.REF_MODE)
    api.order_cache(False)
    api.keep_result(True)
    try:
        img = G(640, 512, 3, 78)
        m = np.zeros((512, 640), np.uint8); m[40:440, 60:560] = 255
        m[100:140, 200:260] = 0           # a hole in the selection: bounding box and rays are not trivial
        n = int((m != 0).sum())
        p = abi.make_params(0, 0, mode, 0.5, 0.117, 4, 2)
        tp = np.ascontiguousarray(np.concatenate([m[:, :, None], img], axis=2))
        cp = np.ascontiguousarray(np.concatenate([(255 - m)[:, :, None], img], axis=2))
        api.set_device_shuffle_min(1 << 15)
        api.set_device_sort_min(1 << 16)
        api.set_seed(1198472)
        _run(p, tp, cp)
        t_gpu, _ = api.last_result()
        ys, xs = np.nonzero(m)
        pts = np.stack([xs, ys], axis=1).astype(np.int32).copy()       # row-major, as the engine collects them
        assert port.port_order(mode, pts.ctypes.data, n, 1198472) == 0
        assert len(t_gpu) == n and (t_gpu == pts).all()
    finally:
        api.keep_result(False)
        api.order_cache(True)
