"""CPU model of k_mt19937_raw (csrc/rs_kernels.cu): the same ring addresses, lane roles, round size, writer lag and jump
combination, statement for statement, in numpy -- against numpy's legacy MT19937.  What the GPU test checks on the
device (`test_gpu_order_cache.py::test_device_prng_stream_is_glib_mt19937`) is checked here for the scheme itself:
every read of a round hits words that are final, writers find the words of the round before still in the ring, a CTA
that jumps ahead continues the stream exactly."""
import ctypes as C

import numpy as np
import pytest

from resynthesizer_b200 import api

U32 = np.uint32
WRITERS = 384


def _seed_ring(seed):
    ring = np.zeros(2048, np.uint32)
    x = seed & 0xFFFFFFFF
    ring[0] = x
    for i in range(1, 624):
        x = (1812433253 * (x ^ (x >> 30)) + i) & 0xFFFFFFFF
        ring[i] = x
    return ring


def _f(a, b):
    y = (a & U32(0x80000000)) | (b & U32(0x7FFFFFFF))
    return (y >> U32(1)) ^ np.where(y & U32(1), U32(0x9908B0DF), U32(0))


def _temper(z):
    z = z ^ (z >> U32(11))
    z = z ^ ((z << U32(7)) & U32(0x9D2C5680))
    z = z ^ ((z << U32(15)) & U32(0xEFC60000))
    return z ^ (z >> U32(18))


def _rounds(ring, n_words, out=None, X=None):
    """rs_mt_rounds: all maker lanes of a round read before any of them writes (the kernel's barrier separates rounds,
    and a round's reads and writes never meet: asserted)."""
    t = np.arange(227)
    third = t < 169
    pb = (t * 4).astype(np.int64)
    rounds = (n_words + 622) // 623
    for r in range(rounds + 1):
        if r < rounds:
            idx = lambda off: ((pb + off) & 8188) >> 2
            reads = np.concatenate([idx(1588), idx(0), idx(4), idx(908), idx(912), idx(1816)[third], idx(1820)[third]])
            writes = np.concatenate([idx(2496), idx(3404), idx(4312)[third]])
            assert not np.intersect1d(reads, writes).size            # no lane overwrites what another still has to read
            c, a0, b0, a1, b1 = ring[idx(1588)], ring[idx(0)], ring[idx(4)], ring[idx(908)], ring[idx(912)]
            a2 = np.where(third, ring[idx(1816)], U32(0))
            b2 = np.where(third, ring[idx(1820)], U32(0))
            v0 = c ^ _f(a0, b0)
            v1 = v0 ^ _f(a1, b1)
            v2 = v1 ^ _f(a2, b2)
        if out is not None and r > 0:                                 # writers: the words of round r - 1, read BEFORE round r's stores land
            base = (r - 1) * 623
            for k in range((623 + WRITERS - 1) // WRITERS):
                o = np.arange(WRITERS) + k * WRITERS
                o = o[(o < 623) & (base + o < n_words)]
                pos = (base + o + 624) & 2047
                if r < rounds:
                    assert not np.intersect1d(pos, writes).size      # ... and round r's stores do not touch them anyway
                out[base + o] = _temper(ring[pos])
        if r < rounds:
            ring[idx(2496)] = v0
            ring[idx(3404)] = v1
            ring[idx(4312)[third]] = v2[third]
            if X is not None:
                i0 = 624 + r * 623 + t
                X[i0] = v0
                X[i0 + 227] = v1
                X[(i0 + 454)[third]] = v2[third]
            pb = (pb + 2492) & 8188


def _want(seed, n):
    return np.random.RandomState(seed).randint(0, 2 ** 32, n, dtype=np.uint32)


@pytest.mark.parametrize("n", [1, 168, 169, 227, 622, 623, 624, 1246, 1247, 5000, 40000])
def test_one_cta_round_scheme(n):
    out = np.zeros(n, np.uint32)
    _rounds(_seed_ring(1198472), n, out=out)
    assert (out == _want(1198472, n)).all()


def test_jump_ahead_cta_continues_the_stream():
    """CTA q of a multi-CTA launch: 19936 untempered words from the seed, the windows its polynomial selects XORed into a
    new state (index list padded to a multiple of eight with the position of 624 zero words), then its own words."""
    jump, q, n_own = 1 << 18, 2, 3000
    L = api.lib()
    L.rs_host_mt_jump_poly.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    L.rs_host_mt_jump_poly.restype = C.c_uint32
    idx = np.zeros(19968, np.uint16)
    cnt = L.rs_host_mt_jump_poly(q, jump, idx.ctypes.data)
    PAD = 624 + 19936
    idx[cnt:(cnt + 7) // 8 * 8] = PAD
    cnt8 = (cnt + 7) // 8
    ring = _seed_ring(1198472)
    X = np.zeros(PAD + 624, np.uint32)
    X[:624] = ring[:624]
    _rounds(ring, 19936, X=X)
    t = np.arange(624)
    acc = np.zeros(624, np.uint32)
    for e in range(cnt8 * 8):
        acc ^= X[int(idx[e]) + t]
    ring[:624] = acc
    out = np.zeros(n_own, np.uint32)
    _rounds(ring, n_own, out=out)
    want = _want(1198472, q * jump + n_own)
    assert (out == want[q * jump:]).all()
