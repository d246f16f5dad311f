"""MT19937 jump-ahead polynomials (csrc/host_prep.cpp: mt_jump_poly) against an independent big-integer restatement:
characteristic polynomial by Berlekamp-Massey, z^J mod phi by square and multiply, and the identity the device kernel
relies on -- the state J words into the stream is the XOR of the windows of the untempered stream selected by the
polynomial.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from resynthesizer_b200 import api

N = 19937


def _untempered(seed, count):
    mt = [0] * 624
    mt[0] = seed
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
    x = list(mt)   # x[624 + i] = word i of the untempered stream
    for _ in range(count):
        y = (x[-624] & 0x80000000) | (x[-623] & 0x7FFFFFFF)
        x.append(x[-227] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0))
    return x


def _phi():
    x = _untempered(4357, 2 * N)
    s = [x[624 + i] & 1 for i in range(2 * N)]
    c, b, length, m, r = 1, 1, 0, 1, 0
    for i, bit in enumerate(s):
        r = (r << 1) | bit
        if (c & r).bit_count() & 1:
            t = c
            c ^= b << m
            if 2 * length <= i:
                length, b, m = i + 1 - length, t, 1
            else:
                m += 1
        else:
            m += 1
    assert length == N
    phi = 0
    for k in range(N + 1):
        if (c >> (N - k)) & 1:
            phi |= 1 << k
    return phi


def _mulmod(a, b, phi):
    r = 0
    while b:
        low = b & -b
        r ^= a << (low.bit_length() - 1)
        b ^= low
    while r.bit_length() > N:
        r ^= phi << (r.bit_length() - 1 - N)
    return r


def _zpow(e, phi):
    res, base = 1, 2
    while e:
        if e & 1:
            res = _mulmod(res, base, phi)
        base = _mulmod(base, base, phi)
        e >>= 1
    return res


@pytest.fixture(scope="module")
def phi():
    return _phi()


def _lib_poly(q, jump):
    L = api.lib()
    L.rs_host_mt_jump_poly.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    L.rs_host_mt_jump_poly.restype = C.c_uint32
    idx = np.zeros(N, np.uint16)
    cnt = L.rs_host_mt_jump_poly(q, jump, idx.ctypes.data)
    return idx[:cnt].astype(np.int64)


@pytest.mark.parametrize("jump,q", [(262144, 1), (262144, 2), (262144, 5), (1000, 1), (1000, 3), (19937, 1), (12345, 2)])
def test_jump_polynomial_equals_big_integer_restatement(phi, jump, q):
    g = _zpow(q * jump, phi)
    want = [k for k in range(N) if (g >> k) & 1]
    got = _lib_poly(q, jump)
    assert len(got) == len(want) and (got == np.array(want)).all()
    assert (np.diff(got) > 0).all()


def test_jump_identity_on_the_stream():
    """x[J + j] = XOR over the polynomial's set bits k of x[k + j] for the 623 full state words (and the top bit of the
    oldest one): what CTA q of k_mt19937_raw computes before it makes its own words."""
    jump, q = 262144, 2
    idx = _lib_poly(q, jump)
    x = np.array(_untempered(1198472, q * jump + 8), dtype=np.uint32)
    for j in range(0, 624):
        acc = np.bitwise_xor.reduce(x[idx + j])
        if j == 0:
            assert (int(acc) >> 31) == (int(x[q * jump]) >> 31)
        else:
            assert int(acc) == int(x[q * jump + j]), j
