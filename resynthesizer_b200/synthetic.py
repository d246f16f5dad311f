"""Deterministic synthetic inputs for tests and bench (generator G).

G(w,h,c,seed): a 32-bit LCG (s = s*1664525 + 1013904223) advanced once per byte in
row-major / channel-innermost order; byte = ((3x+5y+40ch)&255)/2 + (64 if ((x//8+y//8)&1) else 0)
+ (s>>27), clipped to 255.
"""
import numpy as np


def lcg_stream(n, seed, block=1 << 16):
    """s_1..s_n of s = s*1664525 + 1013904223 (mod 2^32) from s_0 = seed, in blocks: the closed form
    s_{j+k} = a^k s_j + c (a^k - 1)/(a - 1) for k <= block, block starts advanced by the block's own (a^B, geo_B)."""
    B = min(block, max(n, 1))
    A = np.cumprod(np.full(B, 1664525, np.uint32), dtype=np.uint32)     # a^k, k = 1..B  (mod 2^32)
    geo = np.empty(B, np.uint32)                                        # sum_{i<k} a^i
    geo[0] = 1
    if B > 1:
        geo[1:] = np.cumsum(A[:-1], dtype=np.uint32) + np.uint32(1)
    gc = geo * np.uint32(1013904223)
    nb = (n + B - 1) // B
    starts = np.empty(nb, np.uint32)
    cur = int(seed) & 0xFFFFFFFF
    aB, gB = int(A[-1]), int(gc[-1])
    for j in range(nb):
        starts[j] = cur
        cur = (aB * cur + gB) & 0xFFFFFFFF
    return (A[None, :] * starts[:, None] + gc[None, :]).reshape(-1)[:n]


def G(w, h, c, seed):
    s = lcg_stream(w * h * c, seed).reshape(h, w, c)
    x = np.arange(w, dtype=np.int32)[None, :, None]
    y = np.arange(h, dtype=np.int32)[:, None, None]
    ch = np.arange(c, dtype=np.int32)[None, None, :]
    v = ((3 * x + 5 * y + 40 * ch) & 255) >> 1                       # (h, w, c) by broadcasting, small integers
    v = v + (((x >> 3) + (y >> 3)) & 1) * 64
    v = v + (s >> np.uint32(27)).astype(np.int32)
    return np.ascontiguousarray(np.minimum(v, 255).astype(np.uint8))


def centered_mask(w, h, mw, mh):
    m = np.zeros((h, w), np.uint8)
    x0, y0 = (w - mw) // 2, (h - mh) // 2
    m[y0:y0 + mh, x0:x0 + mw] = 255
    return m
