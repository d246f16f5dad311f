"""Deterministic synthetic inputs for tests and bench (generator G).

G(w,h,c,seed): a 32-bit LCG (s = s*1664525 + 1013904223) advanced once per byte in
row-major / channel-innermost order; byte = ((3x+5y+40ch)&255)/2 + (64 if ((x//8+y//8)&1) else 0)
+ (s>>27), clipped to 255.
"""
import numpy as np


def lcg_stream(n, seed):
    a = np.full(n, 1664525, np.uint32)
    A = np.cumprod(a, dtype=np.uint32)                      # a^k, k=1..n  (mod 2^32)
    geo = np.empty(n, np.uint32)                            # sum_{i<k} a^i
    geo[0] = 1
    if n > 1:
        geo[1:] = (np.cumsum(A[:-1], dtype=np.uint32) + np.uint32(1))
    return A * np.uint32(seed & 0xFFFFFFFF) + geo * np.uint32(1013904223)


def G(w, h, c, seed):
    s = lcg_stream(w * h * c, seed).reshape(h, w, c)
    y, x, ch = np.meshgrid(np.arange(h), np.arange(w), np.arange(c), indexing="ij")
    v = ((3 * x + 5 * y + 40 * ch) & 255) // 2 + np.where(((x // 8 + y // 8) & 1) != 0, 64, 0) + (s >> np.uint32(27)).astype(np.int64)
    return np.ascontiguousarray(np.minimum(v, 255).astype(np.uint8))


def centered_mask(w, h, mw, mh):
    m = np.zeros((h, w), np.uint8)
    x0, y0 = (w - mw) // 2, (h - mh) // 2
    m[y0:y0 + mh, x0:x0 + mw] = 255
    return m
