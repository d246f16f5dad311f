#!/usr/bin/env python
"""Host-clock phases of a call (RS_DEBUG) for a few jobs of each workload, from page-locked caller buffers and from
malloc'ed ones, and the time of the device's PRNG stream kernel beside the host producer's.
  python tools/phase_times.py cfg3 cfg5 cfg2"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RS_DEBUG"] = "1"
import torch  # noqa: E402
import bench  # noqa: E402
from resynthesizer_b200 import api, build  # noqa: E402

build.build()
L = api.lib()
api.set_device(0)
L.rs_host_draws.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
L.rs_cuda_mt19937_raw.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
n = 4400000
out = np.zeros(n, np.uint32)
t = time.perf_counter(); L.rs_host_draws(1198472, n, n, out.ctypes.data, 1); dt = time.perf_counter() - t
print("host PRNG draws (stream + range reduction, one core): %.2f ms for %d words (%.2f ns/word)" % (dt * 1e3, n, dt * 1e9 / n), flush=True)
for rep in range(3):  # includes cudaMalloc, the D2H copy of the words and cudaFree: an upper bound for the kernel
    t = time.perf_counter(); L.rs_cuda_mt19937_raw(1198472, n, out.ctypes.data); dt = time.perf_counter() - t
    print("device PRNG stream incl. malloc + 17.6 MB D2H: %.2f ms for %d words" % (dt * 1e3, n), flush=True)
api.order_cache(False)
for name in sys.argv[1:]:
    w = bench.workload(name)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    for pinned in (True, False):
        print("==", w["name"], "| caller buffers", "page-locked" if pinned else "malloc'ed", flush=True)
        r = bench.Runner(api, torch, w, fi, pinned=pinned)
        for rep in range(4):
            r.step()
