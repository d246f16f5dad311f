#!/usr/bin/env python
"""Host-clock phases of a call (RS_DEBUG) for a few jobs of each workload, and the speed of the PRNG raw-word producer.
  python tools/phase_times.py cfg3 cfg5 cfg2"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RS_DEBUG"] = "1"
import bench  # noqa: E402
from resynthesizer_b200 import api, build  # noqa: E402

build.build()
L = api.lib()
L.rs_host_draws.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
n = 4400000
out = np.zeros(n, np.uint32)
t = time.perf_counter(); L.rs_host_draws(1198472, n, n, out.ctypes.data, 1); dt = time.perf_counter() - t
print("PRNG raw stream: %.2f ms for %d words (%.2f ns/word)" % (dt * 1e3, n, dt * 1e9 / n), flush=True)
api.order_cache(False)
for name in sys.argv[1:]:
    w = bench.workload(name)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    print("==", w["name"], flush=True)
    for rep in range(4):
        if "simple" in w:
            img = w["tgt"].copy()
            assert api.image_synth(img, w["tmask"], w["simple"], w["params"]) == 0
        else:
            tp, cp = bench.pixmaps(w)
            assert api.engine(w["params"], fi, tp, cp) == 0
