#!/usr/bin/env python
"""Wall time of imageSynth() (simple API, host buffers) vs engine() on the same heal job."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from resynthesizer_b200 import abi, api, build  # noqa: E402

build.build()
api.order_cache(False)
for name in sys.argv[1:]:
    w = bench.workload(name)
    img, mask = w["tgt"], w["tmask"]
    fmt = abi.T_RGBA if img.shape[2] == 4 else abi.T_RGB
    ts = []
    for rep in range(4):
        a = img.copy()
        t0 = time.perf_counter()
        assert api.image_synth(a, mask, fmt, None) == 0
        ts.append(time.perf_counter() - t0)
    st = api.last_stats()
    print("%-8s imageSynth %.2f ms (engine part %.2f ms, kernels %.2f ms)" % (name, 1e3 * min(ts[1:]), st["ms_total"], st["ms_kernels"]))
