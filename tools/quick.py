#!/usr/bin/env python
"""One line per workload: kernel ms of a warm job and per-pass ms (device clock).  For parameter sweeps via env."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from resynthesizer_b200 import api, build  # noqa: E402

build.build()
api.order_cache(False)   # every job orders its own points
for name in sys.argv[1:]:
    w = bench.workload(name)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    best = None
    for rep in range(3):
        tp, cp = bench.pixmaps(w)
        assert api.engine(w["params"], fi, tp, cp) == 0
        st = api.last_stats()
        if rep and (best is None or st["ms_kernels"] < best["ms_kernels"]):
            best = st
    print("%-16s kern %8.3f ms  passes %s  evals %.3g  [prep %.2f stage %.2f total %.2f ms]" % (
        name, best["ms_kernels"], " ".join("%.3f" % x for x in best["ms_pass"][:best["passes_run"]]), best["evals"],
        best["ms_prep"], best["ms_h2d"], best["ms_total"]), flush=True)
