#!/usr/bin/env python
"""Throughput profile of the passes of one job: visits per microsecond in each 4096-visit slice (device clock).

  python tools/pass_timeline.py --workload cfg2 [--out profiles/timeline_cfg2.txt]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from resynthesizer_b200 import api, build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    build.build()
    w = bench.workload(a.workload)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    api.keep_result(True)
    lines = []
    for rep in range(2):   # second run = warm workspace
        tp, cp = bench.pixmaps(w)
        assert api.engine(w["params"], fi, tp, cp) == 0
    st = api.last_stats()
    lines.append("# %s: ms_kernels %.3f, per pass %s" % (w["name"], st["ms_kernels"], ["%.3f" % x for x in st["ms_pass"]]))
    for p in range(st["passes_run"]):
        t = api.last_timeline(p).astype(np.float64)
        if len(t) < 2:
            continue
        dt = np.diff(t) / 1000.0   # us per 4096 visits
        rate = 4096.0 / np.maximum(dt, 1e-9)
        lines.append("pass %d: %d slices; visits/us by slice (first 24, then every 8th):" % (p, len(dt)))
        lines.append("  " + " ".join("%.0f" % r for r in rate[:24]))
        lines.append("  " + " ".join("%.0f" % r for r in rate[24::8]))
        lines.append("  cumulative ms at 1/16, 1/8, 1/4, 1/2, 1: " + " ".join(
            "%.3f" % (t[min(len(t) - 1, int(len(t) * f))] / 1e6) for f in (1 / 16, 1 / 8, 1 / 4, 1 / 2, 1)))
    text = "\n".join(lines)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
