#!/usr/bin/env python
"""Runs `--jobs` identical jobs of one bench workload and nothing else: the command to put under ncu.

  ncu --set full --clock-control none --import-source on -k k_synth_pass -s 7 -c 1 -o gpurun_out/x \
      python tools/ncu_job.py --workload cfg2 --jobs 2
(k_synth_pass launches per job: 6, of which pass 0 is the first; -s 7 = pass 1 of the second job.)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from resynthesizer_b200 import api, build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--jobs", type=int, default=2)
    ap.add_argument("--probes", type=int, default=0)
    a = ap.parse_args()
    bench.PROBES_OVERRIDE = a.probes
    build.build()
    w = bench.workload(a.workload)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    for _ in range(a.jobs):
        if "simple" in w:
            img = w["tgt"].copy()
            assert api.image_synth(img, w["tmask"], w["simple"], w["params"]) == 0
        else:
            tp, cp = bench.pixmaps(w)
            assert api.engine(w["params"], fi, tp, cp) == 0
    st = api.last_stats()
    print("%s: ms_kernels %.3f passes %d ms_pass %s" % (w["name"], st["ms_kernels"], st["passes_run"], ["%.3f" % x for x in st["ms_pass"]]))


if __name__ == "__main__":
    main()
