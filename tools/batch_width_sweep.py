#!/usr/bin/env python
"""cfg5 batch (64 heal jobs 2048x2048, one GPU): jobs in flight x team widths (RS_TEAM_P0 / RS_TEAM_PN; '-' = the plan's own).
  python tools/batch_width_sweep.py [--jobs 64] [--probes 200]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resynthesizer_b200 import abi, api, build  # noqa: E402
from resynthesizer_b200.synthetic import G, centered_mask  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=64)
    ap.add_argument("--probes", type=int, default=200)
    ap.add_argument("--slots", default="4,8,16")
    ap.add_argument("--widths", default="-:-,8:2,4:4,4:2,4:1,2:2,2:1,1:1")
    a = ap.parse_args()
    build.build()
    api.set_device(0)
    m = centered_mask(2048, 2048, 256, 256)
    pristine = [G(2048, 2048, 3, 100 + k) for k in range(a.jobs)]
    work = [p.copy() for p in pristine]
    prm = abi.default_params(); prm.maxProbeCount = a.probes
    n_px = int((m != 0).sum())
    api.order_cache(True)
    ref = None
    for wd in a.widths.split(","):
        p0, pn = wd.split(":")
        for k, v in (("RS_TEAM_P0", p0), ("RS_TEAM_PN", pn)):
            if v == "-": os.environ.pop(k, None)
            else: os.environ[k] = v
        for slots in [int(x) for x in a.slots.split(",")]:
            for rep in range(2):
                for d, s in zip(work, pristine):
                    np.copyto(d, s)
                t0 = time.perf_counter()
                errs = api.image_synth_batch(work, [m] * a.jobs, abi.T_RGB, prm, devices=[0], slots=slots)
                t = time.perf_counter() - t0
                assert not any(errs)
            if ref is None: ref = [w.copy() for w in work]
            same = all((x == y).all() for x, y in zip(work, ref))   # the width of a team never changes a result
            print("widths %s slots %2d : %7.3f ms/job  %.3g px/s  %s" % (wd, slots, 1000 * t / a.jobs, a.jobs * n_px / t,
                                                                         "same images" if same else "IMAGES DIFFER"), flush=True)


if __name__ == "__main__":
    main()
