#!/usr/bin/env python
"""cfg5 batch (64 heal jobs 2048x2048, one GPU) under different dealing policies: jobs in flight x SM sharing.
  python tools/batch_sweep.py [--jobs 64] [--probes 200]
RS_BATCH_SHARE=1: the jobs in flight split the SMs; 0: every job launches full-width grids (pipelining only)."""
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
    a = ap.parse_args()
    build.build()
    api.set_device(0)
    m = centered_mask(2048, 2048, 256, 256)
    pristine = [G(2048, 2048, 3, 100 + k) for k in range(a.jobs)]
    work = [p.copy() for p in pristine]
    prm = abi.default_params(); prm.maxProbeCount = a.probes
    n_px = int((m != 0).sum())
    api.order_cache(True)
    # one at a time through imageSynth(): the loop the reference's callers write
    for rep in range(2):
        for d, s in zip(work, pristine):
            np.copyto(d, s)
        t0 = time.perf_counter()
        for img in work:
            assert api.image_synth(img, m, abi.T_RGB, prm) == 0
        t = time.perf_counter() - t0
    print("loop of imageSynth() calls   : %7.3f ms/job  %.3g px/s" % (1000 * t / a.jobs, a.jobs * n_px / t), flush=True)
    for share in (1, 0):
        for slots in (1, 2, 3, 4, 6):
            os.environ["RS_BATCH_SHARE"] = str(share)
            for rep in range(2):
                for d, s in zip(work, pristine):
                    np.copyto(d, s)
                t0 = time.perf_counter()
                errs = api.image_synth_batch(work, [m] * a.jobs, abi.T_RGB, prm, devices=[0], slots=slots)
                t = time.perf_counter() - t0
                assert not any(errs)
            print("batch share_sms=%d slots=%d    : %7.3f ms/job  %.3g px/s" % (share, slots, 1000 * t / a.jobs, a.jobs * n_px / t), flush=True)


if __name__ == "__main__":
    main()
