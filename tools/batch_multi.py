#!/usr/bin/env python
"""cfg5 batch (64 heal jobs 2048x2048) dealt over 1..all GPUs of the box from ONE process (rs_image_synth_batch with a
device list), and the shared-corpus batch across the same devices (peer copies).  python tools/batch_multi.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resynthesizer_b200 import abi, api, build  # noqa: E402
from resynthesizer_b200.synthetic import G, centered_mask  # noqa: E402


def main():
    build.build()
    n_dev = api.lib().rs_cuda_device_count()
    jobs = 64
    m = centered_mask(2048, 2048, 256, 256)
    pristine = [G(2048, 2048, 3, 100 + k) for k in range(jobs)]
    work = [p.copy() for p in pristine]
    n_px = int((m != 0).sum())
    api.order_cache(True)
    n = 1
    while n <= n_dev:
        devs = list(range(n))
        for rep in range(2):
            for d, s in zip(work, pristine):
                np.copyto(d, s)
            t0 = time.perf_counter()
            errs = api.image_synth_batch(work, [m] * jobs, abi.T_RGB, None, devices=devs, slots=4)
            t = time.perf_counter() - t0
            assert not any(errs)
        print("in-process dealer, %d GPU(s): %7.3f ms/job  %.3g px/s" % (n, 1000 * t / jobs, jobs * n_px / t), flush=True)
        n *= 2
    # one corpus, many targets, all devices: built on one, peer-copied to the others
    cor = G(2048, 2048, 3, 77)
    cp = np.ascontiguousarray(np.concatenate([np.full((2048, 2048, 1), 255, np.uint8), cor], axis=2))
    prm = abi.make_params(0, 0, 0, 0.5, 0.117, 9, 200)
    fi = api.format_indices(3)
    for rep in range(2):
        tps = [np.full((256, 256, 4), 255, np.uint8) for _ in range(64)]
        b0 = api.shared_corpus_stats()
        t0 = time.perf_counter()
        errs = api.engine_batch([(prm, fi, tp, cp) for tp in tps], slots=4, devices=list(range(n_dev)))
        t = time.perf_counter() - t0
        b1 = api.shared_corpus_stats()
        assert not any(errs)
    print("shared corpus, %d GPU(s): %.3f ms/job; corpora built %d, reused %d, peer-copied %d" % (
        n_dev, 1000 * t / 64, b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]), flush=True)


if __name__ == "__main__":
    main()
