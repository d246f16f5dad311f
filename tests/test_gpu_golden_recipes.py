"""-m gpu: the reference's own golden recipes (Test/testResynth.py:216-408, restated in oracle/goldens.py) through the
CUDA engine, held against the reference.

Two claims per recipe:
  (1) exact: at the reference's seed the CUDA engine reproduces, bit for bit, the output of the sequential definition of
      its semantics (oracle GPU mode) on the reference's real test images -- digest committed in
      tests/golden/recipe_ref_spread.json by tests/golden/make_recipe_spread.py;
  (2) statistical, against the REFERENCE algorithm itself (the sequential PRNG stream and the live recentProber map
      cannot be followed in parallel).  Tolerance, as SURVEY.md section 8c proposes it, with the
      reference distribution taken from 8 seeds of the oracle in reference mode (which reproduces the goldens at the
      reference's seed) on the same recipe, and NO extra slack:
          whole-image PSNR against the reference's golden image   >= mean_ref - 2 sigma_ref
          mean best-match distance over the last pass that ran    <= mean_ref + 2 sigma_ref
      held by the MEDIAN of the CUDA engine's GPU_SEEDS runs (the last-pass mean of a 5000-pixel heal is heavy-tailed:
      one seed in six lands a few bad pixels, in the reference's own runs as well) and by at least 4 of the 6 runs
      individually.  Measured table: profiles/r02_quality_goldens.md.
Needs oracle/_ref/recipe_images.npz (packed by build() where /root/reference exists; travels with the snapshot)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import goldens
from resynthesizer_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPREAD = json.load(open(os.path.join(ROOT, "tests", "golden", "recipe_ref_spread.json")))
GPU_SEEDS = [1198472, 7, 99, 2024, 31337, 424242]
TOL_SIGMA = 2.0
MIN_RUNS_WITHIN = 4

needs_images = pytest.mark.skipif(not goldens.available(), reason="oracle/_ref/recipe_images.npz not packed (build() without /root/reference)")


def _psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d ** 2).mean())
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


@needs_images
@pytest.mark.parametrize("name", list(goldens.CASES))
def test_recipe_on_cuda_engine(built_lib, name):
    _exact, fn = goldens.CASES[name]
    ref = SPREAD[name]
    gold = goldens.load_golden(name)
    api.order_cache(False)
    ps, mb = [], []
    for seed in GPU_SEEDS:
        api.set_seed(seed)
        try:
            out = fn(api.lib())
            st = api.last_stats()
        finally:
            api.set_seed(1198472)
        assert out.shape == gold.shape
        if seed == 1198472:   # (1) the sequential definition of the engine, on the reference's images, bit for bit
            assert hashlib.sha1(np.ascontiguousarray(out).tobytes()).hexdigest() == ref["gpu_mode_sha1"]
            assert st["passes_run"] == ref["gpu_mode_passes"]
        p = st["passes_run"] - 1
        ps.append(_psnr(out, gold))
        mb.append(st["sum_best"][p] / max(st["pass_visits"][p], 1))
    # (2) within the reference's own seed-to-seed spread
    rps = [x for x in ref["psnr_vs_golden"] if x is not None]
    rmb = ref["mean_best"]
    psnr_bound = np.mean(rps) - TOL_SIGMA * np.std(rps)
    best_bound = np.mean(rmb) + TOL_SIGMA * np.std(rmb)
    assert np.median(ps) >= psnr_bound, (ps, psnr_bound)
    assert np.median(mb) <= best_bound, (mb, best_bound)
    assert sum(p >= psnr_bound for p in ps) >= MIN_RUNS_WITHIN, (ps, psnr_bound)
    assert sum(b <= best_bound for b in mb) >= MIN_RUNS_WITHIN, (mb, best_bound)
