"""-m gpu: whole-image parity with the REFERENCE semantics, within a stated tolerance.

The CUDA engine cannot follow the reference's sequential PRNG stream (its consumption depends on every earlier
result) and takes a per-pass snapshot of the recentProber map, so against the reference itself whole images are
compared statistically (north star):

  tolerance (written here, used below)
    (i)  mean best-match distance of the last full pass:   gpu <= mean_ref + 2*sigma_ref + 2% of mean_ref
    (ii) PSNR over target pixels, gpu vs any reference run: >= mean(seed-to-seed PSNR of the reference) - 2*sigma - 0.5 dB
  where the reference distribution comes from S=6 runs of the oracle in reference mode (MT19937 stream, live
  recentProber; bit-identical to the compiled reference, tests/test_port_vs_ref.py) with different PRNG seeds.
"""
import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu
SEEDS = [1198472, 7, 99, 2024, 31337, 424242]


def _psnr(a, b, sel):
    d = (a.astype(np.float64) - b.astype(np.float64))[sel]
    mse = (d ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _natural_like(w, h, seed):
    """Smooth random texture + edges: closer to a photograph than G (which is periodic)."""
    rng = np.random.RandomState(seed)
    base = rng.rand(h // 8 + 2, w // 8 + 2, 3)
    img = np.kron(base, np.ones((8, 8, 1)))[:h, :w]
    yy, xx = np.mgrid[0:h, 0:w]
    img += 0.25 * np.sin(xx / 5.0)[:, :, None] + 0.2 * ((xx + 2 * yy) % 37 < 6)[:, :, None]
    img += 0.08 * rng.randn(h, w, 3)
    img = (img - img.min()) / (img.max() - img.min())
    return np.ascontiguousarray((img * 255).astype(np.uint8))


@pytest.mark.parametrize("kind", ["synthetic_G", "natural_like"])
def test_heal_quality_within_reference_spread(built_oracle, built_lib, kind):
    w = h = 160
    img = G(w, h, 3, 12345) if kind == "synthetic_G" else _natural_like(w, h, 5)
    mask = centered_mask(w, h, 48, 40)
    sel = mask != 0
    params = abi.default_params()
    ref_outs, ref_dist = [], []
    for s in SEEDS:
        port = R.load_port(R.REF_MODE, s)
        err, out = R.image_synth(port, img, mask, abi.T_RGB, params)
        assert err == 0
        st = R.port_stats(port)
        last = 1  # pass 1 is the last pass over ALL target points (lib/passes.h:78-91)
        ref_dist.append(st["sum_best"][last] / st["pass_visits"][last])
        ref_outs.append(out)
    pair = [_psnr(ref_outs[i], ref_outs[j], sel) for i in range(len(SEEDS)) for j in range(i + 1, len(SEEDS))]
    gpu_dist, gpu_psnr = [], []
    for s in SEEDS[:3]:
        api.set_seed(s)
        out = img.copy()
        assert api.image_synth(out, mask, abi.T_RGB, params) == 0
        st = api.last_stats()
        gpu_dist.append(st["sum_best"][1] / st["pass_visits"][1])
        gpu_psnr.append(np.mean([_psnr(out, r, sel) for r in ref_outs]))
        assert (out[~sel] == img[~sel]).all()          # context untouched
    m, sd = np.mean(ref_dist), np.std(ref_dist)
    assert np.mean(gpu_dist) <= m + 2 * sd + 0.02 * m, (gpu_dist, ref_dist)
    pm, psd = np.mean(pair), np.std(pair)
    assert np.mean(gpu_psnr) >= pm - 2 * psd - 0.5, (gpu_psnr, pair)


def test_gpu_semantics_oracle_is_within_reference_spread(built_oracle):
    """Same tolerance, CPU only: the oracle in GPU mode (what the CUDA engine equals bit for bit) vs reference mode."""
    w = h = 128
    img = G(w, h, 3, 777)
    mask = centered_mask(w, h, 40, 36)
    params = abi.default_params()
    ref, gpu = [], []
    for s in SEEDS:
        for mode, acc in ((R.REF_MODE, ref), (R.GPU_MODE, gpu)):
            port = R.load_port(mode, s)
            assert R.image_synth(port, img, mask, abi.T_RGB, params)[0] == 0
            st = R.port_stats(port)
            acc.append(st["sum_best"][1] / st["pass_visits"][1])
    m, sd = np.mean(ref), np.std(ref)
    assert np.mean(gpu) <= m + 2 * sd + 0.02 * m, (gpu, ref)
