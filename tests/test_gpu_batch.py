"""-m gpu: the batch dealers.  rs_image_synth_batch / rs_engine_batch_multi run independent jobs from one queue over
the devices given (a job never shards, a batch does); every job's result must equal the same
job run alone -- i.e. the oracle's -- whatever device, slot and order it ran in."""
import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _heal_jobs(n, side, hole):
    return [G(side, side, 3, 100 + k) for k in range(n)], [centered_mask(side, side, hole, hole)] * n


def _oracle_heal(img, mask, params=None):
    port = R.load_port(R.GPU_MODE)
    err, want = R.image_synth(port, img, mask, abi.T_RGB, params)
    assert err == 0
    return want


def _devices():
    return list(range(min(api.lib().rs_cuda_device_count(), 8)))


@pytest.mark.parametrize("slots", [1, 3])
def test_image_synth_batch_equals_oracle_per_job(built_oracle, built_lib, slots):
    imgs, masks = _heal_jobs(7, 96, 24)
    want = [_oracle_heal(i, m) for i, m in zip(imgs, masks)]
    got = [i.copy() for i in imgs]
    errs = api.image_synth_batch(got, masks, abi.T_RGB, None, devices=[0], slots=slots)
    assert errs == [0] * 7
    for g, w in zip(got, want):
        assert (g == w).all()


def test_batch_over_every_device_of_the_box(built_oracle, built_lib):
    """All devices of the box from ONE process (one job on a 1-GPU box still goes through the dealer)."""
    devs = _devices()
    imgs, masks = _heal_jobs(4 * len(devs) + 1, 80, 20)
    want = [_oracle_heal(i, m) for i, m in zip(imgs, masks)]
    got = [i.copy() for i in imgs]
    api.order_cache(True)
    errs = api.image_synth_batch(got, masks, abi.T_RGB, None, devices=devs, slots=2)
    api.order_cache(False)
    assert not any(errs)
    for g, w in zip(got, want):
        assert (g == w).all()
    # the calling thread keeps its device: a plain call afterwards runs where it did before
    one = imgs[0].copy()
    assert api.image_synth(one, masks[0], abi.T_RGB, None) == 0 and (one == want[0]).all()


def test_unequal_jobs_longest_first_and_full_api(built_oracle, built_lib):
    """Jobs of different sizes (the dealer orders the queue by estimated cost) through rs_engine_batch_multi."""
    port = R.load_port(R.GPU_MODE)
    fi = api.format_indices(3)
    jobs, want = [], []
    for k, (side, hole) in enumerate([(48, 8), (120, 60), (64, 16), (96, 40), (40, 6)]):
        img = G(side, side, 3, 30 + k)
        m = centered_mask(side, side, hole, hole)
        tp = np.ascontiguousarray(np.concatenate([m[:, :, None], img], axis=2))
        cp = np.ascontiguousarray(np.concatenate([(255 - m)[:, :, None], img], axis=2))
        p = abi.make_params(0, 0, 1, 0.5, 0.117, 12 + k, 40 + 10 * k)
        ref_t = tp.copy()
        assert R.engine(port, p, R.format_indices(port, 3, 0, False, False, False), ref_t, cp.copy()) == 0
        want.append(ref_t)
        jobs.append((p, fi, tp, cp))
    errs = api.engine_batch(jobs, slots=2, devices=_devices())
    assert not any(errs)
    for (_p, _f, tp, _c), w in zip(jobs, want):
        assert (tp == w).all()


def test_bad_job_reports_its_error_and_the_rest_still_run(built_oracle, built_lib):
    imgs, masks = _heal_jobs(3, 64, 16)
    masks = list(masks)
    masks[1] = np.zeros((64, 64), np.uint8)           # empty selection: IMAGE_SYNTH_ERROR_EMPTY_TARGET for that job only
    got = [i.copy() for i in imgs]
    errs = api.image_synth_batch(got, masks, abi.T_RGB, None, devices=[0], slots=2)
    assert errs[1] == 5 and errs[0] == 0 and errs[2] == 0
    assert (got[1] == imgs[1]).all() and (got[0] != imgs[0]).any()


def test_shared_corpus_is_built_once_per_device_and_results_equal_single_calls(built_oracle, built_lib):
    """One corpus, many targets: the jobs of a batch that pass the SAME corpus pixmap stage and
    prepare it once per device; every target comes out as from its own engine() call (= the oracle's)."""
    port = R.load_port(R.GPU_MODE)
    cor = G(72, 60, 3, 5)
    cmask = np.full((60, 72), 255, np.uint8); cmask[10:20, 30:50] = 0          # a corpus with a hole: point list != identity
    cp = np.ascontiguousarray(np.concatenate([cmask[:, :, None], cor], axis=2))
    p = abi.make_params(1, 0, 0, 0.5, 0.117, 9, 60)
    fi = api.format_indices(3)
    jobs, want = [], []
    for k in range(9):
        side = 24 + 4 * k
        tp = np.ascontiguousarray(np.concatenate([np.full((side, side, 1), 255, np.uint8), G(side, side, 3, 50 + k)], axis=2))
        ref_t = tp.copy()
        assert R.engine(port, p, R.format_indices(port, 3, 0, False, False, False), ref_t, cp.copy()) == 0
        want.append(ref_t)
        jobs.append((p, fi, tp, cp))                                          # the same corpus array in every job
    devs = _devices()
    b0 = api.shared_corpus_stats()
    errs = api.engine_batch(jobs, slots=3, devices=devs)
    b1 = api.shared_corpus_stats()
    assert not any(errs)
    for (_p, _f, tp, _c), w in zip(jobs, want):
        assert (tp == w).all()
    built, reused, peer = (b1[i] - b0[i] for i in range(3))
    assert 1 <= built + peer <= len(devs) and built >= 1       # once per device that took a job: from the host, or from a peer
    assert built + peer + reused == len(jobs)
    # a different corpus per job: nothing is shared
    jobs2 = [(p, fi, j[2].copy(), cp.copy()) for j in jobs[:3]]
    b2 = api.shared_corpus_stats()
    assert not any(api.engine_batch(jobs2, slots=2))
    assert api.shared_corpus_stats() == b2
