"""-m gpu: the CUDA engine, called through the C-ABI, must equal the oracle's sequential definition of its
semantics (oracle/resynth_port.c in GPU_MODE) bit for bit: every output pixel, every counter."""
import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _compare_simple(img, mask, fmt, params, seed=1198472):
    port = R.load_port(R.GPU_MODE, seed)
    e_ref, out_ref = R.image_synth(port, img, mask, fmt, params)
    ps = R.port_stats(port)
    api.set_seed(seed)
    out = img.copy()
    e = api.image_synth(out, mask, fmt, params)
    st = api.last_stats()
    assert e == e_ref == 0
    nd = int((out != out_ref).any(axis=2).sum())
    assert nd == 0, "%d pixels differ from the oracle" % nd
    assert st["passes_run"] == ps["passes_run"]
    assert st["betters"] == ps["betters"]
    assert st["sum_best"] == ps["sum_best"]
    assert st["visits"] == ps["visits"]
    assert st["evals"] == ps["evals"]
    assert st["perfect"] == ps["perfect"]
    return st, ps


@pytest.mark.parametrize("ctx", list(range(1, 9)))
def test_heal_rgb_all_orderings(built_oracle, built_lib, ctx):
    img = G(96, 80, 3, 11)
    mask = centered_mask(96, 80, 32, 24)
    _compare_simple(img, mask, abi.T_RGB, abi.make_params(0, 0, ctx, 0.5, 0.117, 16, 60))


@pytest.mark.parametrize("fmt,c", [(abi.T_RGB, 3), (abi.T_RGBA, 4), (abi.T_Gray, 1), (abi.T_GrayA, 2)])
def test_formats(built_oracle, built_lib, fmt, c):
    img = G(72, 64, c, 5)
    if c in (2, 4):
        img[:, 8:14, c - 1] = 0       # a transparent band: excluded from corpus and context
        img[:, 14:, c - 1] = 255
    mask = centered_mask(72, 64, 24, 20)
    mask[10:14, 40:44] = 100           # partially selected: target, never corpus
    _compare_simple(img, mask, fmt, abi.make_params(0, 0, 1, 0.5, 0.117, 30, 100))


def test_default_params_cfg1_shape(built_oracle, built_lib):
    img = G(128, 128, 3, 12345)
    mask = centered_mask(128, 128, 32, 32)
    _compare_simple(img, mask, abi.T_RGB, None)


def _engine_case(tw, th, cw, ch, n_color, n_map, alpha, params, full_target, seed=1198472):
    rng = np.random.RandomState(3)
    tgt = G(tw, th, n_color, 21)
    cor = G(cw, ch, n_color, 22)
    tmask = np.full((th, tw), 255, np.uint8) if full_target else centered_mask(tw, th, tw // 3, th // 3)
    cmask = np.full((ch, cw), 255, np.uint8)
    cmask[:3, :5] = 0
    ta = ca = None
    if alpha:
        ta = np.full((th, tw), 255, np.uint8); ta[::7, ::5] = 0
        ca = np.full((ch, cw), 255, np.uint8); ca[::6, ::4] = 0
    tmaps = cmaps = None
    if n_map:
        tmaps = G(tw, th, n_map, 31)
        cmaps = G(cw, ch, n_map, 32)
    port = R.load_port(R.GPU_MODE, seed)
    fi = R.format_indices(port, n_color, n_map, alpha, alpha, n_map > 0)
    tp = R.build_pixmap(tmask, tgt, ta, tmaps)
    cp = R.build_pixmap(cmask, cor, ca, cmaps)
    tp2, cp2 = tp.copy(), cp.copy()
    assert R.engine(port, params, fi, tp, cp) == 0
    ps = R.port_stats(port)
    api.set_seed(seed)
    assert api.engine(params, fi, tp2, cp2) == 0
    st = api.last_stats()
    assert int((tp != tp2).any(axis=2).sum()) == 0
    assert st["betters"] == ps["betters"] and st["sum_best"] == ps["sum_best"] and st["evals"] == ps["evals"]


def test_render_texture_no_context_tiled(built_oracle, built_lib):
    _engine_case(64, 48, 32, 32, 3, 0, False, abi.make_params(1, 1, 0, 0.0, 0.117, 9, 50), True)


def test_render_texture_no_context_untiled(built_oracle, built_lib):
    _engine_case(64, 48, 32, 32, 3, 0, False, abi.make_params(0, 0, 0, 0.0, 0.117, 9, 50), True)


def test_map_style_rgb_maps(built_oracle, built_lib):
    _engine_case(48, 40, 40, 36, 3, 3, False, abi.make_params(1, 1, 1, 0.5, 0.117, 9, 60), True)


def test_map_style_gray_map_alpha(built_oracle, built_lib):
    _engine_case(48, 40, 40, 36, 3, 1, True, abi.make_params(0, 0, 1, 0.25, 0.117, 12, 60), False)


def test_gray_with_gray_map(built_oracle, built_lib):
    _engine_case(40, 40, 30, 30, 1, 1, False, abi.make_params(1, 1, 1, 0.4, 0.117, 9, 40), True)


@pytest.mark.parametrize("patch", [0, 1, 2, 64])
def test_patch_size_edges(built_oracle, built_lib, patch):
    img = G(48, 48, 3, 9)
    mask = centered_mask(48, 48, 12, 12)
    _compare_simple(img, mask, abi.T_RGB, abi.make_params(0, 0, 1, 0.5, 0.117, patch, 30))


def test_zero_probes(built_oracle, built_lib):
    img = G(48, 48, 3, 9)
    mask = centered_mask(48, 48, 12, 12)
    _compare_simple(img, mask, abi.T_RGB, abi.make_params(0, 0, 1, 0.5, 0.117, 8, 0))


def test_other_seed(built_oracle, built_lib):
    img = G(64, 64, 3, 77)
    mask = centered_mask(64, 64, 20, 20)
    _compare_simple(img, mask, abi.T_RGB, abi.make_params(0, 0, 2, 0.5, 0.117, 20, 80), seed=4242)


def test_larger_job_is_deterministic_and_exact(built_oracle, built_lib):
    """9216 targets, many epochs in flight: five runs must be identical to each other and to the oracle."""
    img = G(256, 256, 3, 5150)
    mask = centered_mask(256, 256, 96, 96)
    st, ps = _compare_simple(img, mask, abi.T_RGB, None)
    first = None
    for _ in range(4):
        out = img.copy()
        assert api.image_synth(out, mask, abi.T_RGB, None) == 0
        if first is None:
            first = out
        assert (out == first).all()
        assert api.last_stats()["sum_best"] == st["sum_best"]


def test_render_texture_bench_shape_small(built_oracle, built_lib):
    """cfg2's shape at 1/8 scale: 128x128 target from a 64x64 corpus, ctx 0, 9/200 (dependency-heavy pass 0)."""
    _engine_case(128, 128, 64, 64, 3, 0, False, abi.make_params(0, 0, 0, 0.5, 0.117, 9, 200), True)


@pytest.mark.parametrize("case", ["tiled_maps", "heal_alpha", "texture"])
def test_patches_of_later_passes_gathered_up_front(built_oracle, built_lib, case, monkeypatch):
    """Jobs of 2 Mi+ target points gather the patches of the passes >= 1 once (k_gather_later) instead of scanning the
    offsets in every pass; here the same path is forced on small jobs and compared with the oracle bit for bit."""
    monkeypatch.setenv("RS_LATER_LISTS_MIN", "1")
    if case == "tiled_maps":
        _engine_case(48, 40, 40, 36, 3, 3, False, abi.make_params(1, 1, 1, 0.5, 0.117, 9, 60), True)
    elif case == "heal_alpha":
        _engine_case(72, 60, 72, 60, 3, 0, True, abi.make_params(0, 0, 1, 0.5, 0.117, 30, 80), False)
    else:
        _engine_case(64, 48, 32, 32, 3, 0, False, abi.make_params(0, 1, 0, 0.0, 0.117, 16, 50), True)
