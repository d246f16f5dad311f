"""-m gpu: EVERY variant of the pass kernels against the oracle, bit for bit.

The launch plan of a pass (throughput kernel = one warp per visit, or teams of 2/4/8 warps per visit; pass 0 cut into
segments of shrinking width; chunk size 2 or 3; corpus in shared memory or in L2) follows the job size, so small jobs
under the default plan only ever run the 8-wide team kernel.  Here the knobs RS_TEAM_P0 / RS_TEAM_PN / RS_SEG_P0 /
RS_CHUNK / RS_SMEM_CORPUS force each variant on jobs the oracle finishes in seconds, and pixels, sources, betters,
sum of best distances, evals and perfect matches must all equal oracle/resynth_port.c in GPU mode
(lib/synthesize.h:426-642 restated sequentially)."""
import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _case(kind):
    """(params, n_color, n_map, alpha, target pixmap, corpus pixmap)"""
    if kind == "heal30":        # CH=3 kernels, context, K=30
        img = G(112, 96, 3, 41)
        m = centered_mask(112, 96, 44, 40)
        return (abi.make_params(0, 0, 1, 0.5, 0.117, 30, 120), 3, 0, False,
                R.build_pixmap(m, img), R.build_pixmap(255 - m, img))
    if kind == "texture9":      # CH=2 kernels, no context, K=9, small corpus (the render-texture shape)
        tgt = np.full((96, 96, 3), 255, np.uint8)
        cor = G(48, 40, 3, 42)
        cm = np.full((40, 48), 255, np.uint8); cm[:2, :7] = 0
        return (abi.make_params(0, 0, 0, 0.5, 0.117, 9, 150), 3, 0, False,
                R.build_pixmap(np.full((96, 96), 255, np.uint8), tgt), R.build_pixmap(cm, cor))
    if kind == "texture9_tiled":
        tgt = np.full((80, 72, 3), 255, np.uint8)
        cor = G(64, 64, 3, 43)
        return (abi.make_params(1, 1, 0, 0.5, 0.117, 9, 100), 3, 0, False,
                R.build_pixmap(np.full((80, 72), 255, np.uint8), tgt), R.build_pixmap(np.full((64, 64), 255, np.uint8), cor))
    if kind == "maps9":         # MAPS kernels, CH=2
        tgt = G(72, 64, 3, 44); cor = G(56, 60, 3, 45)
        full_t = np.full((64, 72), 255, np.uint8); full_c = np.full((60, 56), 255, np.uint8)
        return (abi.make_params(1, 1, 1, 0.5, 0.117, 9, 100), 3, 3, False,
                R.build_pixmap(full_t, tgt, None, G(72, 64, 3, 46)), R.build_pixmap(full_c, cor, None, G(56, 60, 3, 47)))
    if kind == "maps20_alpha":  # MAPS kernels, CH=3, alpha channel, gray map
        tgt = G(64, 64, 3, 48); cor = G(64, 48, 3, 49)
        tm = centered_mask(64, 64, 40, 36)
        ta = np.full((64, 64), 255, np.uint8); ta[::7, ::5] = 0
        ca = np.full((48, 64), 255, np.uint8); ca[::6, ::4] = 0
        return (abi.make_params(0, 0, 1, 0.3, 0.117, 20, 90), 3, 1, True,
                R.build_pixmap(tm, tgt, ta, G(64, 64, 1, 50)), R.build_pixmap(np.full((48, 64), 255, np.uint8), cor, ca, G(64, 48, 1, 51)))
    if kind == "gray16":
        img = G(90, 70, 1, 52)
        m = centered_mask(90, 70, 30, 30)
        return (abi.make_params(0, 0, 2, 0.5, 0.117, 16, 80), 1, 0, False,
                R.build_pixmap(m, img), R.build_pixmap(255 - m, img))
    if kind == "heal_tiled":    # context AND tiling: the first visits' patches stay with the cooperative scan
        img = G(96, 80, 3, 53)
        m = centered_mask(96, 80, 50, 44)
        return (abi.make_params(1, 1, 1, 0.5, 0.117, 24, 60), 3, 0, False,
                R.build_pixmap(m, img), R.build_pixmap(255 - m, img))
    if kind == "heal64_deep":   # the largest patch (63 neighbours) in a hole whose middle is far from any context
        img = G(150, 140, 3, 54)
        m = centered_mask(150, 140, 110, 100)
        m[20:24, 30:90] = 0     # a strip of context inside the hole, and context that is not usable (ragged mask)
        return (abi.make_params(0, 0, 1, 0.5, 0.117, 64, 40), 3, 0, False,
                R.build_pixmap(m, img), R.build_pixmap(255 - m, img))
    if kind == "texture9_htile":  # wrapped aliases in one direction only, offsets table narrower than the target
        tgt = np.full((70, 90, 3), 255, np.uint8)
        cor = G(40, 36, 3, 55)
        return (abi.make_params(1, 0, 0, 0.5, 0.117, 9, 60), 3, 0, False,
                R.build_pixmap(np.full((70, 90), 255, np.uint8), tgt), R.build_pixmap(np.full((36, 40), 255, np.uint8), cor))
    raise KeyError(kind)


def _check(kind, seed=1198472):
    params, n_color, n_map, alpha, tp, cp = _case(kind)
    port = R.load_port(R.GPU_MODE, seed)
    fi = R.format_indices(port, n_color, n_map, alpha, alpha, n_map > 0)
    want = tp.copy()
    assert R.engine(port, params, fi, want, cp) == 0
    ps = R.port_stats(port)
    t_ref, s_ref = R.port_last_result(port)
    api.set_seed(seed)
    api.order_cache(False)
    api.keep_result(True)
    try:
        got = tp.copy()
        assert api.engine(params, fi, got, cp) == 0
        st = api.last_stats()
        t_gpu, s_gpu = api.last_result()
    finally:
        api.keep_result(False)
        api.order_cache(True)
    assert (t_gpu == t_ref).all(), "visit order differs"
    nd = int((got != want).any(axis=2).sum())
    assert nd == 0, "%d pixels differ from the oracle" % nd
    assert (s_gpu == s_ref).all(), "%d sources differ" % int((s_gpu != s_ref).any(axis=1).sum())
    for k in ("passes_run", "betters", "sum_best", "visits", "pass_visits", "evals", "perfect", "heur_evals"):
        assert st[k] == ps[k], (k, st[k], ps[k])
    return st


KINDS = ["heal30", "texture9", "texture9_tiled", "maps9", "maps20_alpha", "gray16"]


@pytest.mark.parametrize("width", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", KINDS)
def test_every_width_in_every_pass(built_oracle, built_lib, monkeypatch, kind, width):
    """k_synth_pass (width 1) and k_synth_pass_team (2/4/8) for pass 0 AND the later passes."""
    monkeypatch.setenv("RS_TEAM_P0", str(width))
    monkeypatch.setenv("RS_TEAM_PN", str(width))
    _check(kind)


@pytest.mark.parametrize("p0,pn", [(8, 1), (1, 8), (4, 2), (2, 4)])
def test_mixed_widths(built_oracle, built_lib, monkeypatch, p0, pn):
    monkeypatch.setenv("RS_TEAM_P0", str(p0))
    monkeypatch.setenv("RS_TEAM_PN", str(pn))
    _check("heal30")
    _check("texture9")


@pytest.mark.parametrize("plan", ["512:8,2048:4,6000:2,0:1", "100:2,101:8,3000:1,0:4", "64:1,0:8", "4000:4,0:1"])
@pytest.mark.parametrize("kind", ["texture9", "heal30", "maps9"])
def test_multi_segment_pass0(built_oracle, built_lib, monkeypatch, plan, kind):
    """Pass 0 cut into launches of different widths (plan_segments): visits hand over between kernels mid-pass."""
    monkeypatch.setenv("RS_SEG_P0", plan)
    monkeypatch.setenv("RS_TEAM_PN", "1")
    _check(kind)


@pytest.mark.parametrize("chunk", [2, 3])
@pytest.mark.parametrize("width", [1, 4])
def test_chunk_size_override(built_oracle, built_lib, monkeypatch, chunk, width):
    """Both chunk sizes (neighbours per early-out check) on both patch-size classes."""
    monkeypatch.setenv("RS_CHUNK", str(chunk))
    monkeypatch.setenv("RS_TEAM_P0", str(width))
    monkeypatch.setenv("RS_TEAM_PN", str(width))
    _check("heal30")
    _check("texture9")
    _check("maps9")


@pytest.mark.parametrize("width", [1, 2, 8])
def test_smem_corpus_off_equals_on(built_oracle, built_lib, monkeypatch, width):
    """Corpora without map channels that fit are staged into shared memory by TMA bulk copies when the throughput kernel
    runs (width 1) -- whole in one CTA, or split over a 2-CTA cluster and read through distributed shared memory;
    RS_SMEM_CORPUS=0 forces the L2 path, =2 the cluster split.  All must equal the oracle."""
    monkeypatch.setenv("RS_TEAM_P0", str(width))
    monkeypatch.setenv("RS_TEAM_PN", str(width))
    for flag in ("0", "1", "2"):   # L2 path / as it fits (one CTA here) / split over the two CTAs of a cluster (DSMEM)
        monkeypatch.setenv("RS_SMEM_CORPUS", flag)
        _check("texture9")
        _check("texture9_tiled")
        _check("gray16")
        _check("heal30")


@pytest.mark.parametrize("width", [1, 4])
def test_corpus_point_lookup_paths(built_oracle, built_lib, monkeypatch, width):
    """The idx-th corpus point three ways (rs_corpus_point): identity when every corpus pixel is usable (texture9_tiled,
    maps9), bitmap select for dense selections from RS_SELECT_MIN points on (forced here: heal30, gray16, texture9 have
    holes in their corpora), and the point table (RS_NO_CORPUS_BITS).  Same points, so the same images as the oracle."""
    monkeypatch.setenv("RS_TEAM_P0", str(width))
    monkeypatch.setenv("RS_TEAM_PN", str(width))
    monkeypatch.setenv("RS_SELECT_MIN", "0")
    for kind in ("heal30", "gray16", "texture9", "texture9_tiled", "maps9", "maps20_alpha"):
        _check(kind)
    monkeypatch.delenv("RS_SELECT_MIN")
    monkeypatch.setenv("RS_NO_CORPUS_BITS", "1")
    _check("heal30")
    _check("texture9")


@pytest.mark.parametrize("probe", ["0", "64"])
@pytest.mark.parametrize("kind", KINDS + ["heal_tiled", "heal64_deep", "texture9_htile"])
def test_pass0_patches_by_search(built_oracle, built_lib, monkeypatch, kind, probe):
    """k_gather_pass0_sparse: the patches of the first 8192 visits of pass 0 come from a search over the earlier target
    points (+ their wrapped aliases) and the context blocks instead of a scan of the sorted offsets table
    (lib/synthesize.h:189-241).  By default a visit searches only when the first 512 table entries do not fill its
    patch -- never, on jobs this small; RS_SPARSE_PROBE=0 makes every visit search, =64 mixes both."""
    monkeypatch.setenv("RS_SPARSE_PROBE", probe)
    _check(kind)


@pytest.mark.parametrize("kind", ["heal30", "texture9_tiled", "heal_tiled", "heal64_deep", "texture9_htile"])
def test_pass0_patches_by_scan(built_oracle, built_lib, monkeypatch, kind):
    """RS_SPARSE_GATHER=0: the cooperative scan for the first visits, as for jobs with a caller's offsets table."""
    monkeypatch.setenv("RS_SPARSE_GATHER", "0")
    _check(kind)


@pytest.mark.parametrize("mode", ["0", "1", "2"])
@pytest.mark.parametrize("kind", ["texture9", "texture9_tiled", "maps9", "texture9_htile", "gray16"])
def test_two_visits_per_warp(built_oracle, built_lib, monkeypatch, kind, mode):
    """k_synth_pass<..., 16>: patches of at most 16 neighbours run TWO consecutive visits per warp, side by side in its
    halves through geometry / values / candidates / commit and one after the other through the distance loop; a second
    visit whose patch holds the first one's pixel runs after it.  RS_PAIR=0: one visit per warp; 1: the default; 2: pairs
    for spatially sorted orders too (gray16 visits by rows... every second visit then depends on the first)."""
    monkeypatch.setenv("RS_PAIR", mode)
    monkeypatch.setenv("RS_TEAM_P0", "1")
    monkeypatch.setenv("RS_TEAM_PN", "1")
    _check(kind)
    monkeypatch.setenv("RS_SMEM_CORPUS", "2")
    _check(kind)


@pytest.mark.parametrize("width", [1, 2, 4, 8])
def test_later_lists_with_every_width(built_oracle, built_lib, monkeypatch, width):
    monkeypatch.setenv("RS_LATER_LISTS_MIN", "1")
    monkeypatch.setenv("RS_TEAM_P0", str(width))
    monkeypatch.setenv("RS_TEAM_PN", str(width))
    _check("heal30")
    _check("maps9")


def test_mid_size_default_plans(built_oracle, built_lib):
    """Default plans of the size classes between 'always 8 wide' and 'segments + throughput kernel'
    (pass_width: <= 32 Ki: 8/8; <= 200 k: 8/4; <= 600 k: 4/2): a 240x200 hole (48 000 targets) and a 512x512 texture
    (262 144 targets)."""
    img = G(400, 320, 3, 61)
    m = centered_mask(400, 320, 240, 200)
    port = R.load_port(R.GPU_MODE)
    p = abi.make_params(0, 0, 1, 0.5, 0.117, 12, 40)
    e_ref, want = R.image_synth(port, img, m, abi.T_RGB, p)
    ps = R.port_stats(port)
    got = img.copy()
    api.set_seed(1198472)
    assert api.image_synth(got, m, abi.T_RGB, p) == 0 and e_ref == 0
    st = api.last_stats()
    assert (got == want).all() and st["evals"] == ps["evals"] and st["betters"] == ps["betters"] and st["sum_best"] == ps["sum_best"]

    tgt = np.full((512, 512, 3), 255, np.uint8)
    cor = G(128, 128, 3, 62)
    fi = R.format_indices(port, 3, 0, False, False, False)
    tp = R.build_pixmap(np.full((512, 512), 255, np.uint8), tgt)
    cp = R.build_pixmap(np.full((128, 128), 255, np.uint8), cor)
    p = abi.make_params(0, 0, 0, 0.5, 0.117, 9, 60)
    want = tp.copy()
    assert R.engine(port, p, fi, want, cp) == 0
    ps = R.port_stats(port)
    got = tp.copy()
    assert api.engine(p, fi, got, cp) == 0
    st = api.last_stats()
    assert (got == want).all() and st["evals"] == ps["evals"] and st["betters"] == ps["betters"] and st["sum_best"] == ps["sum_best"]
