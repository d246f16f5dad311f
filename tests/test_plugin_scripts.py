"""not-gpu: the GIMP-free mirrors of plug_in_resynthesizer and the PluginScripts (resynthesizer_b200/plugin.py),
run on the compiled reference / the restatement instead of the CUDA engine, must reproduce the reference's own
golden images -- the same goldens Test/testResynth.py checks the real scripts against.  Needs /root/reference."""
import numpy as np
import pytest

from oracle import goldens
from oracle import refdriver as R
from resynthesizer_b200 import plugin

needs_ref = pytest.mark.skipif(not goldens.available(), reason="/root/reference not present")


def _oracle_engine(lib):
    def engine_fn(params, fi, tp, cp, **_kw):
        return R.engine(lib, params, fi, tp, cp)

    def fi_fn(n_color, n_map, at, ac, is_map):
        return R.format_indices(lib, n_color, n_map, at, ac, is_map)
    return dict(engine_fn=engine_fn, format_indices_fn=fi_fn)


def _lib():
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    if os.path.exists(os.path.join(here, "..", "oracle", "_ref", "libref_mt_1t.so")):
        return R.load("ref_mt_1t")
    return R.load_port(R.REF_MODE)


def _sel(shape_hw, rect):
    return goldens.rect_mask(shape_hw, rect)


@needs_ref
def test_heal_selection_script_goldens(built_oracle):
    kw = _oracle_engine(_lib())
    for img_name, golden, direction, order in (("ufo-input", "heal-ufo-input", 1, 1),
                                               ("wander", "healgray-wander", 1, 1),
                                               ("ufo-input", "healaroundrandom-ufo-input", 0, 0)):
        img = goldens.load_png(img_name)
        d = plugin.Drawable(img, _sel(img.shape[:2], goldens.SEL1))
        assert plugin.heal_selection(d, 50, direction, order, **kw) == 0
        assert (d.pixels == goldens.load_golden(golden)).all(), golden


@needs_ref
def test_uncrop_and_render_texture_goldens(built_oracle):
    kw = _oracle_engine(_lib())
    out = plugin.uncrop(plugin.Drawable(goldens.load_png("ufo-input")), 20, **kw)
    assert (out.pixels == goldens.load_golden("uncrop-ufo-input")).all()
    out = plugin.render_texture(plugin.Drawable(goldens.load_png("grass-input")), 2, 1, **kw)
    assert (out.pixels == goldens.load_golden("rendertexture-grass-input")).all()


@needs_ref
def test_plug_in_resynthesizer_goldens(built_oracle):
    kw = _oracle_engine(_lib())
    ufo = goldens.load_png("ufo-input")
    grass = goldens.load_png("grass-input")
    sel = _sel(ufo.shape[:2], goldens.SEL1)
    # testResynth.py:313-328: corpus = same drawable & selection / a second image / tileable without context
    d = plugin.Drawable(ufo.copy(), sel)
    plugin.plug_in_resynthesizer(d, 0, 0, 1, plugin.Drawable(ufo.copy(), sel.copy()), None, None, 0.0, 0.117, 16, 500, **kw)
    assert (d.pixels == goldens.load_golden("resynth-ufo-input")).all()
    d = plugin.Drawable(ufo.copy(), sel)
    plugin.plug_in_resynthesizer(d, 0, 0, 1, plugin.Drawable(grass), None, None, 0.0, 0.117, 16, 500, **kw)
    assert (d.pixels == goldens.load_golden("resynthtwoimages-ufo-input")).all()
    d = plugin.Drawable(ufo.copy(), sel)
    plugin.plug_in_resynthesizer(d, 1, 1, 0, plugin.Drawable(grass), None, None, 0.0, 0.117, 16, 500, **kw)
    assert (d.pixels == goldens.load_golden("resynthtileable-ufo-input")).all()


@needs_ref
def test_map_style_golden(built_oracle):
    kw = _oracle_engine(_lib())
    d = plugin.Drawable(goldens.load_png("wander"))
    plugin.map_style(d, plugin.Drawable(goldens.load_png("wander-texture")), 50, 0, **kw)
    assert (d.pixels == goldens.load_golden("mapstylegraygray-wander")).all()


def test_plugin_error_messages():
    rgb = plugin.Drawable(np.zeros((8, 8, 3), np.uint8))
    gray = plugin.Drawable(np.zeros((8, 8, 1), np.uint8))
    kw = dict(engine_fn=lambda *a, **k: 0, format_indices_fn=lambda *a: None)
    with pytest.raises(plugin.PluginError, match="same number of color channels"):
        plugin.plug_in_resynthesizer(rgb, 0, 0, 1, gray, **kw)
    with pytest.raises(plugin.PluginError, match="same size as the input texture"):
        plugin.plug_in_resynthesizer(rgb, 0, 0, 1, plugin.Drawable(np.zeros((8, 8, 3), np.uint8)),
                                     plugin.Drawable(np.zeros((4, 4, 3), np.uint8)), rgb.copy(), **kw)
    with pytest.raises(plugin.PluginError, match="select a region"):
        plugin.heal_selection(rgb, **kw)
    assert abs(plugin.calculate_map_weight(50) - 0.25012680) < 1e-7
    assert abs(plugin.calculate_map_weight(10) - 0.39778528) < 1e-7
