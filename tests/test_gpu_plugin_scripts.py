"""-m gpu: the plug-in / script mirrors on the CUDA engine equal the same recipes run on the oracle's GPU mode."""
import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import plugin
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _oracle_kw():
    lib = R.load_port(R.GPU_MODE)
    return dict(engine_fn=lambda p, fi, tp, cp, **k: R.engine(lib, p, fi, tp, cp),
                format_indices_fn=lambda *a: R.format_indices(lib, *a))


def test_heal_selection_all_directions(built_oracle, built_lib):
    img = G(140, 120, 3, 9)
    sel = np.zeros((120, 140), np.uint8); sel[50:75, 60:95] = 255
    for direction, order in ((0, 0), (1, 1), (2, 2), (0, 1)):
        a = plugin.Drawable(img.copy(), sel); b = plugin.Drawable(img.copy(), sel)
        assert plugin.heal_selection(a, 20, direction, order) == 0
        assert plugin.heal_selection(b, 20, direction, order, **_oracle_kw()) == 0
        assert (a.pixels == b.pixels).all() and (a.pixels != img).any()


def test_uncrop_render_texture_map_style(built_oracle, built_lib):
    img = G(90, 70, 3, 4)
    a = plugin.uncrop(plugin.Drawable(img.copy()), 20)
    b = plugin.uncrop(plugin.Drawable(img.copy()), 20, **_oracle_kw())
    assert a.pixels.shape == (84, 108, 3) and (a.pixels == b.pixels).all()
    a = plugin.render_texture(plugin.Drawable(G(40, 36, 3, 5)), 2, 1)
    b = plugin.render_texture(plugin.Drawable(G(40, 36, 3, 5)), 2, 1, **_oracle_kw())
    assert a.pixels.shape == (72, 80, 3) and (a.pixels == b.pixels).all()
    t1, t2 = plugin.Drawable(G(64, 48, 3, 6)), plugin.Drawable(G(64, 48, 3, 6))
    src = plugin.Drawable(G(50, 44, 3, 7))
    plugin.map_style(t1, src, 50, 0)
    plugin.map_style(t2, src, 50, 0, **_oracle_kw())
    assert (t1.pixels == t2.pixels).all()


def test_heal_transparency_rgba(built_oracle, built_lib):
    img = G(80, 64, 4, 8); img[:, :, 3] = 255; img[20:30, 30:44, 3] = 0
    a, b = plugin.Drawable(img.copy()), plugin.Drawable(img.copy())
    assert plugin.heal_transparency(a, 16, 2) == 0
    assert plugin.heal_transparency(b, 16, 2, **_oracle_kw()) == 0
    assert (a.pixels == b.pixels).all() and (a.pixels[:, :, 3] == img[:, :, 3]).all()
