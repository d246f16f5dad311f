"""TEST INFRASTRUCTURE (oracle/): the reference's own golden images, without GIMP.

Each recipe restates what one case of Test/testResynth.py (line cited) feeds to
`plug_in_resynthesizer` after the GIMP script/adapter has prepared drawables,
selections and maps (src/resynthesizer/adaptGimp.h:294-322,
PluginScripts/plugin-heal-selection.py:64-148, plugin-uncrop.py:78-129,
plugin-map-style.py:245-359, plugin-render-texture.py:175), and compares the
engine output with Test/reference_out_images/<golden>.ppm.

Needs /root/reference (input PNGs + goldens): used by `-m "not gpu"` tests in
the build container only; never by GPU tests, smoke() or bench.py.
"""
import math
import os

import numpy as np

from resynthesizer_b200 import abi
from . import refdriver as R

REF_ROOT = os.environ.get("REF_ROOT", "/root/reference")
IN_DIR = os.path.join(REF_ROOT, "Test", "in_images")
GOLD_DIR = os.path.join(REF_ROOT, "Test", "reference_out_images")

SEL1 = (100, 90, 100, 50)   # x, y, w, h  (testResynth.py:81)
SEL2 = (90, 175, 135, 100)  # (testResynth.py:82)


# The reference's test images cannot travel in this repository's history; oracle/make_recipe_images.py (run by
# __graft_entry__.build() where /root/reference exists) packs the inputs and goldens of the recipes below into
# oracle/_ref/recipe_images.npz -- git-ignored like the compiled reference beside it, and shipped to the GPU box the
# same way.  The loaders fall back to it when /root/reference is absent.
PACK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "recipe_images.npz")
INPUT_NAMES = ["brick", "ufo-input", "zap-texture", "donkey_original", "grass-input", "grass-input-alpha", "wander",
               "ufo-input-w-alpha-gray", "ufo-input-w-alpha", "angel_target", "angel_texture", "wander-texture"]
_pack = None


def _packed(key):
    global _pack
    if _pack is None:
        _pack = np.load(PACK)
    return np.ascontiguousarray(_pack[key])


def available():
    return (os.path.isdir(IN_DIR) and os.path.isdir(GOLD_DIR)) or os.path.exists(PACK)


def load_png(name):
    if not os.path.isdir(IN_DIR):
        return _packed("in/" + name)
    from PIL import Image
    im = Image.open(os.path.join(IN_DIR, name + ".png"))
    a = np.asarray(im)
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a.astype(np.uint8))


def load_golden(name):
    """ASCII P3/P2 -> (h, w, c) uint8."""
    if not os.path.isdir(GOLD_DIR):
        return _packed("gold/" + name)
    with open(os.path.join(GOLD_DIR, name + ".ppm"), "rb") as f:
        toks = []
        for line in f:
            line = line.split(b"#")[0]
            toks.extend(line.split())
    magic = toks[0]
    w, h, _maxv = int(toks[1]), int(toks[2]), int(toks[3])
    c = 3 if magic == b"P3" else 1
    data = np.array(toks[4:], dtype=np.int32).astype(np.uint8)
    return data.reshape(h, w, c)


def rect_mask(shape_hw, sel):
    h, w = shape_hw
    x, y, sw, sh = sel
    m = np.zeros((h, w), np.uint8)
    m[max(y, 0):min(y + sh, h), max(x, 0):min(x + sw, w)] = 255
    return m


def gimp_grow(mask, r):
    """GIMP selection-grow kernel (circular): (x,y) in grow(R,r) iff some
    (x',y') in R has |y-y'| <= c(|x-x'|), c(0)=r, c(d)=rint(sqrt(r^2-(d-.5)^2))."""
    h, w = mask.shape
    sel = mask > 0
    out = np.zeros_like(sel)
    for d in range(0, r + 1):
        c = r if d == 0 else int(np.rint(math.sqrt(r * r - (d - 0.5) ** 2)))
        # column-wise dilation by c, then shift by +-d in x
        col = np.zeros_like(sel)
        csum = np.cumsum(np.pad(sel, ((c + 1, c), (0, 0))).astype(np.int32), axis=0)
        col = (csum[2 * c + 1:] - csum[:-(2 * c + 1)]) > 0
        col = col[:h]
        if d == 0:
            out |= col
        else:
            out[:, d:] |= col[:, :-d]
            out[:, :-d] |= col[:, d:]
    return out.astype(np.uint8) * 255


def flatten_white(rgba):
    """GIMP flatten onto white background (approximate rounding; +-1 on partial alpha)."""
    a = rgba[:, :, -1:].astype(np.float64) / 255.0
    c = rgba[:, :, :-1].astype(np.float64)
    return np.clip(np.rint(c * a + 255.0 * (1 - a)), 0, 255).astype(np.uint8)


def _run(lib, params, n_color, target_color, target_mask, corpus_color, corpus_mask,
         t_alpha=None, c_alpha=None, t_maps=None, c_maps=None):
    has_alpha = t_alpha is not None or c_alpha is not None
    is_map = t_maps is not None
    n_map = t_maps.shape[2] if is_map else 0
    fi = R.format_indices(lib, n_color, n_map, t_alpha is not None, c_alpha is not None, is_map)
    if has_alpha:
        if t_alpha is None:
            t_alpha = np.full(target_mask.shape, 255, np.uint8)
        if c_alpha is None:
            c_alpha = np.full(corpus_mask.shape, 255, np.uint8)
    tp = R.build_pixmap(target_mask, target_color, t_alpha if has_alpha else None, t_maps)
    cp = R.build_pixmap(corpus_mask, corpus_color, c_alpha if has_alpha else None, c_maps)
    err = R.engine(lib, params, fi, tp, cp)
    assert err == 0, err
    out = tp[:, :, 1:1 + n_color]
    if t_alpha is not None:
        out = np.concatenate([out, tp[:, :, 1 + n_color:2 + n_color]], axis=2)
    return out, (fi, tp, cp)


def map_weight(percent):
    return math.acos((percent / 100.0) * 2 - 1) / (2 * 3.14)  # plugin-map-style.py:256


def heal_corpus(img, sel, radius, direction):
    """plugin-heal-selection.py:64-116: frisket = grow(sel) - sel, cropped by direction."""
    h, w = img.shape[:2]
    selm = rect_mask((h, w), sel)
    grown = gimp_grow(selm, radius)
    fr = (grown > 0) & (selm == 0)
    ys, xs = np.nonzero(grown)
    fx0, fx1, fy0, fy1 = xs.min(), xs.max() + 1, ys.min(), ys.max() + 1
    tys, txs = np.nonzero(selm)
    tx0, tx1, ty0, ty1 = txs.min(), txs.max() + 1, tys.min(), tys.max() + 1
    if direction == 0:
        x0, y0, cw, ch = fx0, fy0, fx1 - fx0, fy1 - fy0
    elif direction == 1:
        x0, y0, cw, ch = fx0, ty0, fx1 - fx0, ty1 - ty0
    else:
        x0, y0, cw, ch = tx0, fy0, tx1 - tx0, fy1 - fy0
    cw = min(w - x0, cw)
    ch = min(h - y0, ch)
    crop = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])
    cmask = np.ascontiguousarray(fr[y0:y0 + ch, x0:x0 + cw].astype(np.uint8) * 255)
    return crop, cmask


# name -> (golden file stem, exact?, builder(lib) -> output image (h,w,c) comparable with the golden)
def _resynthfull(img_name, sel):
    def f(lib):
        img = load_png(img_name)
        m = rect_mask(img.shape[:2], sel)
        p = abi.make_params(0, 0, 1, 0.0, 0.117, 16, 500)
        out, _ = _run(lib, p, 3, img, m, img, 255 - m)
        return out
    return f


def _resynth_same(lib):
    img = load_png("ufo-input")
    m = rect_mask(img.shape[:2], SEL1)
    p = abi.make_params(0, 0, 1, 0.0, 0.117, 16, 500)
    return _run(lib, p, 3, img, m, img, m.copy())[0]


def _resynth_two(h, v, ctx):
    def f(lib):
        img = load_png("ufo-input")
        grass = load_png("grass-input")
        m = rect_mask(img.shape[:2], SEL1)
        p = abi.make_params(h, v, ctx, 0.0, 0.117, 16, 500)
        return _run(lib, p, 3, img, m, grass, np.full(grass.shape[:2], 255, np.uint8))[0]
    return f


def _rendertexture(lib):
    grass = load_png("grass-input")
    h, w = grass.shape[0] * 2, grass.shape[1] * 2
    tgt = np.full((h, w, 3), 255, np.uint8)
    p = abi.make_params(1, 1, 0, 0.0, 0.117, 9, 200)
    return _run(lib, p, 3, tgt, np.full((h, w), 255, np.uint8), grass,
                np.full(grass.shape[:2], 255, np.uint8))[0]


def _rendertexture_alpha(lib):
    grass = load_png("grass-input-alpha")
    h, w = grass.shape[0] * 2, grass.shape[1] * 2
    tgt = np.full((h, w, 3), 255, np.uint8)
    p = abi.make_params(1, 1, 0, 0.0, 0.117, 9, 200)
    out, _ = _run(lib, p, 3, tgt, np.full((h, w), 255, np.uint8), grass[:, :, :3],
                  np.full(grass.shape[:2], 255, np.uint8),
                  t_alpha=np.full((h, w), 255, np.uint8), c_alpha=np.ascontiguousarray(grass[:, :, 3]))
    return flatten_white(out)


def _heal(img_name, n_color, radius, direction, ctx):
    def f(lib):
        img = load_png(img_name)
        color = img[:, :, :n_color]
        alpha = np.ascontiguousarray(img[:, :, n_color]) if img.shape[2] > n_color else None
        m = rect_mask(img.shape[:2], SEL1)
        crop, cmask = heal_corpus(img, SEL1, radius, direction)
        c_alpha = np.ascontiguousarray(crop[:, :, n_color]) if alpha is not None else None
        p = abi.make_params(0, 0, ctx, 0.0, 0.117, 16, 500)
        out, _ = _run(lib, p, n_color, np.ascontiguousarray(color), m,
                      np.ascontiguousarray(crop[:, :, :n_color]), cmask, t_alpha=alpha, c_alpha=c_alpha)
        return flatten_white(out) if alpha is not None else out
    return f


def _mapstyle(target_name, source_name, percent, n_color):
    def f(lib):
        tgt = load_png(target_name)[:, :, :n_color]
        src = load_png(source_name)[:, :, :n_color]
        tgt = np.ascontiguousarray(tgt)
        src = np.ascontiguousarray(src)
        p = abi.make_params(1, 1, 1, map_weight(percent), 0.117, 9, 200)
        return _run(lib, p, n_color, tgt, np.full(tgt.shape[:2], 255, np.uint8), src,
                    np.full(src.shape[:2], 255, np.uint8), t_maps=tgt.copy(), c_maps=src.copy())[0]
    return f


def _uncrop(lib):
    img = load_png("ufo-input")
    h, w = img.shape[:2]
    frac = 20 / 100.0 + 1.0
    nw, nh = int(w * frac), int(h * frac)
    ox, oy = int((w * frac - w) / 2), int((h * frac - h) / 2)
    canvas = np.zeros((nh, nw, 3), np.uint8)
    canvas[oy:oy + h, ox:ox + w] = img
    tm = np.full((nh, nw), 255, np.uint8)
    tm[oy:oy + h, ox:ox + w] = 0
    shrink = int(max(w * 0.2, h * 0.2) / 2)
    cm = np.full((h, w), 255, np.uint8)
    cm[shrink:h - shrink, shrink:w - shrink] = 0
    p = abi.make_params(0, 0, 5, 0.0, 0.117, 16, 500)
    return _run(lib, p, 3, canvas, tm, img, cm)[0]


CASES = {
    # testResynth.py:221-242
    "resynthfull-brick": (True, _resynthfull("brick", SEL1)),
    "resynthfull-ufo-input": (True, _resynthfull("ufo-input", SEL1)),
    "resynthfull-zap-texture": (True, _resynthfull("zap-texture", SEL1)),
    "resynthfull-donkey_original": (True, _resynthfull("donkey_original", SEL2)),
    # :313-328
    "resynth-ufo-input": (True, _resynth_same),
    "resynthtwoimages-ufo-input": (True, _resynth_two(0, 0, 1)),
    "resynthtileable-ufo-input": (True, _resynth_two(1, 1, 0)),
    # :297-305
    "rendertexture-grass-input": (True, _rendertexture),
    "rendertexturealpha-grass-input-alpha": (True, _rendertexture_alpha),
    # :334-366
    "heal-ufo-input": (True, _heal("ufo-input", 3, 50, 1, 3)),
    "healgray-wander": (True, _heal("wander", 1, 50, 1, 3)),
    "healaroundrandom-ufo-input": (True, _heal("ufo-input", 3, 50, 0, 1)),
    "healalphagray-ufo-input-w-alpha-gray": (False, _heal("ufo-input-w-alpha-gray", 1, 50, 0, 2)),
    "healincludedalpha-ufo-input-w-alpha": (False, _heal("ufo-input-w-alpha", 3, 50, 1, 3)),
    # :250-287
    "mapstyle-ufo-input": (True, _mapstyle("ufo-input", "grass-input-alpha", 50, 3)),
    "mapstylealpha-ufo-input": (True, _mapstyle("ufo-input", "grass-input-alpha", 50, 3)),
    "mapstyleangel-angel_target": (True, _mapstyle("angel_target", "angel_texture", 10, 3)),
    "mapstylegraygray-wander": (True, _mapstyle("wander", "wander-texture", 50, 1)),
    # :391-393
    "uncrop-ufo-input": (True, _uncrop),
}


def check(lib, name):
    """Returns (n_differing_pixels, max_abs_diff, exact_expected)."""
    exact, fn = CASES[name]
    out = fn(lib)
    gold = load_golden(name)
    assert out.shape == gold.shape, (out.shape, gold.shape)
    d = np.abs(out.astype(np.int32) - gold.astype(np.int32))
    return int((d.max(axis=2) > 0).sum()), int(d.max()), exact


if __name__ == "__main__":
    import sys
    import time
    lib = R.load(sys.argv[1] if len(sys.argv) > 1 else "ref_mt_1t")
    for name in CASES:
        t0 = time.time()
        n, mx, exact = check(lib, name)
        print("%-45s diff_px=%6d max=%3d %s  %.1fs" % (name, n, mx, "EXACT" if exact else "+-1", time.time() - t0), flush=True)
