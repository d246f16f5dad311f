"""GIMP-free mirror of the callers above the library boundary: the engine plug-in's PDB procedure
`plug_in_resynthesizer` (src/resynthesizer/resynthesizer.c:286-548, argument order of
src/resynth-parameters.h:46-73) and the user-level PluginScripts that call it -- heal selection, heal transparency,
uncrop, render texture, map style -- over numpy arrays instead of GIMP drawables.

A `Drawable` is the pixels of a layer (h, w, c) with c = 1 (gray), 2 (gray+alpha), 3 (RGB) or 4 (RGBA), plus an
optional selection mask (h, w) in layer coordinates; `selection=None` means "no selection intersects this
drawable", for which the plug-in uses the whole drawable (src/resynthesizer/adaptGimp.h:177-256).

The synthesis itself is `engine()` of libresynthesizer_b200.so (resynthesizer_b200.api.engine); `engine_fn` can be
replaced by any callable with the same signature, which is how the tests run these recipes on the oracle.
"""
import math

import numpy as np

from . import abi

MAX_NEIGHBORS = 64  # IMAGE_SYNTH_MAX_NEIGHBORS (lib/imageSynthConstants.h:26)


class PluginError(RuntimeError):
    """The engine plug-in's ERROR_RETURN messages (src/resynthesizer/resynthesizer.c)."""


class Drawable:
    def __init__(self, pixels, selection=None):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        if pixels.ndim == 2:
            pixels = pixels[:, :, None]
        if pixels.ndim != 3 or pixels.shape[2] not in (1, 2, 3, 4):
            raise PluginError("Incompatible image mode.")
        self.pixels = pixels
        self.selection = None if selection is None else np.ascontiguousarray(selection, dtype=np.uint8)
        if self.selection is not None and self.selection.shape != pixels.shape[:2]:
            raise ValueError("selection must have the drawable's height and width")

    @property
    def height(self):
        return self.pixels.shape[0]

    @property
    def width(self):
        return self.pixels.shape[1]

    @property
    def has_alpha(self):
        return self.pixels.shape[2] in (2, 4)

    @property
    def color_channels(self):
        return self.pixels.shape[2] - (1 if self.has_alpha else 0)

    def copy(self):
        return Drawable(self.pixels.copy(), None if self.selection is None else self.selection.copy())


def _default_engine():
    from . import api
    return api.engine, api.format_indices


def _mask_of(drawable):
    """fetch_mask (adaptGimp.h:177-256): the selection where it intersects the drawable, else everything selected."""
    sel = drawable.selection
    if sel is None or not sel.any():
        return np.full((drawable.height, drawable.width), 0xFF, np.uint8)
    return sel.copy()


def _pixmap(drawable, fi, map_drawable):
    """fetch_image_mask_map (adaptGimp.h:294-322): [mask][colour+alpha of the drawable][map colours, alpha dropped]."""
    h, w = drawable.height, drawable.width
    pm = np.zeros((h, w, fi.total_bpp), np.uint8)
    pm[:, :, 0] = _mask_of(drawable)
    pm[:, :, 1:1 + drawable.pixels.shape[2]] = drawable.pixels
    if map_drawable is not None:
        n = map_drawable.color_channels
        pm[:, :, fi.map_start_bip:fi.map_start_bip + n] = map_drawable.pixels[:, :, :n]
    return pm


def plug_in_resynthesizer(drawable, vtile, htile, use_context, corpus, inmask=None, outmask=None, map_weight=0.5,
                          autism=0.117, neighbourhood=30, trys=200, engine_fn=None, format_indices_fn=None,
                          progress=None):
    """The PDB procedure `plug-in-resynthesizer` minus (run_mode, image): synthesises the selected part of
    `drawable` in place from `corpus`.  inmask / outmask are the corpus map and the target map (both or neither).
    Raises PluginError with the plug-in's messages; returns the engine's error code (0)."""
    if engine_fn is None:
        engine_fn, format_indices_fn = _default_engine()
    neighbourhood = min(int(neighbourhood), MAX_NEIGHBORS)            # resynthesizer.c:377-378
    if drawable.color_channels != corpus.color_channels:              # :386-390
        raise PluginError("The input texture and output image must have the same number of color channels.")
    with_map = inmask is not None and outmask is not None             # :392, a single map is ignored quietly
    if with_map:
        if inmask.color_channels != outmask.color_channels:
            raise PluginError("The input and output maps must have the same mode")
        if (inmask.width, inmask.height) != (corpus.width, corpus.height):
            raise PluginError("The input map should be the same size as the input texture image")
        if (outmask.width, outmask.height) != (drawable.width, drawable.height):
            raise PluginError("The output map should be the same size as the output image")
    fi = format_indices_fn(drawable.color_channels, inmask.color_channels if with_map else 0,
                           drawable.has_alpha, corpus.has_alpha, with_map)      # :451-458
    # the internal pixel has ONE alpha slot if either image has alpha (lib/imageFormat.c:170-177); it stays 0 for
    # the image without alpha, which the engine never reads (isAlphaTarget / isAlphaSource)
    tp = _pixmap(drawable, fi, outmask if with_map else None)
    cp = _pixmap(corpus, fi, inmask if with_map else None)
    params = abi.make_params(htile, vtile, use_context, map_weight, autism, neighbourhood, trys)
    kwargs = {"progress": progress} if progress is not None else {}
    err = engine_fn(params, fi, tp, cp, **kwargs)
    if err == abi.IMAGE_SYNTH_ERROR_EMPTY_CORPUS:                     # :515-522
        raise PluginError("The texture source is empty. Does any selection include non-transparent pixels?")
    if err == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET:
        raise PluginError("The output layer is empty. Does any selection have visible pixels in the active layer?")
    if err == 0:                                                      # post_results_to_gimp (:243-253)
        drawable.pixels[:, :, :] = tp[:, :, 1:1 + drawable.pixels.shape[2]]
    return err


# ------------------------------------------------------------------------------------------ GIMP operations
def gimp_selection_grow(mask, r):
    """GIMP's selection grow by r pixels (circular structuring element; reproduces the reference goldens,
    SURVEY.md App. B): (x,y) is selected iff some selected (x',y') has |y-y'| <= c(|x-x'|), c(0)=r,
    c(d)=rint(sqrt(r^2-(d-0.5)^2)) for 1<=d<=r."""
    h, w = mask.shape
    sel = mask > 0
    out = np.zeros_like(sel)
    for d in range(0, r + 1):
        c = r if d == 0 else int(np.rint(math.sqrt(r * r - (d - 0.5) ** 2)))
        csum = np.cumsum(np.pad(sel, ((c + 1, c), (0, 0))).astype(np.int32), axis=0)
        col = ((csum[2 * c + 1:] - csum[:-(2 * c + 1)]) > 0)[:h]
        if d == 0:
            out |= col
        else:
            out[:, d:] |= col[:, :-d]
            out[:, :-d] |= col[:, d:]
    return out.astype(np.uint8) * 255


def _bounds(mask):
    ys, xs = np.nonzero(mask)
    return int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1


# ------------------------------------------------------------------------------------------ PluginScripts
def heal_selection(drawable, sampling_radius=50, direction=0, order=0, **kw):
    """PluginScripts/plugin-heal-selection.py:36-150.  direction 0 all around / 1 sides / 2 above and below;
    order 0 random / 1 inwards / 2 outwards.  The drawable's selection is the region to heal."""
    if drawable.selection is None or not drawable.selection.any():
        raise PluginError("You must first select a region to heal.")
    sel = drawable.selection
    grown = gimp_selection_grow(sel, int(sampling_radius))
    frisket = np.where((grown > 0) & (sel == 0), 255, 0).astype(np.uint8)   # grown minus the original selection
    fx0, fy0, fx1, fy1 = _bounds(grown)
    tx0, ty0, tx1, ty1 = _bounds(sel)
    if direction == 0:
        x0, y0, cw, ch = fx0, fy0, fx1 - fx0, fy1 - fy0
    elif direction == 1:
        x0, y0, cw, ch = fx0, ty0, fx1 - fx0, ty1 - ty0
    else:
        x0, y0, cw, ch = tx0, fy0, tx1 - tx0, fy1 - fy0
    cw = min(drawable.width - x0, cw)
    ch = min(drawable.height - y0, ch)
    corpus = Drawable(drawable.pixels[y0:y0 + ch, x0:x0 + cw].copy(), frisket[y0:y0 + ch, x0:x0 + cw].copy())
    use_border = 1 if not order else (direction + 2 if order == 1 else direction + 5)   # :121-130
    return plug_in_resynthesizer(drawable, 0, 0, use_border, corpus, None, None, 0.0, 0.117, 16, 500, **kw)  # :148


def heal_transparency(drawable, sampling_radius=50, order=2, **kw):
    """PluginScripts/plugin-heal-transparency.py:36-66: select the fully transparent pixels, grow by 1, heal."""
    if not drawable.has_alpha:
        raise PluginError("The active layer has no alpha channel to heal.")
    transparent = np.where(drawable.pixels[:, :, -1] == 0, 255, 0).astype(np.uint8)
    if not transparent.any():
        raise PluginError("There are no transparent pixels to heal.")
    work = Drawable(drawable.pixels, gimp_selection_grow(transparent, 1))
    err = heal_selection(work, sampling_radius, 0, order, **kw)
    return err


def uncrop(drawable, percent_enlarge=10, **kw):
    """PluginScripts/plugin-uncrop.py:78-141: enlarge the canvas by percent, synthesise the new outer band from a
    band of equal width at the edge of the original, outwards (use_border 5).  Returns the new Drawable."""
    h, w = drawable.height, drawable.width
    frac = percent_enlarge / 100.0 + 1.0
    nw, nh = int(w * frac), int(h * frac)
    ox, oy = int((w * frac - w) / 2), int((h * frac - h) / 2)
    canvas = np.zeros((nh, nw, drawable.pixels.shape[2]), np.uint8)
    canvas[oy:oy + h, ox:ox + w] = drawable.pixels
    tsel = np.full((nh, nw), 255, np.uint8)
    tsel[oy:oy + h, ox:ox + w] = 0
    shrink = int(max(w * (percent_enlarge / 100.0), h * (percent_enlarge / 100.0)) / 2)
    csel = np.full((h, w), 255, np.uint8)
    csel[shrink:h - shrink, shrink:w - shrink] = 0
    target = Drawable(canvas, tsel)
    plug_in_resynthesizer(target, 0, 0, 5, Drawable(drawable.pixels.copy(), csel), None, None, 0.0, 0.117, 16, 500, **kw)
    return target


def render_texture(drawable, resize_ratio=2, make_tile=0, **kw):
    """PluginScripts/plugin-render-texture.py:72-190: a new image resize_ratio times the size of the (selected part
    of the) drawable, synthesised from it without context; make_tile => seamlessly tileable.  Returns it."""
    src = drawable
    if drawable.selection is not None and drawable.selection.any():
        x0, y0, x1, y1 = _bounds(drawable.selection)
        src = Drawable(drawable.pixels[y0:y1, x0:x1].copy(), drawable.selection[y0:y1, x0:x1].copy())
    nh, nw = int(src.height * resize_ratio), int(src.width * resize_ratio)
    out = Drawable(np.full((nh, nw, src.pixels.shape[2]), 255, np.uint8))
    tile = 1 if make_tile else 0
    plug_in_resynthesizer(out, tile, tile, 0, src, None, None, 0.0, 0.117, 9, 200, **kw)   # :175
    return out


def calculate_map_weight(percent_transfer):
    return math.acos((percent_transfer / 100.0) * 2 - 1) / (2 * 3.14)      # plugin-map-style.py:245-256 (3.14 sic)


def map_style(drawable, source, percent_transfer=50, map_mode=0, **kw):
    """PluginScripts/plugin-map-style.py:258-370, map_mode 0 (colour and brightness: the maps are the images
    themselves).  map_mode 1 (brightness only) needs GIMP's grayscale/contrast operations and is not mirrored."""
    if map_mode != 0:
        raise NotImplementedError("map_mode 1 needs GIMP's desaturate and brightness-contrast")
    tgt_map = Drawable(drawable.pixels[:, :, :drawable.color_channels].copy())
    src_map = Drawable(source.pixels[:, :, :source.color_channels].copy())
    return plug_in_resynthesizer(drawable, 1, 1, 1, source, src_map, tgt_map, calculate_map_weight(percent_transfer),
                                 0.117, 9, 200, **kw)                                   # :359
