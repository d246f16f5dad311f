"""Command-line front end over resynthesizer_b200.plugin (needs Pillow for image files).

  python -m resynthesizer_b200.cli heal    --image in.png --mask sel.png --out out.png [--radius 50 --direction 0 --order 0]
  python -m resynthesizer_b200.cli uncrop  --image in.png --out out.png [--percent 10]
  python -m resynthesizer_b200.cli texture --image in.png --out out.png [--ratio 2 --tile]
  python -m resynthesizer_b200.cli style   --image target.png --source style.png --out out.png [--percent 50]
  python -m resynthesizer_b200.cli resynth --image t.png --mask tsel.png --corpus c.png [--corpus-mask csel.png]
                                           [--vtile 0 --htile 0 --context 1 --neighbours 30 --trys 200] --out out.png

Masks are 8-bit images of the same size; non-zero = selected.  Output format follows the file extension
(.ppm/.pgm gives binary PNM, like the goldens of Test/testResynth.py apart from their ASCII encoding)."""
import argparse

import numpy as np

from . import plugin


def _load(path):
    from PIL import Image
    a = np.asarray(Image.open(path))
    return a[:, :, None] if a.ndim == 2 else a


def _load_mask(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("L"))


def _save(path, pixels):
    from PIL import Image
    Image.fromarray(pixels[:, :, 0] if pixels.shape[2] == 1 else pixels).save(path)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="resynthesizer_b200.cli")
    ap.add_argument("command", choices=["heal", "uncrop", "texture", "style", "resynth"])
    ap.add_argument("--image", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--mask")
    ap.add_argument("--source")
    ap.add_argument("--corpus")
    ap.add_argument("--corpus-mask")
    ap.add_argument("--radius", type=int, default=50)
    ap.add_argument("--direction", type=int, default=0)
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--percent", type=float, default=None)
    ap.add_argument("--ratio", type=float, default=2)
    ap.add_argument("--tile", action="store_true")
    ap.add_argument("--vtile", type=int, default=0)
    ap.add_argument("--htile", type=int, default=0)
    ap.add_argument("--context", type=int, default=1)
    ap.add_argument("--map-weight", type=float, default=0.5)
    ap.add_argument("--autism", type=float, default=0.117)
    ap.add_argument("--neighbours", type=int, default=30)
    ap.add_argument("--trys", type=int, default=200)
    a = ap.parse_args(argv)
    img = plugin.Drawable(_load(a.image), _load_mask(a.mask) if a.mask else None)
    if a.command == "heal":
        plugin.heal_selection(img, a.radius, a.direction, a.order)
        out = img
    elif a.command == "uncrop":
        out = plugin.uncrop(img, 10 if a.percent is None else a.percent)
    elif a.command == "texture":
        out = plugin.render_texture(img, a.ratio, 1 if a.tile else 0)
    elif a.command == "style":
        plugin.map_style(img, plugin.Drawable(_load(a.source)), 50 if a.percent is None else a.percent, 0)
        out = img
    else:
        corpus = plugin.Drawable(_load(a.corpus or a.image), _load_mask(a.corpus_mask) if a.corpus_mask else None)
        plugin.plug_in_resynthesizer(img, a.vtile, a.htile, a.context, corpus, None, None, a.map_weight, a.autism,
                                     a.neighbours, a.trys)
        out = img
    _save(a.out, out.pixels)


if __name__ == "__main__":
    main()
