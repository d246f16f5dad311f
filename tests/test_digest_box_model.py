"""CPU model of the column range k_target_digest returns (csrc/rs_kernels.cu): per 32-pixel word of the linear pixel
index, exact columns where the word's selected pixels lie in one row, the whole width where they straddle rows.  The box
a page-locked result is copied back through (rs_job_download*, x_min..x_max) must contain every target point, and be the
exact bounding box whenever no word straddles rows."""
import numpy as np
import pytest


def model_box(mask):
    h, w = mask.shape
    sel = np.flatnonzero(mask.reshape(-1) != 0)
    xmin, xmax, ymin, ymax = 2 ** 32 - 1, 0, 2 ** 32 - 1, 0
    for word in np.unique(sel // 32):
        bits = sel[sel // 32 == word]
        first, last = int(bits[0]), int(bits[-1])
        y0, y1 = first // w, last // w
        ymin, ymax = min(ymin, y0), max(ymax, y1)
        xmin = min(xmin, first - y0 * w if y0 == y1 else 0)
        xmax = max(xmax, last - y1 * w if y0 == y1 else w - 1)
    return xmin, xmax, ymin, ymax


@pytest.mark.parametrize("w,h", [(64, 40), (97, 83), (33, 9), (200, 160), (31, 31)])
def test_box_contains_every_target_point(w, h):
    rng = np.random.RandomState(w * 1000 + h)
    for trial in range(40):
        m = np.zeros((h, w), np.uint8)
        for _ in range(rng.randint(1, 4)):
            y0, x0 = rng.randint(0, h), rng.randint(0, w)
            m[y0:y0 + rng.randint(1, 12), x0:x0 + rng.randint(1, 12)] = 255
        ys, xs = np.nonzero(m)
        xmin, xmax, ymin, ymax = model_box(m)
        assert xmin <= xs.min() and xs.max() <= xmax < w
        assert (ymin, ymax) == (ys.min(), ys.max())
        if w % 32 == 0:
            assert (xmin, xmax) == (xs.min(), xs.max())
