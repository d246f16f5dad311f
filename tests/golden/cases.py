"""Small synthetic cases shared by the golden-vector generator (run on the compiled reference) and the tests
(run on the restatement / the CUDA engine).  Inputs are regenerated from seeds, only outputs are stored."""
import numpy as np

from oracle import refdriver as R
from resynthesizer_b200 import abi
from resynthesizer_b200.synthetic import G, centered_mask


def all_cases():
    c = {}
    for ctx in range(9):
        for tile in (0, 1):
            c["simple_rgb_ctx%d_tile%d" % (ctx, tile)] = dict(kind="simple", fmt=abi.T_RGB, w=56, h=44, seed=ctx + 1,
                                                              params=(tile, tile, ctx, 0.5, 0.117, 14, 40))
    for fmt, name in ((abi.T_RGBA, "rgba"), (abi.T_Gray, "gray"), (abi.T_GrayA, "graya")):
        c["simple_%s" % name] = dict(kind="simple", fmt=fmt, w=48, h=40, seed=20, params=(0, 0, 1, 0.5, 0.117, 30, 60))
    c["simple2_rgb"] = dict(kind="simple2", fmt=abi.T_RGB, w=48, h=40, seed=21, params=(0, 0, 1, 0.5, 0.117, 16, 50))
    c["engine_maps_rgb"] = dict(kind="engine", n_color=3, n_map=3, alpha=False, tw=40, th=36, cw=32, ch=30, seed=30,
                                params=(1, 1, 1, 0.5, 0.117, 9, 40), full=True)
    c["engine_maps_gray_alpha"] = dict(kind="engine", n_color=3, n_map=1, alpha=True, tw=40, th=36, cw=32, ch=30,
                                       seed=31, params=(0, 0, 2, 0.25, 0.117, 12, 40), full=False)
    c["engine_render_texture"] = dict(kind="engine", n_color=3, n_map=0, alpha=False, tw=48, th=40, cw=24, ch=24,
                                      seed=32, params=(1, 1, 0, 0.0, 0.117, 9, 40), full=True)
    c["engine_gray_graymap"] = dict(kind="engine", n_color=1, n_map=1, alpha=False, tw=36, th=36, cw=28, ch=28,
                                    seed=33, params=(1, 1, 1, 0.4, 0.117, 9, 30), full=True)
    return c


def simple_inputs(case):
    nch = abi.FORMAT_CHANNELS[case["fmt"]]
    img = G(case["w"], case["h"], nch, case["seed"])
    if nch in (2, 4):
        img[:, 5:9, nch - 1] = 0
        img[:, 9:, nch - 1] = 255
        img[3:6, 20:24, nch - 1] = 77
    mask = centered_mask(case["w"], case["h"], case["w"] // 3, case["h"] // 3)
    mask[4:7, 30:34] = 99
    return img, mask


def engine_inputs(case):
    tw, th, cw, ch = case["tw"], case["th"], case["cw"], case["ch"]
    tgt, cor = G(tw, th, case["n_color"], case["seed"]), G(cw, ch, case["n_color"], case["seed"] + 100)
    tmask = np.full((th, tw), 255, np.uint8) if case["full"] else centered_mask(tw, th, tw // 3, th // 3)
    cmask = np.full((ch, cw), 255, np.uint8)
    cmask[:2, :4] = 0
    ta = ca = None
    if case["alpha"]:
        ta = np.full((th, tw), 255, np.uint8); ta[::7, ::5] = 0
        ca = np.full((ch, cw), 255, np.uint8); ca[::6, ::4] = 0
    tm = G(tw, th, case["n_map"], case["seed"] + 200) if case["n_map"] else None
    cm = G(cw, ch, case["n_map"], case["seed"] + 300) if case["n_map"] else None
    return R.build_pixmap(tmask, tgt, ta, tm), R.build_pixmap(cmask, cor, ca, cm)


def run_case(lib, case):
    """Runs a case on a reference-ABI library (via oracle.refdriver); returns the output array."""
    p = abi.make_params(*case["params"])
    if case["kind"] in ("simple", "simple2"):
        img, mask = simple_inputs(case)
        mask2 = None
        if case["kind"] == "simple2":
            mask2 = np.where(mask == 0, 255, 0).astype(np.uint8)
            mask2[:, :6] = 0
        err, out = R.image_synth(lib, img, mask, case["fmt"], p, row_pad=3, mask2=mask2)
        assert err == 0
        return out
    tp, cp = engine_inputs(case)
    fi = R.format_indices(lib, case["n_color"], case["n_map"], case["alpha"], case["alpha"], case["n_map"] > 0)
    assert R.engine(lib, p, fi, tp, cp) == 0
    return tp
