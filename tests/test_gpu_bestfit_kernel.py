"""-m gpu: the best-fit kernel alone, on candidate lists and patches dumped by the oracle running the
REFERENCE semantics (MT19937 stream, live recentProber): (bestPatchDiff, winning candidate) must be bit-exact
for every visit (north star: "bit-exact against the reference's integer computeBestFit on identical candidate
lists")."""
import ctypes as C

import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def _pack_xy(a):
    a = np.asarray(a, np.int64)
    return ((a[:, 0] & 0xFFFF) | ((a[:, 1] & 0xFFFF) << 16)).astype(np.uint32)


def run_bestfit_on_traces(traces, fi, corpus_pixmap, sens, map_weight):
    L = api.lib()
    c512 = np.zeros(512, np.uint16); m512 = np.zeros(512, np.uint32)
    L.rs_host_metric_tables(sens, map_weight, c512.ctypes.data, m512.ctypes.data)
    c256 = c512[256:].astype(np.uint32); m256 = m512[256:].copy()
    ch, cw, bpp = corpus_pixmap.shape
    d = api.RsJobDesc()
    d.tw = d.th = 1; d.cw = cw; d.ch = ch; d.bpp = bpp
    d.n_color = fi.img_match_bpp; d.n_map = fi.map_match_bpp; d.map_bip = fi.map_start_bip
    nb_begin = np.zeros(len(traces) + 1, np.uint32); cand_begin = np.zeros(len(traces) + 1, np.uint32)
    offs, pix, cands = [], [], []
    for i, t in enumerate(traces):
        nb_begin[i + 1] = nb_begin[i] + t["K"]
        cand_begin[i + 1] = cand_begin[i] + len(t["cands"])
        offs.append(_pack_xy(t["offsets"])); pix.append(t["pixels"]); cands.append(_pack_xy(t["cands"]) if len(t["cands"]) else np.zeros(0, np.uint32))
    offs = np.ascontiguousarray(np.concatenate(offs)); pix = np.ascontiguousarray(np.concatenate(pix))
    cands = np.ascontiguousarray(np.concatenate(cands)) if cands else np.zeros(1, np.uint32)
    best = np.zeros(len(traces), np.uint32); bidx = np.zeros(len(traces), np.int32)
    corpus = np.ascontiguousarray(corpus_pixmap)
    rc = L.rs_bestfit_batch(C.byref(d), corpus.ctypes.data, c256.ctypes.data, m256.ctypes.data, int(m512[0]),
                            len(traces), nb_begin.ctypes.data, offs.ctypes.data, pix.ctypes.data,
                            cand_begin.ctypes.data, cands.ctypes.data, best.ctypes.data, bidx.ctypes.data)
    assert rc == 0, L.rs_cuda_last_error()
    return best, bidx


def _check(traces, best, bidx):
    bad = 0
    for t, b, i in zip(traces, best, bidx):
        if not t["bettered"]:
            assert i == -1
            continue
        ok = int(b) == t["best"] and tuple(t["cands"][i]) == t["best_xy"]
        bad += (not ok)
    assert bad == 0, "%d of %d visits differ" % (bad, len(traces))


@pytest.mark.parametrize("n_map,alpha", [(0, False), (3, False), (1, True)])
def test_bestfit_vs_reference_traces(built_oracle, built_lib, n_map, alpha):
    port = R.load_port(R.REF_MODE)
    tw, th, cw, ch = 64, 56, 48, 40
    tgt, cor = G(tw, th, 3, 41), G(cw, ch, 3, 42)
    tmask = centered_mask(tw, th, 24, 20)
    cmask = np.full((ch, cw), 255, np.uint8); cmask[5:9, 7:12] = 0
    ta = ca = None
    if alpha:
        ta = np.full((th, tw), 255, np.uint8); ca = np.full((ch, cw), 255, np.uint8); ca[::5, ::3] = 0
    tmaps = G(tw, th, n_map, 43) if n_map else None
    cmaps = G(cw, ch, n_map, 44) if n_map else None
    fi = R.format_indices(port, 3, n_map, alpha, alpha, n_map > 0)
    tp = R.build_pixmap(tmask, tgt, ta, tmaps); cp = R.build_pixmap(cmask, cor, ca, cmaps)
    params = abi.make_params(0, 0, 1, 0.5, 0.117, 25, 120)
    port.port_trace_enable(1, 100000)
    assert R.engine(port, params, fi, tp, cp.copy()) == 0
    traces = R.port_trace(port)
    port.port_trace_enable(0, 0)
    assert len(traces) > 1000
    best, bidx = run_bestfit_on_traces(traces, fi, cp, 0.117, 0.5)
    _check(traces, best, bidx)
