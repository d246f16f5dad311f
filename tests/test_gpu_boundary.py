"""-m gpu: behaviour at the drop-in boundary through the CUDA engine: the reference's known answers
(src/testSynth.c), progress-callback sequence, cancellation, row padding, imageSynth2, reentrancy."""
import ctypes as C
import threading

import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

pytestmark = pytest.mark.gpu


def test_testsynth_known_answers_on_gpu(built_lib):
    p = abi.default_params()
    img = np.zeros((3, 3, 4), np.uint8); img[0, 1, 3] = 1; img[1, 1] = [1, 1, 1, 1]; img[2, 2, 3] = 8
    mask = np.zeros((3, 3), np.uint8); mask[1, 1] = 0xFF
    err, percents = api.image_synth(img, mask, abi.T_RGBA, p, return_progress=True)
    assert err == 0 and list(img[1, 1]) == [0, 0, 0, 1]          # testSynth.c:165-179
    assert percents == [204750, 409500]                           # tick at index 0 of each of the two passes
    img = np.array([[[128, 128, 128, 255], [1, 1, 1, 1], [0, 0, 0, 0]]], np.uint8)
    m = np.array([[0, 255, 0]], np.uint8)
    assert api.image_synth(img, m, abi.T_RGBA, p) == 0            # testSynth.c:183
    assert img.reshape(-1).tolist() == [0x80, 0x80, 0x80, 0xFF, 0x80, 0x80, 0x80, 0x01, 0, 0, 0, 0]
    img = np.array([[[128] * 3, [1] * 3, [2] * 3], [[64] * 3, [4] * 3, [3] * 3]], np.uint8)
    m2 = np.array([[0, 0, 0], [0, 255, 0]], np.uint8)
    assert api.image_synth(img, m2, abi.T_RGB, p) == 0            # testSynth.c:186
    assert img.reshape(2, -1).tolist() == [[128] * 3 + [1] * 3 + [2] * 3, [64] * 3 + [1] * 3 + [3] * 3]
    img = np.array([[[128, 255], [64, 1], [1, 0]]], np.uint8)
    assert api.image_synth(img, m, abi.T_GrayA, p) == 0           # testSynth.c:189
    assert img.reshape(-1).tolist() == [0x80, 0xFF, 0x80, 0x01, 0x01, 0x00]
    img = np.array([[[128], [64], [1]]], np.uint8)
    assert api.image_synth(img, m, abi.T_Gray, None) == 0         # NULL parameters -> defaults (testSynth.c:193)
    assert img.reshape(-1).tolist()[0] == 0x80 and img.reshape(-1).tolist()[2] == 0x01
    assert img.reshape(-1).tolist()[1] in (0x80, 0x01)            # the two corpus pixels tie; the PRNG picks


def test_progress_sequence_equals_oracle(built_oracle, built_lib):
    img = G(200, 160, 3, 3)
    mask = centered_mask(200, 160, 120, 100)                      # 12000 targets: several ticks per pass
    port = R.load_port(R.GPU_MODE)
    pr = R.Progress()
    assert R.image_synth(port, img, mask, abi.T_RGB, None, pr)[0] == 0
    out = img.copy()
    err, percents = api.image_synth(out, mask, abi.T_RGB, None, return_progress=True)
    assert err == 0 and percents == pr.percents and len(percents) >= 10
    assert percents == sorted(percents)


def test_cancel_leaves_image_untouched(built_lib):
    img = G(256, 256, 3, 4)
    mask = centered_mask(256, 256, 128, 128)
    out = img.copy()
    err, percents = api.image_synth(out, mask, abi.T_RGB, None, cancel_after=2, return_progress=True)
    assert err == 0                                               # success even when cancelled (engine.c:689)
    assert (out == img).all()                                     # imageSynth skips the write-back (imageSynth.c:108)
    assert 2 <= len(percents) < 20
    # and the library is usable afterwards
    out2 = img.copy()
    assert api.image_synth(out2, mask, abi.T_RGB, None) == 0 and (out2 != img).any()


@pytest.mark.parametrize("cancel_after", [1, 2, 4, 6])
def test_cancelled_progress_sequence_equals_reference(built_oracle, built_lib, cancel_after):
    """After a cancel the reference's pass loop goes on (lib/refiner.h:75-121): the cancelled pass returns its betters so
    far; unless they are under the stop fraction the next pass starts, ticks once at index 0 (synthesize.h:493-497) and
    ends the loop with no betters.  Same callback sequence from the CUDA engine, the port and the compiled reference."""
    img = G(200, 160, 3, 3)
    mask = centered_mask(200, 160, 120, 100)                      # 12000 targets: ticks at 0, 4096, 8192 in pass 0
    seqs = []
    for lib in (R.load_port(R.GPU_MODE), R.load_port(R.REF_MODE)) + ((R.load("ref_mt_1t"),) if R.have("ref_mt_1t") else ()):
        pr = R.Progress(cancel_after=cancel_after)
        err, out = R.image_synth(lib, img, mask, abi.T_RGB, None, pr)
        assert err == 0 and (out == img).all()
        seqs.append(pr.percents)
    got = img.copy()
    err, percents = api.image_synth(got, mask, abi.T_RGB, None, cancel_after=cancel_after, return_progress=True)
    assert err == 0 and (got == img).all()
    for s_ in seqs:
        assert percents == s_, (percents, seqs)
    assert len(percents) in (cancel_after, cancel_after + 1)


def test_row_padding_and_imagesynth2(built_oracle, built_lib):
    L = api.lib()
    img = G(40, 30, 3, 8)
    mask = centered_mask(40, 30, 12, 10)
    mask2 = np.where(mask == 0, 255, 0).astype(np.uint8); mask2[:, :7] = 0
    port = R.load_port(R.GPU_MODE)
    err, want = R.image_synth(port, img, mask, abi.T_RGB, None, row_pad=5, mask2=mask2)
    assert err == 0
    rb = 40 * 3 + 5
    buf = np.full((30, rb), 0xEE, np.uint8); buf[:, :120] = img.reshape(30, 120)
    mbuf = np.zeros((30, 41), np.uint8); mbuf[:, :40] = mask
    ib, _a = abi.image_buffer_padded(buf.reshape(-1), 40, 30, rb)
    mb, _b = abi.image_buffer_padded(mbuf.reshape(-1), 40, 30, 41)
    m2 = np.ascontiguousarray(mask2)
    mb2, _c = abi.image_buffer_padded(m2.reshape(-1), 40, 30, 40)
    cancel = C.c_int(0)
    cb = abi.PROGRESS_CB(lambda p, c: None)
    assert L.imageSynth2(C.byref(ib), C.byref(mb), C.byref(mb2), abi.T_RGB, None, cb, None, C.byref(cancel)) == 0
    assert (buf[:, :120].reshape(30, 40, 3) == want).all() and (buf[:, 120:] == 0xEE).all()


def test_reentrant_from_threads(built_oracle, built_lib):
    """'The engine is reentrant, thread safe' (lib/imageSynth.c:13-15): concurrent jobs on separate images."""
    img = [G(96, 80, 3, 50 + i) for i in range(4)]
    mask = centered_mask(96, 80, 30, 24)
    port = R.load_port(R.GPU_MODE)
    want = [R.image_synth(port, im, mask, abi.T_RGB, None)[1] for im in img]
    outs = [im.copy() for im in img]
    errs = [None] * 4

    def work(i):
        errs[i] = api.image_synth(outs[i], mask, abi.T_RGB, None)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    assert errs == [0, 0, 0, 0]
    for o, w in zip(outs, want):
        assert (o == w).all()


def test_reference_testsynth_binary_runs_on_this_library(built_lib):
    """src/testSynth.c of the reference, compiled unchanged (oracle/build_ref.sh) and linked against
    libresynthesizer_b200.so: every 'Result:' it prints must equal its own 'Expect:' line, and the error cases must
    return the reference's codes (6, 3, 2, 5)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "testSynth_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/testSynth_b200 not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    assert "After\n00 00 00 00  00 00 00 01  00 00 00 00" in out
    assert "00 00 00 00  00 00 00 01  00 00 00 00  \n00 00 00 00  00 00 00 00  00 00 00 08" in out
    blocks = out.split("\n\n")
    errors = [int(l.split(":")[-1]) for l in out.splitlines() if "imageSynth returned error" in l]
    assert errors == [6, 3, 2, 5]
    norm = lambda s: " ".join(s.split())
    checked = 0
    for name in ("Test mix of full transparency and opaque", "Test RGB w/o alpha", "Test Gray w/ alpha"):
        seg = out[out.index(name):]
        expect = seg[seg.index("Expect:\n") + 8:seg.index("Result:\n")]
        result = seg[seg.index("Result:\n") + 8:].split("\n\n")[0]
        assert norm(expect) == norm(result), (name, expect, result)
        checked += 1
    assert checked == 3


def _pinned_like(a):
    """A copy of `a` in page-locked host memory (a numpy view of a pinned torch tensor)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


def test_page_locked_buffers_are_copied_directly(built_lib):
    """A caller's page-locked image / mask / pixmaps are read by the copy engine and written back in place, without the
    staging copy through the workspace (rs_cuda_host_is_pinned, rs_job_result_direct): same result as from malloc'ed
    buffers (the reference's ImageBuffer, lib/imageBuffer.h), padded rows included, and nothing is written on cancel."""
    L = api.lib()
    L.rs_cuda_host_is_pinned.argtypes = [C.c_void_p]
    img = G(200, 160, 3, 77)
    m = centered_mask(200, 160, 70, 60)
    p = abi.make_params(0, 0, 1, 0.5, 0.117, 16, 60)
    want = img.copy()
    assert api.image_synth(want, m, abi.T_RGB, p) == 0
    keep_i, pimg = _pinned_like(img)
    keep_m, pm = _pinned_like(m)
    assert L.rs_cuda_host_is_pinned(pimg.ctypes.data) == 1 and L.rs_cuda_host_is_pinned(img.ctypes.data) == 0
    assert api.image_synth(pimg, pm, abi.T_RGB, p) == 0
    assert (pimg == want).all() and (pimg != img).any()
    # rows wider than the image (ImageBuffer.rowBytes), page-locked: a strided copy in, a strided copy out
    wide = np.zeros((160, 640), np.uint8)
    keep_w, pw = _pinned_like(wide)
    pw[:, :600] = img.reshape(160, 600)
    pw[:, 600:] = 0xAB
    ib, _a = abi.image_buffer_padded(pw.reshape(-1), 200, 160, 640)
    mb, _b = abi.image_buffer_padded(pm.reshape(-1), 200, 160, 200)
    cancel = C.c_int(0)
    cb = abi.PROGRESS_CB(lambda pc, c: None)
    assert L.imageSynth(C.byref(ib), C.byref(mb), abi.T_RGB, C.byref(p), cb, None, C.byref(cancel)) == 0
    assert (pw[:, :600].reshape(160, 200, 3) == want).all() and (pw[:, 600:] == 0xAB).all()
    # engine(): page-locked pixmaps
    tp = R.build_pixmap(m, img)
    cp = R.build_pixmap(255 - m, img)
    fi = api.format_indices(3)
    want_t = tp.copy()
    assert api.engine(p, fi, want_t, cp) == 0
    keep_t, ptp = _pinned_like(tp)
    keep_c, pcp = _pinned_like(cp)
    assert api.engine(p, fi, ptp, pcp) == 0
    assert (ptp == want_t).all()
    # cancelled: the page-locked image stays as it was
    pimg[...] = img
    err = api.image_synth(pimg, pm, abi.T_RGB, p, cancel_after=1)
    assert (pimg == img).all()


def test_empty_target_counted_on_the_device(built_lib, monkeypatch):
    """With a device the empty-target scan of the mask is left to the selection digest (host_engine.cpp: count_on_device);
    error codes and their precedence stay the reference's (lib/engine.c:605-647): empty target before empty corpus before
    the context-type range; padding bytes of a mask with rowBytes > width do not count; an unchanged image on error."""
    L = api.lib()
    img = G(40, 24, 3, 5)
    zero = np.zeros((24, 40), np.uint8)
    p = abi.make_params(0, 0, 1, 0.5, 0.117, 16, 60)
    for force_host in (False, True):
        if force_host:
            monkeypatch.setenv("RS_HOST_TARGET_SCAN", "1")
        work = img.copy()
        assert api_err(lambda: api.image_synth(work, zero, abi.T_RGB, p)) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
        assert (work == img).all()
        # all selected: nothing left for the corpus
        assert api_err(lambda: api.image_synth(work, np.full((24, 40), 255, np.uint8), abi.T_RGB, p)) == abi.IMAGE_SYNTH_ERROR_EMPTY_CORPUS
        # empty target AND a bad context type: the target error comes first; a non-empty one reports the range
        bad = abi.make_params(0, 0, 9, 0.5, 0.117, 16, 60)
        assert api_err(lambda: api.image_synth(work, zero, abi.T_RGB, bad)) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
        one = zero.copy(); one[3, 7] = 1
        assert api_err(lambda: api.image_synth(work, one, abi.T_RGB, bad)) == abi.IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE
        # full API: empty target pixmap, and empty target + fully transparent corpus (both empty: target first)
        fi = api.format_indices(3)
        tp = R.build_pixmap(zero, img); cp = R.build_pixmap(255 - zero, img)
        assert api_err(lambda: api.engine(p, fi, tp, cp)) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
        cp0 = R.build_pixmap(zero, img)
        assert api_err(lambda: api.engine(p, fi, tp, cp0)) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
        # a padded mask whose padding is non-zero is still empty
        wide = np.zeros((24, 64), np.uint8); wide[:, 40:] = 0xFF
        ib, _a = abi.image_buffer_padded(work.reshape(-1), 40, 24, 120)
        mb, _b = abi.image_buffer_padded(wide.reshape(-1), 40, 24, 64)
        cancel = C.c_int(0)
        cb = abi.PROGRESS_CB(lambda pc, c: None)
        assert L.imageSynth(C.byref(ib), C.byref(mb), abi.T_RGB, C.byref(p), cb, None, C.byref(cancel)) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
        # and a job after the errors runs as before
        m = centered_mask(40, 24, 10, 8)
        assert api.image_synth(work, m, abi.T_RGB, p) == 0 and (work != img).any()


def api_err(call):
    """The error code of a call through resynthesizer_b200.api (which raises only for the CUDA layer's code 100)."""
    return call()


@pytest.mark.parametrize("w,h", [(200, 160), (97, 83)])
def test_page_locked_read_back_is_the_target_box(built_lib, w, h):
    """Page-locked caller buffers get back the BOX that holds target points (columns from the selection digest, exact per
    32-pixel word, the whole width where a word straddles rows): selections off centre, at the image edges, in two pieces,
    widths that are no multiple of 32 -- same images as from malloc'ed buffers, through imageSynth() and engine()."""
    img = G(w, h, 3, 91)
    p = abi.make_params(0, 0, 1, 0.5, 0.117, 12, 40)
    masks = []
    m = np.zeros((h, w), np.uint8); m[5:20, w - 14:w] = 255; masks.append(m)                 # right edge, top
    m = np.zeros((h, w), np.uint8); m[h - 9:h, 0:11] = 255; masks.append(m)                  # left edge, bottom
    m = np.zeros((h, w), np.uint8); m[10:22, 8:20] = 255; m[h - 30:h - 20, w - 40:w - 31] = 255; masks.append(m)   # two pieces
    m = np.zeros((h, w), np.uint8); m[h // 2, 31:34] = 255; masks.append(m)                  # three pixels across a word boundary
    fi = api.format_indices(3)
    for m in masks:
        want = img.copy()
        assert api.image_synth(want, m, abi.T_RGB, p) == 0
        keep_i, pimg = _pinned_like(img)
        keep_m, pm = _pinned_like(m)
        assert api.image_synth(pimg, pm, abi.T_RGB, p) == 0
        assert (pimg == want).all() and (pimg != img).any()
        tp = R.build_pixmap(m, img); cp = R.build_pixmap(255 - m, img)
        want_t = tp.copy()
        assert api.engine(p, fi, want_t, cp) == 0
        keep_t, ptp = _pinned_like(tp)
        keep_c, pcp = _pinned_like(cp)
        assert api.engine(p, fi, ptp, pcp) == 0
        assert (ptp == want_t).all()
