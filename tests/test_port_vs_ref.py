"""not-gpu: the C restatement (oracle/_port, reference mode) against
 (a) the committed outputs of the compiled reference (tests/golden/ref_vectors.npz) -- runs anywhere;
 (b) the compiled reference itself (oracle/_ref), live, where it has been built -- both PRNG variants;
 (c) the known-answer cases of the reference's own test harness src/testSynth.c:87-224."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import abi
from tests.golden import cases

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
HAVE_REF = os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libref_mt_1t.so"))


@pytest.mark.parametrize("name", sorted(cases.all_cases()))
def test_port_equals_committed_reference_output(built_oracle, name):
    port = R.load_port(R.REF_MODE)
    out = cases.run_case(port, cases.all_cases()[name])
    assert out.shape == VEC[name].shape and (out == VEC[name]).all()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["simple_rgb_ctx2_tile0", "simple_rgb_ctx8_tile1", "simple_graya", "engine_maps_rgb"])
def test_port_equals_live_reference(built_oracle, name):
    ref = R.load("ref_mt_1t")
    port = R.load_port(R.REF_MODE)
    case = cases.all_cases()[name]
    assert (cases.run_case(ref, case) == cases.run_case(port, case)).all()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_port_rand_mode_equals_standalone_reference(built_oracle):
    """The reference's standalone build draws from libc rand() (glibProxy.c:36-49)."""
    ref = R.load("ref_rand_1t")
    port = R.load_port(R.RAND_MODE)
    case = cases.all_cases()["simple_rgb_ctx5_tile0"]
    assert (cases.run_case(ref, case) == cases.run_case(port, case)).all()


# ---- src/testSynth.c known answers, restated as data -------------------------------------------------
def _kat_libs():
    # testSynth is a standalone (glibProxy, libc rand()) program: its "Expect" strings belong to that PRNG
    libs = [("port", lambda: R.load_port(R.RAND_MODE))]
    if HAVE_REF:
        libs.append(("ref", lambda: R.load("ref_rand_1t")))
    return libs


@pytest.mark.parametrize("which", [n for n, _ in _kat_libs()])
def test_testsynth_known_answers(built_oracle, which):
    lib = dict(_kat_libs())[which]()
    p = abi.default_params()
    # 3x3 RGBA, centre selected: colour from the opaque-ish surround, alpha byte unchanged (testSynth.c:165-179)
    img = np.zeros((3, 3, 4), np.uint8); img[0, 1, 3] = 1; img[1, 1] = [1, 1, 1, 1]; img[2, 2, 3] = 8
    mask = np.zeros((3, 3), np.uint8); mask[1, 1] = 0xFF
    pr = R.Progress()
    err, out = R.image_synth(lib, img, mask, abi.T_RGBA, p, pr, row_pad=2)
    assert err == 0 and list(out[1, 1]) == [0, 0, 0, 1]
    assert pr.percents == [204750, 409500]          # tick at index 0 of each of two passes
    # 1x3 RGBA: opaque, target, transparent (testSynth.c:183)
    img = np.array([[[128, 128, 128, 255], [1, 1, 1, 1], [0, 0, 0, 0]]], np.uint8)
    mask = np.array([[0, 255, 0]], np.uint8)
    err, out = R.image_synth(lib, img, mask, abi.T_RGBA, p, row_pad=2)
    assert err == 0 and out.reshape(-1).tolist() == [0x80, 0x80, 0x80, 0xFF, 0x80, 0x80, 0x80, 0x01, 0, 0, 0, 0]
    # 2x3 RGB (testSynth.c:186)
    img = np.array([[[128] * 3, [1] * 3, [2] * 3], [[64] * 3, [4] * 3, [3] * 3]], np.uint8)
    mask = np.array([[0, 0, 0], [0, 255, 0]], np.uint8)
    err, out = R.image_synth(lib, img, mask, abi.T_RGB, p, row_pad=2)
    assert err == 0 and out.reshape(2, -1).tolist() == [[128] * 3 + [1] * 3 + [2] * 3, [64] * 3 + [1] * 3 + [3] * 3]
    # 1x3 GrayA (testSynth.c:189)
    img = np.array([[[128, 255], [64, 1], [1, 0]]], np.uint8)
    mask = np.array([[0, 255, 0]], np.uint8)
    err, out = R.image_synth(lib, img, mask, abi.T_GrayA, p, row_pad=2)
    assert err == 0 and out.reshape(-1).tolist() == [0x80, 0xFF, 0x80, 0x01, 0x01, 0x00]
    # 1x3 Gray with NULL parameters (testSynth.c:193)
    img = np.array([[[128], [64], [1]]], np.uint8)
    err, out = R.image_synth(lib, img, mask, abi.T_Gray, None, row_pad=2)
    assert err == 0 and out.reshape(-1).tolist() == [0x80, 0x01, 0x01]


@pytest.mark.parametrize("which", [n for n, _ in _kat_libs()])
def test_testsynth_error_cases(built_oracle, which):
    lib = dict(_kat_libs())[which]()
    p = abi.default_params()
    mask = np.array([[0, 255, 0]], np.uint8)
    # all transparent -> empty corpus (testSynth.c:200)
    img = np.array([[[128, 128, 128, 0], [1, 1, 1, 1], [0, 0, 0, 0]]], np.uint8)
    img[0, 1, 3] = 0
    before = img.copy()
    err, out = R.image_synth(lib, img, mask, abi.T_RGBA, p)
    assert err == abi.IMAGE_SYNTH_ERROR_EMPTY_CORPUS and (out == before).all()
    gray = np.array([[[128], [64], [1]]], np.uint8)
    p.patchSize = 65   # testSynth.c:203
    assert R.image_synth(lib, gray, mask, abi.T_Gray, p)[0] == abi.IMAGE_SYNTH_ERROR_PATCH_SIZE_EXCEEDED
    p.patchSize = 10
    assert R.image_synth(lib, gray, np.zeros((1, 1), np.uint8), abi.T_Gray, p)[0] == abi.IMAGE_SYNTH_ERROR_IMAGE_MASK_MISMATCH
    assert R.image_synth(lib, gray, np.zeros((1, 3), np.uint8), abi.T_Gray, p)[0] == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
    p.matchContextType = 9
    assert R.image_synth(lib, gray, mask, abi.T_Gray, p)[0] == abi.IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE
