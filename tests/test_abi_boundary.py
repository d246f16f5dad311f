"""not-gpu: the C-ABI library loads, exports every symbol include/*.h declares, keeps the reference's struct
layouts, returns the reference's parameter/data errors without touching CUDA, and fails LOUDLY (no CPU fallback)
when synthesis is requested without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from resynthesizer_b200 import abi, api
from resynthesizer_b200.synthetic import G, centered_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"^[A-Za-z_][A-Za-z0-9_ \*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src, flags=re.M):
        if "typedef" not in m.group(0) and m.group(1) not in ("defined",):
            names.add(m.group(1))
    return names


@pytest.mark.parametrize("header", ["resynthesizer.h", "rs_cuda.h", "rs_host.h"])
def test_every_declared_symbol_is_exported(built_lib, header):
    L = C.CDLL(api.LIB_PATH)
    names = _declared_functions(header)
    assert len(names) >= 4
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match_reference():
    # x86-64 sizes of the reference's types
    assert C.sizeof(abi.ImageBuffer) == 24 and C.sizeof(abi.TImageSynthParameters) == 40
    assert C.sizeof(abi.TFormatIndices) == 16 and C.sizeof(abi.Map) == 24
    assert abi.TImageSynthParameters.mapWeight.offset == 16 and abi.TImageSynthParameters.patchSize.offset == 32
    assert abi.TFormatIndices.isAlphaTarget.offset == 8


def test_default_params(built_lib):
    p = abi.TImageSynthParameters()
    api.lib().setDefaultParams(C.byref(p))
    assert (p.isMakeSeamlesslyTileableHorizontally, p.isMakeSeamlesslyTileableVertically, p.matchContextType,
            p.mapWeight, p.sensitivityToOutliers, p.patchSize, p.maxProbeCount) == (0, 0, 1, 0.5, 0.117, 30, 200)


def test_reference_error_codes_without_cuda(built_lib):
    img = G(16, 12, 3, 1)
    mask = centered_mask(16, 12, 4, 4)
    p = abi.default_params(); p.patchSize = 65
    assert api.image_synth(img.copy(), mask, abi.T_RGB, p) == abi.IMAGE_SYNTH_ERROR_PATCH_SIZE_EXCEEDED
    assert api.image_synth(img.copy(), np.zeros((12, 16), np.uint8), abi.T_RGB, None) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
    assert api.image_synth(img.copy(), np.zeros((3, 3), np.uint8), abi.T_RGB, None) == abi.IMAGE_SYNTH_ERROR_IMAGE_MASK_MISMATCH
    assert api.image_synth(img.copy(), np.full((12, 16), 255, np.uint8), abi.T_RGB, None) == abi.IMAGE_SYNTH_ERROR_EMPTY_CORPUS
    assert api.image_synth(img.copy(), mask, 666, None) == abi.IMAGE_SYNTH_ERROR_INVALID_IMAGE_FORMAT
    p = abi.default_params(); p.matchContextType = 12
    assert api.image_synth(img.copy(), mask, abi.T_RGB, p) == abi.IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE


def test_selection_scan_finds_a_single_selected_byte_anywhere(built_lib):
    # the empty-target check reads the mask 64 bytes at a time with a byte tail: one selected pixel at any position of
    # a row whose width is not a multiple of 64 must be seen (whatever follows -- success or "no CUDA device" -- it is
    # not EMPTY_TARGET), and the padding of a mask with rowBytes > width must not count
    L = api.lib()
    w, h, row_bytes = 131, 5, 192
    img = G(w, h, 3, 2)
    cb = abi.PROGRESS_CB(lambda pct, ctx: None)

    def call(mask_rows):
        im, cancel = img.copy(), C.c_int(0)
        ib = abi.ImageBuffer(im.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, w * 3)
        mb = abi.ImageBuffer(mask_rows.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, row_bytes)
        return L.imageSynth(C.byref(ib), C.byref(mb), abi.T_RGB, None, cb, None, C.byref(cancel))

    padded = np.zeros((h, row_bytes), np.uint8)
    padded[:, w:] = 255
    assert call(padded) == abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET
    for x in (0, 63, 64, 127, 128, 130):
        m = padded.copy()
        m[h - 1, x] = 1
        assert call(m) != abi.IMAGE_SYNTH_ERROR_EMPTY_TARGET


def test_map_helpers(built_lib):
    L = api.lib()
    pm, bm = abi.Map(), abi.Map()
    L.new_pixmap(C.byref(pm), 5, 4, 3)
    L.new_bytemap(C.byref(bm), 5, 4)
    assert (pm.width, pm.height, pm.depth, bm.depth) == (5, 4, 3, 1)
    L.set_bytemap(C.byref(bm), C.c_ubyte(0x0F))
    L.invert_bytemap(C.byref(bm))
    L.interleave_mask(C.byref(pm), C.byref(bm))
    px = np.ctypeslib.as_array(C.cast(pm.data.contents.data, C.POINTER(C.c_ubyte)), (4, 5, 3))
    assert (px[:, :, 0] == 0xF0).all() and (px[:, :, 1:] == 0).all()
    L.free_map(C.byref(pm)); L.free_map(C.byref(bm))
    assert not pm.data


def test_no_cpu_fallback_when_cuda_is_absent(built_lib):
    """On a box without a GPU a synthesis request must raise, never quietly compute on the CPU."""
    if api.lib().rs_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    img = G(16, 12, 3, 1)
    before = img.copy()
    with pytest.raises(api.ResynthError):
        api.image_synth(img, centered_mask(16, 12, 4, 4), abi.T_RGB, None)
    assert (img == before).all()


def test_host_cores_divides_the_box_among_local_ranks(built_lib):
    """rs_host_cores(): helper threads a process may use = cores / LOCAL_WORLD_SIZE (one process per GPU), >= 1."""
    import subprocess
    import sys
    code = ("import ctypes, sys; L = ctypes.CDLL(sys.argv[1]); L.rs_host_cores.restype = ctypes.c_uint; "
            "print(L.rs_host_cores(), L.rs_host_cores())")
    def cores(**env):
        e = dict(os.environ); e.pop("LOCAL_WORLD_SIZE", None); e.pop("RS_HOST_THREADS", None); e.update(env)
        out = subprocess.run([sys.executable, "-c", code, api.LIB_PATH], env=e, capture_output=True, text=True, timeout=60)
        a, b = out.stdout.split()
        assert a == b
        return int(a)
    hw = os.cpu_count() or 1
    assert cores() == hw
    assert cores(LOCAL_WORLD_SIZE="8") == max(1, hw // 8)
    assert cores(LOCAL_WORLD_SIZE="8", RS_HOST_THREADS="3") == 3


def test_batch_calls_fail_loudly_without_cuda(built_lib):
    """The batch dealers (rs_image_synth_batch / rs_engine_batch_multi) have no CPU path either: without a device every
    job reports the CUDA-layer error, the images stay untouched, and rs_last_error() says why."""
    if api.lib().rs_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    imgs = [G(16, 12, 3, k) for k in range(3)]
    before = [i.copy() for i in imgs]
    masks = [centered_mask(16, 12, 4, 4)] * 3
    with pytest.raises(api.ResynthError) as e:
        api.image_synth_batch(imgs, masks, abi.T_RGB, None, devices=None, slots=2)
    assert "CUDA" in str(e.value)
    with pytest.raises(api.ResynthError):
        api.image_synth_batch(imgs, masks, abi.T_RGB, None, devices=[0, 1], slots=2)   # ordinals out of range
    assert all((a == b).all() for a, b in zip(imgs, before))
    assert api.image_synth_batch([], [], abi.T_RGB) == []                               # an empty batch is a no-op


def test_product_sources_do_not_reference_the_oracle():
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "resynthesizer_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/_" not in text, f
