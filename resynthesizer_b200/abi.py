"""ctypes mirror of the C ABI in include/resynthesizer.h.

Struct layouts follow the reference's public types so that one set of
definitions drives both this library and (in tests) the compiled reference:
ImageBuffer (lib/imageBuffer.h:12-19), TImageSynthParameters
(lib/engineParams.h:31-86), TFormatIndices (lib/imageFormatIndicies.h:46-58),
Map / Coordinates (lib/map.h:28-46), GArray prefix (lib/glibProxy.h:86-92).
"""
import ctypes as C

import numpy as np

# TImageFormat (lib/imageFormat.h:36-42)
T_RGB, T_RGBA, T_Gray, T_GrayA = 0, 1, 2, 3
FORMAT_CHANNELS = {T_RGB: 3, T_RGBA: 4, T_Gray: 1, T_GrayA: 2}

# TImageSynthError (lib/engineParams.h:13-28)
IMAGE_SYNTH_SUCCESS = 0
IMAGE_SYNTH_ERROR_INVALID_IMAGE_FORMAT = 1
IMAGE_SYNTH_ERROR_IMAGE_MASK_MISMATCH = 2
IMAGE_SYNTH_ERROR_PATCH_SIZE_EXCEEDED = 3
IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE = 4
IMAGE_SYNTH_ERROR_EMPTY_TARGET = 5
IMAGE_SYNTH_ERROR_EMPTY_CORPUS = 6


class ImageBuffer(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_ubyte)), ("width", C.c_uint),
                ("height", C.c_uint), ("rowBytes", C.c_size_t)]


class TImageSynthParameters(C.Structure):
    _fields_ = [("isMakeSeamlesslyTileableHorizontally", C.c_int),
                ("isMakeSeamlesslyTileableVertically", C.c_int),
                ("matchContextType", C.c_int),
                ("mapWeight", C.c_double),
                ("sensitivityToOutliers", C.c_double),
                ("patchSize", C.c_uint),
                ("maxProbeCount", C.c_uint)]


class TFormatIndices(C.Structure):
    _fields_ = [("colorEndBip", C.c_ubyte), ("alpha_bip", C.c_ubyte),
                ("map_start_bip", C.c_ubyte), ("map_end_bip", C.c_ubyte),
                ("img_match_bpp", C.c_ubyte), ("map_match_bpp", C.c_ubyte),
                ("total_bpp", C.c_ubyte),
                ("isAlphaTarget", C.c_int), ("isAlphaSource", C.c_int)]


class GArray(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint)]


class Map(C.Structure):
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("depth", C.c_uint),
                ("data", C.POINTER(GArray))]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p)

assert C.sizeof(ImageBuffer) == 24
assert C.sizeof(TImageSynthParameters) == 40
assert C.sizeof(TFormatIndices) == 16
assert C.sizeof(Map) == 24


def default_params():
    """Values of setDefaultParams (lib/engineParams.c:8-19)."""
    p = TImageSynthParameters()
    p.isMakeSeamlesslyTileableHorizontally = 0
    p.isMakeSeamlesslyTileableVertically = 0
    p.matchContextType = 1
    p.mapWeight = 0.5
    p.sensitivityToOutliers = 0.117
    p.patchSize = 30
    p.maxProbeCount = 200
    return p


def make_params(htile=0, vtile=0, ctx=1, map_weight=0.5, sigma=0.117, patch=30, probes=200):
    p = TImageSynthParameters()
    p.isMakeSeamlesslyTileableHorizontally = int(htile)
    p.isMakeSeamlesslyTileableVertically = int(vtile)
    p.matchContextType = int(ctx)
    p.mapWeight = float(map_weight)
    p.sensitivityToOutliers = float(sigma)
    p.patchSize = int(patch)
    p.maxProbeCount = int(probes)
    return p


def image_buffer(arr):
    """Wrap a C-contiguous uint8 array (h, rowBytes) or (h, w, c); returns (ImageBuffer, keepalive)."""
    assert arr.dtype == np.uint8 and arr.flags["C_CONTIGUOUS"]
    if arr.ndim == 2:
        raise ValueError("pass (h, w, c); use image_buffer_padded for explicit rowBytes")
    h, w, c = arr.shape
    ib = ImageBuffer(arr.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, w * c)
    return ib, arr


def image_buffer_padded(flat, width, height, row_bytes):
    assert flat.dtype == np.uint8 and flat.flags["C_CONTIGUOUS"]
    ib = ImageBuffer(flat.ctypes.data_as(C.POINTER(C.c_ubyte)), width, height, row_bytes)
    return ib, flat


def make_map(pixmap):
    """Caller-owned Map over a (h, w, depth) uint8 array (engine() full API)."""
    assert pixmap.dtype == np.uint8 and pixmap.flags["C_CONTIGUOUS"] and pixmap.ndim == 3
    h, w, d = pixmap.shape
    ga = GArray(pixmap.ctypes.data, w * h)
    m = Map(w, h, d, C.pointer(ga))
    return m, (ga, pixmap)


def bind(lib):
    """Declare the argument types of the reference-compatible entry points on a CDLL."""
    lib.imageSynth.argtypes = [C.POINTER(ImageBuffer), C.POINTER(ImageBuffer), C.c_int,
                               C.POINTER(TImageSynthParameters), PROGRESS_CB, C.c_void_p,
                               C.POINTER(C.c_int)]
    lib.imageSynth.restype = C.c_int
    lib.imageSynth2.argtypes = [C.POINTER(ImageBuffer), C.POINTER(ImageBuffer), C.POINTER(ImageBuffer),
                                C.c_int, C.POINTER(TImageSynthParameters), PROGRESS_CB, C.c_void_p,
                                C.POINTER(C.c_int)]
    lib.imageSynth2.restype = C.c_int
    lib.engine.argtypes = [TImageSynthParameters, C.POINTER(TFormatIndices), C.POINTER(Map),
                           C.POINTER(Map), PROGRESS_CB, C.c_void_p, C.POINTER(C.c_int)]
    lib.engine.restype = C.c_int
    lib.setDefaultParams.argtypes = [C.POINTER(TImageSynthParameters)]
    lib.setDefaultParams.restype = None
    lib.prepareImageFormatIndices.argtypes = [C.POINTER(TFormatIndices), C.c_uint, C.c_uint,
                                              C.c_int, C.c_int, C.c_int]
    lib.prepareImageFormatIndices.restype = None
    lib.prepareImageFormatIndicesFromFormatType.argtypes = [C.POINTER(TFormatIndices), C.c_int]
    lib.prepareImageFormatIndicesFromFormatType.restype = C.c_int
    lib.countPixelelsPerPixelForFormat.argtypes = [C.c_int]
    lib.countPixelelsPerPixelForFormat.restype = C.c_uint
    return lib
