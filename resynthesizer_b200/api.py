"""Python mirror of the reference library's API over libresynthesizer_b200.so (ctypes).

Same names, argument meaning and error codes as the reference's C API
(lib/imageSynth.h:31-52, lib/engine.h:3-12): image_synth() / image_synth2() /
engine().  The shared library is the product; this module only marshals numpy
arrays.  If the library is missing or no CUDA device is usable the calls fail
loudly -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# RS_LIB_VARIANT selects a differently compiled build of the same sources (kernel parameter sweeps, see build.py)
LIB_PATH = os.path.join(_HERE, "lib", "libresynthesizer_b200%s.so" % os.environ.get("RS_LIB_VARIANT", ""))
_lib = None


class ResynthError(RuntimeError):
    pass


class RsStats(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in ("visits", "evals", "evals_issued", "compares", "offset_scans",
                                             "heur_evals", "heur_skips", "perfect")] + \
               [("betters", C.c_ulonglong * 6), ("pass_visits", C.c_ulonglong * 6), ("sum_best", C.c_ulonglong * 6),
                ("passes_run", C.c_uint), ("n_targets", C.c_uint), ("n_corpus", C.c_uint),
                ("ms_prep", C.c_float), ("ms_h2d", C.c_float), ("ms_kernels", C.c_float),
                ("ms_d2h", C.c_float), ("ms_total", C.c_float), ("ms_pass", C.c_float * 6), ("ms_synth", C.c_float),
                ("kernel_launches", C.c_uint), ("synth_launches_run", C.c_uint), ("order_cache_hit", C.c_uint)]

    def as_dict(self):
        d = {}
        for name, _t in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class RsJobDesc(C.Structure):
    _fields_ = [("tw", C.c_int32), ("th", C.c_int32), ("cw", C.c_int32), ("ch", C.c_int32), ("bpp", C.c_int32),
                ("n_color", C.c_int32), ("n_map", C.c_int32), ("map_bip", C.c_int32), ("alpha_bip", C.c_int32),
                ("alpha_target", C.c_int32), ("alpha_source", C.c_int32), ("htile", C.c_int32), ("vtile", C.c_int32), ("use_context", C.c_int32),
                ("patch_size", C.c_uint32), ("max_probes", C.c_uint32), ("seed", C.c_uint32),
                ("pass_end", C.c_uint32 * 6), ("n_passes", C.c_uint32), ("terminate_fraction", C.c_double),
                ("ordered_visits", C.c_int32), ("reserved", C.c_int32)]


def lib():
    """Loads the CUDA library (once). Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ResynthError("%s is missing: run `python -m resynthesizer_b200.build` (nvcc, sm_100a)" % LIB_PATH)
        L = abi.bind(C.CDLL(LIB_PATH))
        L.rs_last_error.restype = C.c_char_p
        L.rs_cuda_last_error.restype = C.c_char_p
        L.rs_get_stats.argtypes = [C.POINTER(RsStats)]
        L.rs_set_seed.argtypes = [C.c_uint]
        L.rs_set_device.argtypes = [C.c_int]
        L.rs_cuda_device_count.restype = C.c_int
        L.rs_host_metric_tables.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.rs_host_sorted_offsets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        L.rs_host_sorted_offsets.restype = C.c_uint32
        L.rs_host_order_targets.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
        L.rs_host_pass_schedule.argtypes = [C.c_uint32, C.c_void_p]
        L.rs_host_pass_schedule.restype = C.c_uint32
        L.rs_bestfit_batch.argtypes = [C.POINTER(RsJobDesc)] + [C.c_void_p] * 3 + [C.c_uint32, C.c_uint32] + \
                                      [C.c_void_p] * 7
        _lib = L
    return _lib


def _check(err):
    if err >= 100:
        raise ResynthError("CUDA layer failed (code %d): %s" % (err, lib().rs_last_error().decode()))
    return err


def set_device(ordinal):
    _check(lib().rs_set_device(int(ordinal)))


def set_seed(seed):
    lib().rs_set_seed(int(seed) & 0xFFFFFFFF)


def last_stats():
    s = RsStats()
    lib().rs_get_stats(C.byref(s))
    return s.as_dict()


def order_cache(enabled=True):
    """Enable (default) or drop + disable the device-side cache of visit orders (rs_order_cache)."""
    lib().rs_order_cache(1 if enabled else 0)


def set_device_sort_min(n_points):
    """Point lists of at least n_points are sorted on the device for orderings 2-8 (rs_set_device_sort_min)."""
    lib().rs_set_device_sort_min(int(n_points))


def set_device_shuffle_min(n_points):
    """Shuffling orders (matchContextType 0, 1) of at least n_points are resolved on the device."""
    lib().rs_set_device_shuffle_min(int(n_points))


def total_kernel_launches():
    L = lib()
    L.rs_total_kernel_launches.restype = C.c_ulonglong
    return int(L.rs_total_kernel_launches())


def gather_rate(buffer_bytes, elem_bytes=4, repeats=3):
    """Sustained random-gather rate of this GPU for corpus-pixel-sized loads, loads/s (rs_cuda_gather_rate)."""
    L = lib()
    L.rs_cuda_gather_rate.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.rs_cuda_gather_rate.restype = C.c_int
    out = C.c_double(0.0)
    if L.rs_cuda_gather_rate(int(buffer_bytes), int(elem_bytes), int(repeats), C.byref(out)):
        raise ResynthError("rs_cuda_gather_rate failed")
    return out.value


def plan_pass(n_targets, pass_end, patch_size, pass_index, ordered_visits=False):
    """Launch plan of one pass: [(segment end, warps per visit)], 1 warp = the throughput kernel (rs_cuda_plan_pass)."""
    L = lib()
    L.rs_cuda_plan_pass.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.rs_cuda_plan_pass.restype = C.c_int
    ends, widths = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    k = L.rs_cuda_plan_pass(int(n_targets), int(pass_end), int(bool(ordered_visits)), int(patch_size), int(pass_index), ends, widths)
    return [(int(ends[i]), int(widths[i])) for i in range(k)]


def last_timeline(pass_index):
    """ns from the start of `pass_index` to the claim of its visit 4096*i, for the last engine() call made with
    keep_result(True); empty if the pass did not run."""
    out = np.zeros(512, np.uint64)
    L = lib()
    L.rs_get_timeline.argtypes = [C.c_uint, C.c_void_p, C.c_uint]
    L.rs_get_timeline.restype = C.c_uint
    n = L.rs_get_timeline(int(pass_index), out.ctypes.data, 512)
    return out[:n].copy()


class _Progress:
    def __init__(self, callback, cancel_after):
        self.percents = []
        self.cancel = C.c_int(0)

        def cb(percent, _ctx):
            self.percents.append(percent)
            if callback is not None:
                callback(percent)
            if cancel_after is not None and len(self.percents) >= cancel_after:
                self.cancel.value = 1
        self.cb = abi.PROGRESS_CB(cb)


def image_synth(image, mask, fmt, params=None, progress=None, cancel_after=None, mask2=None, return_progress=False):
    """imageSynth()/imageSynth2(): heal the masked part of `image` (h,w,c uint8) in place.

    Returns the reference's error code (0 = success).  `progress(percent)` is
    the reference's progress callback; `cancel_after=k` raises the cancel flag
    inside the k-th callback.
    """
    L = lib()
    assert image.dtype == np.uint8 and image.ndim == 3 and image.flags["C_CONTIGUOUS"]
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w, c = image.shape
    ib = abi.ImageBuffer(image.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, w * c)
    mb = abi.ImageBuffer(mask.ctypes.data_as(C.POINTER(C.c_ubyte)), mask.shape[1], mask.shape[0], mask.shape[1])
    pr = _Progress(progress, cancel_after)
    pp = C.byref(params) if params is not None else None
    if mask2 is None:
        err = L.imageSynth(C.byref(ib), C.byref(mb), fmt, pp, pr.cb, None, C.byref(pr.cancel))
    else:
        mask2 = np.ascontiguousarray(mask2, dtype=np.uint8)
        mb2 = abi.ImageBuffer(mask2.ctypes.data_as(C.POINTER(C.c_ubyte)), mask2.shape[1], mask2.shape[0], mask2.shape[1])
        err = L.imageSynth2(C.byref(ib), C.byref(mb), C.byref(mb2), fmt, pp, pr.cb, None, C.byref(pr.cancel))
    _check(err)
    return (err, pr.percents) if return_progress else err


def format_indices(n_color, n_map=0, alpha_target=False, alpha_source=False, is_map=False):
    fi = abi.TFormatIndices()
    lib().prepareImageFormatIndices(C.byref(fi), n_color, n_map, int(alpha_target), int(alpha_source), int(is_map))
    return fi


def engine(params, fi, target_pixmap, corpus_pixmap, progress=None, cancel_after=None, return_progress=False):
    """engine(): full API over internal pixmaps [mask][colours][alpha?][maps]; target_pixmap changes in place."""
    L = lib()
    tm, _k1 = abi.make_map(target_pixmap)
    cm, _k2 = abi.make_map(corpus_pixmap)
    pr = _Progress(progress, cancel_after)
    err = L.engine(params, C.byref(fi), C.byref(tm), C.byref(cm), pr.cb, None, C.byref(pr.cancel))
    _check(err)
    return (err, pr.percents) if return_progress else err


def keep_result(yes=True):
    """Make later engine() calls also fetch per-target sources (see last_result())."""
    lib().rs_keep_result(1 if yes else 0)


def last_result():
    """(targets, sources) of the last engine() call: (n,2) int32 arrays in visit order; source (-1,-1) = none."""
    L = lib()
    L.rs_get_last_result.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.rs_get_last_result.restype = C.c_uint
    n = L.rs_get_last_result(None, None, 0)
    t = np.zeros(n, np.uint32); s = np.zeros(n, np.uint32)
    L.rs_get_last_result(t.ctypes.data, s.ctypes.data, n)
    txy = np.stack([t & 0xFFFF, t >> 16], axis=1).astype(np.int32)
    sxy = np.stack([s & 0xFFFF, s >> 16], axis=1).astype(np.int32)
    sxy[s == 0xFFFFFFFF] = -1
    return txy, sxy


def device_sorted_offsets(tw, th, cw, ch):
    """The neighbour-offset table exactly as the CUDA layer builds it on the device ((n,2) int32), for tests."""
    L = lib()
    d = RsJobDesc()
    d.tw, d.th, d.cw, d.ch, d.bpp, d.n_color, d.n_map, d.map_bip, d.alpha_bip = tw, th, cw, ch, 4, 3, 0, 4, -1
    d.patch_size, d.max_probes, d.n_passes = 4, 1, 1
    d.pass_end[0] = 1
    d.terminate_fraction = 0.1
    job = C.c_void_p()
    L.rs_job_create.argtypes = [C.POINTER(RsJobDesc), C.POINTER(C.c_void_p)]
    L.rs_job_upload.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                                                  C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    L.rs_job_read_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.rs_job_destroy.argtypes = [C.c_void_p]
    if L.rs_job_create(C.byref(d), C.byref(job)):
        raise ResynthError(L.rs_cuda_last_error().decode())
    try:
        t = np.zeros((th, tw, 4), np.uint8); t[0, 0, 0] = 255
        c = np.zeros((ch, cw, 4), np.uint8); c[:, :, 0] = 255
        tg = np.zeros(1, np.uint32); cp = np.zeros(1, np.uint32); lut = np.zeros(256, np.uint32)
        if L.rs_job_upload(job, t.ctypes.data, c.ctypes.data, tg.ctypes.data, 1, cp.ctypes.data, 1, None, 0,
                           lut.ctypes.data, lut.ctypes.data, 0):
            raise ResynthError(L.rs_cuda_last_error().decode())
        w, h = min(tw, cw), min(th, ch)
        n = (2 * w - 1) * (2 * h - 1)
        out = np.zeros(n, np.uint32)
        if L.rs_job_read_offsets(job, out.ctypes.data, n):
            raise ResynthError(L.rs_cuda_last_error().decode())
    finally:
        L.rs_job_destroy(job)
    x = (out & 0xFFFF).astype(np.int16).astype(np.int32)
    y = (out >> 16).astype(np.int16).astype(np.int32)
    return np.stack([x, y], axis=1)


def image_synth_batch(images, masks, fmt, params=None, devices=None, slots=4, masks2=None):
    """rs_image_synth_batch(): imageSynth() (or imageSynth2() where masks2[i] is given) for every image of the batch,
    dealt over `devices` (CUDA ordinals; None = this thread's device) from one process; images change in place.
    Returns the list of per-job error codes."""
    L = lib()
    n = len(images)
    IB = (C.POINTER(abi.ImageBuffer) * n)()
    MB = (C.POINTER(abi.ImageBuffer) * n)()
    M2 = (C.POINTER(abi.ImageBuffer) * n)() if masks2 is not None else None
    keep = []

    def buf(a, row_channels):
        b = abi.ImageBuffer(a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.shape[1], a.shape[0], a.shape[1] * row_channels)
        keep.append((a, b))
        return C.pointer(b)
    for i in range(n):
        im = images[i]
        assert im.dtype == np.uint8 and im.ndim == 3 and im.flags["C_CONTIGUOUS"]
        IB[i] = buf(im, im.shape[2])
        MB[i] = buf(np.ascontiguousarray(masks[i], dtype=np.uint8), 1)
        if M2 is not None and masks2[i] is not None:
            M2[i] = buf(np.ascontiguousarray(masks2[i], dtype=np.uint8), 1)
    errs = (C.c_int * n)()
    nd = len(devices) if devices else 0
    dv = (C.c_int * max(nd, 1))(*(devices or [0]))
    L.rs_image_synth_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p]
    rc = L.rs_image_synth_batch(n, IB, MB, M2, fmt, C.byref(params) if params is not None else None, nd,
                                dv if nd else None, int(slots), errs)
    _check(rc)
    return list(errs)


def shared_corpus_stats():
    """(corpora built, reuses, peer copies) by the batch calls of this process (rs_shared_corpus_stats)."""
    a, b, c = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
    lib().rs_shared_corpus_stats(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def engine_batch(jobs, slots=4, devices=None):
    """rs_engine_batch() / rs_engine_batch_multi(): jobs = list of (params, fi, target_pixmap, corpus_pixmap); pixmaps
    change in place.  devices: CUDA ordinals the batch is dealt over (None = this thread's device).
    Returns the list of per-job error codes."""
    L = lib()
    n = len(jobs)
    P = (abi.TImageSynthParameters * n)(*[j[0] for j in jobs])
    keep = []
    FI = (C.POINTER(abi.TFormatIndices) * n)()
    TM = (C.POINTER(abi.Map) * n)()
    CM = (C.POINTER(abi.Map) * n)()
    for i, (_p, fi, tp, cp) in enumerate(jobs):
        tm, k1 = abi.make_map(tp)
        cm, k2 = abi.make_map(cp)
        keep.append((fi, tm, cm, k1, k2))
        FI[i] = C.pointer(fi); TM[i] = C.pointer(tm); CM[i] = C.pointer(cm)
    errs = (C.c_int * n)()
    if devices:
        dv = (C.c_int * len(devices))(*devices)
        L.rs_engine_batch_multi.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_int, C.c_void_p]
        rc = L.rs_engine_batch_multi(n, P, FI, TM, CM, len(devices), dv, int(slots), errs)
    else:
        L.rs_engine_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        rc = L.rs_engine_batch(n, P, FI, TM, CM, int(slots), errs)
    _check(rc)
    return list(errs)
