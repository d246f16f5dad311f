"""TEST INFRASTRUCTURE (oracle/): drives a reference-ABI shared library from numpy.

Works for oracle/_ref/libref_*.so (the compiled reference), for
oracle/_port/libresynth_port.so (the C restatement) and -- because the ABI is
the same -- for the product library; the product never imports this module.
"""
import ctypes as C
import os

import numpy as np

from resynthesizer_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    """name: 'ref_mt_1t' | 'ref_rand_1t' | 'ref_rand_8t' | 'port'."""
    if name == "port":
        path = os.path.join(HERE, "_port", "libresynth_port.so")
    else:
        path = os.path.join(HERE, "_ref", "lib%s.so" % name)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return abi.bind(C.CDLL(path))


def have(name):
    """True when that library has been built (oracle/_ref needs /root/reference at build time)."""
    return os.path.exists(os.path.join(HERE, "_port", "libresynth_port.so") if name == "port"
                          else os.path.join(HERE, "_ref", "lib%s.so" % name))


class Progress:
    def __init__(self, cancel_after=None):
        self.percents = []
        self.cancel = C.c_int(0)
        self.cancel_after = cancel_after

        def cb(percent, ctx):
            self.percents.append(percent)
            if self.cancel_after is not None and len(self.percents) >= self.cancel_after:
                self.cancel.value = 1
        self.cb = abi.PROGRESS_CB(cb)


def image_synth(lib, image, mask, fmt, params=None, progress=None, row_pad=0, mask2=None):
    """imageSynth()/imageSynth2() over (h,w,c) uint8 image and (h,w) uint8 mask.

    Returns (error, result image). row_pad adds trailing bytes per row to
    exercise rowBytes.
    """
    h, w, c = image.shape
    rb = w * c + row_pad
    buf = np.zeros((h, rb), np.uint8)
    buf[:, :w * c] = image.reshape(h, w * c)
    mrb = mask.shape[1] + (1 if row_pad else 0)
    mbuf = np.zeros((mask.shape[0], mrb), np.uint8)
    mbuf[:, :mask.shape[1]] = mask
    ib, _k1 = abi.image_buffer_padded(buf.reshape(-1), w, h, rb)
    mb, _k2 = abi.image_buffer_padded(mbuf.reshape(-1), mask.shape[1], mask.shape[0], mrb)
    pr = progress or Progress()
    pp = C.byref(params) if params is not None else None
    if mask2 is None:
        err = lib.imageSynth(C.byref(ib), C.byref(mb), fmt, pp, pr.cb, None, C.byref(pr.cancel))
    else:
        m2 = np.ascontiguousarray(mask2)
        mb2, _k3 = abi.image_buffer_padded(m2.reshape(-1), m2.shape[1], m2.shape[0], m2.shape[1])
        err = lib.imageSynth2(C.byref(ib), C.byref(mb), C.byref(mb2), fmt, pp, pr.cb, None,
                              C.byref(pr.cancel))
    out = buf[:, :w * c].reshape(h, w, c).copy()
    return err, out


def format_indices(lib, n_color, n_map, alpha_target, alpha_source, is_map):
    fi = abi.TFormatIndices()
    lib.prepareImageFormatIndices(C.byref(fi), n_color, n_map, int(alpha_target), int(alpha_source), int(is_map))
    return fi


def build_pixmap(mask, color, alpha=None, maps=None):
    """Internal interleaved pixel [mask][colours][alpha?][maps] (lib/imageFormat.c:134-155)."""
    h, w = mask.shape
    parts = [mask.reshape(h, w, 1), color.reshape(h, w, -1)]
    if alpha is not None:
        parts.append(alpha.reshape(h, w, 1))
    if maps is not None:
        parts.append(maps.reshape(h, w, -1))
    return np.ascontiguousarray(np.concatenate(parts, axis=2).astype(np.uint8))


def engine(lib, params, fi, target_pixmap, corpus_pixmap, progress=None):
    """engine() over internal pixmaps; target_pixmap is modified in place. Returns error."""
    tm, _k1 = abi.make_map(target_pixmap)
    cm, _k2 = abi.make_map(corpus_pixmap)
    pr = progress or Progress()
    return lib.engine(params, C.byref(fi), C.byref(tm), C.byref(cm), pr.cb, None, C.byref(pr.cancel))


# ---------------------------------------------------------------- helpers for the C restatement (port)
class PortStats(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in ("visits", "evals", "compares", "offset_scans", "heur_evals",
                                             "heur_skips", "perfect")] + \
               [("betters", C.c_ulonglong * 6), ("pass_visits", C.c_ulonglong * 6), ("sum_best", C.c_ulonglong * 6),
                ("passes_run", C.c_uint), ("n_targets", C.c_uint), ("n_corpus", C.c_uint)]


REF_MODE = (0, 0)      # GLib MT19937 stream, live recentProber: the reference product build
RAND_MODE = (1, 0)     # libc rand() proxy: the reference standalone build
GPU_MODE = (2, 2)      # counter hash + epoch-snapshot recentProber: sequential definition of the CUDA engine


def load_port(mode=REF_MODE, seed=1198472):
    lib = load("port")
    lib.port_set_mode.argtypes = [C.c_int, C.c_int]
    lib.port_set_seed.argtypes = [C.c_uint]
    lib.port_get_stats.argtypes = [C.POINTER(PortStats)]
    lib.port_trace_enable.argtypes = [C.c_uint, C.c_uint]
    lib.port_trace_size.restype = C.c_size_t
    lib.port_trace_count.restype = C.c_uint
    lib.port_trace_copy.argtypes = [C.c_void_p]
    lib.port_luts.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.port_offsets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint]
    lib.port_offsets.restype = C.c_uint
    lib.port_order.argtypes = [C.c_int, C.c_void_p, C.c_uint, C.c_uint]
    lib.port_set_mode(*mode)
    lib.port_set_seed(seed)
    return lib


def port_stats(lib):
    s = PortStats()
    lib.port_get_stats(C.byref(s))
    d = {}
    for name, _t in s._fields_:
        v = getattr(s, name)
        d[name] = list(v) if hasattr(v, "__len__") else v
    return d


def port_trace(lib):
    """Parses the per-visit dump (see resynth_port.c 'trace') into a list of dicts."""
    n = lib.port_trace_size()
    buf = np.zeros(n, np.uint8)
    if n:
        lib.port_trace_copy(buf.ctypes.data)
    out, pos = [], 0
    while pos < n:
        hd = buf[pos:pos + 44].view(np.uint32)
        pos += 44
        K, nc = int(hd[4]), int(hd[5])
        nb = buf[pos:pos + 24 * K].reshape(K, 24)
        pos += 24 * K
        cands = buf[pos:pos + 8 * nc].view(np.int32).reshape(nc, 2).copy()
        pos += 8 * nc
        out.append(dict(pass_=int(hd[0]), index=int(hd[1]), x=int(hd[2]), y=int(hd[3]), K=K, best=int(hd[6]),
                        best_xy=(int(np.int32(hd[7])), int(np.int32(hd[8]))), bettered=int(hd[9]), n_heur=int(hd[10]),
                        offsets=nb[:, 0:8].copy().view(np.int32).reshape(K, 2),
                        pixels=nb[:, 8:16].copy(),
                        sources=nb[:, 16:24].copy().view(np.int32).reshape(K, 2),
                        cands=cands))
    return out


def port_last_result(lib):
    lib.port_last_result.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    lib.port_last_result.restype = C.c_uint
    n = lib.port_last_result(None, None, 0)
    t = np.zeros((n, 2), np.int32); s = np.zeros((n, 2), np.int32)
    lib.port_last_result(t.ctypes.data, s.ctypes.data, n)
    return t, s
