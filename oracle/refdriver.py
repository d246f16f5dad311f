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
