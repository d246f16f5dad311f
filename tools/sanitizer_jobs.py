import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from resynthesizer_b200 import api, abi
from resynthesizer_b200.synthetic import G, centered_mask
import bench
# small jobs through both kernels (team for small n; warp via env override), maps + tiling, and the bestfit path
img = G(96,80,3,5); m = centered_mask(96,80,30,24)
out = img.copy(); assert api.image_synth(out, m, abi.T_RGB, None) == 0
w = bench.workload("cfg4", 0, 0.03125)
fi = api.format_indices(3,3,False,False,True); tp,cp = bench.pixmaps(w); assert api.engine(w["params"], fi, tp, cp) == 0
w = bench.workload("cfg2", 0, 0.0625)
fi = api.format_indices(3,0,False,False,False); tp,cp = bench.pixmaps(w); assert api.engine(w["params"], fi, tp, cp) == 0
print("ok", api.last_stats()["visits"])
# the device-side ordering paths and the later-pass patch lists, forced on small jobs
os.environ["RS_LATER_LISTS_MIN"] = "1"
api.order_cache(False)
api.set_device_shuffle_min(1); api.set_device_sort_min(1)
for mode in (1, 2, 5):
    out = img.copy(); p = abi.make_params(0, 0, mode, 0.5, 0.117, 16, 60)
    assert api.image_synth(out, m, abi.T_RGB, p) == 0
os.environ["RS_HOST_PRNG"] = "1"   # the host's producer thread instead of k_mt19937_raw
out = img.copy(); assert api.image_synth(out, m, abi.T_RGB, abi.make_params(0, 0, 1, 0.5, 0.117, 16, 60)) == 0
del os.environ["RS_HOST_PRNG"]
os.environ["RS_NO_RAW_STREAM"] = "1"
out = img.copy(); assert api.image_synth(out, m, abi.T_RGB, abi.make_params(0, 0, 1, 0.5, 0.117, 16, 60)) == 0
api.order_cache(True)
out = img.copy(); assert api.image_synth(out, m, abi.T_RGB, None) == 0
out = img.copy(); assert api.image_synth(out, m, abi.T_RGB, None) == 0 and api.last_stats()["order_cache_hit"] == 1
# the PRNG stream kernel with several CTAs (jump-ahead): 3 CTAs
import ctypes as C
L = api.lib(); L.rs_cuda_mt19937_raw.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
words = np.zeros(600000, np.uint32)
assert L.rs_cuda_mt19937_raw(1198472, 600000, words.ctypes.data) == 0
assert (words == np.random.RandomState(1198472).randint(0, 2 ** 32, 600000, dtype=np.uint32)).all()
print("ordering paths ok")
