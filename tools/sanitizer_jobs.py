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
