#!/usr/bin/env python
"""Algorithmic work counts of the BASELINE configurations, from the sequential oracle.

Runs oracle/resynth_port.c in GPU mode (counter-hash probes, lagged-epoch recentProber: the sequential definition of
what the CUDA engine computes, bit for bit) on the full-size synthetic workloads of bench.py and writes the counters
of the SEQUENTIAL algorithm -- visits, evals, neighbour compares, offset scans, heuristic evaluations -- to
profiles/algorithmic_counts.json.  bench.py's roofline uses these (algorithmic bytes, not the
compares the parallel early-out happens to issue) after checking that the device's own visit / eval counts are
identical.  TEST/MEASUREMENT INFRASTRUCTURE: runs on the CPU, here, once per workload definition.

  python tools/algorithmic_counts.py [cfg1 cfg2 cfg5 cfg5:50 ... cfg4 cfg3]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import refdriver as R  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "algorithmic_counts.json")


def run(key):
    name, _, probes = key.partition(":")
    bench.PROBES_OVERRIDE = int(probes) if probes else 0
    w = bench.workload(name)
    port = R.load_port(R.GPU_MODE, 1198472)
    t0 = time.time()
    if "simple" in w:
        err, _ = R.image_synth(port, w["tgt"], w["tmask"], w["simple"], w["params"])
    else:
        fi = R.format_indices(port, w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
        tp, cp = bench.pixmaps(w)
        err = R.engine(port, w["params"], fi, tp, cp)
    assert err == 0
    st = R.port_stats(port)
    st["oracle_seconds"] = round(time.time() - t0, 2)
    st["workload"] = w["name"]
    st["semantics"] = "oracle/resynth_port.c GPU_MODE (2,2), seed 1198472"
    return st


def main():
    keys = sys.argv[1:] or ["cfg1", "cfg2", "cfg5", "cfg5:50", "cfg5:100", "cfg5:500", "cfg5:1000", "cfg4", "cfg3"]
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for k in keys:
        data[k] = run(k)
        print(k, json.dumps(data[k]), flush=True)
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
