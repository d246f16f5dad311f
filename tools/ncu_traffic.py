#!/usr/bin/env python
"""DRAM traffic of the synthesis launches of ONE job, from an ncu launch list, for bench.py's roofline.traffic.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'k_synth_pass|k_gather|k_ctx' --csv --log-file gpurun_out/traffic_cfg3.csv python tools/ncu_job.py --workload cfg3 --jobs 2
  python tools/ncu_traffic.py gpurun_out/traffic_cfg3.csv cfg3 hbm        -> profiles/traffic_cfg3.json

The last job's launches are taken (the first job also builds the offsets table and warms the workspace).  `bound` is
the label bench.py prints for this workload: "hbm" when the DRAM traffic of the pass kernels is a large fraction of
what the HBM can deliver in their time, "l2_gather" when the working set is L2-resident and DRAM is idle."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, wname, bound):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    iid, iname, imetric, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = {}
    for r in rows[1:]:
        if len(r) <= ival:
            continue
        d = launches.setdefault(int(r[iid]), {"kernel": r[iname].split("(")[0]})
        v = float(r[ival].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
                "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}.get(r[iunit], 1)
        d[r[imetric]] = v * mult
    ids = sorted(launches)
    # split into jobs at the first synthesis launch of a job: the context-block count ahead of the pass-0 patch gathers
    first = "k_ctx_blocks" if any("k_ctx_blocks" in launches[i]["kernel"] for i in ids) else "k_gather_pass0_coop"
    starts = [i for i in ids if first in launches[i]["kernel"]]
    last = [i for i in ids if i >= starts[-1]] if starts else ids
    per = [{"kernel": launches[i]["kernel"], "ms": launches[i].get("gpu__time_duration.sum", 0.0),
            "dram_bytes": launches[i].get("dram__bytes_read.sum", 0.0) + launches[i].get("dram__bytes_write.sum", 0.0)} for i in last]
    per = [p for p in per if p["ms"] > 0.02]   # launches of passes after the stop rule exit at once
    out = {"source": os.path.basename(path), "workload": wname, "bound": bound, "launches_captured": len(per),
           "dram_bytes_per_job": sum(p["dram_bytes"] for p in per), "ms_under_ncu": sum(p["ms"] for p in per),
           "per_launch": per,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of every synthesis launch (pass-0 gathers + pass kernels that did work) "
                   "of one job, ncu --clock-control none; per-launch times under ncu are serialised and cold-cache"}
    dst = os.path.join(ROOT, "profiles", "traffic_%s.json" % wname)
    if os.path.exists(dst):   # the hand-written note on what limits the kernels stays until it is rewritten
        old = json.load(open(dst))
        if "limiter" in old:
            out["limiter"] = old["limiter"]
    json.dump(out, open(dst, "w"), indent=1)
    print(wname, "launches", len(per), "dram GB/job %.3f" % (out["dram_bytes_per_job"] / 1e9), "ms", round(out["ms_under_ncu"], 3))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "hbm")
