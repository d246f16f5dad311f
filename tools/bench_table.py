#!/usr/bin/env python
"""The table of DESIGN.md section 8.1 from one bench line (headline + `configs` sub-records):
  python tools/bench_table.py profiles/r03w_bench.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
print("| config | call | target px/s (kernels) | evals/s | kernel ms per job | per pass ms | e2e px/s | e2e ms per call | host prep / copies queued+digest / read-back ms | roofline.frac (HBM, algorithmic bytes) |")
print("|---|---|---|---|---|---|---|---|---|---|")
e = d["e2e"]
n_px = d["value"] * e["ms_kernels"] / 1000.0 / d["n_gpus"]
print("| %s (headline) | %s | %.3g | %.3g | %.2f | %s | %.3g | %.2f | %.2f / %.2f / %.2f | %.3f |" % (
    d["config"]["workload"].split()[0], d["config"]["api"].split("(")[0], d["value"], d["evals_per_s"], e["ms_kernels"],
    " ".join("%.2f" % x for x in d["ms_pass"][:d["passes_run"]]), e["value"], e["ms_call"], e["ms_prep"], e["ms_h2d"],
    e["ms_d2h"], d["roofline"]["frac"]))
for k, v in (d.get("configs") or {}).items():
    if "ms_kernels" not in v:
        continue
    print("| %s | %s | %.3g | %.3g | %.2f | %s | %.3g | %.2f | %.2f / %.2f / %.2f | %.3f |" % (
        k, v["api"].split("(")[0], v["value"], v["evals_per_s"], v["ms_kernels"],
        " ".join("%.2f" % x for x in v["ms_pass"][:v["passes_run"]]), v["e2e"], v["ms_e2e"], v["ms_prep"], v["ms_h2d"],
        v["ms_d2h"], v["roofline"]["frac"]))
p = d.get("e2e_pageable")
if p:
    print("\nmalloc'ed caller buffers (`e2e_pageable`): %.3g px/s, %.2f ms per call." % (p["value"], p["ms_call"]))
oc = d.get("e2e_order_cached")
if oc:
    print("visit-order cache on (`e2e_order_cached`): %.3g px/s." % oc["value"])
b = d.get("cfg5_batch")
if b:
    print("\ncfg5 batch (%d jobs, %d GPU(s)): " % (b["jobs"], d["n_gpus"]) +
          "; ".join("%s probes %.2f ms per job (%.3g px/s)" % (k, v["ms_per_job"], v["px_per_s"]) for k, v in b["probes"].items()))
sc = (d.get("configs") or {}).get("shared_corpus_batch")
if sc:
    print("shared corpus (%s): loop %.2f ms per job, batch %.2f (%.2fx)." % (sc["workload"], sc["loop_ms_per_job"], sc["batch_ms_per_job"], sc["speedup"]))
c = d.get("cpu_baseline")
if c:
    print("cpu_baseline: %.3g px/s on %d core(s) (%s)" % (c["value"], c["cores"], c["sample"]))
print("clocks:", d.get("clocks"))
