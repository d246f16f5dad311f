#!/usr/bin/env python
"""Turns the artifacts of tools/round_profile.sh (gpurun_out/<tag>_*) into profiles/ copies and the measurement tables
of DESIGN.md section 8.

  python tools/measurements_md.py r01d      -> prints the markdown block, copies the artifacts to profiles/<tag>_*
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(src, "%s_%s" % (tag, name))
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().split("\n") if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


rows = []
for w in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    d = load("bench_%s.json" % w)
    if d is None:
        continue
    e, r = d["e2e"], d["roofline"]
    oc = (d.get("e2e_order_cached") or {}).get("value")
    cpu = (d.get("cpu_baseline") or {}).get("value")
    rows.append("| %s | %.3g | %.3g | %.2f | %s | %.3g | %s | %.2f / %.2f | %.3f | %.3f | %s |" % (
        w, d["value"], d["evals_per_s"], e["ms_kernels"], " ".join("%.2f" % x for x in d["ms_pass"][:d["passes_run"]]),
        e["value"], ("%.3g" % oc) if oc else "-", e["ms_prep"], e["ms_h2d"], r["frac"], r["gather_ceiling"]["frac"],
        ("%.3g" % cpu) if cpu else "-"))
out = []
out.append("| config | target px/s (kernels) | evals/s | kernel ms per job | per pass ms | e2e px/s (order cache off) | e2e px/s (cache on) | host prep / stage+digest ms | roofline.frac (HBM) | gather-ceiling frac | reference, 1 core, px/s |")
out.append("|---|---|---|---|---|---|---|---|---|---|---|")
out += rows
b = load("bench_cfg5_batch64.json")
if b:
    out.append("")
    out.append("Batch (`rs_engine_batch`, cfg5 x 64 jobs per step, order cache on): %.3g px/s = %.2f ms per job." % (
        b["value"], 1000.0 * 65536 / b["value"]))
ref = load("bench_reference.json")
if ref and "value" in ref:
    out.append("")
    out.append("Reference arm (`bench.py --impl reference`, cfg2 sample): %.3g px/s on %d host cores (%s)." % (
        ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"]))
d2 = load("bench_cfg2.json")
if d2:
    g = d2["roofline"]["gather_ceiling"]
    out.append("")
    out.append("cfg2 roofline detail: %.0f GB/s algorithmic vs %.1f GB/s HBM peak (frac %.3f); %.3g compares/s vs %.3g "
               "random 4-byte gathers/s measured in the same run (frac %.3f); %d pass-kernel launches that did work, "
               "%.3f ms each on average; clocks %s." % (
                   d2["roofline"]["achieved"], d2["roofline"]["peak"], d2["roofline"]["frac"], g["compares_per_s"],
                   g["loads_per_s"], g["frac"], d2["roofline"]["launches"], d2["roofline"]["avg_launch_ms"],
                   json.dumps(d2["clocks"])))
print("\n".join(out))

os.makedirs(dst, exist_ok=True)
for f in sorted(os.listdir(src)):
    if f.startswith(tag + "_") and (f.endswith(".json") or f.endswith(".txt") or f.endswith(".csv")):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
