#!/usr/bin/env python
"""bench.py -- synthesized target px/s (and patch-distance evals/s) of the synthesis hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N ...            # the reference's CPU engine (oracle/_ref), same config

One "step" = one complete synthesis job (all passes) of the workload on each rank.  The default workload is the
largest single-GPU configuration of BASELINE.json, configs[2]: large-hole inpaint of a 4096x4096 RGBA image, 25 %
masked (centred 2048x2048), alpha weighting, defaults (30 neighbours / 200 probes), all six refiner passes, through
imageSynth().  It does not sit on the 10 % stop rule's knife edge (both engines run all six passes, so both arms do
the same 4.05 visits per target pixel), which makes px/s comparable between the arms; every line also carries
visits/s and evals/s, the equal-work rates.  The other BASELINE configurations ride along as sub-records
("configs"), and configuration 5 -- the FIXED batch of 64 heal jobs dealt over the ranks, with its probe-count
sweep -- as "cfg5_batch" (strong scaling; the headline is weak scaling: one job stream per GPU).

  value  = target px/s with inputs resident in HBM: n_targets / CUDA-event time of the job's kernels on the job's
           stream (pass-0 patch gather + every pass), whole job: N ranks x K steps, max over ranks.
  e2e    = the same metric through the reference-facing C-ABI call with HOST buffers: host prep, H2D, ordering,
           passes, D2H, write-back all inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from resynthesizer_b200 import abi, sharding  # noqa: E402
from resynthesizer_b200.synthetic import G, centered_mask  # noqa: E402

METRIC = "synthesized target px/s"
UNIT = "px/s"
DTYPE = "u8/u32 integer"
DEFAULT_WORKLOAD = "cfg3"
SEED = 1198472   # the reference seeds its PRNG with this constant in every engine() call (lib/engine.c:643)


# ------------------------------------------------------------------------------------------ workloads
PROBES_OVERRIDE = 0


def workload(name, seed_shift=0, scale=1.0):
    """Returns dict(params, fi_args, target pixmap builder inputs, n_color...)."""
    w = _workload(name, seed_shift, scale)
    if PROBES_OVERRIDE:
        w["params"].maxProbeCount = PROBES_OVERRIDE
        w["name"] = w["name"].replace("probes 200", "probes %d" % PROBES_OVERRIDE)
    return w


def _workload(name, seed_shift=0, scale=1.0):
    if name == "cfg2":      # render-texture 1024^2 from 256^2 corpus, ctx 0, 9/200
        t = int(1024 * scale)
        cor = G(256, 256, 3, 1 + seed_shift)
        tgt = np.full((t, t, 3), 255, np.uint8)
        return dict(name="cfg2 render-texture %dx%d from 256x256 corpus, ctx0, patch 9, probes 200" % (t, t),
                    params=abi.make_params(0, 0, 0, 0.5, 0.117, 9, 200), n_color=3, n_map=0, alpha=False,
                    tmask=np.full((t, t), 255, np.uint8), tgt=tgt, cmask=np.full((256, 256), 255, np.uint8), cor=cor,
                    bpp=4)
    if name == "cfg1":      # heal 64^2 hole in 512^2, defaults 30/200
        s = int(512 * scale)
        img = G(s, s, 3, 12345 + seed_shift)
        m = centered_mask(s, s, s // 8, s // 8)
        return dict(name="cfg1 heal %dx%d, %dx%d hole, ctx1, patch 30, probes 200" % (s, s, s // 8, s // 8),
                    params=abi.default_params(), n_color=3, n_map=0, alpha=False,
                    tmask=m, tgt=img, cmask=255 - m, cor=img, bpp=4, simple=abi.T_RGB)
    if name == "cfg1_brick":  # BASELINE configs[0] on the real image it names (Test/in_images/brick.png, packed by build())
        img = np.load(os.path.join(ROOT, "baseline", "_ref", "brick_512.npy"))
        m = centered_mask(512, 512, 64, 64)
        return dict(name="cfg1 heal brick.png 512x512, 64x64 hole, ctx1, patch 30, probes 200",
                    params=abi.default_params(), n_color=3, n_map=0, alpha=False,
                    tmask=m, tgt=img, cmask=255 - m, cor=img, bpp=4, simple=abi.T_RGB)
    if name == "cfg5":      # one heal job of the batch config: 2048^2, 256^2 hole, 30/200
        s = int(2048 * scale)
        img = G(s, s, 3, 100 + seed_shift)
        m = centered_mask(s, s, s // 8, s // 8)
        return dict(name="cfg5 heal %dx%d, %dx%d hole, ctx1, patch 30, probes 200" % (s, s, s // 8, s // 8),
                    params=abi.default_params(), n_color=3, n_map=0, alpha=False,
                    tmask=m, tgt=img, cmask=255 - m, cor=img, bpp=4, simple=abi.T_RGB)
    if name == "cfg3":      # large-hole inpaint 4096^2 RGBA, 25% masked (centred 2048^2), transparent band, 30/200
        s_ = int(4096 * scale)
        img = G(s_, s_, 4, 3 + seed_shift)
        img[:, :, 3] = 255
        img[:, s_ // 8:s_ // 8 + s_ // 16, 3] = 0       # 256-px transparent band at x in [512,768)
        m = centered_mask(s_, s_, s_ // 2, s_ // 2)
        return dict(name="cfg3 inpaint %dx%d RGBA, %dx%d hole, ctx1, patch 30, probes 200" % (s_, s_, s_ // 2, s_ // 2),
                    params=abi.default_params(), n_color=3, n_map=0, alpha=True,
                    tmask=m, tgt=img, cmask=255 - m, cor=img, bpp=5, simple=abi.T_RGBA)
    if name == "cfg4":      # map-style transfer 2048^2 / 2048^2, RGB maps = the images, mapWeight 0.5, tiling, 9/200
        s_ = int(2048 * scale)
        tgt = G(s_, s_, 3, 4 + seed_shift); cor = G(s_, s_, 3, 5 + seed_shift)
        full = np.full((s_, s_), 255, np.uint8)
        return dict(name="cfg4 map-style %dx%d target / corpus, RGB maps, mapWeight 0.5, tiled, ctx1, patch 9, probes 200" % (s_, s_),
                    params=abi.make_params(1, 1, 1, 0.5, 0.117, 9, 200), n_color=3, n_map=3, alpha=False,
                    tmask=full, tgt=tgt, cmask=full.copy(), cor=cor, bpp=7, tmaps=tgt.copy(), cmaps=cor.copy())
    if name.startswith("heal:"):   # heal:<image side>:<hole side>[:<matchContextType>]  (experiments)
        parts = name.split(":")
        side, hole = int(parts[1]), int(parts[2])
        mode = int(parts[3]) if len(parts) > 3 else 1
        img = G(side, side, 3, 4321 + seed_shift)
        m = centered_mask(side, side, hole, hole)
        return dict(name="heal %dx%d, %dx%d hole, ctx%d, patch 30, probes 200" % (side, side, hole, hole, mode),
                    params=abi.make_params(0, 0, mode, 0.5, 0.117, 30, 200), n_color=3, n_map=0, alpha=False,
                    tmask=m, tgt=img, cmask=255 - m, cor=img, bpp=4, simple=abi.T_RGB)
    raise SystemExit("unknown workload %s" % name)


def pixmaps(w):
    tparts = [w["tmask"][:, :, None], w["tgt"]]
    cparts = [w["cmask"][:, :, None], w["cor"]]
    if "tmaps" in w:
        tparts.append(w["tmaps"]); cparts.append(w["cmaps"])
    return np.ascontiguousarray(np.concatenate(tparts, axis=2)), np.ascontiguousarray(np.concatenate(cparts, axis=2))


def api_name(w):
    return "imageSynth() simple API" if "simple" in w else "engine() full API"


def config_of(w):
    """The `config` object of a line: identical in both arms (what the driver compares)."""
    return {"workload": w["name"], "api": api_name(w)}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def visits_of(n, passes_run):
    """Visits the pass schedule (lib/passes.h:67-93) makes in `passes_run` passes over n target points."""
    ends = [n, n]
    for _ in range(4):
        ends.append(ends[-1] * 3 // 4)
    return sum(ends[:passes_run])


def _ref_worker(args):
    libname, wname, scale, seed_shift, steps = args
    from oracle import refdriver as R
    lib = R.load(libname)
    w = workload(wname, seed_shift, scale)
    fi = R.format_indices(lib, w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    times = []
    n = int((w["tmask"] != 0).sum())
    for _ in range(steps):
        t0 = time.perf_counter()
        if "simple" in w:   # the same entry point our arm times for this workload
            err, _out = R.image_synth(lib, w["tgt"], w["tmask"], w["simple"], w["params"])
        else:
            tp, cp = pixmaps(w)
            t0 = time.perf_counter()
            err = R.engine(lib, w["params"], fi, tp, cp)
        times.append(time.perf_counter() - t0)
        assert err == 0
    return n, times


def cpu_reference_run(wname, scale, steps, warmup, procs, libname="ref_rand_1t"):
    """Runs `procs` identical sample jobs per step, one per host core, with the compiled reference."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        if warmup:
            pool.map(_ref_worker, [(libname, wname, scale, 0, warmup) for _ in range(procs)])
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(libname, wname, scale, 0, steps) for _ in range(procs)])
        wall = time.perf_counter() - t0
    n = res[0][0]
    return dict(px_per_s=procs * steps * n / wall, wall=wall, n=n, per_job_s=float(np.mean([np.mean(r[1]) for r in res])))


def reference_counters(wname, scale):
    """Visits / evals / passes of ONE reference job of the sample: the C restatement in the reference's own semantics
    with the standalone build's libc-rand stream (oracle RAND_MODE, pinned to libref_rand_1t output for output in
    tests/test_port_vs_ref.py) -- the compiled reference itself keeps no counters."""
    from oracle import refdriver as R
    port = R.load_port(R.RAND_MODE)
    w = workload(wname, 0, scale)
    if "simple" in w:
        err, _ = R.image_synth(port, w["tgt"], w["tmask"], w["simple"], w["params"])
    else:
        fi = R.format_indices(port, w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
        tp, cp = pixmaps(w)
        err = R.engine(port, w["params"], fi, tp, cp)
    assert err == 0
    return R.port_stats(port)


def reference_sample_scale(wname):
    # Bounded samples of the same workload.  Sized so that the sample runs the same number of passes as the full job
    # (checked below) and one step is a few seconds per core; cfg2 is run in full (its 1024^2 job stops after 2 passes,
    # its 512^2 sample after 3: not like for like, hence no scaling).
    return {"cfg2": 1.0, "cfg1": 1.0, "cfg5": 0.5, "cfg3": 0.125, "cfg4": 0.125}.get(wname, 1.0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rand_1t.so")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    procs = max(1, min(os.cpu_count() or 1, 64))
    scale = reference_sample_scale(a.workload)
    w_full = workload(a.workload)
    w = workload(a.workload, 0, scale)
    r = cpu_reference_run(a.workload, scale, a.steps, a.warmup, procs)
    ctr = reference_counters(a.workload, scale)
    n = r["n"]
    step_s = r["wall"] / a.steps
    sample = ("%s; each step = %d identical jobs, one per host core, complete (all passes) through the unthreaded "
              "reference build (oracle/_ref/libref_rand_1t.so); counters of one job from the oracle port in the same "
              "semantics" % (w["name"], procs))
    line = {"metric": METRIC, "value": r["px_per_s"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * step_s,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": config_of(w_full), "sample": sample,
            # equal-work keys (the same ones our arm prints): px/s compares like with like when visits_per_px agree
            "passes_run": ctr["passes_run"], "visits_per_px": ctr["visits"] / n, "evals_per_visit": ctr["evals"] / ctr["visits"],
            "visits_per_step": procs * ctr["visits"], "evals_per_step": procs * ctr["evals"],
            "visits_per_s": procs * ctr["visits"] / step_s, "evals_per_s": procs * ctr["evals"] / step_s,
            "compares_per_s": procs * ctr["compares"] / step_s,
            "cpu_baseline": {"value": r["px_per_s"], "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": r["px_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "visits_per_s": procs * ctr["visits"] / step_s, "evals_per_s": procs * ctr["evals"] / step_s},
            "gpu_launches": 0, "per_job_seconds": r["per_job_s"]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def algorithmic_counts(key):
    p = os.path.join(ROOT, "profiles", "algorithmic_counts.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key)


def algorithmic_bytes(c, w):
    """SURVEY.md section 8d, from the counters of the SEQUENTIAL algorithm (oracle): per neighbour compare bpp bytes of
    corpus; per scanned offset 8 + 1 (offset + hasValue); per visit K x (bpp + 8) (neighbour pixel + sourceOf) and the
    commit (colour bytes + 8 + 1); per heuristic candidate 4 + 4 (recentProber read + write); per probe drawn 8
    (corpusPoints lookup)."""
    bpp = w["bpp"]
    K = max(2, w["params"].patchSize)
    probes_drawn = c["evals"] - c["heur_evals"]
    return (c["compares"] * bpp + c["offset_scans"] * 9 + c["visits"] * (K * (bpp + 8) + w["n_color"] + 9) +
            (c["heur_evals"] + c["heur_skips"]) * 8 + probes_drawn * 8)


def roofline_of(w, wname, stats, steps, api, gather=True):
    """Roofline record of the synthesis kernels of `steps` identical jobs (their stats in `stats`)."""
    key = wname + (":%d" % PROBES_OVERRIDE if PROBES_OVERRIDE and PROBES_OVERRIDE != 200 else "")
    oc = algorithmic_counts(key)
    dev = {k: stats[-1][k] for k in ("visits", "evals", "compares", "offset_scans", "heur_evals", "heur_skips")}
    if oc is not None and (oc["visits"], oc["evals"]) == (dev["visits"], dev["evals"]):
        counts, src = oc, "profiles/algorithmic_counts.json (sequential oracle; visits and evals equal the device's)"
    else:
        counts, src = dev, ("the device's own counters (compares as issued by the chunked parallel early-out: an upper "
                            "bound of the sequential algorithm's)" + ("" if oc is None else "; the oracle's visit/eval counts differ!"))
    algo = algorithmic_bytes(counts, w)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "MEASURED_PEAKS.json hbm_gbs (measured, burst)"
    else:
        peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    kern_s = float(np.mean([s["ms_kernels"] for s in stats])) / 1000.0
    launches = int(np.mean([s["synth_launches_run"] for s in stats])) + 2   # pass launches that did work + the two pass-0 gathers
    achieved = algo / kern_s / 1e9
    traffic = bound = traffic_src = limiter = None
    tp_path = os.path.join(ROOT, "profiles", "traffic_%s.json" % wname)
    if os.path.exists(tp_path):   # one `ncu --set full` capture of every synthesis launch of one job (tools/ncu_traffic.py)
        t = json.load(open(tp_path))
        if "dram_bytes_per_job" in t:
            launches = int(t.get("launches_captured", launches))   # every launch that did work, as ncu saw them (incl. k_gather_later)
            traffic = t["dram_bytes_per_job"] / launches
            traffic_src = "%s: dram__bytes_read.sum + dram__bytes_write.sum over the %d synthesis launches of one job / %d" % (
                os.path.basename(tp_path), t.get("launches_captured", launches), launches)
        bound = t.get("bound")
        limiter = t.get("limiter")
    rec = {"bound": bound or "hbm", "limiter": limiter, "kernel": "k_synth_pass / k_synth_pass_team (all pass launches + pass-0 patch gather of one job)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
           "traffic_source": traffic_src, "hbm_frac_of_traffic": (traffic * launches / kern_s / 1e9 / peak) if traffic else None,
           "peak_source": peak_src, "counts_source": src,
           "algorithmic_bytes_per_job": algo, "algorithmic_bytes_per_launch": algo / launches,
           "ms_kernels_per_job": 1000.0 * kern_s, "avg_launch_ms": 1000.0 * kern_s / launches, "launches_per_job": launches,
           "counts": {k: counts[k] for k in ("visits", "evals", "compares", "offset_scans", "heur_evals", "heur_skips")},
           "compares_issued": dev["compares"],
           "formula": "compares*bpp + offset_scans*9 + visits*(K*(bpp+8) + n_color+9) + (heur_evals+heur_skips)*8 + "
                      "(evals-heur_evals)*8, bpp=%d K=%d; achieved = that / ms_kernels" % (w["bpp"], max(2, w["params"].patchSize))}
    elem = 8 if w["n_map"] else 4
    nbytes = w["cor"].shape[0] * w["cor"].shape[1] * elem
    if gather and nbytes <= (32 << 20):
        # L2-resident corpora: the access pattern's own ceiling -- random corpus-pixel gathers from a corpus-sized buffer,
        # measured now.  (Not reported for corpora beyond the L2: there the neighbour compares of a patch share sectors
        # that a uniformly random probe stream does not, and DRAM traffic is what binds -- see `traffic`.)
        g = api.gather_rate(nbytes, elem, 7)
        rec["l2_gather"] = {"loads_per_s": g, "compares_per_s": counts["compares"] / kern_s,
                            "frac": counts["compares"] / kern_s / g,
                            "what": "uniformly random %d-byte loads from a %d-byte buffer (rs_cuda_gather_rate: the throughput "
                                    "kernel's launch shape and L1 carve-out, >= 20 ms launches, median of 7): the access "
                                    "pattern of one neighbour compare" % (elem, nbytes)}
    return rec


class Runner:
    """One workload through the public API, with host buffers, L2 flushed before every job.  The caller's buffers are
    made once; after a job only the rows and columns that hold target points are put back from a pristine copy (the
    engine changes nothing else), so that the bench itself moves as little host memory as possible between the timed
    calls -- at 8 ranks on one host, copying whole 64 MB images per step would compete with the other ranks' staging."""

    def __init__(self, api, torch, w, fi, pinned=True):
        """pinned: the caller's buffers are page-locked (the bench contract's "inputs in pinned host memory"): the engine
        hands them to the copy engine as they are; False: malloc'ed numpy buffers like the reference's callers'
        (lib/imageBuffer.h), which the engine stages through its own pinned workspace."""
        self.api, self.torch, self.w, self.fi = api, torch, w, fi
        self.pinned, self._keep = pinned, []
        self.n = int((w["tmask"] != 0).sum())
        self.tmask = self._host(w["tmask"])
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        rows = np.flatnonzero(w["tmask"].any(axis=1))
        cols = np.flatnonzero(w["tmask"].any(axis=0))
        self.box = (slice(int(rows[0]), int(rows[-1]) + 1), slice(int(cols[0]), int(cols[-1]) + 1))
        n_rows = int(rows[-1] - rows[0] + 1)
        n_cols = int(cols[-1] - cols[0] + 1)
        if "simple" in w:
            self.img = self._host(w["tgt"])
            self.pristine = w["tgt"][self.box].copy()
            self.h2d = w["tgt"].nbytes + w["tmask"].nbytes      # image + mask planes (the PRNG stream of the order is made on the device)
            # what comes back: the box that holds target points (page-locked buffers), their whole rows (staged)
            self.d2h = n_rows * (n_cols if pinned else w["tmask"].shape[1]) * (w["bpp"] - 1)
        else:
            self.tp, self.cp = (self._host(x) for x in pixmaps(w))
            self.pristine = self.tp[self.box].copy()
            self.h2d = self.tp.nbytes + self.cp.nbytes
            self.d2h = n_rows * (n_cols if pinned else w["tmask"].shape[1]) * w["bpp"]

    def _host(self, a):
        """A private copy of `a` for the calls: page-locked (a numpy view of a pinned torch tensor) or malloc'ed."""
        if not self.pinned:
            return a.copy()
        t = self.torch.empty(a.shape, dtype=self.torch.uint8, pin_memory=True)
        self._keep.append(t)
        v = t.numpy()
        v[...] = a
        return v

    def step(self):
        api, w = self.api, self.w
        api.set_seed(SEED)
        if "simple" in w:
            self.img[self.box] = self.pristine
            self.flush.fill_(1)
            self.torch.cuda.synchronize()
            t0 = time.perf_counter()
            err = api.image_synth(self.img, self.tmask, w["simple"], w["params"])
            wall = time.perf_counter() - t0
        else:
            self.tp[self.box] = self.pristine
            self.flush.fill_(1)
            self.torch.cuda.synchronize()
            t0 = time.perf_counter()
            err = api.engine(w["params"], self.fi, self.tp, self.cp)
            wall = time.perf_counter() - t0
        assert err == 0
        return wall, api.last_stats()


def sub_record(api, torch, wname, steps, warmup):
    """A short run of another BASELINE configuration, same definitions as the headline."""
    w = workload(wname)
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    r = Runner(api, torch, w, fi)
    for _ in range(warmup):
        r.step()
    walls, stats = [], []
    for _ in range(steps):
        wall, st = r.step()
        walls.append(wall); stats.append(st)
    kern_s = sum(s["ms_kernels"] for s in stats) / 1000.0
    visits = stats[-1]["visits"]; evals = stats[-1]["evals"]
    rf = roofline_of(w, wname, stats, steps, api, gather=False)
    return {"workload": w["name"], "api": api_name(w), "steps": steps, "warmup": warmup,
            "value": steps * r.n / kern_s, "unit": UNIT, "e2e": steps * r.n / sum(walls),
            "ms_kernels": 1000.0 * kern_s / steps, "ms_e2e": 1000.0 * sum(walls) / steps,
            "ms_e2e_steps": [round(1000.0 * x, 3) for x in walls],
            "ms_prep": float(np.mean([s["ms_prep"] for s in stats])), "ms_h2d": float(np.mean([s["ms_h2d"] for s in stats])),
            "ms_d2h": float(np.mean([s["ms_d2h"] for s in stats])),
            "ms_pass": [float(np.mean([s["ms_pass"][p] for s in stats])) for p in range(6)],
            "passes_run": stats[-1]["passes_run"], "visits_per_px": visits / r.n, "evals_per_visit": evals / max(visits, 1),
            "visits_per_s": steps * visits / kern_s, "evals_per_s": steps * evals / kern_s,
            "roofline": {k: rf[k] for k in ("bound", "achieved", "peak", "frac", "traffic", "counts_source", "algorithmic_bytes_per_job")}}


def cfg5_batch_record(api, torch, dist, rank, local, world, n_jobs, probe_list, slots):
    """BASELINE.json config 5: a FIXED batch of `n_jobs` independent 2048x2048 heal jobs (images G(2048,2048,3,100+k),
    the same centred 256x256 hole) dealt round robin over the ranks (strong scaling), once per probe count.  Every rank
    runs its share through rs_image_synth_batch() -- host buffers in, healed images out -- and the time of a sweep
    point is the slowest rank's wall clock between two barriers."""
    from concurrent.futures import ThreadPoolExecutor
    mine = sharding.deal_round_robin(n_jobs, world, rank)
    m = centered_mask(2048, 2048, 256, 256)
    with ThreadPoolExecutor(4) as ex:
        pristine = list(ex.map(lambda k: G(2048, 2048, 3, 100 + k), mine))
    # the caller's images and mask in page-locked memory (one allocation, a view per job): copied from and to directly
    pin = torch.empty((max(len(mine), 1), 2048, 2048, 3), dtype=torch.uint8, pin_memory=True).numpy()
    pin_m = torch.empty((2048, 2048), dtype=torch.uint8, pin_memory=True).numpy()
    pin_m[...] = m
    work = [pin[i] for i in range(len(mine))]
    masks = [pin_m] * len(mine)
    n_px = int((m != 0).sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    api.order_cache(True)   # same selection in every job of the batch: its target points are ordered once per GPU
    out = {}
    for probes in probe_list:
        prm = abi.default_params()
        prm.maxProbeCount = probes
        for rep in range(2):   # rep 0: warm-up (workspaces, order cache), rep 1: timed
            for dst, src in zip(work, pristine):
                np.copyto(dst, src)
            barrier()
            t0 = time.perf_counter()
            errs = api.image_synth_batch(work, masks, abi.T_RGB, prm, devices=[local], slots=slots) if work else []
            t_rank = time.perf_counter() - t0
            barrier()
            t_all = time.perf_counter() - t0
            assert not any(errs)
        t_all, t_rank = sharding.max_over_ranks([t_all, t_rank], device="cuda")
        changed = all((wk != pr).any() for wk, pr in zip(work, pristine))
        ok = sharding.sum_over_ranks([0.0 if changed else 1.0], device="cuda")[0] == 0.0
        out[str(probes)] = {"px_per_s": n_jobs * n_px / t_all, "ms_per_job": 1000.0 * t_all / n_jobs,
                            "ms_batch": 1000.0 * t_all, "every_image_healed": bool(ok)}
    api.order_cache(False)
    return {"workload": "cfg5: %d heal jobs 2048x2048 RGB, 256x256 hole, ctx1, patch 30, dealt round robin over %d GPU(s)" % (n_jobs, world),
            "api": "rs_image_synth_batch() = imageSynth() per image; page-locked host buffers in and out; visit-order cache on",
            "jobs": n_jobs, "jobs_per_rank": [len(sharding.deal_round_robin(n_jobs, world, r)) for r in range(world)],
            "slots_per_gpu": slots, "scaling": "strong", "unit": UNIT, "timing": "wall clock between barriers, max over ranks",
            "h2d_bytes_per_job": int(pristine[0].nbytes + m.nbytes) if pristine else 0,
            "d2h_bytes_per_job": 256 * 2048 * 3, "probes": out}


def shared_corpus_record(api, n_jobs=32):
    """One corpus, many targets: `n_jobs` 256x256 targets textured from ONE 2048x2048 corpus
    (ctx 0, 9 neighbours, 200 probes) -- as one rs_engine_batch() call, whose jobs share the device-resident corpus,
    against the loop over engine() the reference's callers write."""
    cor = G(2048, 2048, 3, 77)
    cp = np.ascontiguousarray(np.concatenate([np.full((2048, 2048, 1), 255, np.uint8), cor], axis=2))
    prm = abi.make_params(0, 0, 0, 0.5, 0.117, 9, 200)
    fi = api.format_indices(3)

    def targets():
        return [np.ascontiguousarray(np.concatenate([np.full((256, 256, 1), 255, np.uint8), np.full((256, 256, 3), 255, np.uint8)], axis=2))
                for _ in range(n_jobs)]
    api.order_cache(True)
    out = {}
    for rep in range(2):
        tps = targets()
        t0 = time.perf_counter()
        for tp in tps:
            assert api.engine(prm, fi, tp, cp) == 0
        out["loop_ms_per_job"] = 1000.0 * (time.perf_counter() - t0) / n_jobs
    loop_first = tps[0].copy()
    b0 = api.shared_corpus_stats()
    for rep in range(2):
        tps = targets()
        t0 = time.perf_counter()
        errs = api.engine_batch([(prm, fi, tp, cp) for tp in tps], slots=4)
        out["batch_ms_per_job"] = 1000.0 * (time.perf_counter() - t0) / n_jobs
        assert not any(errs)
    b1 = api.shared_corpus_stats()
    api.order_cache(False)
    out.update(workload="%d targets 256x256 from one 2048x2048 corpus, ctx0, patch 9, probes 200" % n_jobs,
               speedup=out["loop_ms_per_job"] / out["batch_ms_per_job"], same_result_as_loop=bool((tps[0] == loop_first).all()),
               corpora_built=b1[0] - b0[0], corpus_reuses=b1[1] - b0[1], h2d_bytes_saved_per_job=int(cp.nbytes))
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from resynthesizer_b200 import api, build

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path")
    torch.cuda.set_device(local)
    api.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = workload(a.workload)   # the same job on every rank: weak scaling compares like with like
    fi = api.format_indices(w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    runner = Runner(api, torch, w, fi)
    n = runner.n

    # The device-side cache of visit orders (a job with the same selection, size, context type and seed reuses the
    # order of the first one) is OFF for the headline numbers: every timed job orders its target points itself.
    api.order_cache(False)
    for i in range(a.warmup):
        runner.step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = api.total_kernel_launches()
    t_begin = time.perf_counter()
    walls, stats = [], []
    for i in range(a.steps):
        wall, st = runner.step()
        walls.append(wall); stats.append(st)
    barrier()
    t_total = time.perf_counter() - t_begin
    n_launches = api.total_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # the same call from malloc'ed buffers, as the reference's callers hold them (lib/imageBuffer.h): the engine stages
    # them through its own pinned workspace with host threads (rank 0's figure is reported; every rank runs it so that the
    # staging threads compete for the host as they would)
    pageable_walls = []
    if not a.quick:
        pr = Runner(api, torch, w, fi, pinned=False)
        pr.step()
        barrier()
        for i in range(max(3, a.steps // 4)):
            pageable_walls.append(pr.step()[0])
        barrier()
        del pr

    # the same job stream with the order cache on (what a batch of same-shaped jobs or the frames of a clip see)
    cached_walls = []
    if not a.quick:
        api.order_cache(True)
        runner.step()
        for i in range(max(3, a.steps // 4)):
            cached_walls.append(runner.step()[0])
        api.order_cache(False)

    kern_s = sum(s["ms_kernels"] for s in stats) / 1000.0
    e2e_s = sum(walls)
    # the only communication of the multi-GPU path: max over ranks of three timing scalars
    kern_s, e2e_s, t_total = sharding.max_over_ranks([kern_s, e2e_s, t_total], device="cuda")
    total_px = world * a.steps * n
    evals = sum(s["evals"] for s in stats); issued = sum(s["evals_issued"] for s in stats)
    compares = sum(s["compares"] for s in stats); visits = sum(s["visits"] for s in stats)

    # BASELINE config 5, dealt over the ranks (all ranks take part), and the other configurations (rank 0, one GPU)
    batch = None
    if not a.quick and a.cfg5_jobs > 0:
        batch = cfg5_batch_record(api, torch, dist, rank, local, world, a.cfg5_jobs,
                                  [int(p) for p in a.cfg5_probes.split(",") if p], a.slots)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    subs = {}
    if not a.quick and world == 1:
        names = ["cfg1", "cfg2", "cfg4", "cfg5"]
        if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "brick_512.npy")):
            names.insert(1, "cfg1_brick")
        for name in names:
            if name != a.workload:
                subs[name] = sub_record(api, torch, name, 5, 2)
        if a.workload != DEFAULT_WORKLOAD:
            subs[DEFAULT_WORKLOAD] = sub_record(api, torch, DEFAULT_WORKLOAD, 5, 2)
        subs["shared_corpus_batch"] = shared_corpus_record(api)

    roofline = roofline_of(w, a.workload, stats, a.steps, api)

    # ---- CPU baseline beside it: the compiled reference on a bounded sample of the same workload (rank 0, N = 1)
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rand_1t.so")):
            scale = reference_sample_scale(a.workload)
            r1 = cpu_reference_run(a.workload, scale, 1, 0, 1, "ref_rand_1t")
            r8 = cpu_reference_run(a.workload, scale, 1, 0, 1, "ref_rand_8t")
            ctr = reference_counters(a.workload, scale)
            best, cores, which = (r1, 1, "unthreaded") if r1["px_per_s"] >= r8["px_per_s"] else (r8, 8, "8-thread refiner")
            cpu = {"value": best["px_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%s, one complete job, reference %s build; unthreaded %.0f px/s, threaded(8) %.0f px/s" %
                             (workload(a.workload, 0, scale)["name"], which, r1["px_per_s"], r8["px_per_s"]),
                   "passes_run": ctr["passes_run"], "visits_per_px": ctr["visits"] / best["n"],
                   "visits_per_s": ctr["visits"] / best["per_job_s"], "evals_per_s": ctr["evals"] / best["per_job_s"]}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    line = {"metric": METRIC, "value": total_px / kern_s, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * t_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_of(w),
            "parallelism": "independent jobs, one job stream per GPU, %d GPU(s), no collective on the data path" % world,
            "l2": "256 MiB flush before every job",
            # equal-work keys (also printed by --impl reference)
            "passes_run": stats[-1]["passes_run"], "visits_per_px": visits / a.steps / n,
            "evals_per_visit": evals / max(visits, 1),
            "visits_per_step": world * visits / a.steps, "evals_per_step": world * evals / a.steps,
            "visits_per_s": world * visits / kern_s, "evals_per_s": world * evals / kern_s,
            "evals_issued_per_s": world * issued / kern_s, "compares_per_s": world * compares / kern_s,
            "compares_per_eval_issued": compares / max(issued, 1),
            "e2e": {"value": total_px / e2e_s, "unit": UNIT,
                    "host_buffers": "page-locked (pinned torch tensors handed to the C-ABI call; copied from and to directly)",
                    "h2d_bytes_per_step": int(runner.h2d), "d2h_bytes_per_step": int(runner.d2h),
                    "visits_per_s": world * visits / e2e_s, "evals_per_s": world * evals / e2e_s,
                    "ms_call": 1000.0 * e2e_s / a.steps,
                    "ms_prep": float(np.mean([s["ms_prep"] for s in stats])), "ms_h2d": float(np.mean([s["ms_h2d"] for s in stats])),
                    "ms_kernels": float(np.mean([s["ms_kernels"] for s in stats])), "ms_d2h": float(np.mean([s["ms_d2h"] for s in stats]))},
            "e2e_pageable": ({"value": world * n / float(np.mean(pageable_walls)), "unit": UNIT,
                              "ms_call": 1000.0 * float(np.mean(pageable_walls)),
                              "what": "same call with malloc'ed caller buffers (rank 0's mean over %d jobs x %d ranks): staged "
                                      "through the workspace's pinned memory by host threads" % (len(pageable_walls), world)}
                             if pageable_walls else None),
            "ms_pass": [float(np.mean([s["ms_pass"][p] for s in stats])) for p in range(6)],
            "e2e_order_cached": ({"value": n / float(np.mean(cached_walls)), "unit": UNIT,
                                  "what": "same call with the visit-order cache on (rank 0, %d jobs): the target order of an "
                                          "identical selection is reused from the device" % len(cached_walls)}
                                 if cached_walls else None),
            "gpu_launches": int(n_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "configs": subs or None, "cfg5_batch": batch}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only: no sub-records, no cfg5 batch, no cached-order leg")
    ap.add_argument("--probes", type=int, default=0, help="override maxProbeCount of the headline workload")
    ap.add_argument("--cfg5-jobs", type=int, default=64, help="jobs of the cfg5 batch record (0 = skip)")
    ap.add_argument("--cfg5-probes", default="50,100,200,500,1000", help="probe counts of the cfg5 sweep")
    ap.add_argument("--slots", type=int, default=4, help="jobs in flight per GPU in the cfg5 batch")
    a = ap.parse_args()
    global PROBES_OVERRIDE
    PROBES_OVERRIDE = a.probes
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
