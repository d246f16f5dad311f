#!/usr/bin/env python
"""bench.py -- synthesized target px/s (and patch-distance evals/s) of the synthesis hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N ...            # the reference's CPU engine (oracle/_ref)

One "step" = one complete synthesis job (all passes) of the workload on each rank.  Default workload is
BASELINE.json configs[1]: render-texture, a 1024x1024 target synthesised from a 256x256 synthetic corpus
tile, no context matching (matchContextType 0, no tiling), the render-texture script's 9 neighbours / 200
probes (PluginScripts/plugin-render-texture.py:175), through the full API engine().

  value  = target px/s with inputs resident in HBM: n_targets / CUDA-event time from "upload complete" to
           "last pass done" on the job's stream (whole job: N ranks x K steps, max over ranks).
  e2e    = the same metric through the reference-facing C-ABI call engine() with HOST buffers: host prep,
           H2D, passes, D2H, write-back all inside the timed region.
Multi-GPU: independent jobs per rank (a job does not shard), no data-path collective, "scaling": "weak".
"""
import argparse
import ctypes as C
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
def _ref_worker(args):
    libname, wname, scale, seed_shift, steps = args
    from oracle import refdriver as R
    lib = R.load(libname)
    w = workload(wname, seed_shift, scale)
    fi = R.format_indices(lib, w["n_color"], w["n_map"], w["alpha"], w["alpha"], w["n_map"] > 0)
    times = []
    n = int((w["tmask"] != 0).sum())
    for _ in range(steps):
        tp, cp = pixmaps(w)
        t0 = time.perf_counter()
        err = R.engine(lib, w["params"], fi, tp, cp)
        times.append(time.perf_counter() - t0)
        assert err == 0
    return n, times


def cpu_reference_run(wname, scale, steps, warmup, procs, libname="ref_rand_1t"):
    """Runs `procs` independent sample jobs per step on the host cores with the compiled reference."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        if warmup:
            pool.map(_ref_worker, [(libname, wname, scale, i, 1) for i in range(procs)])
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(libname, wname, scale, i, steps) for i in range(procs)])
        wall = time.perf_counter() - t0
    n = res[0][0]
    return dict(px_per_s=procs * steps * n / wall, wall=wall, n=n, per_job_s=float(np.mean([np.mean(r[1]) for r in res])))


def reference_sample_scale(wname):
    # bounded samples of the same workload: 10-30 s of CPU work in all (unthreaded + 8-thread build, or one job per core)
    return {"cfg2": 0.5, "cfg1": 1.0, "cfg5": 0.25, "cfg3": 0.0625, "cfg4": 0.125}.get(wname, 1.0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rand_1t.so")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    procs = max(1, min(os.cpu_count() or 1, 64))
    scale = reference_sample_scale(a.workload)
    w = workload(a.workload, 0, scale)
    r = cpu_reference_run(a.workload, scale, a.steps, min(a.warmup, 1), procs)
    sample = "%s; %d independent jobs (one per host core) x %d steps, unthreaded reference build (libref_rand_1t)" % (w["name"], procs, a.steps)
    line = {"metric": METRIC, "value": r["px_per_s"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1000.0 * r["wall"] / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer",
            "data": "synthetic", "config": {"workload": workload(a.workload)["name"], "sample": sample},
            "cpu_baseline": {"value": r["px_per_s"], "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": r["px_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "per_job_seconds": r["per_job_s"]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm

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
    n = int((w["tmask"] != 0).sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    B = max(1, a.batch)

    def one_step(seed):
        api.set_seed(seed)
        if B == 1 and "simple" in w:
            # the heal configurations are imageSynth() jobs (BASELINE.json configs 1, 3, 5): one image + mask, host buffers
            img = w["tgt"].copy()
            flush.fill_(seed & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            err = api.image_synth(img, w["tmask"], w["simple"], w["params"])
            wall = time.perf_counter() - t0
            assert err == 0
            return wall, api.last_stats(), img, w["tmask"]
        if B == 1:
            tp, cp = pixmaps(w)
            flush.fill_(seed & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            err = api.engine(w["params"], fi, tp, cp)
            wall = time.perf_counter() - t0
            assert err == 0
            return wall, api.last_stats(), tp, cp
        # a step = a batch of B independent jobs kept `slots` at a time on this GPU (rs_engine_batch)
        jobs = [(w["params"], fi) + pixmaps(w) for _ in range(B)]
        flush.fill_(seed & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        errs = api.engine_batch(jobs, a.slots)
        wall = time.perf_counter() - t0
        assert not any(errs)
        st = {k: 0 for k in ("evals", "evals_issued", "compares", "visits", "offset_scans")}
        st.update(ms_kernels=wall * 1000.0, ms_synth=wall * 1000.0, ms_prep=0.0, ms_h2d=0.0, ms_d2h=0.0, passes_run=0,
                  ms_pass=[0.0] * 6, kernel_launches=0, synth_launches_run=0,
                  n_corpus=int((w["cmask"] == 255).sum()))
        return wall, st, jobs[0][2], jobs[0][3]

    # The reference seeds its PRNG with the same constant in every engine() call (lib/engine.c:643); so does every step.
    # The device-side cache of visit orders (a job with the same selection, size, context type and seed reuses the
    # order of the first one) is OFF for the headline numbers: every timed job orders its target points itself.
    SEED = 1198472
    api.order_cache(B > 1)   # batch mode IS the cache's use case (same selection in every job of the batch): on, and said so
    for i in range(a.warmup):
        one_step(SEED)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = api.total_kernel_launches()
    t_begin = time.perf_counter()
    walls, stats = [], []
    h2d = d2h = 0
    for i in range(a.steps):
        wall, st, tp, cp = one_step(SEED)
        walls.append(wall); stats.append(st)
        h2d = tp.nbytes + cp.nbytes + 4 * n      # both pixmaps (or image + mask) and the visit order
        d2h = int((np.flatnonzero(w["tmask"].any(axis=1))[[0, -1]] * [-1, 1]).sum() + 1) * w["tmask"].shape[1] * (
            (w["bpp"] - 1) if "simple" in w and B == 1 else w["bpp"])   # the rows that hold target points
    barrier()
    t_total = time.perf_counter() - t_begin
    n_launches = api.total_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # the same job stream with the order cache on (what a batch of same-shaped jobs or the frames of a clip see)
    cached_walls = []
    if B == 1:
        api.order_cache(True)
        one_step(SEED)
        for i in range(max(3, a.steps // 4)):
            cached_walls.append(one_step(SEED)[0])
        api.order_cache(False)

    kern_s = sum(s["ms_kernels"] for s in stats) / 1000.0
    e2e_s = sum(walls)
    # the only communication of the multi-GPU path: max over ranks of three timing scalars
    kern_s, e2e_s, t_total = sharding.max_over_ranks([kern_s, e2e_s, t_total], device="cuda")
    total_px = world * a.steps * n * B
    evals = sum(s["evals"] for s in stats); issued = sum(s["evals_issued"] for s in stats)
    compares = sum(s["compares"] for s in stats); visits = sum(s["visits"] for s in stats)
    scans = sum(s["offset_scans"] for s in stats)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_synth_pass / k_synth_pass_team): algorithmic bytes per SURVEY.md section 8d.
    # Duration = CUDA events on the job's stream around the pass kernels alone; launches = the pass-kernel launches
    # that did work (passes after the 10 % stop rule exit immediately and are not counted).
    bpp = w["bpp"]
    K = max(2, w["params"].patchSize)
    P = w["params"].maxProbeCount
    fixed = scans * (4 + 4) + visits * (K * (8 + (4 if w["n_map"] else 0)) + K * 8 + P * 4 + 8 + 4)
    algo_bytes = compares * bpp + fixed
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    synth_s = sum(s["ms_synth"] for s in stats) / 1000.0
    launches_pass = max(1, sum(s["synth_launches_run"] for s in stats))
    achieved = algo_bytes / synth_s / 1e9
    traffic = None
    tp_path = os.path.join(ROOT, "profiles", "traffic_%s.json" % a.workload)
    if os.path.exists(tp_path):
        traffic = json.load(open(tp_path)).get("dram_bytes_per_launch")
    # the access pattern's own ceiling: random corpus-pixel gathers from a corpus-sized buffer, measured now
    elem = 8 if w["n_map"] else 4
    gather = api.gather_rate(w["cor"].shape[0] * w["cor"].shape[1] * elem, elem)
    roofline = {"bound": "hbm", "kernel": "k_synth_pass(+_team)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes / launches_pass,
                "avg_launch_ms": 1000.0 * synth_s / launches_pass, "launches": launches_pass,
                "gather_ceiling": {"loads_per_s": gather, "compares_per_s": compares / synth_s,
                                   "frac": compares / synth_s / gather,
                                   "what": "uniformly random %d-byte loads from a %d-byte buffer (rs_cuda_gather_rate), "
                                           "the access pattern of one neighbour compare" % (elem, w["cor"].shape[0] * w["cor"].shape[1] * elem)},
                "note": "bytes = neighbour-compares x %d B corpus pixel + per-visit fixed part; the working set is "
                        "L2-resident, so DRAM traffic is far below the algorithmic bytes and the HBM fraction is small "
                        "by construction; the gather ceiling is the bound that applies" % bpp}

    # ---- CPU baseline beside it: the compiled reference on a bounded sample of the same workload
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rand_1t.so")):
            scale = reference_sample_scale(a.workload)
            r1 = cpu_reference_run(a.workload, scale, 1, 0, 1, "ref_rand_1t")
            r8 = cpu_reference_run(a.workload, scale, 1, 0, 1, "ref_rand_8t")
            best, cores, which = (r1, 1, "unthreaded") if r1["px_per_s"] >= r8["px_per_s"] else (r8, 8, "8-thread refiner")
            cpu = {"value": best["px_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%s, one job, reference %s build; unthreaded %.0f px/s, threaded(8) %.0f px/s" %
                             (workload(a.workload, 0, scale)["name"], which, r1["px_per_s"], r8["px_per_s"])}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    line = {"metric": METRIC, "value": total_px / kern_s, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * t_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
            "config": {"workload": w["name"] + (" x %d jobs per step, %d in flight, visit-order cache on" % (B, a.slots) if B > 1 else ""),
                       "parallelism": "independent jobs, %d GPU(s)" % world,
                       "l2": "256 MiB flush between steps",
                       "api": "imageSynth() simple API" if ("simple" in w and B == 1) else "engine() full API"},
            "evals_per_s": world * evals / kern_s, "evals_issued_per_s": world * issued / kern_s,
            "compares_per_s": world * compares / kern_s, "compares_per_eval_issued": compares / max(issued, 1),
            "passes_run": stats[-1]["passes_run"], "visits_per_step": visits / a.steps,
            "e2e": {"value": total_px / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_prep": float(np.mean([s["ms_prep"] for s in stats])), "ms_h2d": float(np.mean([s["ms_h2d"] for s in stats])),
                    "ms_kernels": float(np.mean([s["ms_kernels"] for s in stats])), "ms_d2h": float(np.mean([s["ms_d2h"] for s in stats]))},
            "ms_pass": [float(np.mean([s["ms_pass"][p] for s in stats])) for p in range(6)],
            "e2e_order_cached": ({"value": n / float(np.mean(cached_walls)), "unit": UNIT,
                                  "what": "same call with the visit-order cache on (rank 0, %d jobs): the target order of an "
                                          "identical selection is reused from the device" % len(cached_walls)}
                                 if cached_walls else None),
            "gpu_launches": int(n_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--probes", type=int, default=0, help="override maxProbeCount (the cfg5 sweep of BASELINE.json: 50..1000)")
    ap.add_argument("--batch", type=int, default=1, help="jobs per step (rs_engine_batch); value is then wall-clock based")
    ap.add_argument("--slots", type=int, default=8, help="jobs in flight per GPU in batch mode")
    a = ap.parse_args()
    global PROBES_OVERRIDE
    PROBES_OVERRIDE = a.probes
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
