#!/usr/bin/env python
"""bench.py -- agent-steps/s of the Sugarscape timestep hot path on synthetic NxN lattices.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|<n>] [--impl reference]

One "step" = one whole timestep (View.updateGame, sugarscape.py:506-686) over the workload.
Workloads:
  c5 (default)  16384 x 16384 tiled two-peak landscape, 16,777,216 agents, vision 1-6,
                grow + moveEat + canStarve, keyed RNG seed 18        <- the metric's config
  c4            4096 x 4096, 1,048,576 agents, same rules + pollution (A=B=1, diffusion every step)
  c5u / c4u     the c5 / c4 worlds under the utilicalc rule instead of moveEat (no pollution)
The JSON line carries: value (state resident in HBM, device-timed with CUDA events on the
engine's stream), e2e (through the public Python API with host buffers: upload + K steps
with per-step stats read-back + download), roofline of the dominant kernel class measured
live with CUDA events in a second instrumented run, cpu_baseline (the reference's Python loop
on a bounded sample of the same landscape, with the C oracle port beside it), `secondary`
(short runs of c4 and c5u with their per-kernel rooflines), a decomposition-independent
digest of the final state and the GPU clocks sampled during the timed region.
--gpus N (under torchrun): the same workload as row strips, one rank per GPU
(sugarscape_utilicalc_b200/strips.py).

--impl reference times the reference's own pure-Python loop (the unmodified sources staged under
baseline/_ref by __graft_entry__.build(); headless, one core -- the loop is sequential) on the
shipped presets and on the workload's landscape at 200^2 and 800^2, with the C port
(oracle/ss_oracle.c) beside it.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.clocks import ClockSampler  # noqa: E402

HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def kernels_sha():
    """Hash of the CUDA sources' code (// comments and white space dropped): a DRAM-traffic figure in profiles/traffic.json
    counts only for the kernels it was captured on."""
    import hashlib
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sugarscape_utilicalc_b200", "csrc")
    for f in ("ss_device.cuh", "ss_world.cuh", "ss_lattice.cu", "ss_rounds.cu"):
        with open(os.path.join(d, f), "r") as fh:
            code = "".join(re.sub(r"//.*$", "", line) for line in fh.read().splitlines())
        h.update(re.sub(r"\s+", "", code).encode())
    return h.hexdigest()[:16]


def recorded_traffic(workload_key, kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum of the class's launches of one step, from the committed `ncu --set full`
    capture (profiles/traffic.json, written by tools/ncu_traffic.py); None when there is none for these kernel sources."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t["entries"][workload_key][kernel_class]
        if t.get("kernels_sha") != kernels_sha():
            return None, "profiles/traffic.json was captured on other kernel sources (sha %s): stale" % t.get("kernels_sha")
        return float(e["dram_bytes_per_step"]), e["source"]
    except Exception:
        return None, None


def workload(name):
    if name == "c5":
        return dict(key="c5", n=16384, pollution=False, label="c5: 16384x16384 tiled landscape, 16777216 agents, "
                    "grow+moveEat+canStarve, vision 1-6, keyed RNG seed 18")
    if name == "c4":
        return dict(key="c4", n=4096, pollution=True, label="c4: 4096x4096 tiled landscape, 1048576 agents, "
                    "grow+moveEat+canStarve+pollution(A=B=1, diffusion every step), vision 1-6, keyed RNG seed 18")
    if name in ("c5u", "c4u"):      # the same worlds under the utilicalc rule (agent.py:1025-1141) instead of moveEat
        n = 16384 if name == "c5u" else 4096
        return dict(key=name, n=n, pollution=False, utilicalc=True,
                    label="%s: %dx%d tiled landscape, %d agents, grow+utilicalc+canStarve, vision 1-6, keyed RNG seed 18"
                    % (name, n, n, n * n // 16))
    if name.startswith("preset:"):
        return dict(preset=name.split(":", 1)[1], n=50, pollution=False, label="shipped preset " + name.split(":", 1)[1])
    n = int(name)
    return dict(n=n, pollution=False, label="synthetic %dx%d tiled landscape, %d agents, grow+moveEat+canStarve"
                % (n, n, n * n // 16))


def make_world(wl):
    t0 = time.time()
    w = synthetic.tiled_world(wl["n"])
    s = synthetic.synthetic_settings(wl["n"], pollution=wl["pollution"], utilicalc=wl.get("utilicalc", False))
    sys.stderr.write("[bench] world %dx%d, %d agents generated in %.1fs\n" % (wl["n"], wl["n"], len(w["id"]),
                                                                               time.time() - t0))
    return s, w


ENGINE_OPTS = {}


def upload(world, w):
    for k, v in ENGINE_OPTS.items():
        world.set_option(k, v)
    world.upload_cells(w["sugar"], w["cap"], w["pollution"])
    world.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])


def oracle_sample(settings, w, n_sample, steps, warmup=0):
    """The C oracle (port of the reference loop) on the top-left n_sample^2 sub-lattice of the world."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import OracleWorld
    n = w["cap"].shape[0]
    ns = min(n_sample, n)
    keep = (w["x"] < ns) & (w["y"] < ns)
    s = json.loads(json.dumps(settings))
    s["view"]["grid_size"] = {"x": ns, "y": ns}
    cfg = config_from_settings(s, rng_mode="keyed")
    o = OracleWorld(cfg)
    o.upload_cells(np.ascontiguousarray(w["sugar"][:ns, :ns]), np.ascontiguousarray(w["cap"][:ns, :ns]),
                   np.ascontiguousarray(w["pollution"][:ns, :ns]))
    k = int(keep.sum())
    o.upload_agents(np.arange(1, k + 1, dtype=np.int32), w["x"][keep], w["y"][keep], w["vision"][keep],
                    w["metab"][keep], w["sugar_agent"][keep])
    pop = k
    if warmup:
        pop = int(o.step(warmup)["population"][-1])
    t0 = time.perf_counter()
    st = o.step(steps)
    dt = time.perf_counter() - t0
    pops = np.concatenate([[pop], st["population"][:-1]])
    o.close()
    return dict(agent_steps=float(pops.sum()), seconds=dt, n=ns, agents=k, steps=steps)


def _replica_worker(task):
    """One of P independent replicas of the sequential loop (its own 1024^2 world), for the box-throughput figure."""
    wl, seed, steps = task
    w = synthetic.tiled_world(1024, seed=1234 + seed)
    s = synthetic.synthetic_settings(1024, pollution=wl["pollution"], utilicalc=wl.get("utilicalc", False))
    r = oracle_sample(s, w, 1024, steps, 1)
    return r["agent_steps"], r["seconds"]


def replicas_throughput(wl, steps=3):
    """The reference's loop is sequential, so its faithful figure is one core; this is the aggregate of one independent
    replica per host core, reported beside it and labelled as such."""
    import multiprocessing as mp
    procs = os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_replica_worker, [(wl, k, steps) for k in range(procs)])
    return {"processes": procs, "value": sum(a for a, _ in res) / max(t for _, t in res), "unit": "agent-steps/s",
            "what": "aggregate of %d independent replicas (1024x1024 each, %d steps after 1 warm-up), one per host core; "
                    "not a parallel run of one world" % (procs, steps), "wall_s": time.perf_counter() - t0}


REF_STAGE = os.path.join(ROOT, "baseline", "_ref")


def python_reference_figures(args, wl):
    """The reference's own pure-Python loop (View.updateGame of the UNMODIFIED sources staged under baseline/_ref by
    __graft_entry__.build()), headless (Tk / matplotlib stubbed; oracle/ref_harness.py), one core -- the loop is
    sequential.  Shipped presets in full, and the bench workload's landscape and rules at 200^2 and 800^2."""
    os.environ["SS_REFERENCE"] = REF_STAGE
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness as rh
    out = {"presets": {}, "synthetic": {}}

    def timed(run, steps):
        pops, t0 = [], time.perf_counter()
        for _ in range(steps):
            pops.append(len(run.view.agents))
            run.step()
            if run.exited_at is not None:
                break
        return float(sum(pops)), time.perf_counter() - t0, len(pops)

    presets = [("settings.json as shipped (utilicalc)", None, {}, 100), ("settings.json with moveEat", None, {"moveEat": True, "utilicalc": False}, 100),
               ("examples/seasons", "seasons", {}, 500), ("examples/pollution", "pollution", {}, 500),
               ("examples/utilicalc_sugar", "utilicalc_sugar", {}, 500),
               ("examples/evolution_metabolism (limitedLifespan + procreate)", "evolution_metabolism", {}, 300),
               ("examples/trade_price_volume without trade (spice)", "trade_price_volume", {"trade": False}, 300)]
    for label, preset, rules, steps in presets:
        s = rh.load_settings(preset)
        s["rules"].update(rules)
        r = rh.ReferenceRun(s, rng="mt")
        a, dt, done = timed(r, steps)
        out["presets"][label] = {"agent_steps_per_s": a / dt, "steps": done, "seconds": dt}
    for n, steps in ((200, 10), (800, max(1, min(args.steps, 6)))):
        w = synthetic.tiled_world(n)
        s = synthetic.synthetic_settings(n, pollution=wl["pollution"], utilicalc=wl.get("utilicalc", False))
        r = rh.ReferenceRun(s, rng="keyed", world=w)
        if n == 800:
            timed(r, min(args.warmup, 1))
        a, dt, done = timed(r, steps)
        out["synthetic"]["%dx%d" % (n, n)] = {"agent_steps_per_s": a / dt, "cell_updates_per_s": n * n * done / dt, "agents": len(w["id"]),
                                              "steps": done, "seconds": dt}
    return out


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    py = None
    why_not = None
    if os.path.isfile(os.path.join(REF_STAGE, "sugarscape.py")):
        try:
            py = python_reference_figures(args, wl)
        except Exception as ex:                          # reported, never hidden: the C port below still gives a line
            why_not = "%s: %s" % (type(ex).__name__, ex)
    else:
        why_not = "baseline/_ref/sugarscape.py is missing (staged by __graft_entry__.build() where /root/reference exists)"
    settings, w = make_world(dict(wl, n=min(wl["n"], 4096)))
    ns = 2048 if wl["pollution"] else 4096
    ns = min(ns, w["cap"].shape[0])
    r = oracle_sample(settings, w, ns, args.steps, args.warmup)
    port_value = r["agent_steps"] / r["seconds"]
    port = {"value": port_value, "unit": "agent-steps/s", "cores": 1, "kind": "port",
            "sample": "C port of the reference's sequential loop (oracle/ss_oracle.c), 1 thread, %dx%d sub-lattice of the workload's "
                      "landscape with %d agents, %d steps after %d warm-up" % (r["n"], r["n"], r["agents"], r["steps"], args.warmup),
            "replicas": replicas_throughput(wl)}
    if py is not None:
        big = py["synthetic"]["800x800"]
        value, ms_step = big["agent_steps_per_s"], 1e3 * big["seconds"] / big["steps"]
        cpu = {"value": value, "unit": "agent-steps/s", "cores": 1, "kind": "reference-python", "host_cpus": os.cpu_count(),
               "sample": "the reference's pure-Python loop (View.updateGame of the unmodified sources under baseline/_ref, headless, "
                         "CPython %d.%d, 1 core: the loop is sequential) on the workload's landscape and rules at 800x800 / %d agents, "
                         "%d steps; shipped presets and the 200x200 world under `python_reference`" % (
                             sys.version_info[0], sys.version_info[1], big["agents"], big["steps"]),
               "python_reference": py, "c_port": port}
    else:
        value, ms_step = port_value, 1e3 * r["seconds"] / r["steps"]
        cpu = dict(port, host_cpus=os.cpu_count(), python_reference_unavailable=why_not)
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "timing": "host wall clock, bounded sample (cpu_baseline.sample)"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def algorithmic_bytes(n, pollution, acted, mean_vision=3.5):
    """grow 24 B/cell, diffusion 16 B/cell, move 100+80v (pollution) / 76+48v B/agent-step."""
    per_agent = (100 + 80 * mean_vision) if pollution else (76 + 48 * mean_vision)
    return dict(grow=24.0 * n * n, diffuse=16.0 * n * n if pollution else 0.0, move=per_agent * acted)


def secondary_run(wl2, w2, peak, steps=5, warm=3):
    """A short device-timed run of another scale configuration, with the per-kernel-class rooflines."""
    from sugarscape_utilicalc_b200.engine import Engine
    settings2 = synthetic.synthetic_settings(wl2["n"], pollution=wl2["pollution"], utilicalc=wl2.get("utilicalc", False))
    if w2 is None:
        w2 = synthetic.tiled_world(wl2["n"])
    eng = Engine(config_from_settings(settings2, rng_mode="keyed"))
    upload(eng, w2)
    pop0 = int(eng.step(warm)["population"][-1])
    st = eng.step(steps)
    ms = eng.kernel_times()["last_call"]
    pops = np.concatenate([[pop0], st["population"][:-1]])
    eng.set_option("profile", 1)
    eng.set_option("reset_times", 0)
    st2 = eng.step(steps)
    kt = eng.kernel_times()
    cnt = eng.counters()
    health = eng.health()
    eng.close()
    ab = algorithmic_bytes(wl2["n"], wl2["pollution"], float(st2["acted"].sum()) / steps)
    agent_ms = (kt["prepare"] + kt["move"] + kt["settle"]) / steps
    out = {"workload": wl2["label"], "value": float(pops.sum()) / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms / steps,
           "steps": steps, "warmup": warm, "pass_impl": cnt["pass_impl"], "passes_per_step": cnt["passes"] / (warm + 2.0 * steps),
           "health": health, "kernels": {}}
    for name, t, nbytes in (("agent phase", agent_ms, ab["move"]), ("grow", kt["grow"] / steps, ab["grow"]),
                            ("diffuse", kt["diffuse"] / steps, ab["diffuse"])):
        if t > 0:
            tr, src = recorded_traffic(wl2.get("key"), name.split()[0])
            out["kernels"][name] = {"ms_per_step": t, "algorithmic_bytes_per_step": nbytes, "achieved_gbs": nbytes / (t * 1e-3) / 1e9,
                                    "frac": nbytes / (t * 1e-3) / 1e9 / peak, "traffic": tr, "traffic_source": src}
    return out


def run_engine_single(args, wl):
    from sugarscape_utilicalc_b200.engine import Engine
    settings, w = make_world(wl)
    n = wl["n"]
    cfg = config_from_settings(settings, rng_mode="keyed")
    peak, peak_src = measured_peak()
    K, W = args.steps, args.warmup
    # ---------------- value: state resident in HBM, device-timed ----------------
    eng = Engine(cfg)
    t0 = time.perf_counter()
    upload(eng, w)
    t_upload = time.perf_counter() - t0
    pop0 = len(w["id"])
    if W:
        st = eng.step(W)
        pop0 = int(st["population"][-1])
    eng.set_option("reset_counters", 0)
    sampler = ClockSampler(0)
    sampler.start()
    st = eng.step(K)
    ms = eng.kernel_times()["last_call"]
    clocks = sampler.stop()
    cnt = eng.counters()
    pops = np.concatenate([[pop0], st["population"][:-1]])
    agent_steps = float(pops.sum())
    value = agent_steps / (ms * 1e-3)
    acted = float(st["acted"].sum())
    # ---------------- instrumented run: per-kernel-class device time (CUDA events) ----------------
    eng.set_option("profile", 1)
    eng.set_option("reset_times", 0)
    st2 = eng.step(K)
    kt = eng.kernel_times()
    eng.set_option("profile", 0)
    acted2 = float(st2["acted"].sum())
    ab = algorithmic_bytes(n, wl["pollution"], acted2 / K)
    kt["move"] += kt["settle"]          # the agent phase = move passes + the deferred settle pass
    if cnt["pass_impl"] == 3:           # dependency rounds: "prepare" is the conflict-graph kernel -- part of the agent phase
        kt["build"] = kt["prepare"]
        kt["move"] += kt["prepare"]
        kt["prepare"] = 0.0
    classes = {"move": ("move", ab["move"]), "grow": ("grow", ab["grow"]), "diffuse": ("diffuse", ab["diffuse"])}
    kernels = {}
    tot = sum(kt[k] for k in ("prepare", "move", "stats", "grow", "diffuse"))
    for name, (key, nbytes) in classes.items():
        t = kt[key] / K
        if t > 0:
            kernels[name] = {"ms_per_step": t, "share": kt[key] / tot, "algorithmic_bytes_per_step": nbytes,
                             "achieved_gbs": nbytes / (t * 1e-3) / 1e9, "frac": nbytes / (t * 1e-3) / 1e9 / peak}
    kernels["settle (part of move)"] = {"ms_per_step": kt["settle"] / K}
    kernels["prepare+stats"] = {"ms_per_step": (kt["prepare"] + kt["stats"]) / K,
                                "share": (kt["prepare"] + kt["stats"]) / tot}
    dom = max(("move", "grow", "diffuse"), key=lambda k: kernels.get(k, {"ms_per_step": 0})["ms_per_step"])
    traffic, traffic_src = recorded_traffic(wl.get("key"), dom)
    move_kernel = ["ss_move_pass_kernel", "ss_move_solo_kernel", "ss_move_solo_kernel (lanes)",
                   "ss_dr_build_kernel + ss_dr_rounds_kernel"][cnt["pass_impl"]]
    roofline = {"bound": "hbm", "kernel": {"move": move_kernel + (" (conflict graph + all dependency rounds of a step)" if cnt["pass_impl"] == 3
                                                                  else " (all passes of a step) + ss_settle_kernel"),
                                           "grow": "ss_dr_grow_kernel" if cnt["pass_impl"] == 3 else "ss_grow_kernel", "diffuse": "ss_diffuse_col_kernel"}[dom],
                "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "ms_per_launch_set": kernels[dom]["ms_per_step"]}
    whole_bytes = ab["move"] + ab["grow"] + ab["diffuse"]
    step_frac = whole_bytes / (ms / K * 1e-3) / 1e9 / peak
    if "build" in kt:
        kernels["conflict graph (part of move)"] = {"ms_per_step": kt["build"] / K}
        kernels["dependency rounds (part of move)"] = {"ms_per_step": (kt["move"] - kt["build"]) / K, "rounds_per_step": eng.rounds() / (W + 2.0 * K)}
    # digest of the state after W + 2K steps: a multi-GPU run of the same workload and step counts prints the same one
    from sugarscape_utilicalc_b200.strips import state_digest
    dga = eng.download_agents()
    dg = state_digest(dga["id"], dga["x"], dga["y"], dga["sugar"], n)
    health = eng.health()
    del dga
    eng.close()
    # ---------------- e2e: public API with host buffers ----------------
    # The caller's arrays live in page-locked host memory (ss_host_alloc), allocated and filled before the clock starts;
    # the timed region is create + upload (H2D) + K steps with per-step stats read back + download (D2H) of the state.
    from sugarscape_utilicalc_b200.engine import pinned_copy, pinned_empty
    Ke = min(K, args.e2e_steps)
    cell_keys, agent_keys = ("sugar", "cap", "pollution"), ("id", "x", "y", "vision", "metab")
    pol_on = bool(wl["pollution"])      # without the pollution rule the slot is all zeros: passed as NULL, not downloaded
    hw = {k: (pinned_copy(w[k], np.float64) if (k != "pollution" or pol_on) else None) for k in cell_keys}
    hw.update({k: pinned_copy(w[k], np.int32) for k in agent_keys})
    hw["sugar_agent"] = pinned_copy(w["sugar_agent"], np.float64)
    hout = dict(sugar=pinned_empty((n, n)), occ=pinned_empty((n, n), np.int32))
    if pol_on:
        hout["pollution"] = pinned_empty((n, n))
    aout = {k: pinned_empty(len(w["id"]), np.int32) for k in agent_keys}
    aout["sugar"] = pinned_empty(len(w["id"]), np.float64)
    for a in list(hout.values()) + list(aout.values()):       # the result buffers are the caller's: touched before the clock starts
        a.fill(0)
    e2e_runs = []
    for _ in range(2):                       # the whole end-to-end sequence twice (first uses of page-locked buffers and of the
        t0 = time.perf_counter()             # download kernels have cost up to 1.6 s on a fresh box); the faster one is reported,
        eng = Engine(cfg)                    # both are listed
        t1 = time.perf_counter()
        for k, v in ENGINE_OPTS.items():
            eng.set_option(k, v)
        eng.upload_cells(hw["sugar"], hw["cap"], hw["pollution"])
        t1b = time.perf_counter()
        eng.upload_agents(hw["id"], hw["x"], hw["y"], hw["vision"], hw["metab"], hw["sugar_agent"])
        t2 = time.perf_counter()
        ste = eng.step(Ke)                       # per-step stats are read back to the host
        t3 = time.perf_counter()
        cells = eng.download_cells(cap=False, pollution=pol_on, out=hout)
        ags = eng.download_agents(out=aout)
        t4 = time.perf_counter()
        eng.close()
        e2e_runs.append((t4 - t0, (t0, t1, t1b, t2, t3, t4)))
    t_e2e, (t0, t1, t1b, t2, t3, t4) = min(e2e_runs, key=lambda r: r[0])
    popse = np.concatenate([[len(w["id"])], ste["population"][:-1]])
    e2e_value = float(popse.sum()) / t_e2e
    ncell_arrays = 3 if pol_on else 2                             # sugar, cap (, pollution) + six agent columns
    h2d = (ncell_arrays * 8 * n * n + 28 * len(w["id"])) / Ke
    d2h = (((ncell_arrays - 1) * 8 + 4) * n * n + 28 * len(ags["id"]) + 48 * Ke) / Ke       # sugar (, pollution), occupant ids
    e2e_parts = {"create": t1 - t0, "upload_cells": t1b - t1, "upload_agents": t2 - t1b, "steps": t3 - t2, "download": t4 - t3}
    del cells, ags, hw, hout, aout
    # ---------------- cpu baseline: oracle port on a bounded sample ----------------
    cpu = None
    if not args.no_cpu:
        ns = 1024 if wl["pollution"] else 2048
        r = oracle_sample(settings, w, ns, 5)
        cpu = {"value": r["agent_steps"] / r["seconds"], "unit": "agent-steps/s", "cores": 1, "kind": "port",
               "sample": "oracle/ss_oracle.c (C port of the reference loop), 1 thread, %dx%d sub-lattice of the same "
                         "world, %d agents, 5 steps (%.1f s)" % (r["n"], r["n"], r["agents"], r["seconds"]),
               "host_cpus": os.cpu_count()}
        # ... and the reference's own Python loop on a smaller sample of the same landscape (the --impl reference arm
        # runs the presets and the 800^2 world in full)
        if os.path.isfile(os.path.join(REF_STAGE, "sugarscape.py")):
            try:
                os.environ["SS_REFERENCE"] = REF_STAGE
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import ref_harness as rh
                nr = 400
                wr = synthetic.tiled_world(nr)
                rr = rh.ReferenceRun(synthetic.synthetic_settings(nr, pollution=wl["pollution"], utilicalc=wl.get("utilicalc", False)),
                                     rng="keyed", world=wr)
                t0 = time.perf_counter()
                acted = 0
                for _ in range(6):
                    acted += len(rr.view.agents)
                    rr.step()
                dt = time.perf_counter() - t0
                cpu = {"value": acted / dt, "unit": "agent-steps/s", "cores": 1, "kind": "reference", "host_cpus": os.cpu_count(),
                       "sample": "the reference's pure-Python loop (unmodified sources under baseline/_ref, headless, 1 core) on the same "
                                 "landscape and rules at %dx%d / %d agents, 6 steps (%.1f s)" % (nr, nr, len(wr["id"]), dt),
                       "c_port": cpu}
            except Exception as ex:
                cpu["python_reference_unavailable"] = "%s: %s" % (type(ex).__name__, ex)
    # ---------------- the other scale configurations, short (BASELINE configs 3/4: utilicalc, pollution + diffusion) ----------------
    secondary = {}
    if not args.no_secondary and wl.get("key") == "c5":
        for key in ("c4", "c5u"):
            try:
                secondary[key] = secondary_run(workload(key), w if key == "c5u" else None, peak)
            except Exception as ex:
                secondary[key] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "grid": n, "agents": len(w["id"]),
                   "l2": "working set %.1f GB >> 126 MB L2, no flush needed" % ((28.0 if wl["pollution"] else 20.0) * n * n / 1e9),
                   "timing": "CUDA events on the engine stream around ss_step(K)"},
        "cell_updates_per_s": n * n * K / (ms * 1e-3),
        "population": [int(pop0), int(st["population"][-1])],
        "passes_per_step": cnt["passes"] / K, "gpu_launches": cnt["launches"],
        "move_pass": {"impl": ["cta_per_window", "warp_per_window", "lane_per_agent", "dependency_rounds"][cnt["pass_impl"]], "window": cnt["window"]},
        "digest": {"xor": dg[0], "sum": dg[1], "agents": dg[2], "after_steps": W + 2 * K,
                   "what": "order-independent digest of (id, cell, sugar) of the living agents; the --gpus N lines of the same "
                           "workload and step counts carry the same digest"},
        "health": health,
        "host_syncs_per_step": cnt["syncs"] / K,
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "seconds": t_e2e, "seconds_by_phase": e2e_parts, "seconds_of_both_runs": [r[0] for r in e2e_runs],
                "what": "the faster of two identical runs of: Engine(cfg); upload_cells/agents from page-locked host numpy arrays; step(K) with stats "
                        "read-back; download_cells+download_agents into page-locked host arrays (capacity is not read "
                        "back; the pollution slot crosses the bus only when the pollution rule is on)"},
        "roofline": roofline,
        "step_roofline": {"algorithmic_bytes_per_step": whole_bytes, "frac_of_peak": step_frac},
        "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks, "upload_s": t_upload,
        "secondary": secondary,
    }
    print(json.dumps(line), flush=True)


def run_preset(args, wl):
    """Shipped 50x50 presets through the serial MT-exact kernel, from the golden vectors' initial state
    (tests/golden/<name>.npz, recorded from the unmodified reference): agent-steps/s over the golden run."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from sugarscape_utilicalc_b200.engine import Engine
    g = common.load_golden(wl["preset"])
    cfg = config_from_settings(g["settings"], rng_mode=g["rng"])
    T = g["steps"]
    best = None
    for rep in range(max(1, args.warmup) + 1):
        eng = Engine(cfg)
        common.init_world(eng, common.golden_init(g), g["rng"], g["init_mt"], int(g["init_mt_index"]))
        st = eng.step(T)
        ms = eng.kernel_times()["last_call"]
        ok = bool(np.array_equal(st["population"], g["population"]) and np.array_equal(st["gini"], g["gini"]))
        eng.close()
        best = ms if best is None else min(best, ms)
    pops = np.concatenate([[len(g["init_order"])], g["population"][:-1]])
    line = {"metric": "agent-steps/sec", "value": float(pops.sum()) / (best * 1e-3), "unit": "agent-steps/s", "n_gpus": 1,
            "steps": T, "warmup": args.warmup, "ms_per_step": best / T, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "reference preset (golden initial state)",
            "config": {"workload": wl["label"], "rng": g["rng"], "kernel": "ss_serial_kernel, %d steps in one launch" % T,
                       "timing": "CUDA events on the engine stream"},
            "matches_reference_series": ok, "gpu_launches": 1,
            "reference_python_loop": "BASELINE.md section 2 (measured in the build container): 4.7k-36k agent-steps/s"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SS_BENCH_WORKLOAD", "c5"))
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short c4 / c5u runs of the N=1 line")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine option (ss_set_option) for A/B runs, e.g. --opt pass_impl=0")
    args = ap.parse_args()
    ENGINE_OPTS.update({kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt})
    wl = workload(args.workload)
    if "preset" in wl:
        run_preset(args, wl)
        return
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from sugarscape_utilicalc_b200 import strips
        strips.bench_main(args, wl)
        return
    run_engine_single(args, wl)


if __name__ == "__main__":
    main()
