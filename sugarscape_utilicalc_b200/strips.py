"""Multi-GPU timestep as row strips: one process per GPU, one strip of lattice rows per rank.

BASELINE config 5 ("row-strip sharded with NVLink halo exchange and agent migration").  Rank r of P holds the
rows [r N / P, (r+1) N / P) of the lattice -- cells, occupancy and the records of the agents standing on them -- so the
memory per rank shrinks with P.  Movement is toroidal (agent.py:362-368), so the strips form a ring.  Nothing crosses
the host during a step: the dependency-round kernels (csrc/ss_rounds.cu) read the two neighbours' border rows for the
halo of the conflict graph, scan and move across the border, carry the records of migrating agents and decrement the
dependency counters of agents on other strips through peer-mapped device pointers (CUDA IPC over NVLink); the strips
meet at flag words in peer memory between the phases of a step and between its dependency rounds.  `torch.distributed`
only carries the IPC handles at set-up and the barrier around the timed region.

`VirtualStrips` runs P strips inside one process on one GPU with the same kernels (raw device pointers instead of IPC
handles, the host as the barrier between phases and rounds): the GPU parity test of the N>1 path.
"""
import ctypes as C

import numpy as np

from . import capi, outputs
from .capi import STATS_DTYPE, SsConfig
from .engine import Engine, EngineError, _ptr

NPTR = 11
HANDLE_BYTES = 72


def strip_rows(n, rank, world):
    """Rows [lo, hi) of the lattice that rank `rank` of `world` holds (same rule as ss_create_strip)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def strip_of_row(x, n, world):
    """Rank that holds lattice row(s) x."""
    return ((np.asarray(x, dtype=np.int64) + 1) * world - 1) // n


def split_world(x0, x1, cells, agents, pinned=True):
    """What the rank holding rows [x0, x1) uploads, from the strip's cell rows and the WHOLE world's agent columns: the
    strip's own agents and the four columns the replicated directories are built from -- page-locked copies when
    `pinned`."""
    from .engine import pinned_copy
    cp = (lambda a, dt: pinned_copy(a, dt)) if pinned else (lambda a, dt: np.ascontiguousarray(a, dtype=dt))
    ids, x, y, vis, met, sug = (agents[k] for k in ("id", "x", "y", "vision", "metab", "sugar_agent"))
    m = (x >= x0) & (x < x1)
    return dict(sugar=cp(cells["sugar"], np.float64), cap=cp(cells["cap"], np.float64),
                local=[cp(a[m], np.int32) for a in (ids, x, y, vis, met)] + [cp(sug[m], np.float64)],
                directory=[cp(ids, np.int32), cp(x, np.int32), cp(met, np.int32), cp(sug, np.float64)],
                max_id=int(ids.max()) if len(ids) else 1, vmax=int(vis.max()) if len(vis) else 1)


class StripEngine(Engine):
    """An Engine over one strip of rows (ss_create_strip): uploads / downloads take the strip's share."""

    def __init__(self, cfg, rank, world, device=0):
        if not isinstance(cfg, SsConfig):
            cfg = SsConfig.from_buffer_copy(bytes(cfg))
        self._lib = capi.lib()
        self.cfg = cfg
        self.n = cfg.grid_size
        self.rank, self.world = rank, world
        self.x0, self.x1 = strip_rows(self.n, rank, world)
        self.rows = self.x1 - self.x0
        self._h = C.c_void_p()
        rc = self._lib.ss_create_strip(C.byref(cfg), int(device), int(rank), int(world), C.byref(self._h))
        if rc != 0:
            raise EngineError("ss_create_strip: " + self._lib.ss_last_error(None).decode())

    def upload_world(self, w):
        """w: the WHOLE world (sugar, cap (N,N); id, x, y, vision, metab, sugar_agent) -- this rank takes its strip."""
        n = self.n
        sl = slice(self.x0, self.x1)
        s = np.ascontiguousarray(np.asarray(w["sugar"], dtype=np.float64).reshape(n, n)[sl])
        c = np.ascontiguousarray(np.asarray(w["cap"], dtype=np.float64).reshape(n, n)[sl])
        self._check(self._lib.ss_upload_cells(self._h, _ptr(s), _ptr(c), None), "ss_upload_cells")
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        ids, x, y, vis, met = map(i32, (w["id"], w["x"], w["y"], w["vision"], w["metab"]))
        sug = np.ascontiguousarray(w["sugar_agent"], dtype=np.float64)
        self.set_option("global_slots", int(ids.max()) if len(ids) else 1)
        self.set_option("global_vmax", int(vis.max()) if len(vis) else 1)
        m = (x >= self.x0) & (x < self.x1)
        loc = [np.ascontiguousarray(a[m]) for a in (ids, x, y, vis, met, sug)]
        self._check(self._lib.ss_upload_agents(self._h, int(m.sum()), *[_ptr(a) for a in loc]), "ss_upload_agents")
        self._check(self._lib.ss_strip_init_directory(self._h, len(ids), _ptr(ids), _ptr(x), _ptr(met), _ptr(sug)),
                    "ss_strip_init_directory")
        self.n_slots = int(ids.max()) if len(ids) else 1

    def upload_split(self, sp):
        self._check(self._lib.ss_upload_cells(self._h, _ptr(sp["sugar"]), _ptr(sp["cap"]), None), "ss_upload_cells")
        self.set_option("global_slots", sp["max_id"])
        self.set_option("global_vmax", sp["vmax"])
        loc = sp["local"]
        self._check(self._lib.ss_upload_agents(self._h, len(loc[0]), *[_ptr(a) for a in loc]), "ss_upload_agents")
        d = sp["directory"]
        self._check(self._lib.ss_strip_init_directory(self._h, len(d[0]), *[_ptr(a) for a in d]), "ss_strip_init_directory")
        self.n_slots = sp["max_id"]
        return len(loc[0])

    def upload_strip(self, cells, agents):
        """cells: the strip's rows of sugar and cap; agents: id, x, y, vision, metab, sugar_agent of the WHOLE world."""
        return self.upload_split(split_world(self.x0, self.x1, cells, agents, pinned=False))

    def export_ptrs(self):
        p = np.zeros(NPTR, dtype=np.uint64)
        self._check(self._lib.ss_strip_export(self._h, _ptr(p)), "ss_strip_export")
        return p

    def connect(self, peer_rank, ptrs):
        p = np.ascontiguousarray(ptrs, dtype=np.uint64)
        self._check(self._lib.ss_strip_connect(self._h, int(peer_rank), _ptr(p)), "ss_strip_connect")

    def export_ipc(self):
        h = np.zeros(NPTR * HANDLE_BYTES, dtype=np.uint8)
        self._check(self._lib.ss_strip_ipc_export(self._h, _ptr(h)), "ss_strip_ipc_export")
        return h

    def connect_ipc(self, peer_rank, handles):
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        self._check(self._lib.ss_strip_ipc_connect(self._h, int(peer_rank), _ptr(h)), "ss_strip_ipc_connect")

    def phase(self, phase, seg=(0, 0), want_stats=False):
        tail = C.c_uint32()
        out = np.zeros(1, dtype=STATS_DTYPE) if want_stats else None
        self._check(self._lib.ss_strip_phase(self._h, int(phase), int(seg[0]), int(seg[1]), C.byref(tail), _ptr(out)),
                    "ss_strip_phase")
        return tail.value, out

    def step(self, nsteps=1, want_stats=True):
        """nsteps timesteps, synchronised with the other strips on the device (every rank must call it)."""
        out = np.zeros(nsteps, dtype=STATS_DTYPE) if want_stats else None
        self._check(self._lib.ss_strip_step(self._h, int(nsteps), _ptr(out)), "ss_strip_step")
        return out

    def download_cells(self, sugar=True, cap=True, pollution=True, occ=True, out=None):
        n, r = self.n, self.rows
        mk = lambda on, dt: np.empty((r, n), dtype=dt) if on else None
        s, c, p, o = mk(sugar, np.float64), mk(cap, np.float64), mk(pollution, np.float64), mk(occ, np.int32)
        self._check(self._lib.ss_download_cells(self._h, _ptr(s), _ptr(c), _ptr(p), _ptr(o)), "ss_download_cells")
        return dict(sugar=s, cap=c, pollution=p, occ=o)


def merge_agents(parts, seed, iteration, n_slots):
    """The ranks' agent lists (each in ascending keyed priority of the last step) as the reference's one list."""
    cat = {k: np.concatenate([p[k] for p in parts]) for k in ("id", "x", "y", "vision", "metab", "sugar")}
    if len(cat["id"]) == 0 or iteration <= 0:
        order = np.argsort(cat["id"], kind="stable")
    else:
        order = np.argsort(outputs.keyed_priorities(seed, iteration - 1, cat["id"], n_slots), kind="stable")
    return {k: v[order] for k, v in cat.items()}


class VirtualStrips:
    """P strips on one GPU in one process (the parity test of the multi-GPU path): the same kernels as the real run,
    the neighbours' device pointers handed over directly, the host as the barrier between phases and between rounds."""

    def __init__(self, cfg, world, device=0, options=None):
        self.cfg = SsConfig.from_buffer_copy(bytes(cfg))
        self.n = self.cfg.grid_size
        self.world = world
        self.ranks = [StripEngine(self.cfg, r, world, device=device) for r in range(world)]
        for e in self.ranks:
            for k, v in (options or {}).items():
                e.set_option(k, v)
        self.iteration = int(self.cfg.iteration)
        self.rounds_per_step = []

    @property
    def eng(self):
        return self.ranks[0]

    def upload_cells(self, sugar, cap, pollution=None):
        self._cells = dict(sugar=sugar, cap=cap)

    def upload_agents(self, id, x, y, vision, metab, sugar):
        w = dict(self._cells, id=id, x=x, y=y, vision=vision, metab=metab, sugar_agent=sugar)
        for e in self.ranks:
            e.upload_world(w)
        ptrs = [e.export_ptrs() for e in self.ranks]
        for e in self.ranks:
            for j in range(self.world):
                if j != e.rank:
                    e.connect(j, ptrs[j])
        self.n_slots = max(int(np.max(id)) if len(id) else 1, 1)

    def step(self, nsteps=1):
        out = np.zeros(nsteps, dtype=STATS_DTYPE)
        for t in range(nsteps):
            tails = [e.phase(0)[0] for e in self.ranks]
            begins = [0] * self.world
            rounds = 0
            while any(tl > b for tl, b in zip(tails, begins)):
                segs = list(zip(begins, tails))
                for e, sg in zip(self.ranks, segs):
                    e.phase(1, sg)
                begins = tails
                tails = [e.phase(1, (0, 0))[0] for e in self.ranks]      # (empty segment: only reads the tail)
                rounds += 1
                if rounds > 100000:
                    raise EngineError("dependency rounds did not terminate")
            self.rounds_per_step.append(rounds)
            for e in self.ranks:
                e.phase(2)
            stats = [e.phase(3, want_stats=True)[1] for e in self.ranks]
            for s in stats[1:]:
                if s.tobytes() != stats[0].tobytes():
                    raise EngineError("the strips disagree on the statistics: %r vs %r" % (stats[0], s))
            out[t] = stats[0][0]
            self.iteration += 1
        return out

    def download_cells(self):
        parts = [e.download_cells() for e in self.ranks]
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in ("sugar", "cap", "pollution", "occ")}

    def download_agents(self):
        parts = [e.download_agents() for e in self.ranks]
        return merge_agents(parts, int(self.cfg.keyed_seed), self.iteration, self.n_slots)

    def agent_count(self):
        return sum(e.agent_count() for e in self.ranks)

    def close(self):
        for e in self.ranks:
            e.close()


def connect_over_torch(eng, dist):
    """One process per GPU: exchange the strips' CUDA IPC handles through torch.distributed and connect every peer."""
    import torch
    mine = torch.from_numpy(eng.export_ipc().copy())
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    allh = [torch.empty_like(mine, device=dev) for _ in range(eng.world)]
    dist.all_gather(allh, mine.to(dev))
    for j in range(eng.world):
        if j != eng.rank:
            eng.connect_ipc(j, allh[j].cpu().numpy())
    dist.barrier()


def state_digest(ids, x, y, sugar, n):
    """Order-independent digest of a population (ids, cells, sugar): equal on every decomposition of the same run."""
    ids = np.asarray(ids, dtype=np.uint64)
    cell = np.asarray(x, dtype=np.uint64) * np.uint64(n) + np.asarray(y, dtype=np.uint64)
    h = (ids * np.uint64(0x9E3779B97F4A7C15)) ^ ((cell + np.uint64(1)) * np.uint64(0xD1342543DE82EF95))
    h ^= np.round(np.asarray(sugar, dtype=np.float64) * 100.0).astype(np.int64).astype(np.uint64) * np.uint64(0xA24BAED4963EE407)
    return int(np.bitwise_xor.reduce(h)) if len(h) else 0, int(np.add.reduce(h, dtype=np.uint64)) if len(h) else 0, int(len(h))


def bench_main(args, wl):
    """bench.py --gpus N under torchrun: one rank per GPU, one strip of rows each (strong scaling of the workload)."""
    import json
    import os
    import sys
    import time
    import torch
    import torch.distributed as dist
    from . import config_from_settings, synthetic
    from .clocks import ClockSampler
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = wl["n"]
    if wl["pollution"] or wl.get("utilicalc"):
        raise SystemExit("row strips run rule M without the pollution rule (workload c5 or <n>)")
    settings = synthetic.synthetic_settings(n, pollution=False)
    cfg = config_from_settings(settings, rng_mode="keyed")
    x0, x1 = strip_rows(n, rank, world)
    t0 = time.time()
    w = synthetic.tiled_world_strip(n, x0, x1)
    if rank == 0:
        sys.stderr.write("[bench] strip world %dx%d rows [%d,%d), %d agents generated in %.1fs\n" % (n, n, x0, x1, len(w["id"]), time.time() - t0))
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local)

    sp = split_world(x0, x1, w, w, pinned=True)     # host-side preparation, outside every clock

    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in getattr(args, "opt", [])}

    def run(k_steps, warm):
        eng = StripEngine(cfg, rank, world, device=local)
        for k_, v_ in opts.items():
            eng.set_option(k_, v_)
        t_a = time.perf_counter()
        mine = eng.upload_split(sp)
        t_b = time.perf_counter()
        connect_over_torch(eng, dist)
        pops0 = len(w["id"])
        if warm:
            pops0 = int(eng.step(warm)["population"][-1])
        dist.barrier()
        torch.cuda.synchronize()
        t_c = time.perf_counter()
        st = eng.step(k_steps)
        torch.cuda.synchronize()
        t_d = time.perf_counter()
        ms = torch.tensor([eng.kernel_times()["last_call"]], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return eng, st, pops0, float(ms.item()), dict(upload=t_b - t_a, steps=t_d - t_c), mine

    sampler = ClockSampler(local)
    eng, st, pop0, ms, _, mine = run(K, W)
    sampler.start()
    eng.set_option("reset_times", 0)
    eng.set_option("profile", 1)                   # the timed steps once more, instrumented: device time per phase on this rank
    st2 = eng.step(K)
    eng.set_option("profile", 0)
    clocks = sampler.stop()
    kt = eng.kernel_times()
    ms2 = torch.tensor([eng.kernel_times()["last_call"]], dtype=torch.float64, device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    cnt = eng.counters()
    health = eng.health()
    eng_rounds = eng.rounds()
    # digest of the state after W + 2K steps: the same on every decomposition (compare the N = 1 line)
    ags = eng.download_agents()
    x_, s_, c_ = state_digest(ags["id"], ags["x"], ags["y"], ags["sugar"], n)
    parts = [None] * world
    dist.all_gather_object(parts, (x_, s_, c_, health, int(mine)))
    eng.close()
    # end to end through the public API: strip upload from host arrays, K steps with the statistics read back, download
    dist.barrier()
    t_e0 = time.perf_counter()
    eng = StripEngine(cfg, rank, world, device=local)
    eng.upload_split(sp)
    connect_over_torch(eng, dist)
    ste = eng.step(K)
    cells = eng.download_cells(cap=False, pollution=False)
    ags2 = eng.download_agents()
    torch.cuda.synchronize()
    t_e = torch.tensor([time.perf_counter() - t_e0], dtype=torch.float64, device=dev)
    dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([16.0 * (x1 - x0) * n + 20.0 * len(w["id"]) + 28.0 * mine], dtype=torch.float64, device=dev)
    d2h = torch.tensor([12.0 * (x1 - x0) * n + 28.0 * len(ags2["id"]) + 48.0 * K], dtype=torch.float64, device=dev)
    dist.all_reduce(h2d); dist.all_reduce(d2h)
    eng.close()
    del cells, ags2
    if rank == 0:
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        pops = np.concatenate([[pop0], st["population"][:-1]])
        agent_steps = float(pops.sum())
        value = agent_steps / (ms * 1e-3)
        popse = np.concatenate([[len(w["id"])], ste["population"][:-1]])
        acted = float(st["acted"].sum()) / K
        alg = (24.0 * n * n + (76 + 48 * 3.5) * acted) / world          # per rank and step
        V = int(w["vision"].max())
        halo = 2 * ((2 * V + 3) // 4 * 4) * n * 5                          # border rows of occupancy words + cell bytes, both neighbours
        xd, sd, cd = 0, 0, 0
        for p in parts:
            xd ^= p[0]; sd = (sd + p[1]) % (1 << 64); cd += p[2]
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["label"], "grid": n, "agents": len(w["id"]),
                       "decomposition": "row strips: rank r holds rows [r N / P, (r+1) N / P) and the agents on them; halo reads, "
                                        "moves, migrating records and dependency counters over peer memory (CUDA IPC / NVLink); "
                                        "device-side barriers between phases and dependency rounds",
                       "l2": "working set per rank %.2f GB >> 126 MB L2, no flush needed" % (20.0 * n * n / world / 1e9),
                       "timing": "CUDA events on every rank's stream around ss_strip_step(K), max over ranks"},
            "cell_updates_per_s": n * n * K / (ms * 1e-3),
            "population": [int(pop0), int(st["population"][-1])],
            "ms_per_step_instrumented_run": float(ms2.item()) / K,
            "rank0_ms_per_step": {"conflict_graph": kt["prepare"] / K, "barrier+dependency_rounds": kt["move"] / K,
                                  "merge_stats+delta_lists+barrier": kt["serial"] / K, "statistics": kt["stats"] / K,
                                  "growback": kt["grow"] / K, "end_of_step_barrier": kt["settle"] / K,
                                  "rounds_per_step": eng_rounds / float(W + 2 * K)},
            "passes_per_step": 1.0, "host_syncs_per_step": cnt["syncs"] / max(cnt["steps"], 1),
            "gpu_launches": int(cnt["launches"]),
            "move_pass": {"impl": "dependency_rounds (ss_dr_build_kernel + ss_dr_rounds_kernel, strips)"},
            "e2e": {"value": float(popse.sum()) / float(t_e.item()), "unit": "agent-steps/s",
                    "h2d_bytes_per_step": float(h2d.item()) / K, "d2h_bytes_per_step": float(d2h.item()) / K,
                    "steps": K, "seconds": float(t_e.item()),
                    "what": "per rank: StripEngine; upload of the strip's cells and agents (+ four columns of the whole world's agents "
                            "for the replicated directories) from page-locked host arrays; IPC connect; step(K) with the statistics "
                            "read back; download of the strip's cells and agents; max over ranks"},
            "roofline": {"bound": "hbm", "kernel": "whole step of one rank (conflict graph + dependency rounds + growback)",
                         "achieved": alg / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / K * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
            "nvlink": {"halo_bytes_per_rank_and_step": halo, "achieved_gbs": halo / (ms / K * 1e-3) / 1e9, "peak_gbs": 770.0,
                       "what": "border rows read from the two ring neighbours by the conflict-graph kernel; border agents' cells, "
                               "migrating records and cross-strip counters add a few 100 KB"},
            "digest": {"xor": xd, "sum": sd, "agents": cd, "after_steps": W + 2 * K,
                       "what": "order-independent digest of (id, cell, sugar) of the living agents, merged over the ranks: equal to the "
                               "N=1 line's digest"},
            "strips": [{"rank": i, "agents_at_upload": p[4], "health": p[3]} for i, p in enumerate(parts)],
            "clocks": clocks, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
