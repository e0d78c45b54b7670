"""Multi-GPU timestep of ROUND 1 (kept with its tests; `bench.py --gpus N` runs the row strips of strips.py): one process
per GPU, `torch.distributed` (NCCL over NVLink) for the plumbing.

Decomposition.  Every rank holds a replica of the lattice and the agent store and, in every move
pass of the window kernels, resolves only the agents of ITS contiguous share of the window rows -- windows of one launch are
disjoint, so ranks never touch the same cells.  The turns a rank resolved are appended on the device to two logs: 8-byte
entries for the agents that cannot starve (slot | harvest, new cell) and 48-byte entries for synchronously settled turns and
Q1 skips.  After the pass the ranks all-gather both slices in one buffer and replay the other ranks' entries on their replica
(`ss_apply_clog_kernel` / `ss_apply_log_kernel`: same stores, same arithmetic -> replicas stay bit-identical).  What crosses
NVLink per step is about 10 B per agent instead of lattice halos; agent migration is implicit in the log.  Growback, settle,
statistics and diffusion run on every replica, and every replica replays all other ranks' turns: the scheme scaled 1.40x at
8 GPUs (13.6 ms per C5 step), which is why round 2 replaced it by row strips (4.04x, 2.77 ms).

`VirtualRanks` runs P replicas inside one process on one GPU with the same code path (device
pointers instead of NCCL) -- the GPU parity test of the N>1 path; the gloo test
(tests/test_sharded_gloo.py) covers the host-side collective helpers on CPU.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import capi
from .engine import Engine, EngineError
from .capi import STATS_DTYPE

ENTRY_BYTES = 48         # synchronously settled turns, Q1 skips (ShardLogEntry)
COMPACT_BYTES = 8        # turns of agents that cannot starve: slot | harvest << 24, new cell


def window_rows_for_rank(nwin, rank, world):
    """Contiguous share [lo, hi) of the `nwin` window rows (same rule as ss_shard_pass)."""
    return (rank * nwin) // world, ((rank + 1) * nwin) // world


def gather_logs(local_bytes, n_local, local2, n2, world, gather_counts, gather_padded):
    """Both logs of a pass through one counts collective and one payload collective.  Returns the raw all-gathered
    buffer and its layout: dict(buf=(world, stride) uint8 tensor or None, stride, off2, counts_dev=(world, 2) int64 device
    tensor, mx, mx2) -- what ss_shard_apply_gathered takes."""
    import torch
    counts_t = torch.tensor([int(n_local), int(n2)], dtype=torch.int64, device=local_bytes.device)
    counts_dev = gather_counts(counts_t).reshape(world, 2)
    host = counts_dev.tolist()
    mx, mx2 = max(int(c[0]) for c in host), max(int(c[1]) for c in host)
    if mx == 0 and mx2 == 0:
        return dict(buf=None, stride=0, off2=0, counts_dev=counts_dev, mx=0, mx2=0)
    off2 = (mx * ENTRY_BYTES + 15) // 16 * 16
    buf = torch.empty(off2 + mx2 * COMPACT_BYTES, dtype=torch.uint8, device=local_bytes.device)
    if n_local:
        buf[: n_local * ENTRY_BYTES] = local_bytes[: n_local * ENTRY_BYTES]
    if n2:
        buf[off2: off2 + n2 * COMPACT_BYTES] = local2[: n2 * COMPACT_BYTES]
    allb = gather_padded(buf)
    return dict(buf=allb, stride=buf.numel(), off2=off2, counts_dev=counts_dev, mx=mx, mx2=mx2)


def pad_and_gather(local_bytes, n_local, world, gather_counts, gather_padded, entry_bytes=ENTRY_BYTES, local2=None, n2=0,
                   entry_bytes2=COMPACT_BYTES):
    """Variable-length all-gather through two fixed-size collectives.
    local_bytes: 1-D uint8 tensor with n_local*entry_bytes valid bytes; optionally a second log (local2, n2 entries of
    entry_bytes2) travels in the same buffer.  Returns (counts, parts) or, with a second log, (counts, parts, counts2,
    parts2): per-rank uint8 tensors trimmed to their length."""
    import torch
    two = local2 is not None
    counts_t = torch.tensor([int(n_local), int(n2)], dtype=torch.int64, device=local_bytes.device)
    all_counts = gather_counts(counts_t).reshape(world, 2).tolist()           # (world, 2) int64
    counts = [int(c[0]) for c in all_counts]
    counts2 = [int(c[1]) for c in all_counts]
    mx, mx2 = max(counts), max(counts2)
    if mx == 0 and mx2 == 0:
        empty = [local_bytes[:0] for _ in range(world)]
        return (counts, empty, counts2, list(empty)) if two else (counts, empty)
    off2 = (mx * entry_bytes + 15) // 16 * 16
    buf = torch.empty(off2 + mx2 * entry_bytes2, dtype=torch.uint8, device=local_bytes.device)
    buf[: n_local * entry_bytes] = local_bytes[: n_local * entry_bytes]
    if two and n2:
        buf[off2: off2 + n2 * entry_bytes2] = local2[: n2 * entry_bytes2]
    allb = gather_padded(buf)                                 # (world, len(buf))
    parts = [allb[r, : counts[r] * entry_bytes] for r in range(world)]
    if not two:
        return counts, parts
    return counts, parts, counts2, [allb[r, off2: off2 + counts2[r] * entry_bytes2] for r in range(world)]


class _DevView:
    """Zero-copy torch view of engine device memory (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardRank:
    """One rank's replica: an Engine plus the shard entry points of the C-ABI."""

    def __init__(self, cfg, rank, world, device=0, options=None):
        self.eng = Engine(cfg, device=device)
        for k, v in (options or {}).items():
            self.eng.set_option(k, v)
        self.rank, self.world = rank, world
        self._lib = capi.lib()

    def upload(self, w):
        self.eng.upload_cells(w["sugar"], w["cap"], w["pollution"])
        self.eng.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
        if self.eng.counters()["path"] != 2:
            raise EngineError("multi-GPU needs the lattice path (keyed RNG, moveEat)")
        self._check(self._lib.ss_shard_enable(self.eng._h, self.rank, self.world), "ss_shard_enable")

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s: %s" % (what, self._lib.ss_last_error(self.eng._h).decode()))

    def begin_step(self):
        self._check(self._lib.ss_shard_begin_step(self.eng._h), "ss_shard_begin_step")

    def run_pass(self):
        """One pass over this rank's window rows -> (ptr, n) of its 48-byte entries, (ptr, n) of its 8-byte entries."""
        b, e, base = C.c_uint64(), C.c_uint64(), C.c_void_p()
        cb, ce, cbase = C.c_uint64(), C.c_uint64(), C.c_void_p()
        self._check(self._lib.ss_shard_pass(self.eng._h, C.byref(b), C.byref(e), C.byref(base), C.byref(cb), C.byref(ce),
                                            C.byref(cbase)), "ss_shard_pass")
        return (int(base.value) + b.value * ENTRY_BYTES, int(e.value - b.value),
                int(cbase.value) + cb.value * COMPACT_BYTES, int(ce.value - cb.value))

    def apply_gathered(self, g):
        """Replay every other rank's slices of a gather_logs() result."""
        if g["buf"] is None:
            return
        self._check(self._lib.ss_shard_apply_gathered(self.eng._h, C.c_void_p(g["buf"].data_ptr()), int(g["stride"]),
                                                      int(g["off2"]), C.c_void_p(g["counts_dev"].data_ptr()), self.world,
                                                      int(g["mx"]), int(g["mx2"])), "ss_shard_apply_gathered")

    def apply(self, dev_ptr, n, compact_ptr=0, n_compact=0):
        self._check(self._lib.ss_shard_apply(self.eng._h, C.c_void_p(int(dev_ptr)), int(n), C.c_void_p(int(compact_ptr)),
                                             int(n_compact)), "ss_shard_apply")

    def pending(self):
        p = C.c_uint64()
        self._check(self._lib.ss_shard_pending(self.eng._h, C.byref(p)), "ss_shard_pending")
        return int(p.value)

    def end_step(self):
        out = np.zeros(1, dtype=STATS_DTYPE)
        self._check(self._lib.ss_shard_end_step(self.eng._h, out.ctypes.data_as(C.c_void_p)), "ss_shard_end_step")
        return out

    def close(self):
        self.eng.close()


class VirtualRanks:
    """P replicas in one process on one GPU: the N>1 code path without NCCL (log slices are handed over
    as device pointers).  Same interface as Engine for the parity tests."""

    def __init__(self, cfg, world, device=0, options=None, gathered=False):
        """gathered: hand the logs over in the all-gather layout (torch copies on the one GPU) and replay them with
        ss_shard_apply_gathered, as DistributedEngine does; default: per-rank device pointers and ss_shard_apply."""
        self.ranks = [ShardRank(cfg, r, world, device, options) for r in range(world)]
        self.world = world
        self.passes = 0
        self.gathered = gathered
        self.device = device

    def upload_cells(self, sugar, cap, pollution=None):
        self._cells = (sugar, cap, pollution)

    def upload_agents(self, id, x, y, vision, metab, sugar):
        s, c, p = self._cells
        for r in self.ranks:
            r.upload(dict(sugar=s, cap=c, pollution=p, id=id, x=x, y=y, vision=vision, metab=metab, sugar_agent=sugar))

    def step(self, nsteps=1):
        out = np.zeros(nsteps, dtype=STATS_DTYPE)
        for t in range(nsteps):
            for r in self.ranks:
                r.begin_step()
            while True:
                logs = [r.run_pass() for r in self.ranks]
                self.passes += 1
                if self.gathered:
                    self._apply_gathered(logs)
                for i, r in enumerate(self.ranks if not self.gathered else []):
                    for j, (ptr, n, cptr, cn) in enumerate(logs):
                        if j != i:
                            r.apply(ptr, n, cptr, cn)
                pend = [r.pending() for r in self.ranks]
                assert len(set(pend)) == 1, "replicas disagree on the pending count: %s" % pend
                if pend[0] == 0:
                    break
                if self.passes > 100000:
                    raise EngineError("move passes did not converge")
            stats = [r.end_step() for r in self.ranks]
            for s in stats[1:]:
                assert s.tobytes() == stats[0].tobytes(), "replicas disagree on the step statistics"
            out[t] = stats[0][0]
        return out

    def _apply_gathered(self, logs):
        import torch
        dev = "cuda:%d" % self.device
        locs = [(torch.as_tensor(_DevView(p, max(n, 1) * ENTRY_BYTES), device=dev), n,
                 torch.as_tensor(_DevView(cp, max(cn, 1) * COMPACT_BYTES), device=dev), cn) for p, n, cp, cn in logs]
        counts = torch.tensor([[n, cn] for _, n, _, cn in locs], dtype=torch.int64, device=dev)
        bufs = []

        def gather_counts(t):
            return counts.reshape(-1)

        def gather_padded(buf):
            bufs.append(buf)
            return None
        for loc, n, loc2, cn in locs:                       # every "rank" packs its buffer exactly as gather_logs does
            g = gather_logs(loc, n, loc2, cn, self.world, gather_counts, gather_padded)
        if g["mx"] == 0 and g["mx2"] == 0:
            return
        g["buf"] = torch.stack(bufs)
        torch.cuda.synchronize()
        for r in self.ranks:
            r.apply_gathered(g)
        torch.cuda.synchronize()

    def download_cells(self):
        cells = [r.eng.download_cells() for r in self.ranks]
        for c in cells[1:]:
            for k in ("sugar", "pollution", "occ"):
                assert np.array_equal(c[k], cells[0][k]), "replicas diverged in " + k
        return cells[0]

    def download_agents(self):
        return self.ranks[0].eng.download_agents()

    def close(self):
        for r in self.ranks:
            r.close()


class DistributedEngine:
    """One rank of a torchrun job (backend nccl).  Interface of Engine.step()."""

    def __init__(self, cfg, device=None, options=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = int(os.environ.get("LOCAL_RANK", self.rank)) if device is None else device
        torch.cuda.set_device(self.device)
        self.r = ShardRank(cfg, self.rank, self.world, device=self.device, options=options)
        self.passes = 0
        self.bytes_sent = 0
        self.t = {"begin": 0.0, "pass": 0.0, "gather": 0.0, "apply+pending": 0.0, "end": 0.0}

    def upload(self, w):
        self.r.upload(w)

    def _gather_counts(self, t):
        out = self.torch.empty(self.world * t.numel(), dtype=self.torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        return out

    def _gather_padded(self, buf):
        out = self.torch.empty((self.world, buf.numel()), dtype=self.torch.uint8, device=buf.device)
        self.dist.all_gather_into_tensor(out.view(-1), buf)
        return out

    def step(self, nsteps=1):
        torch = self.torch
        out = np.zeros(nsteps, dtype=STATS_DTYPE)
        pc = time.perf_counter
        for t in range(nsteps):
            t0 = pc()
            self.r.begin_step()
            t1 = pc(); self.t["begin"] += t1 - t0
            while True:
                t1 = pc()
                ptr, n, cptr, cn = self.r.run_pass()         # engine stream is synchronised on return
                t2 = pc(); self.t["pass"] += t2 - t1
                self.passes += 1
                dev = "cuda:%d" % self.device
                local = torch.as_tensor(_DevView(ptr, max(n, 1) * ENTRY_BYTES), device=dev)
                local2 = torch.as_tensor(_DevView(cptr, max(cn, 1) * COMPACT_BYTES), device=dev)
                g = gather_logs(local, n, local2, cn, self.world, self._gather_counts, self._gather_padded)
                torch.cuda.current_stream().synchronize()    # NCCL results before the engine's stream reads them
                t3 = pc(); self.t["gather"] += t3 - t2
                self.bytes_sent += n * ENTRY_BYTES + cn * COMPACT_BYTES
                self.r.apply_gathered(g)                     # every other rank's slices: two launches per log
                done = self.r.pending() == 0
                self.t["apply+pending"] += pc() - t3
                if done:
                    break
            t4 = pc()
            out[t] = self.r.end_step()[0]
            self.t["end"] += pc() - t4
        return out

    def close(self):
        self.r.close()


# ------------------------------------------------------------------------------ bench ----
def bench_main(args, wl):
    """`torchrun ... bench.py --gpus N`: strong scaling of the same workload over N replicas."""
    # NCCL writes its debug output (the version banner included, whether asked for through the environment or
    # /etc/nccl.conf) to stdout, where the one JSON line belongs
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug_%h_%p.log")
    sys.stdout.flush()
    json_fd = os.dup(1)             # ... and whatever else a library prints: stdout proper is kept for the JSON line only
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from . import config_from_settings, synthetic
    dist.init_process_group(backend="nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    t0 = time.time()
    w = synthetic.tiled_world(wl["n"])
    settings = synthetic.synthetic_settings(wl["n"], pollution=wl["pollution"], utilicalc=wl.get("utilicalc", False))
    cfg = config_from_settings(settings, rng_mode="keyed")
    if rank == 0:
        sys.stderr.write("[bench] world %d^2 generated in %.1fs on every rank\n" % (wl["n"], time.time() - t0))
    K, W = args.steps, args.warmup

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in getattr(args, "opt", [])}
    # page-locked copies of the world (both uploads read them; the pageable originals are dropped: N ranks share one host)
    from .engine import pinned_copy, pinned_empty
    pol_on = bool(wl["pollution"])
    hw = {k: (pinned_copy(w[k], np.float64) if (k != "pollution" or pol_on) else None) for k in ("sugar", "cap", "pollution")}
    hw.update({k: pinned_copy(w[k], np.int32) for k in ("id", "x", "y", "vision", "metab")})
    hw["sugar_agent"] = pinned_copy(w["sugar_agent"], np.float64)
    n_agents0 = len(w["id"])
    w = {"id": hw["id"]}
    de = DistributedEngine(cfg, device=local, options=opts)
    de.upload(hw)
    pop0 = n_agents0
    if W:
        pop0 = int(de.step(W)["population"][-1])
    sync_all()
    from .clocks import ClockSampler
    p0, l0, b0 = de.passes, de.r.eng.counters()["launches"], de.bytes_sent
    for k in de.t:
        de.t[k] = 0.0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # device clock: CUDA events on the (idle) torch stream, the first right after the barrier, the second once
    # every stream of this rank has drained -- the engine stream (kernels) and the torch stream (collectives)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    st = de.step(K)
    torch.cuda.synchronize()
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    wallt = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device="cuda")
    dist.all_reduce(wallt, op=dist.ReduceOp.MAX)
    step_ms = 1e3 * float(wallt.item()) / K
    pops = np.concatenate([[pop0], st["population"][:-1]])
    value = float(pops.sum()) / (step_ms * K * 1e-3)
    # e2e: upload on every rank (page-locked host arrays, as in the single-GPU line) + K steps + download on rank 0
    hout = aout = None
    if rank == 0:
        hout = dict(sugar=pinned_empty((wl["n"], wl["n"])), occ=pinned_empty((wl["n"], wl["n"]), np.int32))
        if pol_on:
            hout["pollution"] = pinned_empty((wl["n"], wl["n"]))
        aout = {k: pinned_empty(len(w["id"]), np.int32) for k in ("id", "x", "y", "vision", "metab")}
        aout["sugar"] = pinned_empty(len(w["id"]), np.float64)
    sync_all()
    t2 = time.perf_counter()
    de2 = DistributedEngine(cfg, device=local, options=opts)
    de2.upload(hw)
    Ke = min(K, args.e2e_steps)
    ste = de2.step(Ke)
    if rank == 0:
        de2.r.eng.download_cells(cap=False, pollution=pol_on, out=hout)
        de2.r.eng.download_agents(out=aout)
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t2], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    popse = np.concatenate([[len(w["id"])], ste["population"][:-1]])
    n = wl["n"]
    if rank == 0:
        # dominant kernel: the move passes of rank 0 (its share of the windows) against the HBM roofline, algorithmic bytes
        # as in the single-GPU line (76+48v / 100+80v B per agent-step, E[v] = 3.5); the exchange against NVLink
        try:
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        per_agent = (100 + 80 * 3.5) if wl["pollution"] else (76 + 48 * 3.5)
        move_bytes = per_agent * float(st["acted"].sum()) / K / world
        pass_s = max(de.t["pass"] / K, 1e-12)
        roofline = {"bound": "hbm", "kernel": "ss_move_solo_kernel, rank 0's share of every pass (host-timed incl. launch + sync)",
                    "achieved": move_bytes / pass_s / 1e9, "peak": peak, "unit": "GB/s", "frac": move_bytes / pass_s / 1e9 / peak,
                    "traffic": None, "peak_source": peak_src, "ms_per_launch_set": 1e3 * pass_s}
        recv = (world - 1) * (de.bytes_sent - b0) / K
        gather_s = max(de.t["gather"] / K, 1e-12)
        nvlink = {"what": "all-gather of the turn logs: bytes received per rank and step / time in the collectives (rank 0)",
                  "achieved": recv / gather_s / 1e9, "peak": 900.0, "unit": "GB/s", "frac": recv / gather_s / 1e9 / 900.0}
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["label"], "grid": n, "agents": len(w["id"]),
                       "parallelism": "replicated lattice, window rows sharded over %d ranks, turn logs all-gathered (NCCL)" % world,
                       "timing": "max over ranks of the CUDA-event interval (device clock) from the barrier before the K steps to the drain of all of the rank's streams after them",
                       "l2": "working set >> 126 MB L2"},
            "cell_updates_per_s": n * n * K / (step_ms * K * 1e-3),
            "population": [int(pop0), int(st["population"][-1])],
            "passes_per_step": (de.passes - p0) / K, "gpu_launches": de.r.eng.counters()["launches"] - l0,
            "move_pass": {"impl": ["cta_per_window", "warp_per_window", "lane_per_agent"][de.r.eng.counters()["pass_impl"]],
                          "window": de.r.eng.counters()["window"]},
            "nccl_bytes_sent_per_rank_per_step": (de.bytes_sent - b0) / K,
            "rank0_ms_per_step": {k: 1e3 * v / K for k, v in de.t.items()},
            "e2e": {"value": float(popse.sum()) / float(e2e_s.item()), "unit": "agent-steps/s",
                    "h2d_bytes_per_step": ((2 + pol_on) * 8 * n * n + 28 * len(w["id"])) / Ke,
                    "d2h_bytes_per_step": (((1 + pol_on) * 8 + 4) * n * n + 28 * len(w["id"]) + 48 * Ke) / Ke, "steps": Ke,
                    "what": "per rank: Engine + upload from page-locked host arrays; K sharded steps with per-step stats read-back; rank 0 downloads cells+agents"},
            "roofline": roofline, "nvlink": nvlink, "cpu_baseline": None,
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    de.close()
    de2.close()
    dist.destroy_process_group()
