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
HANDLE_BYTES = 64


def strip_rows(n, rank, world):
    """Rows [lo, hi) of the lattice that rank `rank` of `world` holds (same rule as ss_create_strip)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def strip_of_row(x, n, world):
    """Rank that holds lattice row(s) x."""
    return ((np.asarray(x, dtype=np.int64) + 1) * world - 1) // n


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
