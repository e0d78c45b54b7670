"""Host-side mirror of the reference's timestep interface over the C-ABI.

`Engine.step()` is `View.updateGame()` (sugarscape.py:506-686): one call = one timestep; it
returns what the reference appends to `View.population / metabolismMean / visionMean / gini`.
State lives in HBM between calls; `download_*` are the reads the GUI would do.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import STATS_DTYPE, SsConfig
from .settings import config_from_settings

PATH_AUTO, PATH_SERIAL, PATH_LATTICE = -1, 0, 2


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EngineError(RuntimeError):
    pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked host memory (ss_host_alloc): uploads and downloads on such arrays run at PCIe
    speed.  The memory is released when the array (and every view of it) is garbage-collected."""
    import weakref
    lib = capi.lib()
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = lib.ss_host_alloc(max(nbytes, 1))
    if not p:
        raise EngineError("ss_host_alloc(%d) failed (no CUDA device or out of pinned memory)" % nbytes)
    buf = (C.c_char * max(nbytes, 1)).from_address(p)
    weakref.finalize(buf, lib.ss_host_free, C.c_void_p(p))
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pinned_copy(a, dtype=None):
    a = np.asarray(a)
    out = pinned_empty(a.shape, dtype or a.dtype)
    out[...] = a
    return out


class Engine:
    def __init__(self, cfg, device=0, path=PATH_AUTO):
        if not isinstance(cfg, SsConfig):
            cfg = SsConfig.from_buffer_copy(bytes(cfg))
        self._lib = capi.lib()
        self.cfg = cfg
        self.n = cfg.grid_size
        self._h = C.c_void_p()
        rc = self._lib.ss_create(C.byref(cfg), int(device), C.byref(self._h))
        if rc != 0:
            raise EngineError("ss_create: " + self._lib.ss_last_error(None).decode())
        if path != PATH_AUTO:
            self.set_option("path", path)

    @classmethod
    def from_settings(cls, settings, rng="mt", keyed_seed=None, device=0, path=PATH_AUTO, iteration=0, env_time=0):
        return cls(config_from_settings(settings, rng_mode=rng, keyed_seed=keyed_seed, iteration=iteration,
                                        env_time=env_time), device=device, path=path)

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s: %s" % (what, self._lib.ss_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ss_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- uploads ---------------------------------------------------------------------------
    def upload_cells(self, sugar, cap, pollution=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        s, c, p = f(sugar), f(cap), f(pollution)
        if s.size != self.n * self.n or c.size != self.n * self.n or (p is not None and p.size != self.n * self.n):
            raise ValueError("cell arrays must be (N,N)")
        self._check(self._lib.ss_upload_cells(self._h, _ptr(s), _ptr(c), _ptr(p)), "ss_upload_cells")

    def upload_agents(self, id, x, y, vision, metab, sugar):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        id, x, y, vision, metab = map(i32, (id, x, y, vision, metab))
        sugar = np.ascontiguousarray(sugar, dtype=np.float64)
        if not (len(id) == len(x) == len(y) == len(vision) == len(metab) == len(sugar)):
            raise ValueError("agent arrays must have equal length")
        self._check(self._lib.ss_upload_agents(self._h, len(id), _ptr(id), _ptr(x), _ptr(y), _ptr(vision),
                                               _ptr(metab), _ptr(sugar)), "ss_upload_agents")

    def set_rng_mt19937(self, state, index):
        st = np.ascontiguousarray(state, dtype=np.uint32)
        if st.size != 624:
            raise ValueError("MT19937 state has 624 words")
        self._check(self._lib.ss_set_rng_mt19937(self._h, _ptr(st), int(index)), "ss_set_rng_mt19937")

    def get_rng_mt19937(self):
        st = np.empty(624, dtype=np.uint32)
        idx = C.c_int32()
        self._check(self._lib.ss_get_rng_mt19937(self._h, _ptr(st), C.byref(idx)), "ss_get_rng_mt19937")
        return st, idx.value

    def seed_mt19937(self, seed):
        """`random.seed(seed)` of sugarscape.py:162 (non-negative int) on the engine's MT19937 stream."""
        self._check(self._lib.ss_seed_mt19937(self._h, int(seed)), "ss_seed_mt19937")

    def init_world(self, params):
        """The world of the reference's `__main__` (sugarscape.py:1485-1555), built on the device: sugar sites,
        growback to capacity and the agents, every draw from the MT19937 stream.  `params`: SsInitParams
        (settings.init_params_from_settings)."""
        self._check(self._lib.ss_init_world(self._h, C.byref(params)), "ss_init_world")

    def set_rng_keyed(self, seed):
        self._check(self._lib.ss_set_rng_keyed(self._h, int(seed) & ((1 << 64) - 1)), "ss_set_rng_keyed")

    # -- the hot path ----------------------------------------------------------------------
    def step(self, nsteps=1, want_stats=True):
        out = np.zeros(nsteps, dtype=STATS_DTYPE) if want_stats else None
        self._check(self._lib.ss_step(self._h, int(nsteps), _ptr(out)), "ss_step")
        return out

    # -- downloads -------------------------------------------------------------------------
    def agent_count(self):
        return int(self._lib.ss_agent_count(self._h))

    def download_cells(self, sugar=True, cap=True, pollution=True, occ=True, out=None):
        """`out` may hold preallocated (N,N) arrays (e.g. from pinned_empty) under 'sugar', 'cap', 'pollution', 'occ'."""
        n = self.n
        out = out or {}

        def mk(on, key, dt):
            if not on:
                return None
            a = out.get(key)
            if a is None:
                return np.empty((n, n), dtype=dt)
            if a.dtype != dt or a.size != n * n or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a contiguous (N,N) %s array" % (key, np.dtype(dt).name))
            return a
        s, c, p, o = (mk(sugar, "sugar", np.float64), mk(cap, "cap", np.float64), mk(pollution, "pollution", np.float64),
                      mk(occ, "occ", np.int32))
        self._check(self._lib.ss_download_cells(self._h, _ptr(s), _ptr(c), _ptr(p), _ptr(o)), "ss_download_cells")
        return dict(sugar=s, cap=c, pollution=p, occ=o)

    def download_agents(self, out=None):
        """`out` may hold preallocated 1-D arrays of at least agent_count() entries (e.g. from pinned_empty) under
        'id', 'x', 'y', 'vision', 'metab' (int32) and 'sugar' (float64); the result are views of their first entries."""
        k = self.agent_count()
        out = out or {}

        def mk(key, dt):
            a = out.get(key)
            if a is None:
                return np.empty(k, dtype=dt)
            if a.dtype != dt or a.ndim != 1 or a.size < k or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a contiguous 1-D %s array of >= %d entries" % (key, np.dtype(dt).name, k))
            return a[:k]
        id, x, y, vision, metab = (mk(key, np.int32) for key in ("id", "x", "y", "vision", "metab"))
        sugar = mk("sugar", np.float64)
        self._check(self._lib.ss_download_agents(self._h, _ptr(id), _ptr(x), _ptr(y), _ptr(vision), _ptr(metab),
                                                 _ptr(sugar)), "ss_download_agents")
        return dict(id=id, x=x, y=y, vision=vision, metab=metab, sugar=sugar)

    # -- limitedLifespan + procreate ---------------------------------------------
    LIFE_COLUMNS = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")

    def upload_agents_life(self, id_increment, age, max_age, sex, fert_lo, fert_hi, endowment, tags, alive=None):
        """Agent.age / maxAge / sex / fertility[0] / fertility[1] / sugarEndowment / tags / alive of the agents of the last
        upload_agents (same order) and Environment.idIncrement -- required before step() when rules.limitedLifespan or
        rules.procreate is on (agent.py:134-151, 204-206; environment.py:202-204)."""
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        age, max_age, sex, fert_lo, fert_hi, tags, alive = map(i32, (age, max_age, sex, fert_lo, fert_hi, tags, alive))
        endowment = np.ascontiguousarray(endowment, dtype=np.float64)
        if not (len(age) == len(max_age) == len(sex) == len(fert_lo) == len(fert_hi) == len(tags) == len(endowment)):
            raise ValueError("life columns must have equal length")
        self._check(self._lib.ss_upload_agents_life(self._h, len(age), int(id_increment), _ptr(age), _ptr(max_age), _ptr(sex),
                                                    _ptr(fert_lo), _ptr(fert_hi), _ptr(endowment), _ptr(tags), _ptr(alive)),
                    "ss_upload_agents_life")

    def download_agents_life(self):
        """The same columns in View.agents order (an agent that reached maxAge is still listed, alive = 0, for one more
        step -- as in the reference), plus `id_increment`."""
        k = self.agent_count()
        cols = {c: np.empty(k, dtype=np.float64 if c == "endowment" else np.int32) for c in self.LIFE_COLUMNS}
        idinc = C.c_int32()
        self._check(self._lib.ss_download_agents_life(self._h, *[_ptr(cols[c]) for c in self.LIFE_COLUMNS], C.byref(idinc)),
                    "ss_download_agents_life")
        cols["id_increment"] = idinc.value
        return cols

    def events(self):
        """(k, 4) int32 event records of the last step() call in execution order: kind | step << 8, a, b, c (kind 1: a
        mated with b and created c; 2: a died of starvation; 3: a died of old age); `outputs.event_lines` formats them."""
        cnt = C.c_int32()
        self._check(self._lib.ss_event_count(self._h, C.byref(cnt)), "ss_event_count")
        out = np.zeros((cnt.value, 4), dtype=np.int32)
        if cnt.value:
            self._check(self._lib.ss_get_events(self._h, cnt.value, _ptr(out)), "ss_get_events")
        return out

    # -- spice --------------------------------------------------------------------
    def upload_cells_spice(self, spice, spice_cap):
        """Environment.grid slots [1] spice and [3] spice capacity, (N,N) like the sugar arrays."""
        a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (spice, spice_cap))
        if a.size != self.n * self.n or b.size != self.n * self.n:
            raise ValueError("cell arrays must be (N,N)")
        self._check(self._lib.ss_upload_cells_spice(self._h, _ptr(a), _ptr(b)), "ss_upload_cells_spice")

    def download_cells_spice(self):
        a, b = np.empty((self.n, self.n)), np.empty((self.n, self.n))
        self._check(self._lib.ss_download_cells_spice(self._h, _ptr(a), _ptr(b)), "ss_download_cells_spice")
        return dict(spice=a, spice_cap=b)

    def upload_agents_spice(self, spice_metab, spice, endowment=None):
        """Agent.spiceMetabolism / spice / spiceEndowment of the agents of the last upload_agents, same order."""
        m = np.ascontiguousarray(spice_metab, dtype=np.int32)
        sp = np.ascontiguousarray(spice, dtype=np.float64)
        en = None if endowment is None else np.ascontiguousarray(endowment, dtype=np.float64)
        if len(m) != len(sp) or (en is not None and len(en) != len(m)):
            raise ValueError("spice columns must have equal length")
        self._check(self._lib.ss_upload_agents_spice(self._h, len(m), _ptr(m), _ptr(sp), _ptr(en)), "ss_upload_agents_spice")

    def download_agents_spice(self):
        k = self.agent_count()
        m, sp, en = np.empty(k, dtype=np.int32), np.empty(k), np.empty(k)
        self._check(self._lib.ss_download_agents_spice(self._h, _ptr(m), _ptr(sp), _ptr(en)), "ss_download_agents_spice")
        return dict(spice_metab=m, spice=sp, spice_endowment=en)

    # -- instrumentation -------------------------------------------------------------------
    def counters(self):
        out = np.zeros(6, dtype=np.int64)
        self._check(self._lib.ss_get_counters(self._h, _ptr(out), 6), "ss_get_counters")
        return dict(launches=int(out[0]), passes=int(out[1]), steps=int(out[2]), path=int(out[3]),
                    syncs=int(out[4]), window=int(out[5]) & 0xffff, pass_impl=int(out[5]) >> 16)

    def rounds(self):
        """Dependency rounds (grid barriers of ss_dr_rounds_kernel) since the engine's round state was allocated."""
        out = np.zeros(12, dtype=np.int64)
        self._check(self._lib.ss_get_counters(self._h, _ptr(out), 12), "ss_get_counters")
        return int(out[11])

    def n_slots(self):
        """The largest Agent.id the engine is sized for (the keyed execution order is a bijection on a domain of that size:
        a resumed run must keep it, see checkpoint.py)."""
        out = np.zeros(13, dtype=np.int64)
        self._check(self._lib.ss_get_counters(self._h, _ptr(out), 13), "ss_get_counters")
        return int(out[12])

    def health(self):
        """Device-side consistency counters of the lattice path: watchdog firings of the spin waits and internal
        faults (a draw loop that did not terminate, more tied candidates than the geometry allows).  Both must be 0."""
        out = np.zeros(11, dtype=np.int64)
        self._check(self._lib.ss_get_counters(self._h, _ptr(out), 11), "ss_get_counters")
        return dict(watchdog=int(out[6]), fault=int(out[10]))

    def kernel_times(self):
        out = np.zeros(8, dtype=np.float64)
        self._check(self._lib.ss_get_kernel_times(self._h, _ptr(out), 8), "ss_get_kernel_times")
        return dict(zip(("prepare", "move", "stats", "grow", "diffuse", "serial", "last_call", "settle"), out.tolist()))

    def set_option(self, name, value):
        self._check(self._lib.ss_set_option(self._h, name.encode(), int(value)), "ss_set_option")
