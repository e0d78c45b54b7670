"""ctypes wrapper of the CPU oracle (oracle/libss_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libss_oracle.so")


class SsoConfig(C.Structure):
    _fields_ = [
        ("grid_size", C.c_int32), ("rule_grow", C.c_int32), ("rule_seasons", C.c_int32),
        ("rule_move_eat", C.c_int32), ("rule_utilicalc", C.c_int32), ("rule_can_starve", C.c_int32),
        ("rule_pollution", C.c_int32), ("season_period", C.c_int32),
        ("season_north", C.c_int32 * 4), ("season_south", C.c_int32 * 4),
        ("diffusion_rate", C.c_int32), ("pollution_start_time", C.c_int32),
        ("diffusion_start_time", C.c_int32), ("rng_mode", C.c_int32),
        ("grow_factor", C.c_double), ("season_on_factor", C.c_double), ("season_off_factor", C.c_double),
        ("pollution_a", C.c_double), ("pollution_b", C.c_double), ("max_capacity", C.c_double),
        ("keyed_seed", C.c_uint64), ("iteration", C.c_int64), ("env_time", C.c_int64),
        ("rule_limited_lifespan", C.c_int32), ("rule_procreate", C.c_int32), ("tags_length", C.c_int32),
        ("foresight_lo", C.c_int32), ("foresight_hi", C.c_int32), ("rule_spice", C.c_int32),
    ]


class SsoStepStats(C.Structure):
    _fields_ = [("population", C.c_double), ("metabolism_mean", C.c_double), ("vision_mean", C.c_double),
                ("gini", C.c_double), ("acted", C.c_double), ("deaths", C.c_double)]


STATS_DTYPE = np.dtype([("population", "f8"), ("metabolism_mean", "f8"), ("vision_mean", "f8"),
                        ("gini", "f8"), ("acted", "f8"), ("deaths", "f8")])


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ss_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libss_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB_PATH):
                raise
        L = C.CDLL(_LIB_PATH)
        L.sso_last_error.restype = C.c_char_p
        L.sso_create.argtypes = [C.POINTER(SsoConfig), C.POINTER(C.c_void_p)]
        L.sso_destroy.argtypes = [C.c_void_p]
        L.sso_destroy.restype = None
        vp = C.c_void_p
        L.sso_upload_cells.argtypes = [vp, vp, vp, vp]
        L.sso_upload_agents.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
        L.sso_set_rng_mt19937.argtypes = [vp, vp, C.c_int32]
        L.sso_set_min_slots.argtypes = [vp, C.c_int32]
        L.sso_get_slots.argtypes = [vp]
        L.sso_get_slots.restype = C.c_int32
        L.sso_get_rng_mt19937.argtypes = [vp, vp, C.POINTER(C.c_int32)]
        L.sso_step.argtypes = [vp, C.c_int32, vp]
        L.sso_download_cells.argtypes = [vp, vp, vp, vp, vp]
        L.sso_agent_count.argtypes = [vp]
        L.sso_download_agents.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.sso_upload_agents_life.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
        L.sso_download_agents_life.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.sso_upload_cells_spice.argtypes = [vp, vp, vp]
        L.sso_download_cells_spice.argtypes = [vp, vp, vp]
        L.sso_upload_agents_spice.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.sso_download_agents_spice.argtypes = [vp, vp, vp, vp]
        L.sso_id_increment.argtypes = [vp]
        L.sso_id_increment.restype = C.c_int32
        L.sso_event_count.argtypes = [vp]
        L.sso_event_count.restype = C.c_int32
        L.sso_get_events.argtypes = [vp, C.c_int32, vp]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleWorld:
    """CPU oracle world: same call sequence as the engine's C-ABI."""

    def __init__(self, cfg):
        if not isinstance(cfg, SsoConfig):
            cfg = SsoConfig.from_buffer_copy(bytes(cfg))
        self.cfg = cfg
        self.n = cfg.grid_size
        self._h = C.c_void_p()
        self._check(lib().sso_create(C.byref(cfg), C.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + lib().sso_last_error().decode())

    def close(self):
        if self._h:
            lib().sso_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_cells(self, sugar, cap, pollution=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        s, c, p = f(sugar), f(cap), f(pollution)
        assert s.size == self.n * self.n and c.size == self.n * self.n
        self._check(lib().sso_upload_cells(self._h, _ptr(s), _ptr(c), _ptr(p)))

    def upload_agents(self, id, x, y, vision, metab, sugar):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        id, x, y, vision, metab = map(i32, (id, x, y, vision, metab))
        sugar = np.ascontiguousarray(sugar, dtype=np.float64)
        self._check(lib().sso_upload_agents(self._h, len(id), _ptr(id), _ptr(x), _ptr(y), _ptr(vision),
                                            _ptr(metab), _ptr(sugar)))

    def set_option(self, name, value):
        """Mirror of Engine.set_option for the one option the oracle knows: "global_slots" (order width of a resumed run)."""
        if name != "global_slots":
            raise ValueError("oracle: unknown option " + name)
        self._check(lib().sso_set_min_slots(self._h, int(value)))

    def n_slots(self):
        return int(lib().sso_get_slots(self._h))

    def set_rng_mt19937(self, state, index):
        st = np.ascontiguousarray(state, dtype=np.uint32)
        assert st.size == 624
        self._check(lib().sso_set_rng_mt19937(self._h, _ptr(st), int(index)))

    def get_rng_mt19937(self):
        st = np.empty(624, dtype=np.uint32)
        idx = C.c_int32()
        self._check(lib().sso_get_rng_mt19937(self._h, _ptr(st), C.byref(idx)))
        return st, idx.value

    def step(self, nsteps=1):
        out = np.zeros(nsteps, dtype=STATS_DTYPE)
        self._check(lib().sso_step(self._h, nsteps, _ptr(out)))
        return out

    def download_cells(self):
        n = self.n
        sugar, cap, poll = (np.empty((n, n), dtype=np.float64) for _ in range(3))
        occ = np.empty((n, n), dtype=np.int32)
        self._check(lib().sso_download_cells(self._h, _ptr(sugar), _ptr(cap), _ptr(poll), _ptr(occ)))
        return dict(sugar=sugar, cap=cap, pollution=poll, occ=occ)

    def download_agents(self):
        k = lib().sso_agent_count(self._h)
        id, x, y, vision, metab = (np.empty(k, dtype=np.int32) for _ in range(5))
        sugar = np.empty(k, dtype=np.float64)
        self._check(lib().sso_download_agents(self._h, _ptr(id), _ptr(x), _ptr(y), _ptr(vision), _ptr(metab),
                                              _ptr(sugar)))
        return dict(id=id, x=x, y=y, vision=vision, metab=metab, sugar=sugar)

    # limitedLifespan / procreate
    LIFE_COLUMNS = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")

    def upload_agents_life(self, id_increment, age, max_age, sex, fert_lo, fert_hi, endowment, tags, alive=None):
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        age, max_age, sex, fert_lo, fert_hi, tags, alive = map(i32, (age, max_age, sex, fert_lo, fert_hi, tags, alive))
        endowment = np.ascontiguousarray(endowment, dtype=np.float64)
        self._check(lib().sso_upload_agents_life(self._h, len(age), int(id_increment), _ptr(age), _ptr(max_age), _ptr(sex),
                                                 _ptr(fert_lo), _ptr(fert_hi), _ptr(endowment), _ptr(tags), _ptr(alive)))

    def download_agents_life(self):
        k = lib().sso_agent_count(self._h)
        cols = {c: np.empty(k, dtype=np.float64 if c == "endowment" else np.int32) for c in self.LIFE_COLUMNS}
        self._check(lib().sso_download_agents_life(self._h, *[_ptr(cols[c]) for c in self.LIFE_COLUMNS]))
        cols["id_increment"] = int(lib().sso_id_increment(self._h))
        return cols

    def events(self):
        """(k, 4) int32: kind | step << 8, a, b, c of the last step() call (see ss_oracle.h)."""
        k = int(lib().sso_event_count(self._h))
        out = np.zeros((k, 4), dtype=np.int32)
        if k:
            self._check(lib().sso_get_events(self._h, k, _ptr(out)))
        return out

    # spice
    def upload_cells_spice(self, spice, spice_cap):
        a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (spice, spice_cap))
        self._check(lib().sso_upload_cells_spice(self._h, _ptr(a), _ptr(b)))

    def download_cells_spice(self):
        n = self.n
        a, b = np.empty((n, n)), np.empty((n, n))
        self._check(lib().sso_download_cells_spice(self._h, _ptr(a), _ptr(b)))
        return dict(spice=a, spice_cap=b)

    def upload_agents_spice(self, spice_metab, spice, endowment=None):
        m = np.ascontiguousarray(spice_metab, dtype=np.int32)
        sp = np.ascontiguousarray(spice, dtype=np.float64)
        en = None if endowment is None else np.ascontiguousarray(endowment, dtype=np.float64)
        self._check(lib().sso_upload_agents_spice(self._h, len(m), _ptr(m), _ptr(sp), _ptr(en)))

    def download_agents_spice(self):
        k = lib().sso_agent_count(self._h)
        m, sp, en = np.empty(k, dtype=np.int32), np.empty(k), np.empty(k)
        self._check(lib().sso_download_agents_spice(self._h, _ptr(m), _ptr(sp), _ptr(en)))
        return dict(spice_metab=m, spice=sp, spice_endowment=en)

    def pow_faults(self):
        return int(lib().sso_pow_faults())
