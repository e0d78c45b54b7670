"""CPU tests of the host layer: settings loader / rule validation, the C-ABI surface, the
keyed RNG definition, the synthetic landscape and small arithmetic helpers (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import common
from sugarscape_utilicalc_b200 import (ConflictingRuleException, MissingRuleException, UnsupportedRuleError,
                                       config_from_settings, rule_check, synthetic)
from sugarscape_utilicalc_b200 import capi

ROOT = common.ROOT
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import keyed_rng as kr  # noqa: E402


def _settings():
    return synthetic.synthetic_settings(64)


# ---------------------------------------------------------------- settings / ruleCheck ----
def test_rule_check_matches_reference_order_and_messages():
    s = _settings()
    s["rules"]["utilicalc"] = True                       # moveEat + utilicalc  (sugarscape.py:186-187)
    with pytest.raises(ConflictingRuleException) as e:
        rule_check(s)
    assert e.value.message == "moveEat and utilicalc cannot be used together"
    s = _settings()
    s["rules"]["moveEat"], s["rules"]["combat"] = False, True
    with pytest.raises(MissingRuleException) as e:       # combat needs tags  (sugarscape.py:190-191)
        rule_check(s)
    assert e.value.message == "combat is enabled, but needs tags to function properly"
    s = _settings()
    s["rules"]["pollution"], s["rules"]["spice"] = True, True
    with pytest.raises(ConflictingRuleException):        # pollution x spice  (sugarscape.py:195-197)
        rule_check(s)
    for r in ("trade", "foresight", "credit", "inheritance"):
        s = _settings()
        s["rules"][r] = True
        with pytest.raises(MissingRuleException):        # need spice  (sugarscape.py:199-207)
            rule_check(s)


def test_out_of_scope_rules_are_refused_loudly():
    for r in ("tags", "replacement", "transmit", "disease"):
        s = _settings()
        s["rules"][r] = True
        with pytest.raises((UnsupportedRuleError, ConflictingRuleException, MissingRuleException)):
            config_from_settings(s)
    # spice runs with rule M (with utilicalc the reference crashes: KeyError 'proximity')
    s = _settings()
    s["rules"]["spice"] = True
    if s["rules"]["utilicalc"]:
        with pytest.raises(UnsupportedRuleError):
            config_from_settings(s)
    s["rules"]["moveEat"], s["rules"]["utilicalc"] = True, False
    assert config_from_settings(s, rng_mode="mt").rule_spice == 1
    # limitedLifespan / procreate run with rule M in the MT19937-exact mode -- and nowhere else
    for r in ("limitedLifespan", "procreate"):
        s = _settings()
        s["rules"][r] = True
        if s["rules"]["utilicalc"]:
            with pytest.raises(UnsupportedRuleError):    # (with utilicalc the reference needs spice and crashes)
                config_from_settings(s)
        s["rules"]["moveEat"], s["rules"]["utilicalc"] = True, False
        cfg = config_from_settings(s, rng_mode="mt")
        assert (cfg.rule_limited_lifespan, cfg.rule_procreate) == (int(r == "limitedLifespan"), int(r == "procreate"))
        assert cfg.tags_length == s["rule_params"]["tags"]["tags_length"]
        with pytest.raises(UnsupportedRuleError):
            config_from_settings(s, rng_mode="keyed")
    s = _settings()
    s["rules"]["moveEat"], s["rules"]["utilicalc"], s["rules"]["pollution"] = False, True, True
    with pytest.raises(UnsupportedRuleError):            # the reference raises KeyError 'duration' (agent.py:880)
        config_from_settings(s)
    s = _settings()
    s["view"]["grid_size"]["y"] = 65
    with pytest.raises(UnsupportedRuleError):            # Q12: square grids only
        config_from_settings(s)


@pytest.mark.parametrize("name", ["c1_default_utilicalc_mt", "c2_seasons_mt", "c2_pollution_mt"])
def test_config_mapping_of_shipped_presets(name):
    g = common.load_golden(name)
    s = g["settings"]
    cfg = config_from_settings(s, rng_mode="mt")
    assert cfg.grid_size == 50 and cfg.max_capacity == 10.0 and cfg.grow_factor == 1.0
    assert cfg.rule_utilicalc == int(s["rules"]["utilicalc"]) and cfg.rule_move_eat == int(s["rules"]["moveEat"])
    assert cfg.season_off_factor == 1.0 / 8 and cfg.season_period == 50          # sugarscape.py:83-85
    assert list(cfg.season_north) == [0, 0, 49, 24] and list(cfg.season_south) == [0, 25, 49, 49]
    if s["rules"]["pollution"]:
        assert (cfg.pollution_start_time, cfg.diffusion_start_time, cfg.diffusion_rate) == (50, 100, 1)
    assert cfg.keyed_seed == 18 and cfg.rng_mode == capi.RNG_MT19937


# ------------------------------------------------------------------------------ C-ABI ----
def _header_functions():
    src = open(os.path.join(ROOT, "include", "sugarscape_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_exactly_the_bound_symbols():
    assert _header_functions() == sorted(capi.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
    L = capi.lib()                                       # built by __graft_entry__.build(); no GPU needed to load
    for name in _header_functions():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.ss_version()


def test_struct_layouts_match_the_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sugarscape_b200.h"\n#include "ss_oracle.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ss_config), sizeof(ss_step_stats),'
                    ' offsetof(ss_config, grow_factor), offsetof(ss_config, keyed_seed), offsetof(ss_config, env_time),'
                    ' sizeof(sso_config), offsetof(sso_config, keyed_seed));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-o", str(exe),
                           str(prog)])
    out = list(map(int, subprocess.check_output([str(exe)]).split()))
    assert out[0] == C.sizeof(capi.SsConfig) == out[5]
    assert out[1] == C.sizeof(capi.SsStepStats) == capi.STATS_DTYPE.itemsize
    assert out[2] == capi.SsConfig.grow_factor.offset and out[3] == capi.SsConfig.keyed_seed.offset == out[6]
    assert out[4] == capi.SsConfig.env_time.offset


def test_init_params_layout_and_mapping(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sugarscape_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(ss_init_params), offsetof(ss_init_params, group_count),'
                    ' offsetof(ss_init_params, childbearing), offsetof(ss_init_params, foresight_hi));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = list(map(int, subprocess.check_output([str(exe)]).split()))
    P = capi.SsInitParams
    assert out == [C.sizeof(P), P.group_count.offset, P.childbearing.offset, P.foresight_hi.offset]
    from sugarscape_utilicalc_b200 import init_params_from_settings
    import common
    p = init_params_from_settings(common.load_golden("c1_default_moveeat_mt")["settings"])     # the shipped settings.json
    assert (p.n_sites, list(p.site_x)[:2], list(p.site_y)[:2], list(p.site_radius)[:2]) == (2, [35, 15], [15, 35], [18, 18])
    assert (p.n_groups, list(p.group_count)[:2], p.grow_to_capacity) == (2, [150, 150], 1)
    assert (p.max_metabolism, p.max_vision, p.endowment_min, p.endowment_max, p.age_min, p.age_max) == (5, 5, 20, 30, 60, 100)
    assert [list(r) for r in p.childbearing] == [[12, 15, 50, 60], [12, 15, 40, 50]]


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sugarscape_utilicalc_b200.engine import Engine, EngineError
    with pytest.raises(EngineError, match="no CUDA device"):
        Engine(config_from_settings(_settings(), rng_mode="keyed"))


# --------------------------------------------------------------------------- keyed RNG ----
@pytest.mark.parametrize("n_slots", [1, 2, 5, 300, 4096, 70001])
def test_priority_is_a_keyed_bijection(n_slots):
    half = kr.order_half_bits(n_slots)
    dom = 1 << (2 * half)
    assert dom >= n_slots
    rk = kr.round_keys(kr.step_key(18, 7))
    m = min(dom, 5000)
    pr = [kr.priority(rk, half, s) for s in range(m)]
    assert len(set(pr)) == m and all(0 <= p < dom for p in pr)
    assert all(kr.priority_inverse(rk, half, p) == s for s, p in enumerate(pr))
    rk2 = kr.round_keys(kr.step_key(18, 8))
    assert [kr.priority(rk2, half, s) for s in range(min(m, 64))] != pr[:min(m, 64)] or m < 4


@pytest.mark.parametrize("n_slots", [1, 7, 300, 4097, 16777216])
def test_host_side_keyed_priorities_equal_the_keyed_rng_definition(n_slots):
    """outputs.keyed_priorities (numpy; orders a step's death events in keyed runs) against oracle/keyed_rng.py."""
    from sugarscape_utilicalc_b200 import outputs
    rng = np.random.default_rng(n_slots)
    ids = np.unique(rng.integers(1, n_slots + 1, size=min(n_slots, 500)))
    for seed, t in ((18, 0), (18, 41), (2**40 + 3, 7)):
        skey = kr.step_key(seed, t)
        rk, half = kr.round_keys(skey), kr.order_half_bits(n_slots)
        exp = [kr.priority(rk, half, int(i) - 1) for i in ids]
        assert outputs.keyed_priorities(seed, t, ids, n_slots).tolist() == exp


def test_keyed_stream_uses_cpython_randbelow_and_shuffle():
    import random

    class Probe(random.Random):                          # CPython's own algorithms over the keyed words
        def __init__(self, akey):
            super().__init__(0)
            self.st = kr.KeyedStream(akey)

        def getrandbits(self, k):
            return self.st.getrandbits(k)

    akey = kr.agent_key(kr.step_key(18, 3), 42)
    a, b = list(range(23)), list(range(23))
    kr.KeyedStream(akey).shuffle(a)
    Probe(akey).shuffle(b)                               # random.py:350
    assert a == b
    s1, p = kr.KeyedStream(akey), Probe(akey)
    assert [s1.randbelow(n) for n in (1, 2, 3, 7, 24, 1000)] == [p._randbelow(n) for n in (1, 2, 3, 7, 24, 1000)]


def test_mt19937_known_answers_of_the_survey():
    import random
    random.seed(18)
    assert [random.getrandbits(32) for _ in range(4)] == [778526663, 527488476, 2840822562, 1927690951]


# --------------------------------------------------------------------------- synthetic ----
def test_two_peak_capacity_equals_the_reference_landscape():
    g = common.load_golden("c1_default_moveeat_mt")      # cap produced by Environment.addSugarSite (environment.py:114-122)
    assert np.array_equal(synthetic.two_peak_capacity(50), g["init_cap"])


def test_tiled_world_shape_and_uniqueness():
    w = synthetic.tiled_world(130)
    n = 130 * 130 // 16
    assert len(w["id"]) == n and len(set(zip(w["x"].tolist(), w["y"].tolist()))) == n
    assert w["vision"].min() >= 1 and w["vision"].max() <= 6 and w["metab"].max() <= 4
    assert np.array_equal(w["cap"][:50, :50], synthetic.two_peak_capacity(50))
    idx = np.arange(80, 130) % 50
    assert np.array_equal(w["cap"][50:100, 80:130], synthetic.two_peak_capacity(50)[:, idx])
    g = common.load_golden("c4_synth130_nopoll_keyed")
    assert np.array_equal(w["cap"], g["init_cap"]) and np.array_equal(w["x"], g["init_x"])


# --------------------------------------------------------------------------- arithmetic ----
def test_fma_division_by_three_is_correctly_rounded(tmp_path):
    exe = tmp_path / "div3"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "ctests", "div3_check.c"), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_oracle_rejects_bad_input():
    cfg = config_from_settings(_settings(), rng_mode="keyed")
    o = common.oracle_world(cfg)
    z = np.zeros((64, 64))
    o.upload_cells(z, z)
    with pytest.raises(RuntimeError):
        o.upload_agents([1, 1], [0, 1], [0, 1], [1, 1], [1, 1], [5.0, 5.0])      # duplicate id
    with pytest.raises(RuntimeError):
        o.upload_agents([1, 2], [3, 3], [4, 4], [1, 1], [1, 1], [5.0, 5.0])      # two agents on one cell
    with pytest.raises(RuntimeError):
        o.upload_agents([1], [0], [0], [40], [1], [5.0])                         # 2*vision+1 > N
