"""GPU parity tests: the CUDA engine, called through the C-ABI, against the golden vectors of
the unmodified reference and against the CPU oracle on the same seeded inputs.

Bar: cells, occupancy, agent positions / sugar / list order, population, means and the
MT19937 state are bit-exact; the Gini coefficient of the lattice path (histogram formula
instead of a sequential float sum) agrees to 1e-9 relative."""
import numpy as np
import pytest

import common
from sugarscape_utilicalc_b200 import config_from_settings

pytestmark = pytest.mark.gpu

MT_CASES = [c for c in common.GOLDEN_CASES if c.endswith("_mt")]
KEYED_CASES = [c for c in common.GOLDEN_CASES if c.endswith("_keyed")]
KEYED_MOVE_CASES = [c for c in KEYED_CASES if "utilicalc" not in c]


@pytest.mark.parametrize("name", MT_CASES)
def test_mt_exact_presets(name):
    """Shipped presets, CPython-exact MT19937 stream: bit-exact trajectories, every step."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "engine")
    assert w.counters()["path"] == 0
    assert common.run_against_golden(g, w, exact_gini=True) == g["steps"]
    w.close()


@pytest.mark.parametrize("name", common.LIFE_CASES)
def test_lifespan_and_procreation_match_reference_golden(name):
    """SURVEY 8 f3 on the GPU (ss_serial_kernel, MT19937-exact): examples/evolution_metabolism (= examples/tag_flipping) as
    shipped and procreation without the lifespan rule, against the goldens recorded from the unmodified reference -- cells,
    agents in list order (newborns act in the step they are born in, an agent that reached maxAge stays for one more turn),
    age / sex / fertility / endowment / tags, Environment.idIncrement, the event lines, the statistics and the final
    MT19937 state."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "engine")
    assert w.counters()["path"] == 0
    assert common.run_against_golden(g, w, exact_gini=True) == g["steps"]
    w.close()


def test_lifespan_and_procreation_single_steps_and_store_growth():
    """... stepping one step at a time gives the same run, and so does a population that outgrows the room the agent store
    was allocated with (the kernel stops before such a step, the host enlarges the store and continues)."""
    g = common.load_golden("f3_evolution_metabolism_mt")
    w = common.world_from_golden(g, "engine")
    common.run_against_golden(g, w, exact_gini=True, chunk=1, max_steps=60)
    w.close()
    cfg = config_from_settings(g["settings"], rng_mode="mt")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    init = common.golden_init(g)
    for w in (e, o):
        common.init_world(w, init, "mt", g["init_mt"], int(g["init_mt_index"]))
    e.set_option("life_capacity_slack", 16)                 # (test hook: next to no room for births)
    common.init_world(e, init, "mt", g["init_mt"], int(g["init_mt_index"]))
    for _ in range(4):
        se, so = e.step(25), o.step(25)
        for f in ("population", "metabolism_mean", "vision_mean", "gini", "acted", "deaths"):
            assert np.array_equal(se[f], so[f]), f
        assert np.array_equal(e.events(), o.events())
        ge, go = e.download_agents(), o.download_agents()
        for k in ("id", "x", "y", "vision", "metab", "sugar"):
            assert np.array_equal(ge[k], go[k]), k
        le, lo = e.download_agents_life(), o.download_agents_life()
        for k in common.LIFE_COLUMNS + ("id_increment",):
            assert np.array_equal(le[k], lo[k]), k
    e.close()
    o.close()


@pytest.mark.parametrize("name", common.SPICE_CASES)
def test_spice_matches_reference_golden(name):
    """SURVEY 8 f2 on the GPU (ss_serial_kernel): examples/trade_price_volume and examples/readme_example without the trade
    rule, against the goldens recorded from the unmodified reference -- spice on the cells and in the agents, the arg-max of the
    Cobb-Douglas welfare with the reference's libm pow restated bit for bit (csrc/ss_pow.h), spice metabolism / starvation,
    statistics, and with the life rules createChild / isFertile in both resources; final MT19937 state."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "engine")
    assert w.counters()["path"] == 0
    assert common.run_against_golden(g, w, exact_gini=True) == g["steps"]
    w.close()


def test_spice_with_keyed_words_equals_oracle():
    """... and with the keyed word source (serial kernel, keyed order by the CTA-wide sort): engine against the oracle."""
    g = common.load_golden("f2_spice_mt")
    cfg = config_from_settings(g["settings"], rng_mode="keyed")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    init = common.golden_init(g)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 0
    for _ in range(6):
        se, so = e.step(10), o.step(10)
        for f in ("population", "metabolism_mean", "vision_mean", "gini", "acted", "deaths"):
            assert np.array_equal(se[f], so[f]), f
        ge, go = e.download_agents(), o.download_agents()
        for k in ("id", "x", "y", "sugar"):
            assert np.array_equal(ge[k], go[k]), k
        assert np.array_equal(e.download_agents_spice()["spice"], o.download_agents_spice()["spice"])
        assert np.array_equal(e.download_cells_spice()["spice"], o.download_cells_spice()["spice"])
    e.close()
    o.close()


def test_spice_single_steps_and_missing_uploads():
    from sugarscape_utilicalc_b200.engine import EngineError
    g = common.load_golden("f2_spice_mt")
    w = common.world_from_golden(g, "engine")
    common.run_against_golden(g, w, exact_gini=True, chunk=1, max_steps=50)
    w.close()
    cfg = config_from_settings(g["settings"], rng_mode="mt")
    e = common.engine_world(cfg)
    common.init_world(e, dict(common.golden_init(g), spice=None), "mt", g["init_mt"], int(g["init_mt_index"]))
    with pytest.raises(EngineError):
        e.step(1)                                           # ss_upload_agents_spice first
    e.close()


@pytest.mark.parametrize("seed,grid,blue,red,spice,lifespan,procreate,ages,start", [
    (3, 20, 60, 60, True, True, True, (4, 10), 0), (77, 14, 90, 90, False, True, True, (60, 100), 1),
    (123, 36, 120, 120, True, False, True, (60, 100), 12), (9, 24, 150, 10, True, True, False, (4, 10), 12),
    (2024, 30, 200, 200, False, False, True, (60, 100), 0), (5, 12, 70, 70, True, True, True, (60, 100), 1),
    (11, 300, 1500, 1500, False, True, True, (60, 100), 12),        # N > 256: serial agents + the lattice's environment kernels
    (12, 280, 900, 900, True, True, True, (60, 100), 12)])          # ... with spice: the serial kernel's own environment phase
def test_random_worlds_with_spice_and_life_rules_engine_vs_oracle(seed, grid, blue, red, spice, lifespan, procreate, ages, start):
    """SURVEY 8 f2 + f3 on small random worlds (short lives, fertility from birth, crowded lattices): the world of `__main__`
    built on the device, handed to the oracle as is, then both stepped -- cells, agents, the rules' fields, event records,
    statistics, MT19937 stream.  (The oracle itself is checked on hundreds of such worlds against the live reference:
    tests/test_oracle_vs_reference.py.)"""
    import copy
    from sugarscape_utilicalc_b200 import init_params_from_settings
    s = copy.deepcopy(common.load_golden("f2_spice_life_mt")["settings"])
    s["rules"].update(spice=spice, limitedLifespan=lifespan, procreate=procreate)
    s["view"]["grid_size"] = {"x": grid, "y": grid}
    s["env"]["random_seed"] = seed
    s["env"]["total_agents"] = {"blue": blue, "red": red}
    for site in s["env"]["sites"].values():
        site["x"], site["y"], site["radius"] = site["x"] % grid, site["y"] % grid, max(2, site["radius"] * grid // 50)
    s["agents"]["max_agent_vision"] = min(s["agents"]["max_agent_vision"], (grid - 1) // 2)
    s["agents"]["death_age_range"] = {"min": ages[0], "max": ages[1]}
    for who in ("male", "female"):
        s["agents"]["fertility_age_ranges"][who].update(min_start=start, max_start=start + 3)
    cfg = config_from_settings(s, rng_mode="mt")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    e.seed_mt19937(seed)
    e.init_world(init_params_from_settings(s))
    cells, ags = e.download_cells(), e.download_agents()
    init = dict(sugar=cells["sugar"], cap=cells["cap"], pollution=cells["pollution"], id=ags["id"], x=ags["x"], y=ags["y"],
                vision=ags["vision"], metab=ags["metab"], sugar_agent=ags["sugar"])
    if spice:
        cs, sp = e.download_cells_spice(), e.download_agents_spice()
        init["spice"] = dict(spice=cs["spice"], spice_cap=cs["spice_cap"], spice_metab=sp["spice_metab"], spice_agent=sp["spice"],
                             spice_endowment=sp["spice_endowment"])
    if lifespan or procreate:
        lf = e.download_agents_life()
        init["life"], init["id_increment"] = {k: lf[k] for k in common.LIFE_COLUMNS}, lf["id_increment"]
    mt, idx = e.get_rng_mt19937()
    common.init_world(o, init, "mt", mt, idx)
    for t in range(15):
        se, so = e.step(2), o.step(2)
        for f in ("population", "metabolism_mean", "vision_mean", "gini", "acted", "deaths"):
            assert np.array_equal(se[f], so[f]), (t, f)
        ge, go, ce, co = e.download_agents(), o.download_agents(), e.download_cells(), o.download_cells()
        for k in ("id", "x", "y", "vision", "metab", "sugar"):
            assert np.array_equal(ge[k], go[k]), (t, k)
        assert np.array_equal(ce["occ"], co["occ"]) and np.array_equal(ce["sugar"], co["sugar"]), t
        if spice:
            a, b = e.download_agents_spice(), o.download_agents_spice()
            for k in ("spice_metab", "spice", "spice_endowment"):
                assert np.array_equal(a[k], b[k]), (t, k)
            assert np.array_equal(e.download_cells_spice()["spice"], o.download_cells_spice()["spice"]), t
        if lifespan or procreate:
            a, b = e.download_agents_life(), o.download_agents_life()
            for k in common.LIFE_COLUMNS + ("id_increment",):
                assert np.array_equal(a[k], b[k]), (t, k)
            assert np.array_equal(e.events(), o.events()), t
        me, mo = e.get_rng_mt19937(), o.get_rng_mt19937()
        assert me[1] == mo[1] and np.array_equal(me[0], mo[0]), t
    e.close()
    o.close()


def test_life_rules_need_their_columns_and_the_mt_mode():
    from sugarscape_utilicalc_b200.engine import EngineError
    g = common.load_golden("f3_evolution_metabolism_mt")
    cfg = config_from_settings(g["settings"], rng_mode="mt")
    e = common.engine_world(cfg)
    init = dict(common.golden_init(g), life=None)
    common.init_world(e, init, "mt", g["init_mt"], int(g["init_mt_index"]))
    with pytest.raises(EngineError):
        e.step(1)                                           # ss_upload_agents_life first
    e.close()
    cfg.rng_mode = 1
    with pytest.raises(EngineError):
        common.engine_world(cfg)


@pytest.mark.parametrize("name", common.OUTPUT_CASES)
def test_on_disk_outputs_match_reference_golden(name, tmp_path):
    """SURVEY 8 f1 on the GPU: event log in execution order, plateau exit, data.txt and createLogFile text."""
    common.check_outputs_against_golden(name, "engine", tmp_path)


@pytest.mark.parametrize("name", MT_CASES + common.LIFE_CASES + common.SPICE_CASES)
def test_world_initialisation_on_device_equals_reference(name):
    """`random.seed(seed)` + the reference's `__main__` set-up (sugar sites, growback to capacity,
    Agent() + initAgent per agent) run on the device -- capacity / sugar fields, every agent's id, cell, vision,
    metabolism and endowment and the MT19937 state afterwards equal what the unmodified reference built (the golden
    initial state); the run that follows stays on the golden trajectory."""
    from sugarscape_utilicalc_b200 import init_params_from_settings
    g = common.load_golden(name)
    cfg = config_from_settings(g["settings"], rng_mode="mt")
    e = common.engine_world(cfg)
    e.seed_mt19937(g["settings"]["env"]["random_seed"])
    e.init_world(init_params_from_settings(g["settings"]))
    cells, ags = e.download_cells(), e.download_agents()
    assert np.array_equal(cells["cap"], g["init_cap"]) and np.array_equal(cells["sugar"], g["init_sugar"])
    assert np.array_equal(cells["pollution"], g["init_pollution"])
    for key, gk in (("id", "init_order"), ("x", "init_x"), ("y", "init_y"), ("vision", "init_vision"), ("metab", "init_metab"),
                    ("sugar", "init_sugar_agent")):
        assert np.array_equal(ags[key], g[gk]), key
    mt, idx = e.get_rng_mt19937()
    assert np.array_equal(mt, g["init_mt"]) and idx == int(g["init_mt_index"])
    if name in common.SPICE_CASES:      # the spice peaks and every agent's spiceMetabolism / spice endowment
        cs, sp = e.download_cells_spice(), e.download_agents_spice()
        assert np.array_equal(cs["spice"], g["init_spice"]) and np.array_equal(cs["spice_cap"], g["init_spice_cap"])
        assert np.array_equal(sp["spice_metab"], g["init_spice_metab"]) and np.array_equal(sp["spice"], g["init_spice_agent"])
        assert np.array_equal(sp["spice_endowment"], g["init_spice_endowment"])
    if "init_life_age" in g:            # the drawn maxAge, sex and fertility and the groups' tags are kept
        lf = e.download_agents_life()
        for c in common.LIFE_COLUMNS:
            assert np.array_equal(lf[c], g["init_life_" + c]), c
        assert lf["id_increment"] == int(g["init_id_increment"])
    common.run_against_golden(g, e, exact_gini=True, max_steps=40)
    e.close()


def test_headless_run_with_births_equals_the_reference_log(tmp_path, monkeypatch):
    """`python -m sugarscape_utilicalc_b200 examples/evolution_metabolism/settings.json --steps 120`: the event log of the
    run (mating lines and deaths in execution order) equals the reference's."""
    import json
    from sugarscape_utilicalc_b200 import __main__ as runner, outputs
    g = common.load_golden("f3_evolution_metabolism_mt")
    (tmp_path / "settings.json").write_text(json.dumps(g["settings"]))
    monkeypatch.chdir(tmp_path)
    seen = {}
    real = outputs.RunRecorder

    class Spy(real):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            seen["rec"] = self
    monkeypatch.setattr(outputs, "RunRecorder", Spy)
    assert runner.main(["settings.json", "--quiet", "--steps", "120"]) == 0
    rec = seen["rec"]
    assert rec.population == [int(p) for p in g["population"][:120]]
    assert rec.log == str(g["log_text"]).split("\n")[:int(g["cp100_log_lines"])] + rec.log[int(g["cp100_log_lines"]):]
    assert len(rec.log) > int(g["cp100_log_lines"])


def test_headless_run_of_settings_json_equals_reference_outputs(tmp_path, monkeypatch):
    """`python -m sugarscape_utilicalc_b200 settings.json`: seed, device-side world, steps until the plateau exit --
    data.txt column and event log equal the reference's run of the same file (tests/golden/outputs.json)."""
    import json
    import os
    from sugarscape_utilicalc_b200 import __main__ as runner
    g = common.load_golden("c1_default_moveeat_mt")
    with open(os.path.join(common.GOLDEN, "outputs.json")) as f:
        exp = json.load(f)["c1_default_moveeat_mt"]
    s = dict(g["settings"])
    s["view"] = {**s["view"], "create_log_on_exit": True}
    (tmp_path / "settings.json").write_text(json.dumps(s))
    monkeypatch.chdir(tmp_path)
    assert runner.main(["settings.json", "--quiet"]) == 0
    assert (tmp_path / "data_collection" / "data.txt").read_text() == exp["data_txt"]
    logs = list((tmp_path / "logs").glob("log_*.txt"))
    assert len(logs) == 1
    text = logs[0].read_text()
    assert text[text.index("\nStat lists:"):] == exp["log_file_text"][exp["log_file_text"].index("\nStat lists:"):]


def test_world_initialisation_fills_the_lattice_and_large_grids():
    """More agents than free cells of range(0, N-1)^2: the surplus agents keep their ids and are dropped
    (initAgent returns False); and a 1000^2 lattice goes through the two-level free-cell search."""
    from sugarscape_utilicalc_b200 import init_params_from_settings
    s = common.load_golden("c1_default_moveeat_mt")["settings"]
    s = {**s, "view": {**s["view"], "grid_size": {"x": 8, "y": 8}}, "env": {**s["env"], "total_agents": {"blue": 40, "red": 20}}}
    for k in ("ne", "sw"):
        s["env"]["sites"] = {**s["env"]["sites"], k: {"x": 3, "y": 4, "radius": 3}}
    s["agents"] = {**s["agents"], "max_agent_vision": 3}
    cfg = config_from_settings(s, rng_mode="mt")
    e = common.engine_world(cfg)
    e.seed_mt19937(7)
    e.init_world(init_params_from_settings(s))
    ags = e.download_agents()
    assert len(ags["id"]) == 49 and ags["id"].max() == 49 and ags["x"].max() == 6 and ags["y"].max() == 6
    assert len(set(zip(ags["x"].tolist(), ags["y"].tolist()))) == 49
    e.close()
    s = common.load_golden("c1_default_moveeat_mt")["settings"]
    s = {**s, "view": {**s["view"], "grid_size": {"x": 1000, "y": 1000}}, "env": {**s["env"], "total_agents": {"blue": 30000, "red": 20000}}}
    cfg = config_from_settings(s, rng_mode="keyed")
    e = common.engine_world(cfg)
    import random
    for seed in (0, 18, 2**32 - 1, 2**40 + 5):            # random.seed(int): init_by_array over the 32-bit words
        e.seed_mt19937(seed)
        st = random.Random(seed).getstate()[1]
        mt, idx = e.get_rng_mt19937()
        assert idx == st[624] and np.array_equal(mt, np.array(st[:624], dtype=np.uint32)), seed
    e.init_world(init_params_from_settings(s))
    ags = e.download_agents()
    assert len(ags["id"]) == 50000 and len(set((ags["x"].astype(np.int64) * 1000 + ags["y"]).tolist())) == 50000
    assert ags["x"].max() <= 998 and ags["vision"].min() >= 1 and ags["vision"].max() <= 5
    assert e.counters()["path"] == 2
    e.step(3)
    e.close()


@pytest.mark.parametrize("name", ["c2_pollution_mt", "c2_seasons_keyed", "c1_default_utilicalc_mt", "c4_synth130_nopoll_keyed", "f3_evolution_metabolism_mt", "f2_spice_life_mt"])
def test_checkpoint_resume(name, tmp_path):
    """checkpoint.save / restore over the engine: stop at step 7, resume from the file in a fresh engine, same end state."""
    common.check_checkpoint_resume(name, "engine", tmp_path, total=min(15, common.load_golden(name)["steps"]))


def test_checkpoint_resume_keeps_the_order_width_when_the_top_id_died(tmp_path):
    common.check_resume_after_top_id_died("engine", tmp_path)


def test_mt_exact_single_steps_equal_fused_steps():
    g = common.load_golden("c2_pollution_mt")
    w = common.world_from_golden(g, "engine")
    common.run_against_golden(g, w, exact_gini=True, chunk=1, max_steps=120)
    w.close()


@pytest.mark.parametrize("name", KEYED_CASES)
def test_keyed_serial_path(name):
    g = common.load_golden(name)
    w = common.world_from_golden(g, "engine", path=0)
    assert w.counters()["path"] == 0
    common.run_against_golden(g, w, exact_gini=True)
    w.close()


@pytest.mark.parametrize("impl", [3, 0])
@pytest.mark.parametrize("name", KEYED_MOVE_CASES)
def test_keyed_lattice_path(name, impl):
    """Parallel dependency rounds reproduce the sequential loop of the reference (keyed words): the whole-lattice
    rounds of ss_rounds.cu (pass_impl 3, the default) and the window passes (pass_impl 0) on the reference's goldens."""
    g = common.load_golden(name)
    cfg = config_from_settings(g["settings"], rng_mode=g["rng"])
    w = common.engine_world(cfg, path=2, options={"pass_impl": impl})
    common.init_world(w, common.golden_init(g), g["rng"], g["init_mt"], int(g["init_mt_index"]))
    assert w.counters()["path"] == 2 and w.counters()["pass_impl"] == impl
    common.run_against_golden(g, w, exact_gini=False)
    w.close()


@pytest.mark.parametrize("n,pollution,steps,vision", [
    (64, True, 12, (1, 6)), (177, True, 10, (1, 6)), (200, False, 10, (1, 6)), (257, True, 8, (1, 6)),
    (300, True, 8, (1, 10)), (512, True, 8, (1, 6)), (1000, True, 5, (1, 6)), (1024, False, 6, (1, 6)),
])
@pytest.mark.parametrize("impl", ["dependency_rounds", "dependency_rounds_no_tma", "dependency_rounds_with_barriers", "cta_per_window", "warp_per_window",
                                  "lane_per_agent"])
def test_lattice_vs_oracle_synthetic(n, pollution, steps, vision, impl):
    """The agent-phase kernels against the CPU oracle: the dependency rounds over the whole lattice (ss_dr_build_kernel +
    ss_dr_rounds_kernel, the default; also with the TMA tile load switched off) and the move-pass kernels
    (ss_move_pass_kernel: a CTA per window, a warp per region; ss_move_solo_kernel: a warp per window, here with windows
    of at most 96 cells so that small lattices have several, taking its agents one at a time or -- pass_impl 2 -- 32 at
    a time, one lane each)."""
    s, init = common.synthetic_case(n, pollution, vision=vision)
    cfg = config_from_settings(s, rng_mode="keyed")
    solo = {"warp_per_window": 1, "lane_per_agent": 2}.get(impl, 0)
    if impl.startswith("dependency_rounds"):
        # (default: the rounds kernel as a work list; "with_barriers": one grid barrier per dependency round, as the strips run it)
        opts = {"pass_impl": 3, "tma": 0 if impl.endswith("no_tma") else 1, "async_queue": 0 if impl.endswith("with_barriers") else 1}
    else:
        opts = {"pass_impl": 0} if not solo else {"pass_impl": solo, "solo_window": 96 if vision[1] <= 6 else 120}
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 2
    assert e.counters()["pass_impl"] == (3 if "pass_impl" in opts and opts["pass_impl"] == 3 else (solo if n >= 177 else 0))
    # (the default path in calls of two steps: statistics / growback / diffusion of a step then run beside the conflict graph
    # and -- the sweep -- beside the agent phase of the next one)
    common.compare_worlds(e, o, steps, every=2 if impl == "dependency_rounds" else 1, what="synthetic %d %s" % (n, impl))
    e.close()
    o.close()


@pytest.mark.parametrize("opts", [{"pass_impl": 3}, {"pass_impl": 0}, {"pass_impl": 1, "solo_window": 72}, {"pass_impl": 1, "solo_window": 128},
                                  {"pass_impl": 2, "solo_window": 72}])
def test_lattice_dense_population(opts):
    """density 1/3: deep dependency chains, many ties, many deaths (Q1 skips); density 0.9 overflows the key lists
    of the warp-per-window kernel (priority-threshold trimming)."""
    s, init = common.synthetic_case(150, True, density=1.0 / 3.0)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(e, o, 15, what="dense 150")
    s, init = common.synthetic_case(400, False, density=1.0 / 3.0)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["pass_impl"] == opts["pass_impl"]
    common.compare_worlds(e, o, 10, what="dense 400")
    s, init = common.synthetic_case(256, True, density=0.9, vision=(1, 4))
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(e, o, 4, what="dense 256 (0.9)")


@pytest.mark.parametrize("impl", ["dependency_rounds", "warp_per_window"])
@pytest.mark.parametrize("n,density,vision,steps,window", [(260, 1 / 16, (1, 6), 10, 96), (300, 1 / 8, (1, 6), 8, 96),
                                                            (512, 1 / 16, (1, 10), 6, 128), (257, 1 / 3, (1, 4), 8, 96),
                                                            (1000, 1 / 16, (1, 6), 4, 0), (64, 1 / 4, (1, 15), 6, 0)])
def test_lattice_utilicalc_vs_oracle(n, density, vision, steps, window, impl):
    """The utilicalc rule (agent.py:1025-1141), keyed words: the welfare evaluation fused into the movement scan, every
    agent settled on the spot -- as dependency rounds over the whole lattice (one thread per turn, the default) and in the
    warp-per-window kernel (one warp per turn); bit-exact against the CPU oracle (positions, sugar, deaths incl. the Q1
    skips, statistics)."""
    s, init = common.synthetic_case(n, False, density=density, vision=vision, utilicalc=True)
    cfg = config_from_settings(s, rng_mode="keyed")
    if impl == "dependency_rounds":
        opts, want = {"pass_impl": 3}, 3
    else:
        if n < 177:
            pytest.skip("a single window")
        opts, want = dict({"pass_impl": 1}, **({"solo_window": window} if window else {})), 1
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 2 and e.counters()["pass_impl"] == want
    common.compare_worlds(e, o, steps, what="utilicalc %d %s" % (n, impl))
    e.close()
    o.close()


@pytest.mark.parametrize("name", [c for c in KEYED_CASES if "utilicalc" in c])
def test_keyed_utilicalc_rounds_on_reference_golden(name):
    """The utilicalc turn of the dependency rounds on the goldens recorded from the unmodified reference (keyed words
    served through the shim): the shipped 50 x 50 presets run on the lattice path too."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "engine")
    assert w.counters()["path"] == 2 and w.counters()["pass_impl"] == 3
    common.run_against_golden(g, w, exact_gini=False)
    w.close()


def test_lattice_utilicalc_equals_serial_kernel():
    s, init = common.synthetic_case(400, False, utilicalc=True)
    cfg = config_from_settings(s, rng_mode="keyed")
    for opts, impl in (({}, 3), ({"pass_impl": 1}, 1)):
        a, b = common.engine_world(cfg, options=opts), common.engine_world(cfg, path=0)
        common.init_world(a, init, "keyed")
        common.init_world(b, init, "keyed")
        assert (a.counters()["path"], b.counters()["path"], a.counters()["pass_impl"]) == (2, 1, impl)
        common.compare_worlds(a, b, 6 if impl == 3 else 3, exact_gini=False, what="utilicalc lattice (%d) vs serial" % impl)
        a.close()
        b.close()


@pytest.mark.parametrize("opts", [{"pass_impl": 3}, {"pass_impl": 0}, {"pass_impl": 1, "solo_window": 72}, {"pass_impl": 2, "solo_window": 72}])
def test_lattice_seasons(opts):
    s, init = common.synthetic_case(260, False, seasons=True)
    s["rule_params"]["seasons"]["period"] = 3
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options=opts), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(e, o, 9, what="seasons 260")


@pytest.mark.parametrize("n", [700, 129, 177, 193, 256, 257, 1025])
def test_diffusion_wavefront_equals_diagonal_sweep(n):
    """The three diffusion kernels agree bit for bit: 0 = column-chain wavefront (128-row bands, a compute warp and
    three staging warps per CTA; default); 2 = row-wise wavefront (32-row bands, one warp); 1 = anti-diagonal sweep of
    one CTA (the simple restatement).  Sizes: several bands, partial last bands of 1..65 rows, ring-sized lattices."""
    s, init = common.synthetic_case(n, True)
    cfg = config_from_settings(s, rng_mode="keyed")
    worlds = [common.engine_world(cfg, options={"diffuse_impl": k}) for k in (0, 1, 2)]
    for w in worlds:
        common.init_world(w, init, "keyed")
        w.step(4)
    ref = worlds[1].download_cells()["pollution"]
    assert np.array_equal(worlds[0].download_cells()["pollution"], ref)
    assert np.array_equal(worlds[2].download_cells()["pollution"], ref)
    for w in worlds:
        w.close()


def test_mt_serial_env_path_large_grid():
    """MT19937 mode on a grid too large for the fused kernel: serial agents + lattice env kernels."""
    s, init = common.synthetic_case(300, True, density=1.0 / 64.0)
    cfg = config_from_settings(s, rng_mode="mt")
    mt = np.arange(624, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(12345)
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    common.init_world(e, init, "mt", mt, 624)
    common.init_world(o, init, "mt", mt, 624)
    assert e.counters()["path"] == 1
    common.compare_worlds(e, o, 5, exact_gini=True, what="mt 300")
    me, ie = e.get_rng_mt19937()
    mo, io = o.get_rng_mt19937()
    assert ie == io and np.array_equal(me, mo)


def test_unsupported_configurations_fail_loudly():
    from sugarscape_utilicalc_b200.engine import Engine, EngineError
    s, init = common.synthetic_case(64, True)
    cfg = config_from_settings(s, rng_mode="keyed")
    cfg.rule_utilicalc, cfg.rule_move_eat = 1, 0
    with pytest.raises(EngineError):
        Engine(cfg)                       # utilicalc + pollution: the reference crashes
    cfg = config_from_settings(s, rng_mode="keyed")
    e = Engine(cfg)
    with pytest.raises(EngineError):
        e.step(1)                         # nothing uploaded
    bad = dict(init)
    bad["x"] = init["x"].copy(); bad["y"] = init["y"].copy()
    bad["x"][1], bad["y"][1] = bad["x"][0], bad["y"][0]
    e.upload_cells(init["sugar"], init["cap"], init["pollution"])
    with pytest.raises(EngineError):
        e.upload_agents(bad["id"], bad["x"], bad["y"], bad["vision"], bad["metab"], bad["sugar_agent"])


@pytest.mark.parametrize("impl", [3, 0, 1])
def test_gini_with_wealth_outside_the_histogram(impl):
    """fractional endowments (non-integer wealth) and wealth >= 65536 have no histogram bin: the statistics stay
    exact (outliers merged in value order) instead of reporting NaN."""
    s, init = common.synthetic_case(200, False)
    init = dict(init)
    sa = init["sugar_agent"].copy()
    sa[::7] += 0.375
    sa[5] = 70000.0
    sa[11] = 123456.5
    init["sugar_agent"] = sa
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options={"pass_impl": impl, "solo_window": 72}), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    a, b = e.step(6), o.step(6)
    assert np.isfinite(a["gini"]).all()
    np.testing.assert_allclose(a["gini"], b["gini"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(a["population"], b["population"])
    e.close()
    o.close()


def test_empty_population_and_extinction():
    s, init = common.synthetic_case(64, False)
    cfg = config_from_settings(s, rng_mode="keyed")
    for path, impl in ((0, -1), (2, 3), (2, 0)):
        e, o = common.engine_world(cfg, path, options={"pass_impl": impl}), common.oracle_world(cfg)
        for w in (e, o):
            w.upload_cells(np.zeros((64, 64)), np.zeros((64, 64)), None)      # barren: everyone starves
            w.upload_agents(init["id"][:40], init["x"][:40], init["y"][:40], init["vision"][:40], init["metab"][:40],
                            np.full(40, 3.0))
        common.compare_worlds(e, o, 14, what="extinction path %d" % path)
        assert e.agent_count() == len(o.download_agents()["id"]) == 0


@pytest.mark.parametrize("opts", [{"pass_impl": 0}, {"pass_impl": 1, "solo_window": 72}, {"pass_impl": 2, "solo_window": 72}])
@pytest.mark.parametrize("world,n,pollution", [(2, 300, True), (3, 512, False), (4, 130, True)])
def test_multi_gpu_replicas_equal_single_gpu_and_oracle(world, n, pollution, opts):
    """The N>1 code path (window rows sharded over replicas, turn logs replayed) on one GPU with
    virtual ranks: replicas stay identical to each other and to the CPU oracle."""
    from sugarscape_utilicalc_b200.sharded import VirtualRanks
    s, init = common.synthetic_case(n, pollution)
    cfg = config_from_settings(s, rng_mode="keyed")
    v, o = VirtualRanks(cfg, world, options=opts), common.oracle_world(cfg)
    common.init_world(v, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(v, o, 6, what="virtual ranks %d" % world)
    v.close()
    o.close()


@pytest.mark.parametrize("world,n,pollution", [(4, 300, True), (8, 512, False)])
def test_multi_gpu_gathered_replay(world, n, pollution):
    """The replay path of the real multi-GPU run: all ranks' log slices in the all-gather layout, replayed by
    ss_shard_apply_gathered (one launch per log and phase for all other ranks)."""
    from sugarscape_utilicalc_b200.sharded import VirtualRanks
    s, init = common.synthetic_case(n, pollution)
    cfg = config_from_settings(s, rng_mode="keyed")
    v, o = VirtualRanks(cfg, world, options={"solo_window": 72}, gathered=True), common.oracle_world(cfg)
    common.init_world(v, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(v, o, 6, what="virtual ranks %d gathered" % world)
    v.close()
    o.close()


def test_multi_gpu_replicas_utilicalc():
    """The utilicalc rule through the sharded path: every turn travels as a fully settled log entry."""
    from sugarscape_utilicalc_b200.sharded import VirtualRanks
    s, init = common.synthetic_case(300, False, utilicalc=True)
    cfg = config_from_settings(s, rng_mode="keyed")
    v, o = VirtualRanks(cfg, 3, options={"solo_window": 96}), common.oracle_world(cfg)
    common.init_world(v, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(v, o, 6, what="virtual ranks utilicalc")
    v.close()
    o.close()


def test_multi_gpu_eight_replicas():
    """World size 8 (the largest the bench is launched with): 16 window rows, two per rank."""
    from sugarscape_utilicalc_b200.sharded import VirtualRanks
    s, init = common.synthetic_case(1024, False)
    cfg = config_from_settings(s, rng_mode="keyed")
    v, o = VirtualRanks(cfg, 8, options={"pass_impl": 1, "solo_window": 64}), common.oracle_world(cfg)
    common.init_world(v, init, "keyed")
    common.init_world(o, init, "keyed")
    assert v.ranks[0].eng.counters()["pass_impl"] == 1
    common.compare_worlds(v, o, 4, what="virtual ranks 8")
    v.close()
    o.close()


def test_multi_gpu_replicas_on_reference_golden():
    from sugarscape_utilicalc_b200.sharded import VirtualRanks
    g = common.load_golden("c2_pollution0_keyed")
    cfg = config_from_settings(g["settings"], rng_mode="keyed")
    v = VirtualRanks(cfg, 2)
    common.init_world(v, common.golden_init(g), "keyed")
    common.run_against_golden(g, v, exact_gini=False, max_steps=50)
    v.close()


@pytest.mark.parametrize("world,n,density,vision,seasons", [(2, 300, 1 / 16, (1, 6), False), (3, 512, 1 / 16, (1, 6), False),
                                                         (4, 600, 1 / 8, (1, 6), True), (8, 1024, 1 / 16, (1, 6), False),
                                                         (2, 260, 1 / 3, (1, 4), False), (4, 640, 1 / 16, (1, 10), False)])
def test_row_strips_equal_oracle(world, n, density, vision, seasons):
    """The multi-GPU decomposition (BASELINE config 5: row strips, halo reads, agent migration over peer memory) with
    the strips of `world` ranks on one GPU: each strip holds only its rows and its agents, the kernels are the ones of
    the real run (ss_dr_build_kernel / ss_dr_rounds_kernel with STRIPS, delta lists, merged statistics), the host stands
    in for the device-side barriers.  Bit-exact against the CPU oracle, every step."""
    from sugarscape_utilicalc_b200.strips import VirtualStrips
    s, init = common.synthetic_case(n, False, density=density, vision=vision, seasons=seasons)
    if seasons:
        s["rule_params"]["seasons"]["period"] = 3
    cfg = config_from_settings(s, rng_mode="keyed")
    v, o = VirtualStrips(cfg, world), common.oracle_world(cfg)
    common.init_world(v, init, "keyed")
    common.init_world(o, init, "keyed")
    common.compare_worlds(v, o, 8, what="row strips %d" % world)
    assert sum(e.rows for e in v.ranks) == n and v.agent_count() == len(o.download_agents()["id"])
    v.close()
    o.close()


def test_row_strips_on_reference_golden():
    from sugarscape_utilicalc_b200.strips import VirtualStrips
    g = common.load_golden("c4_synth130_nopoll_keyed")
    cfg = config_from_settings(g["settings"], rng_mode="keyed")
    v = VirtualStrips(cfg, 2)
    common.init_world(v, common.golden_init(g), "keyed")
    common.run_against_golden(g, v, exact_gini=False)
    v.close()


def test_full_size_c4_against_oracle():
    """BASELINE config 4 at full size (4096 x 4096, 1,048,576 agents, pollution with diffusion every step):
    the lattice path against the CPU oracle, bit-exact, three whole timesteps."""
    s, init = common.synthetic_case(4096, True)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 2 and e.counters()["pass_impl"] == 3      # dependency rounds over the whole lattice
    common.compare_worlds(e, o, 3, every=3, what="C4 full size")       # one call: the sweep of a step beside the next agent phase
    e.close()
    o.close()


@pytest.mark.timeout(900)
def test_full_size_c5_against_oracle():
    """BASELINE config 5 at full size (16384 x 16384, 16,777,216 agents) -- the configuration the bench line is quoted on:
    the engine's default path against the CPU oracle, bit-exact on cell sugar, occupancy, the agents (list order,
    positions, sugar) and the statistics, two whole timesteps (the oracle needs ~15 s per step)."""
    n = 16384
    s, init = common.synthetic_case(n, False)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 2 and e.counters()["pass_impl"] == 3
    for t in range(2):
        sa, sb = e.step(1), o.step(1)
        for f in ("population", "metabolism_mean", "vision_mean", "acted", "deaths"):
            assert np.array_equal(sa[f], sb[f]), (f, t, sa[f], sb[f])
        np.testing.assert_allclose(sa["gini"], sb["gini"], rtol=1e-9, atol=1e-12)
        ga, gb = e.download_agents(), o.download_agents()
        for k in ("id", "x", "y", "vision", "metab", "sugar"):
            assert np.array_equal(ga[k], gb[k]), "step %d: agents differ in %s" % (t + 1, k)
        del ga, gb
        cb = o.download_cells()
        ca = e.download_cells(cap=False, pollution=False)
        assert np.array_equal(ca["occ"], cb["occ"]), "step %d: occupancy differs" % (t + 1)
        assert np.array_equal(ca["sugar"], cb["sugar"]), "step %d: cell sugar differs" % (t + 1)
        del ca, cb
    assert e.health() == dict(watchdog=0, fault=0)
    e.close()
    o.close()


@pytest.mark.timeout(600)
def test_full_size_c4_utilicalc_against_oracle():
    """The utilicalc rule at the C4 size (4096 x 4096, 1,048,576 agents; `c4u` of bench.py): two whole timesteps against
    the CPU oracle, bit-exact."""
    s, init = common.synthetic_case(4096, False, utilicalc=True)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    assert e.counters()["path"] == 2
    common.compare_worlds(e, o, 2, what="c4u full size")
    e.close()
    o.close()


def test_full_size_c5_invariants():
    """BASELINE config 5 at full size (16384 x 16384, 16,777,216 agents): size-independent properties after
    two steps -- every living agent stands on the cell that carries its id, cells are unique, the population
    series equals the number of living agents, sugar never exceeds capacity, harvested cells are empty, and
    the run is reproducible (two engines, same inputs -> identical state)."""
    n = 16384
    s, init = common.synthetic_case(n, False)
    cfg = config_from_settings(s, rng_mode="keyed")
    digests = []
    for rep in range(2):
        e = common.engine_world(cfg)
        common.init_world(e, init, "keyed")
        st = e.step(2)
        ags = e.download_agents()
        cells = e.download_cells(cap=False, pollution=False)
        assert len(ags["id"]) == int(st["population"][-1]) == e.agent_count()
        assert st["acted"][0] == len(init["id"]) and st["deaths"].sum() == len(init["id"]) - len(ags["id"])
        lin = ags["x"].astype(np.int64) * n + ags["y"]
        assert len(np.unique(lin)) == len(lin)
        occ = cells["occ"].ravel()
        assert np.array_equal(occ[lin], ags["id"]) and int((occ != 0).sum()) == len(lin)
        assert (cells["sugar"] <= init["cap"]).all() and (cells["sugar"] >= 0).all()
        assert (ags["sugar"] > 0.1).all()
        digests.append((int(lin.sum()), float(ags["sugar"].sum()), float(cells["sugar"].sum()), st.tobytes()))
        e.close()
        del cells, ags
    assert digests[0] == digests[1]


def test_page_locked_buffers_and_out_arguments():
    """ss_host_alloc buffers through upload/download give the same state as ordinary numpy arrays, and the
    device-side export of View.agents (keyed order) equals the oracle's list."""
    from sugarscape_utilicalc_b200.engine import pinned_copy, pinned_empty
    n = 300
    s, init = common.synthetic_case(n, True)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg), common.oracle_world(cfg)
    pinned = {k: pinned_copy(v) for k, v in init.items() if isinstance(v, np.ndarray)}
    common.init_world(e, dict(init, **pinned), "keyed")
    common.init_world(o, init, "keyed")
    e.step(4)
    o.step(4)
    k = e.agent_count()
    aout = {key: pinned_empty(k + 7, np.int32) for key in ("id", "x", "y", "vision", "metab")}
    aout["sugar"] = pinned_empty(k + 7, np.float64)
    cout = dict(sugar=pinned_empty((n, n)), pollution=pinned_empty((n, n)), occ=pinned_empty((n, n), np.int32))
    a, b = e.download_agents(out=aout), o.download_agents()
    for key in ("id", "x", "y", "vision", "metab", "sugar"):
        assert a[key].base is not None and len(a[key]) == k
        assert np.array_equal(a[key], b[key]), key
    c, d = e.download_cells(cap=False, out=cout), o.download_cells()
    assert c["sugar"] is cout["sugar"] and c["cap"] is None
    for key in ("sugar", "pollution", "occ"):
        assert np.array_equal(c[key], d[key]), key
    with pytest.raises(ValueError):
        e.download_agents(out={"id": np.empty(3, np.int32)})
    e.close()
    o.close()
