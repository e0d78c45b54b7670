"""CPU, build container only: the C oracle against the UNMODIFIED reference run headless
(oracle/ref_harness.py).  Skipped where /root/reference is absent (the GPU box)."""
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import common
from sugarscape_utilicalc_b200 import config_from_settings, synthetic

pytestmark = pytest.mark.reference
sys.path.insert(0, os.path.join(common.ROOT, "oracle"))


def _compare(s, rng, world, steps):
    import ref_harness as rh
    r = rh.ReferenceRun(s, rng=rng, world=world)
    s0 = r.snapshot()
    o = common.oracle_world(config_from_settings(s, rng_mode=rng))
    o.upload_cells(s0["sugar"], s0["cap"], s0["pollution"])
    o.upload_agents(s0["order"], s0["x"], s0["y"], s0["vision"], s0["metab"], s0["sugar_agent"])
    if rng == "mt":
        o.set_rng_mt19937(s0["mt"], s0["mt_index"])
    for t in range(steps):
        r.step()
        st_ = o.step(1)[0]
        rs = r.stats()
        assert st_["population"] == rs["population"][-1] and st_["gini"] == rs["gini"][-1]
        assert st_["metabolism_mean"] == rs["metabolismMean"][-1] and st_["vision_mean"] == rs["visionMean"][-1]
    sn, cells, ags = r.snapshot(), o.download_cells(), o.download_agents()
    exp = {k: sn[k] for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent")}
    common.assert_state_equal(cells, ags, exp, ordered=True, what="reference vs oracle")
    if rng == "mt":
        mt, idx = o.get_rng_mt19937()
        assert idx == sn["mt_index"] and np.array_equal(mt, sn["mt"])


def test_live_default_preset_mt():
    import ref_harness as rh
    _compare(rh.load_settings(), "mt", None, 12)


def test_live_pollution_preset_from_t0_keyed():
    import ref_harness as rh
    s = rh.load_settings("pollution")
    s["rule_params"]["pollution"]["pollution_start_time"] = 0
    s["rule_params"]["pollution"]["diffusion_start_time"] = 2
    _compare(s, "keyed", None, 15)


def test_utilicalc_spice_preset_crashes_in_the_reference():
    """parity for this preset is unpinned; the engine refuses it."""
    import ref_harness as rh
    s = rh.load_settings("utilicalc_spice")
    r = rh.ReferenceRun(s)
    with pytest.raises(KeyError):
        r.view.updateGame()
    from sugarscape_utilicalc_b200 import UnsupportedRuleError
    with pytest.raises(UnsupportedRuleError):
        config_from_settings(s)


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(n=st.integers(16, 40), seed=st.integers(0, 10 ** 6), pollution=st.booleans(), utilicalc=st.booleans(),
       seasons=st.booleans(), rng=st.sampled_from(["mt", "keyed"]), dens=st.sampled_from([0.05, 0.15, 0.4]))
def test_random_small_worlds(n, seed, pollution, utilicalc, seasons, rng, dens):
    if utilicalc:
        pollution = False
    w = synthetic.tiled_world(n, density=dens, seed=seed, vision=(1, min(6, (n - 1) // 2)))
    s = synthetic.synthetic_settings(n, pollution=pollution, seasons=seasons, utilicalc=utilicalc, seed=seed % 1000)
    if seasons:
        s["rule_params"]["seasons"]["period"] = 2
    world = dict(cap=w["cap"], sugar=w["sugar"], pollution=w["pollution"], x=w["x"], y=w["y"], vision=w["vision"],
                 metab=w["metab"], sugar_agent=w["sugar_agent"])
    _compare(s, rng, world, 6)


def test_dropin_install_glue_with_reference_view():
    """`dropin.install` on a live reference View (oracle injected as backend: no GPU here):
    stepping through the installed updateGame and syncing back equals the reference run."""
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import dropin
    s = rh.load_settings("pollution")
    s["rule_params"]["pollution"]["pollution_start_time"] = 0
    ref = rh.ReferenceRun(s)
    ref.run(9)
    exp = ref.snapshot()
    exp_stats = ref.stats()
    run = rh.ReferenceRun(s)            # re-imports the reference: same seed, same initial state and MT stream
    h = dropin.install(run.ss, run.view, rng="mt", world_factory=common.oracle_world)
    for _ in range(9):
        run.view.updateGame()
    h.sync_back()
    got = run.snapshot()
    for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent", "mt"):
        assert np.array_equal(got[k], exp[k]), k
    assert got["mt_index"] == exp["mt_index"] and got["iteration"] == 9 and got["time"] == 9
    assert run.stats() == exp_stats
    # hand the View back: the reference's own loop continues from the synced state and RNG stream
    import random
    state9 = random.getstate()
    run.view.updateGame = h.original_update
    run.step()
    random.setstate(state9)                          # both runs share the one global `random`
    ref.step()
    a, b = run.snapshot(), ref.snapshot()
    for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent", "mt"):
        assert np.array_equal(a[k], b[k]), k


LIFE_KEYS = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")


def test_lifespan_and_procreation_oracle_vs_live_reference():
    """examples/tag_flipping (limitedLifespan + procreate) step by step against the unmodified reference --
    cells, agents in list order, the fields the two rules carry, Environment.idIncrement, the event lines, the MT19937
    stream."""
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import outputs
    s = rh.load_settings("tag_flipping")
    ref = rh.ReferenceRun(s)
    s0 = ref.snapshot()
    o = common.oracle_world(config_from_settings(s, rng_mode="mt"))
    init = dict(sugar=s0["sugar"], cap=s0["cap"], pollution=s0["pollution"], id=s0["order"], x=s0["x"], y=s0["y"],
                vision=s0["vision"], metab=s0["metab"], sugar_agent=s0["sugar_agent"], life={k: s0[k] for k in LIFE_KEYS},
                id_increment=s0["id_increment"])
    common.init_world(o, init, "mt", s0["mt"], s0["mt_index"])
    log = []
    for t in range(80):
        ref.step()
        st = o.step(1)
        log += outputs.event_lines(o.events())
        sn, ga, gl, gc = ref.snapshot(), o.download_agents(), o.download_agents_life(), o.download_cells()
        assert np.array_equal(ga["id"], sn["order"]) and np.array_equal(ga["x"], sn["x"]) and np.array_equal(ga["y"], sn["y"]), t
        assert np.array_equal(ga["sugar"], sn["sugar_agent"]) and np.array_equal(gc["occ"], sn["occ"]) and np.array_equal(gc["sugar"], sn["sugar"]), t
        for k in LIFE_KEYS:
            assert np.array_equal(gl[k], sn[k]), (t, k)
        assert gl["id_increment"] == sn["id_increment"] and st["population"][0] == ref.view.population[-1] and st["gini"][0] == ref.view.gini[-1]
        mt, idx = o.get_rng_mt19937()
        assert idx == sn["mt_index"] and np.array_equal(mt, sn["mt"]), t
    assert log == ref.view.log and len(log) > 100
    o.close()


@settings(max_examples=int(os.environ.get("SS_LIFE_EXAMPLES", "40")), deadline=None, suppress_health_check=list(HealthCheck))
@given(grid=st.integers(12, 36), seed=st.integers(0, 10 ** 6), blue=st.integers(5, 120), red=st.integers(5, 120),
       spice=st.booleans(), lifespan=st.booleans(), procreate=st.booleans(), seasons=st.booleans(),
       endow=st.sampled_from([(5, 25), (50, 100)]), ages=st.sampled_from([(4, 10), (60, 100)]), start=st.sampled_from([0, 1, 12]))
def test_random_worlds_with_spice_and_life_rules(grid, seed, blue, red, spice, lifespan, procreate, seasons, endow, ages, start):
    """SURVEY 8 f2 + f3 on random small worlds built by the reference's own initialisation: any combination of spice,
    limitedLifespan and procreate (short lives, early fertility and crowded lattices included), step by step -- cells, agents,
    the fields of the rules, event lines, MT19937 stream."""
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import outputs
    s = rh.load_settings("readme_example")
    s["rules"].update(trade=False, spice=spice, limitedLifespan=lifespan, procreate=procreate, seasons=seasons)
    s["view"]["grid_size"] = {"x": grid, "y": grid}
    s["env"]["random_seed"] = seed
    s["env"]["total_agents"] = {"blue": blue, "red": red}
    for k, site in s["env"]["sites"].items():
        site["x"], site["y"], site["radius"] = site["x"] % grid, site["y"] % grid, max(2, site["radius"] * grid // 50)
    for d in s["env"]["agent_distributions"].values():
        d["x_min"], d["x_max"], d["y_min"], d["y_max"] = 0, grid - 1, 0, grid - 1
    s["agents"]["max_agent_vision"] = min(s["agents"]["max_agent_vision"], (grid - 1) // 2)
    s["agents"]["initial_endowments"] = {"min": endow[0], "max": endow[1]}
    s["agents"]["death_age_range"] = {"min": ages[0], "max": ages[1]}
    for who in ("male", "female"):
        s["agents"]["fertility_age_ranges"][who].update(min_start=start, max_start=start + 3)
    if seasons:
        s["rule_params"]["seasons"]["period"] = 3
        for reg in s["rule_params"]["seasons"]["regions"].values():
            reg["start_x"], reg["end_x"], reg["start_y"], reg["end_y"] = 0, grid - 1, reg["start_y"] % grid, min(reg["end_y"], grid - 1)
    ref = rh.ReferenceRun(s)
    s0 = ref.snapshot()
    cfg = config_from_settings(s, rng_mode="mt")
    o = common.oracle_world(cfg)
    init = dict(sugar=s0["sugar"], cap=s0["cap"], pollution=s0["pollution"], id=s0["order"], x=s0["x"], y=s0["y"],
                vision=s0["vision"], metab=s0["metab"], sugar_agent=s0["sugar_agent"])
    if spice:
        init["spice"] = {k: s0[k] for k in ("spice", "spice_cap", "spice_metab", "spice_agent", "spice_endowment")}
    if lifespan or procreate:
        init["life"], init["id_increment"] = {k: s0[k] for k in LIFE_KEYS}, s0["id_increment"]
    common.init_world(o, init, "mt", s0["mt"], s0["mt_index"])
    log = []
    for t in range(25):
        ref.step()
        st_ = o.step(1)
        if lifespan or procreate:
            log += outputs.event_lines(o.events())
        sn, ga, gc = ref.snapshot(), o.download_agents(), o.download_cells()
        assert np.array_equal(ga["id"], sn["order"]) and np.array_equal(ga["x"], sn["x"]) and np.array_equal(ga["y"], sn["y"]), t
        assert np.array_equal(ga["sugar"], sn["sugar_agent"]) and np.array_equal(gc["occ"], sn["occ"]) and np.array_equal(gc["sugar"], sn["sugar"]), t
        assert np.array_equal(ga["vision"], sn["vision"]) and np.array_equal(ga["metab"], sn["metab"]), t
        if spice:
            gs = o.download_agents_spice()
            assert np.array_equal(gs["spice"], sn["spice_agent"]) and np.array_equal(gs["spice_metab"], sn["spice_metab"]), t
            assert np.array_equal(gs["spice_endowment"], sn["spice_endowment"]) and np.array_equal(o.download_cells_spice()["spice"], sn["spice"]), t
        if lifespan or procreate:
            gl = o.download_agents_life()
            for k in LIFE_KEYS:
                assert np.array_equal(gl[k], sn[k]), (t, k)
            assert gl["id_increment"] == sn["id_increment"], t
        v = ref.view
        assert st_["population"][0] == v.population[-1] and st_["gini"][0] == v.gini[-1], t
        assert st_["metabolism_mean"][0] == v.metabolismMean[-1] and st_["vision_mean"][0] == v.visionMean[-1], t
        mt, idx = o.get_rng_mt19937()
        assert idx == sn["mt_index"] and np.array_equal(mt, sn["mt"]), t
        if not len(sn["order"]):
            break
    if lifespan or procreate:
        assert log == ref.view.log
    assert o.pow_faults() == 0
    o.close()


def test_dropin_with_births_hands_a_working_view_back():
    """`dropin.install` under limitedLifespan + procreate: the installed updateGame keeps View.log complete (mating lines and
    deaths in execution order), sync_back gives the children Agent objects, and the reference's own loop continues from the
    synced View exactly as its uninterrupted run does."""
    import random
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import dropin
    s = rh.load_settings("evolution_metabolism")
    ref = rh.ReferenceRun(s)
    ref.run(30)
    exp, exp_log, exp_stats = ref.snapshot(), list(ref.view.log), ref.stats()
    run = rh.ReferenceRun(s)
    h = dropin.install(run.ss, run.view, rng="mt", world_factory=common.oracle_world)
    for _ in range(30):
        run.view.updateGame()
    h.sync_back()
    got = run.snapshot()
    for k in ("sugar", "occ", "order", "x", "y", "sugar_agent", "vision", "metab", "mt") + LIFE_KEYS:
        assert np.array_equal(got[k], exp[k]), k
    assert got["id_increment"] == exp["id_increment"] and run.view.log == exp_log and run.stats() == exp_stats
    state = random.getstate()
    run.view.updateGame = h.original_update
    for _ in range(5):
        run.step()
    random.setstate(state)
    for _ in range(5):
        ref.step()
    a, b = run.snapshot(), ref.snapshot()
    for k in ("sugar", "occ", "order", "x", "y", "sugar_agent", "mt") + LIFE_KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert run.view.log == ref.view.log


def test_spice_with_keyed_words_oracle_vs_live_reference():
    """SURVEY 8 f2 with the keyed word source (the reference consumes it through the shim): examples/trade_price_volume
    without trade, step by step."""
    import ref_harness as rh
    s = rh.load_settings("trade_price_volume")
    s["rules"]["trade"] = False
    ref = rh.ReferenceRun(s, rng="keyed")
    s0 = ref.snapshot()
    o = common.oracle_world(config_from_settings(s, rng_mode="keyed"))
    init = dict(sugar=s0["sugar"], cap=s0["cap"], pollution=s0["pollution"], id=s0["order"], x=s0["x"], y=s0["y"],
                vision=s0["vision"], metab=s0["metab"], sugar_agent=s0["sugar_agent"],
                spice={k: s0[k] for k in ("spice", "spice_cap", "spice_metab", "spice_agent", "spice_endowment")})
    common.init_world(o, init, "keyed")
    for t in range(40):
        ref.step()
        st = o.step(1)
        sn, ga, gs, gc = ref.snapshot(), o.download_agents(), o.download_agents_spice(), o.download_cells()
        assert np.array_equal(ga["id"], sn["order"]) and np.array_equal(ga["x"], sn["x"]) and np.array_equal(ga["y"], sn["y"]), t
        assert np.array_equal(ga["sugar"], sn["sugar_agent"]) and np.array_equal(gs["spice"], sn["spice_agent"]), t
        assert np.array_equal(gc["occ"], sn["occ"]) and np.array_equal(o.download_cells_spice()["spice"], sn["spice"]), t
        assert st["population"][0] == ref.view.population[-1] and st["gini"][0] == ref.view.gini[-1], t
    assert o.pow_faults() == 0
    o.close()


def test_dropin_with_spice_and_births_hands_a_working_view_back():
    """SURVEY 8 f2 + f3 through `dropin.install` on examples/readme_example without trade: spice on the cells and in the agents,
    births, the event log -- and the reference's own loop continues from the synced View."""
    import random
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import dropin
    s = rh.load_settings("readme_example")
    s["rules"]["trade"] = False
    keys = ("sugar", "spice", "occ", "order", "x", "y", "sugar_agent", "spice_agent", "spice_metab", "spice_endowment", "mt") + LIFE_KEYS
    ref = rh.ReferenceRun(s)
    ref.run(25)
    exp, exp_log, exp_stats = ref.snapshot(), list(ref.view.log), ref.stats()
    run = rh.ReferenceRun(s)
    h = dropin.install(run.ss, run.view, rng="mt", world_factory=common.oracle_world)
    for _ in range(25):
        run.view.updateGame()
    h.sync_back()
    got = run.snapshot()
    for k in keys:
        assert np.array_equal(got[k], exp[k]), k
    assert run.view.log == exp_log and run.stats() == exp_stats
    state = random.getstate()
    run.view.updateGame = h.original_update
    for _ in range(5):
        run.step()
    random.setstate(state)
    for _ in range(5):
        ref.step()
    a, b = run.snapshot(), ref.snapshot()
    for k in keys:
        assert np.array_equal(a[k], b[k]), k
    assert run.view.log == ref.view.log


def _moveeat_settings(rh):
    s = rh.load_settings(None)
    s["rules"]["moveEat"], s["rules"]["utilicalc"] = True, False
    return s


@pytest.mark.parametrize("rng", ["mt", "keyed"])
def test_dropin_event_log_plateau_rule_and_log_file(rng, tmp_path):
    """the installed updateGame leaves the same `View.log` (death events in execution order), step-100
    sample, plateau exit with the same `data_collection/data.txt`, and `createLogFile` text as the reference."""
    import glob
    import os
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import dropin, outputs
    s = _moveeat_settings(rh)
    limit = 330
    ref = rh.ReferenceRun(s, rng=rng)
    ref.run(limit)                  # MT stream: the reference exits at iteration 307 through the plateau rule
    assert rng != "mt" or ref.exited_at is not None
    with rh._cwd(ref.scratch):
        ref.view.createLogFile()
        ref_log_text = open(glob.glob(os.path.join(ref.scratch, "logs", "log_*.txt"))[0]).read()
    data_path = os.path.join(ref.scratch, "data_collection", "data.txt")
    ref_data = open(data_path).read() if os.path.exists(data_path) else None
    exp_log, exited_at = list(ref.view.log), ref.exited_at

    run = rh.ReferenceRun(s, rng="mt")                  # a fresh reference View; the engine side draws its own words
    h = dropin.install(run.ss, run.view, rng=rng, world_factory=common.oracle_world, gui=(rng == "mt"))
    got_exit = None
    with rh._cwd(run.scratch):
        for _ in range(limit):
            try:
                run.view.updateGame()
            except outputs.PlateauReached:
                got_exit = run.view.iteration
                break
        data_path = os.path.join(run.scratch, "data_collection", "data.txt")
        got_data = open(data_path).read() if os.path.exists(data_path) else None
    assert got_exit == exited_at
    assert run.view.log == exp_log[:len(run.view.log)] and len(run.view.log) > 50
    if ref_data is None:
        assert got_data is None and run.view.log == exp_log
    else:       # the reference (exit() swallowed) appends one more column per further step: ours is its first column
        assert got_data == "\n".join(l.split(", ")[0] for l in ref_data.strip().split("\n")) + "\n"
    assert (run.view.pop_at_100, run.view.gini_at_100) == (ref.view.population[99], ref.view.gini[99])
    # the reference kept stepping after its swallowed exit(): compare the log file over the common prefix
    h.sync_back()
    k = got_exit if got_exit is not None else limit
    stats = outputs.ending_stats(k, run.view.population, run.view.metabolismMean, run.view.visionMean, run.view.gini,
                                 [a.getSugar() for a in run.view.agents])
    text = outputs.log_file_text(run.ss.settings, run.ss.seed, run.view.metabolismMean, run.view.visionMean, run.view.gini,
                                 stats, run.view.log)
    ref2 = rh.ReferenceRun(s, rng=rng)
    ref2.run(k)
    with rh._cwd(ref2.scratch):
        ref2.view.createLogFile()
        ref2_text = open(glob.glob(os.path.join(ref2.scratch, "logs", "log_*.txt"))[0]).read()
    assert text == ref2_text
    assert stats == ref2.view.stats
    assert ref_log_text.startswith(text[:text.index("\nStat lists:")])


def test_run_recorder_headless_equals_reference(tmp_path):
    """`outputs.RunRecorder` (no reference objects at all) on the shipped pollution preset: event log, series and
    data file equal the reference's."""
    import os
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import outputs
    s = rh.load_settings("pollution")
    ref = rh.ReferenceRun(s)
    init = ref.snapshot()
    mt0, idx0 = ref.mt_state()
    ref.run(120)
    cfg = config_from_settings(s, rng_mode="mt")
    w = common.oracle_world(cfg)
    w.upload_cells(init["sugar"], init["cap"], init["pollution"])
    w.upload_agents(init["order"], init["x"], init["y"], init["vision"], init["metab"], init["sugar_agent"])
    w.set_rng_mt19937(mt0, idx0)
    rec = outputs.RunRecorder(s, w, rng="mt", data_file=str(tmp_path / "data.txt"))
    for _ in range(120):
        rec.step()
    assert rec.log == ref.view.log and len(rec.log) > 20
    assert rec.population == ref.view.population and rec.gini == ref.view.gini
    assert (rec.pop_at_100, rec.gini_at_100) == (ref.view.pop_at_100, ref.view.gini_at_100)
    ref.view.makeStatsDict()
    assert rec.stats() == ref.view.stats
    assert not os.path.exists(str(tmp_path / "data.txt"))
