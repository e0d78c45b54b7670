"""Shared test helpers: golden fixtures, world construction for the CPU oracle and the CUDA
engine (same call sequence), and the comparison routines."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402

GOLDEN_CASES = [
    "c1_default_utilicalc_mt", "c1_default_moveeat_mt", "c2_seasons_mt", "c2_pollution_mt",
    "c3_utilicalc_sugar_mt", "c1_default_moveeat_keyed", "c1_default_utilicalc_keyed", "c2_seasons_keyed",
    "c2_pollution0_keyed", "c4_synth96_poll_keyed", "c4_synth130_nopoll_keyed",
]


LIFE_CASES = ["f3_evolution_metabolism_mt", "f3_procreate_only_mt"]      # SURVEY 8 f3, recorded by make_golden.py f3
LIFE_COLUMNS = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")
SPICE_CASES = ["f2_spice_mt", "f2_spice_life_mt"]                         # SURVEY 8 f2, recorded by make_golden.py f2


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    g["settings"] = json.loads(str(g["settings_json"]))
    g["rng"] = str(g["rng"])
    g["steps"] = int(g["steps"])
    return g


def oracle_world(cfg):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import OracleWorld
    return OracleWorld(cfg)


def engine_world(cfg, path=-1, options=None):
    """options: ss_set_option settings applied before the upload, e.g. {"pass_impl": 1, "solo_window": 64}."""
    from sugarscape_utilicalc_b200.engine import Engine
    e = Engine(cfg, path=path)
    for k, v in (options or {}).items():
        e.set_option(k, v)
    return e


def init_world(world, init, rng, mt=None, mt_index=None):
    world.upload_cells(init["sugar"], init["cap"], init["pollution"])
    world.upload_agents(init["id"], init["x"], init["y"], init["vision"], init["metab"], init["sugar_agent"])
    if init.get("spice") is not None:
        sp = init["spice"]
        world.upload_cells_spice(sp["spice"], sp["spice_cap"])
        world.upload_agents_spice(sp["spice_metab"], sp["spice_agent"], sp["spice_endowment"])
    if init.get("life") is not None:
        world.upload_agents_life(init["id_increment"], **init["life"])
    if rng == "mt":
        world.set_rng_mt19937(mt, mt_index)


def golden_init(g):
    init = dict(sugar=g["init_sugar"], cap=g["init_cap"], pollution=g["init_pollution"], id=g["init_order"],
                x=g["init_x"], y=g["init_y"], vision=g["init_vision"], metab=g["init_metab"],
                sugar_agent=g["init_sugar_agent"])
    if "init_life_age" in g:
        init["life"] = {k: g["init_life_" + k] for k in LIFE_COLUMNS}
        init["id_increment"] = int(g["init_id_increment"])
    if "init_spice" in g:
        init["spice"] = {k: g["init_" + k] for k in ("spice", "spice_cap", "spice_metab", "spice_agent", "spice_endowment")}
    return init


def world_from_golden(g, kind, path=-1):
    cfg = config_from_settings(g["settings"], rng_mode=g["rng"])
    w = oracle_world(cfg) if kind == "oracle" else engine_world(cfg, path)
    init_world(w, golden_init(g), g["rng"], g["init_mt"], int(g["init_mt_index"]))
    return w


def assert_state_equal(cells, agents, exp, ordered=True, what=""):
    """exp: dict with sugar, pollution, occ, order, x, y, sugar_agent (a reference snapshot)."""
    assert np.array_equal(cells["occ"], exp["occ"]), what + ": occupancy differs"
    assert np.array_equal(cells["sugar"], exp["sugar"]), what + ": cell sugar differs"
    assert np.array_equal(cells["pollution"], exp["pollution"]), what + ": pollution differs"
    ids, eids = agents["id"], exp["order"]
    if ordered:
        assert np.array_equal(ids, eids), what + ": agent list order differs"
        o = eo = slice(None)
    else:
        o, eo = np.argsort(ids, kind="stable"), np.argsort(eids, kind="stable")
        assert np.array_equal(ids[o], eids[eo]), what + ": set of living agents differs"
    assert np.array_equal(agents["x"][o], exp["x"][eo]) and np.array_equal(agents["y"][o], exp["y"][eo]), \
        what + ": agent positions differ"
    assert np.array_equal(agents["sugar"][o], exp["sugar_agent"][eo]), what + ": agent sugar differs"


def run_against_golden(g, world, exact_gini=True, chunk=None, max_steps=None):
    """Step `world` through the golden run; compare the stat series every step and the full
    state at every checkpoint."""
    from sugarscape_utilicalc_b200 import outputs
    steps = g["steps"] if max_steps is None else min(max_steps, g["steps"])
    cps = sorted(int(k[2:].split("_")[0]) for k in g if k.startswith("cp") and k.endswith("_occ"))
    cps = [c for c in cps if c <= steps]
    life = "init_life_age" in g
    log = []
    t = 0
    for cp in cps:
        while t < cp:
            k = cp - t if chunk is None else min(chunk, cp - t)
            st = world.step(k)
            if life:
                log += outputs.event_lines(world.events())
            sl = slice(t, t + k)
            assert np.array_equal(st["population"], g["population"][sl]), "population differs in steps %d..%d" % (t, t + k)
            assert np.array_equal(st["metabolism_mean"], g["metabolismMean"][sl]), "metabolismMean differs"
            assert np.array_equal(st["vision_mean"], g["visionMean"][sl]), "visionMean differs"
            if exact_gini:
                assert np.array_equal(st["gini"], g["gini"][sl]), "gini differs in steps %d..%d" % (t, t + k)
            else:
                np.testing.assert_allclose(st["gini"], g["gini"][sl], rtol=1e-9, atol=1e-12)
            t += k
        exp = {k2: g["cp%d_%s" % (cp, k2)] for k2 in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent")}
        ags = world.download_agents()
        assert_state_equal(world.download_cells(), ags, exp, ordered=True, what="step %d" % cp)
        if "init_spice" in g:       # the second resource on the cells and in the agents
            assert np.array_equal(world.download_cells_spice()["spice"], g["cp%d_spice" % cp]), "step %d: cell spice differs" % cp
            sp = world.download_agents_spice()
            assert np.array_equal(sp["spice"], g["cp%d_spice_agent" % cp]), "step %d: agent spice differs" % cp
            assert np.array_equal(sp["spice_metab"], g["cp%d_spice_metab" % cp]) and np.array_equal(sp["spice_endowment"], g["cp%d_spice_endowment" % cp])
        if life:            # the fields the lifespan / procreation rules carry, and the event lines so far
            lf = world.download_agents_life()
            for c in LIFE_COLUMNS:
                assert np.array_equal(lf[c], g["cp%d_life_%s" % (cp, c)]), "step %d: Agent.%s differs" % (cp, c)
            assert np.array_equal(ags["vision"], g["cp%d_life_vision" % cp]) and np.array_equal(ags["metab"], g["cp%d_life_metab" % cp])
            assert lf["id_increment"] == int(g["cp%d_id_increment" % cp]), "Environment.idIncrement differs"
            assert log == str(g["log_text"]).split("\n")[:int(g["cp%d_log_lines" % cp])], "event log differs by step %d" % cp
    if g["rng"] == "mt" and steps == g["steps"]:
        mt, idx = world.get_rng_mt19937()
        assert idx == int(g["final_mt_index"]) and np.array_equal(mt, g["final_mt"]), "MT19937 state differs"
    return t


def synthetic_case(n, pollution, seed=1234, density=1.0 / 16.0, vision=(1, 6), rng="keyed", seasons=False,
                   utilicalc=False):
    w = synthetic.tiled_world(n, density=density, seed=seed, vision=vision)
    s = synthetic.synthetic_settings(n, pollution=pollution, seasons=seasons, utilicalc=utilicalc)
    init = dict(sugar=w["sugar"], cap=w["cap"], pollution=w["pollution"], id=w["id"], x=w["x"], y=w["y"],
                vision=w["vision"], metab=w["metab"], sugar_agent=w["sugar_agent"])
    return s, init


def compare_worlds(a, b, steps, exact_gini=False, every=1, what=""):
    """Step two worlds (e.g. engine and oracle) side by side and compare."""
    for t in range(0, steps, every):
        k = min(every, steps - t)
        sa, sb = a.step(k), b.step(k)
        for f in ("population", "metabolism_mean", "vision_mean", "acted", "deaths"):
            assert np.array_equal(sa[f], sb[f]), "%s: %s differs at steps %d..%d: %s vs %s" % (what, f, t, t + k, sa[f], sb[f])
        if exact_gini:
            assert np.array_equal(sa["gini"], sb["gini"]), what + ": gini differs"
        else:
            np.testing.assert_allclose(sa["gini"], sb["gini"], rtol=1e-9, atol=1e-12)
        ca, cb = a.download_cells(), b.download_cells()
        ga, gb = a.download_agents(), b.download_agents()
        exp = dict(sugar=cb["sugar"], pollution=cb["pollution"], occ=cb["occ"], order=gb["id"], x=gb["x"], y=gb["y"],
                   sugar_agent=gb["sugar"])
        assert_state_equal(ca, ga, exp, ordered=True, what="%s step %d" % (what, t + k))
    for w in (a, b):            # engines: no watchdog firing, no internal fault on the device
        engines = [getattr(r, "eng", r) for r in w.ranks] if hasattr(w, "ranks") else [w] if hasattr(w, "health") else []
        for e in engines:
            assert e.health() == dict(watchdog=0, fault=0), what


def smoke_check():
    """__graft_entry__.smoke(): shipped default preset (MT-exact, utilicalc) for 20 steps and a
    small synthetic keyed lattice run, both checked against the CPU oracle / golden vectors."""
    g = load_golden("c1_default_utilicalc_mt")
    w = world_from_golden(g, "engine")
    run_against_golden(g, w, exact_gini=True, max_steps=20)
    w.close()
    # the kernels of the bench's step at a small size: ss_dr_build_kernel (TMA tile load) + ss_dr_rounds_kernel,
    # ss_dr_grow_kernel, ss_diffuse_col_kernel, ss_stats_kernel
    s, init = synthetic_case(512, pollution=True)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = engine_world(cfg), oracle_world(cfg)
    init_world(e, init, "keyed")
    init_world(o, init, "keyed")
    compare_worlds(e, o, 6, what="synthetic 512x512")
    assert e.counters()["path"] == 2 and e.counters()["pass_impl"] == 3, "dependency rounds were not selected"
    e.close()
    o.close()
    # spice with the restated libm pow, limited lifespan + procreation (serial kernel, goldens from the reference)
    for name in ("f2_spice_life_mt", "f3_evolution_metabolism_mt"):
        g = load_golden(name)
        w = world_from_golden(g, "engine")
        run_against_golden(g, w, exact_gini=True, max_steps=20)
        w.close()


OUTPUT_CASES = ["c1_default_moveeat_mt", "c1_default_moveeat_keyed", "c2_pollution_mt"]


def check_outputs_against_golden(name, kind, tmp_dir, path=-1):
    """SURVEY 8 f1 without the reference at hand: `outputs.RunRecorder` over `kind` ("oracle" | "engine") reproduces
    the event log, the plateau exit, the data.txt column and the createLogFile text recorded from the unmodified
    reference (tests/golden/outputs.json, made by tests/golden/make_outputs_golden.py)."""
    from sugarscape_utilicalc_b200 import outputs
    with open(os.path.join(GOLDEN, "outputs.json")) as f:
        exp = json.load(f)[name]
    g = load_golden(name)
    w = world_from_golden(g, kind, path)
    data_file = os.path.join(str(tmp_dir), "data.txt")
    rec = outputs.RunRecorder(g["settings"], w, rng=g["rng"], data_file=data_file)
    exited = None
    for _ in range(exp["steps"]):
        try:
            rec.step()
        except outputs.PlateauReached:
            exited = rec.iteration
            break
    assert exited == exp["exited_at"]
    assert rec.log == exp["log"]
    assert (rec.pop_at_100, rec.gini_at_100) == (exp["pop_at_100"], exp["gini_at_100"]) or kind == "engine"
    assert rec.pop_at_100 == exp["pop_at_100"] and abs(rec.gini_at_100 - exp["gini_at_100"]) <= 1e-9
    got_data = open(data_file).read() if os.path.exists(data_file) else None
    assert got_data == exp["data_txt"]
    stats = rec.stats()
    assert stats == exp["stats"]
    text = outputs.log_file_text(g["settings"], g["settings"]["env"]["random_seed"], rec.metabolismMean, rec.visionMean,
                                 rec.gini, stats, rec.log)
    assert text == exp["log_file_text"]
    w.close()


def check_checkpoint_resume(name, kind, tmp_dir, stop=7, total=15):
    if name.startswith("f3_") or name.startswith("f2_"):
        stop, total = 70, 90        # past the first births, deaths and agents at maxAge
    import numpy as np
    from sugarscape_utilicalc_b200 import checkpoint
    g = load_golden(name)
    a = world_from_golden(g, kind)
    sa = a.step(total)
    b = world_from_golden(g, kind)
    sb1 = b.step(stop)
    path = os.path.join(str(tmp_dir), "ck.npz")
    checkpoint.save(path, b, g["settings"], g["rng"], stop)
    b.close()
    factory = oracle_world if kind == "oracle" else None
    c, settings, rng, it = checkpoint.restore(path, world_factory=factory)
    assert (settings, rng, it) == (g["settings"], g["rng"], stop)
    sc = c.step(total - stop)
    for f in ("population", "metabolism_mean", "vision_mean", "gini", "acted", "deaths"):
        assert np.array_equal(np.concatenate([sb1[f], sc[f]]), sa[f]), f
    ca, cc, ga, gc = a.download_cells(), c.download_cells(), a.download_agents(), c.download_agents()
    for k in ("sugar", "cap", "pollution", "occ"):
        assert np.array_equal(ca[k], cc[k]), k
    for k in ("id", "x", "y", "vision", "metab", "sugar"):
        assert np.array_equal(ga[k], gc[k]), k
    if "init_spice" in g:
        assert np.array_equal(a.download_cells_spice()["spice"], c.download_cells_spice()["spice"])
        sa_, sc_ = a.download_agents_spice(), c.download_agents_spice()
        for k in ("spice_metab", "spice", "spice_endowment"):
            assert np.array_equal(sa_[k], sc_[k]), k
    if "init_life_age" in g:
        la, lc = a.download_agents_life(), c.download_agents_life()
        for k in LIFE_COLUMNS + ("id_increment",):
            assert np.array_equal(la[k], lc[k]), k
    if g["rng"] == "mt":
        ma, mc = a.get_rng_mt19937(), c.get_rng_mt19937()
        assert ma[1] == mc[1] and np.array_equal(ma[0], mc[0])
    a.close()
    c.close()


def check_resume_after_top_id_died(kind, tmp_dir):
    """the keyed execution order is sized by the largest Agent.id of the original upload.  Here that agent
    starves before the checkpoint (ids up to 257 -> 5 bits per half; the survivors' largest id is < 256 -> 4 bits if the
    width were recomputed): the resumed run must still equal the uninterrupted one."""
    from sugarscape_utilicalc_b200 import checkpoint
    n = 64
    s, init = synthetic_case(n, False)
    init = {k: (v[:200].copy() if k in ("id", "x", "y", "vision", "metab", "sugar_agent") else v) for k, v in init.items()}
    init["id"] = init["id"].copy()
    init["id"][-1] = 257                                    # the largest id ...
    init["sugar_agent"][-1] = 1.0                           # ... starves in the first steps
    init["metab"][-1] = 4
    s["env"]["max_capacity"] = 10
    cfg = config_from_settings(s, rng_mode="keyed")
    mk = (lambda c: oracle_world(c)) if kind == "oracle" else (lambda c: engine_world(c))
    a = mk(cfg)
    barren = np.zeros((n, n))
    a.upload_cells(barren, init["cap"] * 0.0, None)
    a.upload_agents(init["id"], init["x"], init["y"], init["vision"], init["metab"], init["sugar_agent"])
    sa = a.step(8)
    b = mk(cfg)
    b.upload_cells(barren, init["cap"] * 0.0, None)
    b.upload_agents(init["id"], init["x"], init["y"], init["vision"], init["metab"], init["sugar_agent"])
    sb1 = b.step(3)
    assert 257 not in b.download_agents()["id"].tolist(), "the agent with the largest id should have starved by now"
    path = os.path.join(str(tmp_dir), "ck.npz")
    checkpoint.save(path, b, s, "keyed", 3)
    b.close()
    c, _, _, _ = checkpoint.restore(path, world_factory=(oracle_world if kind == "oracle" else None))
    sc = c.step(5)
    for f in ("population", "metabolism_mean", "vision_mean", "acted", "deaths"):
        assert np.array_equal(np.concatenate([sb1[f], sc[f]]), sa[f]), f
    ga, gc = a.download_agents(), c.download_agents()
    for k in ("id", "x", "y", "sugar"):
        assert np.array_equal(ga[k], gc[k]), k
    a.close()
    c.close()
