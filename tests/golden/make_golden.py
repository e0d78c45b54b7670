"""Generate the golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (or $SS_REFERENCE)

Each .npz holds: the settings.json text, the RNG mode, the state right after the reference's
own initialisation (cells, agents in list order, MT19937 state), the four stat series
`View.population/metabolismMean/visionMean/gini` for every step, cheap per-step digests, full
state checkpoints at a few steps and the final MT19937 state.  tests/ compare the C oracle
(CPU) and the CUDA engine (GPU) against these; nothing else reads the reference at test time.
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness as rh  # noqa: E402
from sugarscape_utilicalc_b200 import synthetic  # noqa: E402

CHECKPOINTS = (1, 2, 5, 10, 20, 50, 100, 200, 300, 500)


def pos_digest(snap):
    o = np.argsort(snap["order"], kind="stable")
    arr = np.stack([snap["order"][o], snap["x"][o], snap["y"][o]], axis=1).astype(np.int32)
    return zlib.crc32(arr.tobytes())


LIFE = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")


SPICE_CELLS = ("spice", "spice_cap")
SPICE_AGENTS = ("spice_metab", "spice_agent", "spice_endowment")


def record(name, settings, steps, rng, world=None, life=False, spice=False):
    r = rh.ReferenceRun(settings, rng=rng, world=world)
    out = {"settings_json": json.dumps(settings), "rng": rng, "steps": steps}
    s0 = r.snapshot()
    for k in ("sugar", "cap", "pollution", "order", "x", "y", "vision", "metab", "sugar_agent", "mt"):
        out["init_" + k] = s0[k]
    out["init_mt_index"] = s0["mt_index"]
    if life:
        for k in LIFE:
            out["init_life_" + k] = s0[k]
        out["init_id_increment"] = s0["id_increment"]
    if spice:
        for k in SPICE_CELLS + SPICE_AGENTS:
            out["init_" + k] = s0[k]
    dig = np.zeros((steps, 5), dtype=np.float64)  # sum agent sugar, sum cell sugar, sum pollution, pos crc, mt idx
    for t in range(steps):
        r.step()
        sn = r.snapshot()
        dig[t] = (sn["sugar_agent"].sum(), sn["sugar"].sum(), sn["pollution"].sum(), pos_digest(sn),
                  sn["mt_index"])
        if (t + 1) in CHECKPOINTS or t + 1 == steps:
            for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent"):
                out["cp%d_%s" % (t + 1, k)] = sn[k]
            if spice:
                for k in ("spice",) + SPICE_AGENTS:
                    out["cp%d_%s" % (t + 1, k)] = sn[k]
            if life:
                for k in LIFE + ("vision", "metab"):
                    out["cp%d_life_%s" % (t + 1, k)] = sn[k]
                out["cp%d_id_increment" % (t + 1)] = sn["id_increment"]
                out["cp%d_log_lines" % (t + 1)] = len(r.view.log)
    st = r.stats()
    out["population"] = np.array(st["population"], dtype=np.float64)
    out["metabolismMean"] = np.array(st["metabolismMean"], dtype=np.float64)
    out["visionMean"] = np.array(st["visionMean"], dtype=np.float64)
    out["gini"] = np.array(st["gini"], dtype=np.float64)
    out["digest"] = dig
    out["final_mt"], out["final_mt_index"] = sn["mt"], sn["mt_index"]
    out["exited_at"] = -1 if r.exited_at is None else r.exited_at
    if life:
        out["log_text"] = "\n".join(r.view.log)         # the event lines of the run (sugarscape.py:551, 597-600)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s rng=%-5s steps=%-4d pop %d -> %d  (%d KB)" % (
        name, rng, steps, len(s0["order"]), st["population"][-1], os.path.getsize(path) // 1024))


def main_life():
    """examples/evolution_metabolism as shipped (grow + moveEat + canStarve + limitedLifespan + procreate;
    examples/tag_flipping has the same settings), and the same world with procreation alone."""
    record("f3_evolution_metabolism_mt", rh.load_settings("evolution_metabolism"), 300, "mt", life=True)
    s = rh.load_settings("evolution_metabolism")
    s["rules"]["limitedLifespan"] = False
    record("f3_procreate_only_mt", s, 120, "mt", life=True)


def main_spice():
    """examples/trade_price_volume without the trade rule (grow + moveEat + canStarve + spice), and
    examples/readme_example without trade (+ limitedLifespan + procreate)."""
    s = rh.load_settings("trade_price_volume")
    s["rules"]["trade"] = False
    record("f2_spice_mt", s, 300, "mt", spice=True)
    s = rh.load_settings("readme_example")
    s["rules"]["trade"] = False
    record("f2_spice_life_mt", s, 120, "mt", life=True, spice=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "f3":
        return main_life()
    if len(sys.argv) > 1 and sys.argv[1] == "f2":
        return main_spice()
    default = rh.load_settings()
    move_eat = rh.load_settings()
    move_eat["rules"]["moveEat"], move_eat["rules"]["utilicalc"] = True, False
    poll0 = rh.load_settings("pollution")            # pollution + diffusion active from t=0
    poll0["rule_params"]["pollution"]["pollution_start_time"] = 0
    poll0["rule_params"]["pollution"]["diffusion_start_time"] = 0
    record("c1_default_utilicalc_mt", default, 100, "mt")
    record("c1_default_moveeat_mt", move_eat, 100, "mt")
    record("c2_seasons_mt", rh.load_settings("seasons"), 500, "mt")
    record("c2_pollution_mt", rh.load_settings("pollution"), 500, "mt")
    record("c3_utilicalc_sugar_mt", rh.load_settings("utilicalc_sugar"), 500, "mt")
    record("c1_default_moveeat_keyed", move_eat, 100, "keyed")
    record("c1_default_utilicalc_keyed", default, 50, "keyed")
    record("c2_seasons_keyed", rh.load_settings("seasons"), 100, "keyed")
    record("c2_pollution0_keyed", poll0, 100, "keyed")
    # synthetic tiled landscapes (C4/C5 shape at oracle-feasible sizes), reference fed through
    # Environment/Agent objects built from the same host arrays
    for n, steps, pol in ((96, 12, True), (130, 8, False)):
        w = synthetic.tiled_world(n)
        s = synthetic.synthetic_settings(n, pollution=pol)
        record("c4_synth%d_%s_keyed" % (n, "poll" if pol else "nopoll"), s, steps, "keyed", world=w)
    main_life()
    main_spice()


if __name__ == "__main__":
    main()
