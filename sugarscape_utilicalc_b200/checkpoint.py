"""Checkpoint / resume of a run (the reference never serialises its state; a long engine run
should be able to).  One `.npz` holds everything a step depends on: the settings.json dict, the RNG mode and stream
position (the 624 MT19937 words + index, or the keyed seed), `View.iteration` / `Environment.time`, the three cell
arrays and the agents in list order (with the lifespan / procreation rules also their age, maxAge, sex, fertility, endowment,
tags and alive flag, and Environment.idIncrement).  `restore` builds a fresh world from it; continuing there is bit-identical to
never having stopped (tests/test_engine_gpu.py::test_checkpoint_resume, tests/test_oracle_golden.py on the CPU oracle).
"""
import json

import numpy as np

from .settings import config_from_settings

LIFE_COLUMNS = ("age", "max_age", "sex", "fert_lo", "fert_hi", "endowment", "tags", "alive")


def save(path, world, settings, rng, iteration, keyed_seed=None):
    """world: Engine (or anything with its interface); iteration: timesteps taken since iteration 0 (= env time)."""
    cells, ags = world.download_cells(), world.download_agents()
    # the keyed execution order is sized by the largest Agent.id of the ORIGINAL upload: saved, so that the resumed run
    # keeps the order even if that agent has died since
    n_slots = int(world.n_slots()) if hasattr(world, "n_slots") else int(max(ags["id"], default=1))
    out = dict(settings_json=np.array(json.dumps(settings)), rng=np.array(rng), iteration=np.int64(iteration),
               n_slots=np.int64(n_slots),
               keyed_seed=np.int64(-1 if keyed_seed is None else keyed_seed),
               sugar=cells["sugar"], cap=cells["cap"], pollution=cells["pollution"],
               **{"agent_" + k: ags[k] for k in ("id", "x", "y", "vision", "metab", "sugar")})
    if settings["rules"].get("limitedLifespan") or settings["rules"].get("procreate"):
        lf = world.download_agents_life()
        out.update({"life_" + k: lf[k] for k in LIFE_COLUMNS})
        out["id_increment"] = np.int64(lf["id_increment"])
    if settings["rules"].get("spice"):
        cs, sp = world.download_cells_spice(), world.download_agents_spice()
        out.update(spice=cs["spice"], spice_cap=cs["spice_cap"], agent_spice_metab=sp["spice_metab"], agent_spice=sp["spice"],
                   agent_spice_endowment=sp["spice_endowment"])
    if rng == "mt":
        mt, idx = world.get_rng_mt19937()
        out["mt"], out["mt_index"] = np.asarray(mt, dtype=np.uint32), np.int64(idx)
    np.savez_compressed(path, **out)


def restore(path, world_factory=None, **factory_kwargs):
    """-> (world, settings, rng, iteration).  world_factory(cfg, **factory_kwargs) builds the backend (default: Engine)."""
    z = np.load(path, allow_pickle=False)
    settings, rng, iteration = json.loads(str(z["settings_json"])), str(z["rng"]), int(z["iteration"])
    keyed_seed = int(z["keyed_seed"])
    cfg = config_from_settings(settings, rng_mode=rng, keyed_seed=None if keyed_seed < 0 else keyed_seed,
                               iteration=iteration, env_time=iteration)
    if world_factory is None:
        from .engine import Engine
        world_factory = Engine
    world = world_factory(cfg, **factory_kwargs)
    if "n_slots" in z.files and hasattr(world, "set_option"):
        world.set_option("global_slots", int(z["n_slots"]))
    world.upload_cells(z["sugar"], z["cap"], z["pollution"])
    world.upload_agents(z["agent_id"], z["agent_x"], z["agent_y"], z["agent_vision"], z["agent_metab"], z["agent_sugar"])
    if "spice" in z.files:
        world.upload_cells_spice(z["spice"], z["spice_cap"])
        world.upload_agents_spice(z["agent_spice_metab"], z["agent_spice"], z["agent_spice_endowment"])
    if "id_increment" in z.files:
        world.upload_agents_life(int(z["id_increment"]), **{k: z["life_" + k] for k in LIFE_COLUMNS})
    if rng == "mt":
        world.set_rng_mt19937(z["mt"], int(z["mt_index"]))
    return world, settings, rng, iteration
