"""Drop-in for the reference's timestep: `install(ss, view)` replaces `View.updateGame`
(sugarscape.py:506-686) of a live reference `View` with the engine.

    import sugarscape as ss                       # the unmodified reference module
    ... build env / agents / view as sugarscape.py:1485-1555 does ...
    from sugarscape_utilicalc_b200 import dropin
    handle = dropin.install(ss, view)             # uploads Environment.grid and View.agents
    view.updateGame()                             # now one engine step; stat lists grow as before
    handle.sync_back()                            # write HBM state into Environment / Agent objects (GUI draw)

The state lives in HBM between calls; `sync_back()` is the read the GUI needs before `draw()`.
`rng="mt"` continues the reference's global `random` stream bit-exactly (and writes the MT19937
state back on `sync_back`); `rng="keyed"` selects the parallel path.
"""
import copy
import random as _pyrandom

import numpy as np

from . import outputs
from .settings import config_from_settings


def _snapshot_view(view):
    env = view.env
    n = env.gridWidth
    g = env.grid
    sugar = np.array([[g[i][j][0] for j in range(n)] for i in range(n)], dtype=np.float64)
    cap = np.array([[g[i][j][2] for j in range(n)] for i in range(n)], dtype=np.float64)
    poll = np.array([[g[i][j][5] for j in range(n)] for i in range(n)], dtype=np.float64)
    ags = view.agents
    return dict(
        sugar=sugar, cap=cap, pollution=poll,
        id=np.array([a.id for a in ags], dtype=np.int32), x=np.array([a.x for a in ags], dtype=np.int32),
        y=np.array([a.y for a in ags], dtype=np.int32), vision=np.array([a.vision for a in ags], dtype=np.int32),
        metab=np.array([a.sugarMetabolism for a in ags], dtype=np.int32),
        sugar_agent=np.array([a.sugar for a in ags], dtype=np.float64),
        # what the lifespan / procreation rules read (agent.py:134-151, 204-206)
        life=dict(age=np.array([a.age for a in ags], dtype=np.int32), max_age=np.array([a.maxAge for a in ags], dtype=np.int32),
                  sex=np.array([a.sex for a in ags], dtype=np.int32),
                  fert_lo=np.array([a.fertility[0] for a in ags], dtype=np.int32),
                  fert_hi=np.array([a.fertility[1] for a in ags], dtype=np.int32),
                  endowment=np.array([a.sugarEndowment for a in ags], dtype=np.float64),
                  tags=np.array([a.tags for a in ags], dtype=np.int32),
                  alive=np.array([1 if a.alive else 0 for a in ags], dtype=np.int32)),
        id_increment=int(env.idIncrement),
        # spice
        spice=dict(spice=np.array([[g[i][j][1] for j in range(n)] for i in range(n)], dtype=np.float64),
                   spice_cap=np.array([[g[i][j][3] for j in range(n)] for i in range(n)], dtype=np.float64),
                   spice_metab=np.array([a.spiceMetabolism for a in ags], dtype=np.int32),
                   spice_agent=np.array([a.spice for a in ags], dtype=np.float64),
                   spice_endowment=np.array([a.spiceEndowment for a in ags], dtype=np.float64)))


class DropinHandle:
    def __init__(self, ss, view, world, rng, keyed_seed=None, events=True, gui=False):
        self.ss, self.view, self.world, self.rng = ss, view, world, rng
        self.by_id = {a.id: a for a in view.agents}
        self.original_update = view.updateGame
        self.events, self.gui = events, gui
        self.keyed_seed = ss.seed if keyed_seed is None else keyed_seed
        self.list_ids = [a.id for a in view.agents]
        self.life = bool(ss.rules["limitedLifespan"] or ss.rules["procreate"])
        self.spice = bool(ss.rules["spice"])
        self.n_slots = max(max(self.list_ids, default=1), int(world.n_slots()) if hasattr(world, "n_slots") else 1)

    def _order_before_step(self):
        """The order `updateGame` walks the agents in (sugarscape.py:517-520), recomputed on the host."""
        if self.rng == "mt":
            mt, idx = self.world.get_rng_mt19937()
            return outputs.execution_order_mt(self.list_ids, mt, idx)
        pr = outputs.keyed_priorities(self.keyed_seed, self.view.iteration, self.list_ids, self.n_slots)
        return [self.list_ids[k] for k in np.argsort(pr, kind="stable")]

    def update_game(self):
        """One `updateGame()`: same observable effects on the View -- stat series (sugarscape.py:603-619), death
        events in execution order (:598-600), iteration / env.time (:670-671), the GUI refresh calls (:665-675, with
        `gui=True`: the state is synced back first), the step-100 sample and the plateau rule (:676-686)."""
        v = self.view
        order = self._order_before_step() if self.events and not self.life else None
        st = self.world.step(1)[0]
        if self.events and self.life:         # births and deaths in execution order, reported by the engine
            for line in outputs.event_lines(self.world.events()):
                v.appendToLog(line.strip(), indent="    ")
        elif self.events:
            self.list_ids = [int(i) for i in self.world.download_agents()["id"]]
            for i in outputs.deaths_in_order(order, self.list_ids):
                v.appendToLog("Agent " + str(i) + " died of starvation", indent="    ")
        v.population.append(int(st["population"]))
        v.metabolismMean.append(float(st["metabolism_mean"]))
        v.visionMean.append(float(st["vision_mean"]))
        v.gini.append(float(st["gini"]))
        if self.gui:
            self.sync_back()
            if v.iteration % self.ss.graphUpdateFrequency == 0 and len(v.onGraphs) > 0:
                v.updateGraphs()
        v.iteration += 1
        v.env.incrementTime()
        if self.gui:
            v.updateStatsWindow()
            v.updateLocationStatsWindow()
            v.drawLocationCanvas()
        if v.iteration == 100:
            v.pop_at_100 = v.population[-1]
            v.gini_at_100 = v.gini[-1]
        if v.iteration >= 200:
            v.population_200_ago = v.population[len(v.population) - 200]
        if v.population[-1] == v.population_200_ago:
            data_to_write = [self.ss.seed, v.pop_at_100, round(v.gini_at_100, 2), str(v.population[-1])]
            outputs.update_data_file("data_collection/data.txt", data_to_write)
            raise outputs.PlateauReached()

    def sync_back(self):
        """HBM -> Environment.grid slots 0/4/5 and Agent.x/y/sugar/alive; View.agents in list order."""
        v, env = self.view, self.view.env
        cells, ags = self.world.download_cells(), self.world.download_agents()
        n = env.gridWidth
        sugar, poll = cells["sugar"], cells["pollution"]
        spice = self.world.download_cells_spice()["spice"] if self.spice else None
        sp = self.world.download_agents_spice() if self.spice else None
        for i in range(n):
            row = env.grid[i]
            for j in range(n):
                c = row[j]
                c[0] = float(sugar[i, j])
                if spice is not None:
                    c[1] = float(spice[i, j])
                c[4] = None
                c[5] = float(poll[i, j])
        lf = self.world.download_agents_life() if self.life else None
        if lf is not None:
            env.idIncrement = int(lf["id_increment"])
        alive = []
        for k in range(len(ags["id"])):
            a = self.by_id.get(int(ags["id"][k]))
            if a is None:                     # born on the device: an Agent object for the GUI (no constructor: it draws)
                a = copy.copy(alive[0] if alive else next(iter(self.by_id.values())))
                a.id = int(ags["id"][k])
                a.sugarMetabolism, a.vision = int(ags["metab"][k]), int(ags["vision"][k])
                a.tradeLog, a.sugarLog, a.spiceLog, a.loanLog, a.log = [], [], [], [], []
                a.lent, a.borrowed, a.childIds, a.diseases = [], [], [], {}
                self.by_id[a.id] = a
            a.x, a.y = int(ags["x"][k]), int(ags["y"][k])
            s = float(ags["sugar"][k])
            a.sugar = int(s) if s == int(s) else s
            if sp is not None:
                v2, e2 = float(sp["spice"][k]), float(sp["spice_endowment"][k])
                a.spice = int(v2) if v2 == int(v2) else v2
                a.spiceMetabolism = int(sp["spice_metab"][k])
                a.spiceEndowment = a.spiceCostForMating = int(e2) if e2 == int(e2) else e2
            if lf is not None:
                a.age, a.maxAge, a.sex = int(lf["age"][k]), int(lf["max_age"][k]), int(lf["sex"][k])
                a.fertility = (int(lf["fert_lo"][k]), int(lf["fert_hi"][k]))
                e = float(lf["endowment"][k])
                a.sugarEndowment = a.sugarCostForMating = int(e) if e == int(e) and isinstance(a.sugarEndowment, int) else e
                a.setTags((int(lf["tags"][k]), a.tagsLength))
                a.alive = bool(lf["alive"][k])
            env.grid[a.x][a.y][4] = a
            alive.append(a)
        living = set(id(a) for a in alive)
        for a in v.agents:
            if id(a) not in living:
                a.setAlive(False)
        v.agents[:] = alive
        if self.rng == "mt":
            mt, idx = self.world.get_rng_mt19937()
            _pyrandom.setstate((3, tuple(int(w) for w in mt) + (int(idx),), None))

    def uninstall(self):
        self.sync_back()
        self.view.updateGame = self.original_update


def install(ss, view, rng="mt", keyed_seed=None, world_factory=None, events=True, gui=False):
    """Replace `view.updateGame` by the engine.  `ss` is the imported reference module (its
    `settings` dict is the settings.json the run was started with).  `world_factory(cfg)` builds the
    backend; default = the CUDA engine (tests inject the CPU oracle to check this glue without a GPU).
    `events`: keep `view.log` complete (one id read-back per step -- GUI-scale worlds; switch off for large ones).
    `gui`: sync the state back and call the reference's per-step window refreshes, as `updateGame` does."""
    cfg = config_from_settings(ss.settings, rng_mode=rng, keyed_seed=keyed_seed, iteration=view.iteration,
                               env_time=view.env.getTime())
    if world_factory is None:
        from .engine import Engine
        world_factory = Engine
    world = world_factory(cfg)
    s = _snapshot_view(view)
    world.upload_cells(s["sugar"], s["cap"], s["pollution"])
    world.upload_agents(s["id"], s["x"], s["y"], s["vision"], s["metab"], s["sugar_agent"])
    if cfg.rule_spice:
        sp = s["spice"]
        world.upload_cells_spice(sp["spice"], sp["spice_cap"])
        world.upload_agents_spice(sp["spice_metab"], sp["spice_agent"], sp["spice_endowment"])
    if cfg.rule_limited_lifespan or cfg.rule_procreate:
        world.upload_agents_life(s["id_increment"], **s["life"])
    if rng == "mt":
        st = _pyrandom.getstate()[1]
        world.set_rng_mt19937(np.array(st[:624], dtype=np.uint32), int(st[624]))
    h = DropinHandle(ss, view, world, rng, keyed_seed=keyed_seed, events=events, gui=gui)
    view.updateGame = h.update_game
    return h
