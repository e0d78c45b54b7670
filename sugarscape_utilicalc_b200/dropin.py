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
import random as _pyrandom

import numpy as np

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
        sugar_agent=np.array([a.sugar for a in ags], dtype=np.float64))


class DropinHandle:
    def __init__(self, ss, view, world, rng):
        self.ss, self.view, self.world, self.rng = ss, view, world, rng
        self.by_id = {a.id: a for a in view.agents}
        self.original_update = view.updateGame

    def update_game(self):
        """One `updateGame()`: same observable effects on the View (sugarscape.py:603-619, 670-671)."""
        v = self.view
        st = self.world.step(1)[0]
        v.population.append(int(st["population"]))
        v.metabolismMean.append(float(st["metabolism_mean"]))
        v.visionMean.append(float(st["vision_mean"]))
        v.gini.append(float(st["gini"]))
        v.iteration += 1
        v.env.incrementTime()

    def sync_back(self):
        """HBM -> Environment.grid slots 0/4/5 and Agent.x/y/sugar/alive; View.agents in list order."""
        v, env = self.view, self.view.env
        cells, ags = self.world.download_cells(), self.world.download_agents()
        n = env.gridWidth
        sugar, poll = cells["sugar"], cells["pollution"]
        for i in range(n):
            row = env.grid[i]
            for j in range(n):
                c = row[j]
                c[0] = float(sugar[i, j])
                c[4] = None
                c[5] = float(poll[i, j])
        alive = []
        for k in range(len(ags["id"])):
            a = self.by_id[int(ags["id"][k])]
            a.x, a.y = int(ags["x"][k]), int(ags["y"][k])
            s = float(ags["sugar"][k])
            a.sugar = int(s) if s == int(s) else s
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


def install(ss, view, rng="mt", keyed_seed=None, world_factory=None):
    """Replace `view.updateGame` by the engine.  `ss` is the imported reference module (its
    `settings` dict is the settings.json the run was started with).  `world_factory(cfg)` builds the
    backend; default = the CUDA engine (tests inject the CPU oracle to check this glue without a GPU)."""
    cfg = config_from_settings(ss.settings, rng_mode=rng, keyed_seed=keyed_seed, iteration=view.iteration,
                               env_time=view.env.getTime())
    if world_factory is None:
        from .engine import Engine
        world_factory = Engine
    world = world_factory(cfg)
    s = _snapshot_view(view)
    world.upload_cells(s["sugar"], s["cap"], s["pollution"])
    world.upload_agents(s["id"], s["x"], s["y"], s["vision"], s["metab"], s["sugar_agent"])
    if rng == "mt":
        st = _pyrandom.getstate()[1]
        world.set_rng_mt19937(np.array(st[:624], dtype=np.uint32), int(st[624]))
    h = DropinHandle(ss, view, world, rng)
    view.updateGame = h.update_game
    return h
