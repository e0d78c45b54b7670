"""Headless driver for the UNMODIFIED reference -- TEST INFRASTRUCTURE (oracle/).

Runs the reference's own hot path (`View.updateGame`, sugarscape.py:506-686) in
this container so that (a) the C restatement in oracle/ss_oracle.c can be pinned
against it and (b) golden vectors can be generated (tests/golden/make_golden.py).
It needs `/root/reference` (or $SS_REFERENCE) and therefore never runs on the GPU
box; nothing in the product package imports it.

How:
  * `tkinter` / `matplotlib` are absent here -> stubbed with MagicMock before the
    reference's `sugarscape.py:17-21` imports them.
  * `settings.json` is opened relative to the CWD at import (sugarscape.py:26-27)
    -> the reference is imported from a scratch directory holding the preset.
  * The setup block under `if __name__ == '__main__'` (sugarscape.py:1485-1555) is
    restated here through the reference's own functions, in the same order (the
    order matters for MT19937 consumption).
  * `updateGame` reads a module-global `env` (sugarscape.py:496,671) and
    `self.onGraphs` (667) -> both are set by the harness.
  * The plateau hack calls `exit()` (sugarscape.py:676-686) after the state is fully
    updated -> `SystemExit` is swallowed.

RNG modes:
  * "mt"    -- the reference's own global `random` (MT19937), untouched.
  * "keyed" -- `agent.random` and `sugarscape.random` are rebound to a shim with the
    `random` module's API that serves the keyed word source of oracle/keyed_rng.py;
    the reference code itself is not modified.
"""
import contextlib
import copy
import json
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import keyed_rng as kr  # noqa: E402

REF = os.environ.get("SS_REFERENCE", "/root/reference")
_REF_MODULES = ("sugarscape", "agent", "environment", "felicific_calculus")


def reference_available():
    return os.path.isfile(os.path.join(REF, "sugarscape.py"))


def load_settings(preset=None):
    """Return the settings dict of a shipped preset ('' / None = default settings.json)."""
    path = os.path.join(REF, "settings.json") if not preset else os.path.join(
        REF, "examples", preset, "settings.json")
    with open(path) as f:
        return json.load(f)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class _KeyedShim:
    """Object with the `random` module's API, serving keyed words (see module doc)."""

    def __init__(self, seed, n_slots):
        self.seed = seed
        self.half = kr.order_half_bits(n_slots)
        self.t = 0
        self.skey = kr.step_key(seed, 0)
        self.rk = kr.round_keys(self.skey)
        self.cur = None  # KeyedStream of the acting agent, None outside an agent's turn

    def begin_step(self, t):
        self.t = t
        self.skey = kr.step_key(self.seed, t)
        self.rk = kr.round_keys(self.skey)
        self.cur = None

    def begin_agent(self, agent_id):
        self.cur = kr.KeyedStream(kr.agent_key(self.skey, agent_id))

    def end_agent(self):
        self.cur = None

    # --- random-module API used on the hot path ---------------------------------
    def shuffle(self, x):
        if self.cur is None:
            # sugarscape.py:517 -- execution order of the step: ascending keyed priority
            x.sort(key=lambda a: kr.priority(self.rk, self.half, a.id - 1))
        else:
            self.cur.shuffle(x)

    def choice(self, seq):
        if not len(seq):
            raise IndexError("Cannot choose from an empty sequence")
        return self.cur.choice(seq)

    def getrandbits(self, k):
        return self.cur.getrandbits(k)

    def randrange(self, start, stop=None):
        if stop is None:
            return self.cur.randbelow(start)
        return start + self.cur.randbelow(stop - start)

    def randint(self, a, b):
        return self.randrange(a, b + 1)

    def random(self):  # pragma: no cover - not on the in-scope path
        raise NotImplementedError("random() is not on the in-scope hot path")


class ReferenceRun:
    """One headless run of the reference.

    settings : dict in the reference's settings.json schema (kept verbatim)
    rng      : "mt" | "keyed"
    world    : None -> the reference's own initialisation (sugarscape.py:1485-1555);
               else dict(cap, sugar, pollution : (N,N) float arrays indexed [x][y];
                         x, y, vision, metab, sugar_agent : 1-D arrays; agent ids are 1..n)
    keyed_seed : seed of the keyed word source (default: settings env.random_seed)
    """

    def __init__(self, settings, rng="mt", world=None, keyed_seed=None):
        assert reference_available(), "reference sources not found at " + REF
        self.settings = copy.deepcopy(settings)
        self.rng = rng
        self.scratch = tempfile.mkdtemp(prefix="ss_ref_")
        with open(os.path.join(self.scratch, "settings.json"), "w") as f:
            json.dump(self.settings, f)
        os.makedirs(os.path.join(self.scratch, "data_collection"), exist_ok=True)
        for m in _REF_MODULES:
            sys.modules.pop(m, None)
        for m in ("tkinter", "matplotlib", "matplotlib.pyplot"):
            if m not in sys.modules or not isinstance(sys.modules[m], MagicMock):
                try:
                    __import__(m)
                except Exception:
                    sys.modules[m] = MagicMock()
        if REF not in sys.path:
            sys.path.insert(0, REF)
        with _cwd(self.scratch):
            import sugarscape as ss  # seeds `random` at import (sugarscape.py:162)
        import agent as agent_mod
        import random as pyrandom
        self.ss, self.agent_mod, self.pyrandom = ss, agent_mod, pyrandom
        if world is None:
            env, agents = self._init_like_main()
        else:
            env, agents = self._init_custom(world)
        ss.env = env                      # sugarscape.py:496,671 read the module global
        self.env = env
        self.view = ss.View(ss.screenSize, env, agents)
        self.view.onGraphs = []           # sugarscape.py:667
        self.exited_at = None
        self.acted_last_step = 0
        self.shim = None
        if rng == "keyed":
            seed = ss.seed if keyed_seed is None else keyed_seed
            n_slots = max([a.id for a in agents], default=1)
            self.shim = _KeyedShim(seed, n_slots)
            ss.random = self.shim
            agent_mod.random = self.shim
            self._wrap_agent_turns()

    # -- sugarscape.py:1485-1555 restated through the reference's own functions ----
    def _init_like_main(self):
        ss = self.ss
        rules = ss.rules
        ss.ruleCheck()
        env = ss.Environment(ss.gridSize)
        env.addSugarSite(ss.sites["northeast"], ss.maxCapacity)
        env.addSugarSite(ss.sites["southwest"], ss.maxCapacity)
        env.setMaxCapacity(ss.maxCapacity)
        if rules["combat"]:
            env.setHasCombat(True)
            env.setCombatAlpha(ss.combatAlpha)
        if rules["pollution"]:
            env.setPollutionRules(ss.pA, ss.pB, ss.pDiffusionRate, ss.pollutionStartTime,
                                  ss.diffusionStartTime)
        if rules["spice"]:
            env.addSpiceSite(ss.sites["southeast"], ss.maxCapacity)
            env.addSpiceSite(ss.sites["northwest"], ss.maxCapacity)
        env.setHasLimitedLifespan(rules["limitedLifespan"])
        if rules["foresight"]:
            env.setHasForesight(True)
            env.setForesightRange(ss.foresightRange)
        if rules["grow"]:
            env.grow(ss.maxCapacity)
        if rules["disease"]:
            env.setHasDisease(True)
            env.setImmuneSystemSize(ss.immuneSystemSize)
            env.setDiseaseLength(ss.diseaseLength)
            for _ in range(ss.numDiseases):
                env.generateDisease()
        if rules["credit"]:
            env.setHasCredit(True)
            env.setLoanRate(ss.loanRate)
            env.setLoanDuration(ss.loanDuration)
        agents = []
        for (n_agents, tags, distribution) in ss.distributions:
            for _ in range(n_agents):
                a = ss.Agent(env)
                if ss.initAgent(a, tags, distribution):
                    env.setAgent(a.getLocation(), a)
                    if rules["disease"]:
                        a.createImmuneSystem()
                        for _ in range(ss.numStartingDiseases):
                            a.addRandomDisease()
                    agents.append(a)
        if rules["utilicalc"]:
            env.setSelfInterestScale(ss.selfInterestScale)
        return env, agents

    def _init_custom(self, w):
        ss = self.ss
        rules = ss.rules
        ss.ruleCheck()
        n = ss.gridSize[0]
        assert ss.gridSize[0] == ss.gridSize[1] == w["cap"].shape[0] == w["cap"].shape[1]
        env = ss.Environment(ss.gridSize)
        env.setMaxCapacity(ss.maxCapacity)
        if rules["pollution"]:
            env.setPollutionRules(ss.pA, ss.pB, ss.pDiffusionRate, ss.pollutionStartTime,
                                  ss.diffusionStartTime)
        env.setHasLimitedLifespan(rules["limitedLifespan"])
        cap, sugar, poll = w["cap"], w["sugar"], w["pollution"]
        for i in range(n):
            row = env.grid[i]
            for j in range(n):
                c = row[j]
                c[0] = float(sugar[i, j])
                c[2] = float(cap[i, j])
                c[5] = float(poll[i, j])
        agents = []
        for k in range(len(w["x"])):
            a = ss.Agent(env, int(w["x"][k]), int(w["y"][k]), int(w["metab"][k]), 4,
                         int(w["vision"][k]), int(w["sugar_agent"][k]))
            assert a.id == k + 1
            assert env.isLocationFree((a.x, a.y))
            env.setAgent((a.x, a.y), a)
            agents.append(a)
        if rules["utilicalc"]:
            env.setSelfInterestScale(ss.selfInterestScale)
        return env, agents

    def _wrap_agent_turns(self):
        shim, Agent = self.shim, self.agent_mod.Agent
        for name in ("move", "utilicalcMove"):
            orig = getattr(Agent, name)

            def wrapped(agent_self, _orig=orig):
                shim.begin_agent(agent_self.id)
                try:
                    return _orig(agent_self)
                finally:
                    shim.end_agent()
            setattr(Agent, name, wrapped)

    # ------------------------------------------------------------------------------
    def step(self):
        """One `View.updateGame()` call (sugarscape.py:506-686)."""
        if self.shim is not None:
            self.shim.begin_step(self.view.iteration)
        with _cwd(self.scratch):
            try:
                self.view.updateGame()
            except SystemExit:           # Q11: state is already fully updated
                if self.exited_at is None:
                    self.exited_at = self.view.iteration

    def run(self, steps):
        for _ in range(steps):
            self.step()

    # ------------------------------------------------------------------------------
    def mt_state(self):
        """(624 uint32 words, index) of the global MT19937 (`random.getstate()`)."""
        st = self.pyrandom.getstate()[1]
        return np.array(st[:624], dtype=np.uint32), int(st[624])

    def snapshot(self):
        env, n = self.env, self.env.gridWidth
        g = env.grid
        sugar = np.array([[g[i][j][0] for j in range(n)] for i in range(n)], dtype=np.float64)
        spice = np.array([[g[i][j][1] for j in range(n)] for i in range(n)], dtype=np.float64)
        cap = np.array([[g[i][j][2] for j in range(n)] for i in range(n)], dtype=np.float64)
        poll = np.array([[g[i][j][5] for j in range(n)] for i in range(n)], dtype=np.float64)
        occ = np.array([[(g[i][j][4].id if g[i][j][4] is not None else 0) for j in range(n)]
                        for i in range(n)], dtype=np.int32)
        ags = self.view.agents
        snap = dict(
            n=n, time=env.time, iteration=self.view.iteration,
            sugar=sugar, spice=spice, cap=cap, pollution=poll, occ=occ,
            order=np.array([a.id for a in ags], dtype=np.int32),
            x=np.array([a.x for a in ags], dtype=np.int32),
            y=np.array([a.y for a in ags], dtype=np.int32),
            vision=np.array([a.vision for a in ags], dtype=np.int32),
            metab=np.array([a.sugarMetabolism for a in ags], dtype=np.int32),
            sugar_agent=np.array([a.sugar for a in ags], dtype=np.float64),
            # what the lifespan / procreation rules read (agent.py:134-151, 204-206)
            age=np.array([a.age for a in ags], dtype=np.int32),
            max_age=np.array([a.maxAge for a in ags], dtype=np.int32),
            sex=np.array([a.sex for a in ags], dtype=np.int32),
            fert_lo=np.array([a.fertility[0] for a in ags], dtype=np.int32),
            fert_hi=np.array([a.fertility[1] for a in ags], dtype=np.int32),
            endowment=np.array([a.sugarEndowment for a in ags], dtype=np.float64),
            tags=np.array([a.tags for a in ags], dtype=np.int32),
            alive=np.array([1 if a.alive else 0 for a in ags], dtype=np.int32),
            id_increment=int(env.idIncrement),
            # spice
            spice_cap=np.array([[g[i][j][3] for j in range(n)] for i in range(n)], dtype=np.float64),
            spice_metab=np.array([a.spiceMetabolism for a in ags], dtype=np.int32),
            spice_agent=np.array([a.spice for a in ags], dtype=np.float64),
            spice_endowment=np.array([a.spiceEndowment for a in ags], dtype=np.float64),
        )
        mt, idx = self.mt_state()
        snap["mt"], snap["mt_index"] = mt, idx
        return snap

    def stats(self):
        v = self.view
        return dict(population=list(v.population), metabolismMean=list(v.metabolismMean),
                    visionMean=list(v.visionMean), gini=list(v.gini))
