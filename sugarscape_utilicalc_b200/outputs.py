"""On-disk outputs and per-step bookkeeping of the reference's `View`, for runs driven by the engine.

What `View.updateGame` leaves behind besides the state itself:
  * the event log -- one line per death, in execution order: "    Agent N died of starvation"
    (`appendToLog`, sugarscape.py:1406-1410, called at :598-600);
  * `pop_at_100`, `gini_at_100`, `population_200_ago` and the plateau rule: once the population equals the
    one 200 steps earlier, `[seed, pop_at_100, round(gini_at_100, 2), str(population)]` is appended column-wise
    to `data_collection/data.txt` and the run exits (sugarscape.py:673-709);
  * the log file of `createLogFile` (sugarscape.py:1412-1478): settings dump, seed, the stat series rounded to
    four places, the ending statistics of `makeStatsDict` (:1046-1081) and the event log.
All of it is host-side text handling; the engine supplies the numbers.  The execution order of a step -- needed
to put a step's deaths in the reference's order -- is recomputed on the host: CPython's own `random.shuffle`
over the MT19937 state the engine reports (MT-exact mode), or the keyed priority (keyed mode; same constants as
csrc/ss_device.cuh).
"""
import datetime
import os
import random as _pyrandom

import numpy as np

_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15
_C_ROUND = 0xA24BAED4963EE407


def event_lines(events):
    """The `appendToLog` lines of the engine's event records (`Engine.events()`: kind | step << 8, a, b, c per event, in
    execution order): sugarscape.py:551 "Agent A mated with agent B and created agent C", :597-600 "Agent A
    died of old age" / "... of starvation" -- each with the four-space indent of `indent="    "`."""
    out = []
    for e in np.asarray(events).reshape(-1, 4):
        kind = int(e[0]) & 255
        if kind == 1:
            out.append("    Agent %d mated with agent %d and created agent %d" % (e[1], e[2], e[3]))
        else:
            out.append("    Agent %d died of %s" % (e[1], "old age" if kind == 3 else "starvation"))
    return out


def _mix64(z):
    z &= _M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


def _fmix32(h):
    h = h & np.uint32(0xFFFFFFFF)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)) & np.uint32(0xFFFFFFFF)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)) & np.uint32(0xFFFFFFFF)
    h ^= h >> np.uint32(16)
    return h


def keyed_priorities(seed, iteration, ids, n_slots):
    """Execution priority (ascending = earlier) of agents `ids` (1-based Agent.id) in step `iteration` of a
    keyed-RNG run with `n_slots` = the largest id uploaded: ss_priority of csrc/ss_device.cuh in numpy."""
    kb = max(1, (max(int(n_slots), 1) - 1).bit_length())
    half = max(1, (kb + 1) // 2)
    skey = _mix64((int(seed) + _GOLDEN * (int(iteration) + 1)) & _M64)
    rk = [np.uint32(_mix64((skey + (r + 1) * _C_ROUND) & _M64) >> 32) for r in range(4)]
    slot = np.asarray(ids, dtype=np.int64).astype(np.uint32) - np.uint32(1)
    mask = np.uint32((1 << half) - 1)
    L, R = (slot >> np.uint32(half)) & mask, slot & mask
    with np.errstate(over="ignore"):
        for r in range(4):
            L, R = R, L ^ (_fmix32(R ^ rk[r]) & mask)
    return (L.astype(np.uint64) << np.uint64(half)) | R.astype(np.uint64)


def execution_order_mt(list_ids, mt_words, mt_index):
    """The agent list after `random.shuffle(self.agents)` (sugarscape.py:517) for the given list and MT19937 state."""
    r = _pyrandom.Random()
    r.setstate((3, tuple(int(w) for w in mt_words) + (int(mt_index),), None))
    ids = [int(i) for i in list_ids]
    r.shuffle(ids)
    return ids


def deaths_in_order(order_ids, surviving_ids):
    """Ids that acted in a step's order and are gone afterwards, in that order."""
    alive = set(int(i) for i in surviving_ids)
    return [i for i in order_ids if i not in alive]


def death_event(agent_id):
    """The log line of sugarscape.py:600 (`indent="    "`)."""
    return "    " + "Agent " + str(int(agent_id)) + " died of starvation"


def update_data_file(filename, data):
    """Column-append of `data` to `filename` (sugarscape.py:688-709): value k goes to the end of line k, separated
    by ", " from what the line already holds; lines beyond len(data) are kept as they are."""
    rows = []
    if os.path.exists(filename):
        with open(filename, "r") as f:
            rows = f.readlines()
    rows += [""] * max(0, len(data) - len(rows))
    for k, value in enumerate(data):
        held = rows[k].strip()
        rows[k] = (held + ", " if held else "") + str(value) + "\n"
    with open(filename, "w") as f:
        f.writelines(rows)


def ending_stats(iteration, population, metabolism_mean, vision_mean, gini, agent_wealth):
    """`View.makeStatsDict` (sugarscape.py:1046-1081) for the rules in scope (no trade/foresight/tags/disease)."""
    if iteration == 0:
        return {"Iteration": 0, "Average Population": 0, "Metabolism Mean": 0, "Vision Mean": 0, "Gini Coefficient": 0,
                "Average Gini": 0, "Average Wealth": 0}
    wealth = list(agent_wealth)
    return {
        "Iteration": iteration,
        "Average Population": round((sum(population) / len(population)) if len(population) > 0 else 0, 2),
        "Metabolism Mean": round(sum(metabolism_mean) / len(metabolism_mean), 2),
        "Vision Mean": round(sum(vision_mean) / len(vision_mean), 2),
        "Gini Coefficient": round(gini[-1], 2),
        "Average Gini": round(sum(gini) / len(gini), 2),
        "Average Wealth": round((sum(wealth) / len(wealth)) if len(wealth) > 0 else 0, 2),
    }


def log_file_text(settings, seed, metabolism_mean, vision_mean, gini, stats, log):
    """The text `View.createLogFile` writes (sugarscape.py:1419-1478)."""
    out = ["\nSettings:\n"]
    for key, value in settings.items():
        out.append(str(key) + ": " + str(value) + "\n")
    out.append("Seed: " + str(seed) + "\n")
    out.append("\nStat lists:\n")
    out.append("Metabolism Mean:\n")
    out.append(", ".join(str(round(x, 4)) for x in metabolism_mean))
    out.append("\n")
    out.append("\nVision Mean:\n")
    out.append(", ".join(str(round(x, 4)) for x in vision_mean))
    out.append("\n")
    out.append("\nGini:\n")
    out.append(", ".join(str(round(x, 4)) for x in gini))
    out.append("\n")
    out.append("\nEnding Stats:\n")
    for key, value in stats.items():
        out.append(str(key) + ": " + str(value) + "\n")
    out.append("\nLog:\n")
    for item in log:
        out.append(item + "\n")
    return "".join(out)


def create_log_file(settings, seed, metabolism_mean, vision_mean, gini, stats, log, directory="./logs", now=None):
    """`View.createLogFile`: ./logs/log_<timestamp>.txt; returns the path."""
    if not os.path.exists(directory):
        os.makedirs(directory)
    now = now or datetime.datetime.now()
    path = os.path.join(directory, "log_" + str(now.strftime("%Y-%m-%d_%H-%M-%S")) + ".txt")
    with open(path, "w") as f:
        f.write(log_file_text(settings, seed, metabolism_mean, vision_mean, gini, stats, log))
    return path


class PlateauReached(SystemExit):
    """The reference calls `exit()` when the population equals the one 200 steps earlier (sugarscape.py:679-682)."""


class RunRecorder:
    """Headless stand-in for the `View` bookkeeping around `updateGame`: the four stat series, the event log,
    the step-100 sample and the plateau rule.  Feed it one engine step at a time:

        rec = RunRecorder(settings, world)          # world: Engine (or anything with its interface)
        while True: rec.step()                      # raises PlateauReached after writing data_collection/data.txt
    """

    def __init__(self, settings, world, rng="mt", keyed_seed=None, list_ids=None, iteration=0,
                 data_file="data_collection/data.txt", events=True):
        self.settings, self.world, self.rng = settings, world, rng
        self.seed = settings["env"]["random_seed"]
        self.keyed_seed = self.seed if keyed_seed is None else keyed_seed
        self.iteration = iteration
        self.population, self.metabolismMean, self.visionMean, self.gini = [], [], [], []
        self.log = []
        self.population_200_ago = 0
        self.pop_at_100 = 0
        self.gini_at_100 = 0
        self.data_file = data_file
        self.events = events
        # with births the engine reports the step's events itself (mating lines, deaths) in execution order
        self.life = bool(settings["rules"].get("limitedLifespan") or settings["rules"].get("procreate"))
        self.list_ids = list(int(i) for i in (world.download_agents()["id"] if list_ids is None else list_ids))
        self.n_slots = max(max(self.list_ids, default=1), int(world.n_slots()) if hasattr(world, "n_slots") else 1)

    def _order_before_step(self):
        if self.rng == "mt":
            mt, idx = self.world.get_rng_mt19937()
            return execution_order_mt(self.list_ids, mt, idx)
        pr = keyed_priorities(self.keyed_seed, self.iteration, self.list_ids, self.n_slots)
        return [self.list_ids[k] for k in np.argsort(pr, kind="stable")]

    def step(self):
        """One timestep with the observable effects of sugarscape.py:575-619 and 670-682."""
        order = self._order_before_step() if self.events and not self.life else None
        st = self.world.step(1)[0]
        if self.events and self.life:
            self.log.extend(event_lines(self.world.events()))
        elif self.events:
            self.list_ids = [int(i) for i in self.world.download_agents()["id"]]
            for i in deaths_in_order(order, self.list_ids):
                self.log.append(death_event(i))
        self.record(st)
        return st

    def record(self, st):
        self.population.append(int(st["population"]))
        self.metabolismMean.append(float(st["metabolism_mean"]))
        self.visionMean.append(float(st["vision_mean"]))
        self.gini.append(float(st["gini"]))
        self.iteration += 1
        if self.iteration == 100:
            self.pop_at_100 = self.population[-1]
            self.gini_at_100 = self.gini[-1]
        if self.iteration >= 200 and len(self.population) >= 200:      # (a recorder started on a resumed run has no history yet)
            self.population_200_ago = self.population[len(self.population) - 200]
        if self.population[-1] == self.population_200_ago:
            update_data_file(self.data_file, [self.seed, self.pop_at_100, round(self.gini_at_100, 2), str(self.population[-1])])
            raise PlateauReached()

    def stats(self):
        wealth = self.world.download_agents()["sugar"]
        return ending_stats(self.iteration, self.population, self.metabolismMean, self.visionMean, self.gini,
                            [int(s) if s == int(s) else float(s) for s in wealth])

    def create_log_file(self, directory="./logs", now=None):
        return create_log_file(self.settings, self.seed, self.metabolismMean, self.visionMean, self.gini, self.stats(),
                               self.log, directory, now)
