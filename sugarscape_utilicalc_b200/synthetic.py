"""Synthetic worlds for the scale configurations.

Host-side initialisation only (component 4 / 14 of the survey stay on the host): the
two-peak capacity map of the shipped 50x50 landscape (`Environment.addSiteHelper`,
environment.py:114-122, called from sugarscape.py:1494-1497) tiled over an NxN lattice, and
a uniform random placement of N^2/16 agents.  The same host arrays feed the oracle and
the engine, so the generator is not parity-relevant.
"""
import copy
import math

import numpy as np

DEFAULT_SITES = ((35, 15, 18), (15, 35, 18))      # settings.json env.sites ne / sw


def site_capacity(n, site, max_capacity):
    """Capacity contribution of one radial site on an n x n grid (environment.py:114-122)."""
    si, sj, r = site
    D = math.sqrt(max(si, n - si) ** 2 + max(sj, n - sj) ** 2) * (r / float(n))
    i = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(n, dtype=np.float64)[None, :]
    di, dj = si - i, sj - j
    dist = np.sqrt(di * di + dj * dj)
    return np.minimum(1 + max_capacity * (1 - dist / D), max_capacity)


def two_peak_capacity(n=50, sites=DEFAULT_SITES, max_capacity=10):
    """The shipped landscape: max over sites, floor 0 (grid starts at 0; `if c > grid`)."""
    cap = np.zeros((n, n), dtype=np.float64)
    for s in sites:
        cap = np.maximum(cap, site_capacity(n, s, max_capacity))
    return cap


def tiled_world(n, density=1.0 / 16.0, seed=1234, vision=(1, 6), metab=(1, 4), endowment=(5, 25),
                tile=50, max_capacity=10):
    """cap[i][j] = cap50[i % 50][j % 50]; sugar = cap; pollution = 0; agents uniform w/o replacement."""
    cap50 = two_peak_capacity(tile, DEFAULT_SITES, max_capacity)
    idx = np.arange(n) % tile
    cap = np.ascontiguousarray(cap50[np.ix_(idx, idx)])
    return dict(_agents(n, density, seed, vision, metab, endowment), cap=cap, sugar=cap.copy(),
                pollution=np.zeros((n, n), dtype=np.float64))


def _agents(n, density, seed, vision, metab, endowment):
    rng = np.random.Generator(np.random.PCG64(seed))
    n_agents = int(n * n * density)
    cells = np.empty(0, dtype=np.int64)
    while cells.size < n_agents:                       # exact count, uniform without replacement
        cand = rng.integers(0, n * n, size=int((n_agents - cells.size) * 1.15) + 16, dtype=np.int64)
        allc = np.concatenate([cells, cand])
        _, first = np.unique(allc, return_index=True)
        cells = allc[np.sort(first)]
    cells = cells[:n_agents]
    world = dict(
        id=np.arange(1, n_agents + 1, dtype=np.int32),
        x=(cells // n).astype(np.int32), y=(cells % n).astype(np.int32),
        vision=rng.integers(vision[0], vision[1] + 1, size=n_agents, dtype=np.int32),
        metab=rng.integers(metab[0], metab[1] + 1, size=n_agents, dtype=np.int32),
        sugar_agent=rng.integers(endowment[0], endowment[1] + 1, size=n_agents).astype(np.float64),
    )
    return world


def tiled_world_strip(n, x0, x1, **kw):
    """The same world as tiled_world(n), but only the lattice rows [x0, x1) of the cell arrays (one rank's strip of a
    multi-GPU run); the agent arrays are those of the whole world."""
    tile, max_capacity = kw.get("tile", 50), kw.get("max_capacity", 10)
    cap50 = two_peak_capacity(tile, DEFAULT_SITES, max_capacity)
    cap = np.ascontiguousarray(cap50[np.ix_(np.arange(x0, x1) % tile, np.arange(n) % tile)])
    ag = _agents(n, kw.get("density", 1.0 / 16.0), kw.get("seed", 1234), kw.get("vision", (1, 6)), kw.get("metab", (1, 4)),
                 kw.get("endowment", (5, 25)))
    return dict(ag, cap=cap, sugar=cap.copy())


#: settings.json of the reference with only the fields the hot path reads filled in
_BASE_SETTINGS = {
    "rules": {"grow": True, "seasons": False, "moveEat": True, "canStarve": True, "pollution": False,
              "tags": False, "combat": False, "limitedLifespan": False, "replacement": False,
              "procreate": False, "transmit": False, "spice": False, "trade": False, "foresight": False,
              "credit": False, "inheritance": False, "disease": False, "utilicalc": False},
    "rule_params": {
        "seasons": {"grow_factor_on_season": 1, "grow_factor_off_season_divisor": 8, "period": 50,
                    "regions": {"north": {"start_x": 0, "start_y": 0, "end_x": 49, "end_y": 24},
                                "south": {"start_x": 0, "start_y": 25, "end_x": 49, "end_y": 49}}},
        "tags": {"tags_length": 11, "tags0": 0, "randomly_distribute_agents": True},
        "combat": {"alpha": 1000000}, "loans": {"rate": 1.05, "duration": 10},
        "disease": {"immune_system_size": 10, "disease_length": 7, "num_diseases": 5,
                    "num_starting_diseases": 2},
        "pollution": {"A": 1, "B": 1, "diffusion_rate": 1, "pollution_start_time": 0,
                      "diffusion_start_time": 0},
        "foresight": {"range_start": 0, "range_end": 10},
        "utilicalc": {"self_interest_scale": None}},
    "view": {"start_paused": True, "screen_size": {"x": 800, "y": 800}, "grid_size": {"x": 50, "y": 50},
             "background_colors": {"R": 250, "G": 250, "B": 250},
             "env_color_hashes": {"sugar": "#F2FA00", "spice": "#9B4722", "both": "#BF8232", "red": "#FA3232",
                                  "pink": "#FA32FA", "blue": "#3232FA", "pollution": "#88C641"},
             "graph_update_frequency": 1, "bound_graph_data": True, "bound_to_n_deviations": 2,
             "location_canvas_grid_size": 5, "create_log_on_exit": False},
    "env": {"is_random": False, "random_seed": 18, "total_agents": {"blue": 0, "red": 0},
            "agent_distributions": {"blue": {"x_min": 0, "x_max": 25, "y_min": 25, "y_max": 49},
                                    "red": {"x_min": 25, "x_max": 49, "y_min": 0, "y_max": 25}},
            "max_capacity": 10,
            "sites": {"ne": {"x": 35, "y": 15, "radius": 18}, "sw": {"x": 15, "y": 35, "radius": 18},
                      "se": {"x": 40, "y": 40, "radius": 14}, "nw": {"x": 10, "y": 10, "radius": 14}},
            "growth_factor": {"grow_factor": 1}},
    "agents": {"max_agent_metabolism": 4, "max_agent_vision": 6, "initial_endowments": {"min": 5, "max": 25},
               "death_age_range": {"min": 60, "max": 100},
               "fertility_age_ranges": {"male": {"min_start": 12, "max_start": 15, "min_end": 50, "max_end": 60},
                                        "female": {"min_start": 12, "max_start": 15, "min_end": 40,
                                                   "max_end": 50}}},
}


def synthetic_settings(n, pollution=False, seasons=False, utilicalc=False, seed=18):
    """settings.json (reference schema) for a synthetic NxN run: grow+moveEat+canStarve [+pollution]."""
    s = copy.deepcopy(_BASE_SETTINGS)
    s["view"]["grid_size"] = {"x": int(n), "y": int(n)}
    s["rules"]["pollution"] = bool(pollution)
    s["rules"]["seasons"] = bool(seasons)
    if utilicalc:
        s["rules"]["utilicalc"], s["rules"]["moveEat"] = True, False
    if seasons:
        r = s["rule_params"]["seasons"]["regions"]
        r["north"] = {"start_x": 0, "start_y": 0, "end_x": n - 1, "end_y": n // 2 - 1}
        r["south"] = {"start_x": 0, "start_y": n // 2, "end_x": n - 1, "end_y": n - 1}
    s["env"]["random_seed"] = int(seed)
    return s
