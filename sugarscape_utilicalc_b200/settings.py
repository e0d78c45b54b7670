"""settings.json -> engine configuration (the reference's schema, kept verbatim).

Mirrors the import-time flattening of `sugarscape.py:26-162` and the rule validation
of `ruleCheck` (`sugarscape.py:182-207`): same exception names, same messages.  Rules
that are outside the accelerated hot path are refused loudly with
`UnsupportedRuleError` -- there is no CPU fallback.
"""
import json

from .capi import SsConfig, RNG_MT19937, RNG_KEYED


class ConflictingRuleException(Exception):
    """sugarscape.py:165-168"""

    def __init__(self, rule1, rule2):
        self.message = str(rule1) + " and " + str(rule2) + " cannot be used together"
        super().__init__(self.message)


class MissingRuleException(Exception):
    """sugarscape.py:171-174"""

    def __init__(self, missingRule, enabledRule):
        self.message = str(enabledRule) + " is enabled, but needs " + str(missingRule) + " to function properly"
        super().__init__(self.message)


class UnsupportedRuleError(Exception):
    """A rule (or rule combination) the B200 engine does not accelerate."""


#: rules.* keys the engine implements; everything else must be false
SUPPORTED_RULES = ("grow", "seasons", "moveEat", "canStarve", "pollution", "utilicalc", "limitedLifespan", "procreate", "spice")
UNSUPPORTED_RULES = ("tags", "combat", "replacement", "transmit",
                     "trade", "foresight", "credit", "inheritance", "disease")


def load_settings(path):
    with open(path) as f:
        return json.load(f)


def rule_check(settings):
    """`ruleCheck()` of sugarscape.py:182-207, same order of tests."""
    rules = settings["rules"]
    if rules["moveEat"]:
        if rules["combat"]:
            raise ConflictingRuleException("moveEat", "combat")
        if rules["utilicalc"]:
            raise ConflictingRuleException("moveEat", "utilicalc")
    if rules["combat"]:
        if not rules["tags"]:
            raise MissingRuleException("tags", "combat")
        if rules["spice"]:
            raise ConflictingRuleException("combat", "spice")
    if rules["pollution"]:
        if rules["spice"]:
            raise ConflictingRuleException("pollution", "spice")
    if not rules["spice"]:
        if rules["trade"]:
            raise MissingRuleException("spice", "trade")
        if rules["foresight"]:
            raise MissingRuleException("spice", "foresight")
        if rules["credit"]:
            raise MissingRuleException("spice", "credit")
        if rules["inheritance"]:
            raise MissingRuleException("spice", "inheritance")


def check_supported(settings):
    rules = settings["rules"]
    for r in UNSUPPORTED_RULES:
        if rules.get(r, False):
            raise UnsupportedRuleError(
                "rule '%s' is outside the accelerated hot path (SURVEY.md section 8); refusing -- no CPU fallback" % r)
    if rules["utilicalc"] and rules["pollution"]:
        raise UnsupportedRuleError(
            "utilicalc+pollution raises KeyError: 'duration' in the reference (agent.py:880); parity is unpinned")
    if (rules.get("limitedLifespan", False) or rules.get("procreate", False)) and not rules["moveEat"]:
        raise UnsupportedRuleError("limitedLifespan / procreate are built with rule M (moveEat) only: with utilicalc the "
                                   "reference needs spice (examples/utilicalc_spice raises KeyError 'proximity')")


def config_from_settings(settings, rng_mode="mt", keyed_seed=None, iteration=0, env_time=0):
    """Fill an `SsConfig` from a settings dict (sugarscape.py:32-160)."""
    rule_check(settings)
    check_supported(settings)
    rules, rp, env = settings["rules"], settings["rule_params"], settings["env"]
    gx, gy = settings["view"]["grid_size"]["x"], settings["view"]["grid_size"]["y"]
    if gx != gy:
        raise UnsupportedRuleError("non-square grids are broken in the reference (environment.py:34,119-122)")
    cfg = SsConfig()
    cfg.grid_size = gx
    cfg.rule_grow = int(bool(rules["grow"]))
    cfg.rule_seasons = int(bool(rules["seasons"]))
    cfg.rule_move_eat = int(bool(rules["moveEat"]))
    cfg.rule_utilicalc = int(bool(rules["utilicalc"]))
    cfg.rule_can_starve = int(bool(rules["canStarve"]))
    cfg.rule_pollution = int(bool(rules["pollution"]))
    s = rp["seasons"]
    cfg.season_period = int(s["period"])
    for k, name in enumerate(("start_x", "start_y", "end_x", "end_y")):
        cfg.season_north[k] = int(s["regions"]["north"][name])
        cfg.season_south[k] = int(s["regions"]["south"][name])
    on = s["grow_factor_on_season"]
    cfg.season_on_factor = float(on)
    cfg.season_off_factor = float(on) / s["grow_factor_off_season_divisor"]     # sugarscape.py:85
    p = rp["pollution"]
    cfg.diffusion_rate = int(p["diffusion_rate"])
    cfg.pollution_start_time = int(p["pollution_start_time"])
    cfg.diffusion_start_time = int(p["diffusion_start_time"])
    cfg.pollution_a = float(p["A"])
    cfg.pollution_b = float(p["B"])
    cfg.grow_factor = float(env["growth_factor"]["grow_factor"])
    cfg.max_capacity = float(env["max_capacity"])
    cfg.rng_mode = RNG_MT19937 if rng_mode == "mt" else RNG_KEYED
    cfg.keyed_seed = int(env["random_seed"] if keyed_seed is None else keyed_seed) & ((1 << 64) - 1)
    cfg.iteration = int(iteration)
    cfg.env_time = int(env_time)
    cfg.rule_limited_lifespan = int(bool(rules.get("limitedLifespan", False)))
    cfg.rule_procreate = int(bool(rules.get("procreate", False)))
    if cfg.rule_limited_lifespan or cfg.rule_procreate:
        if rng_mode != "mt":
            raise UnsupportedRuleError("limitedLifespan / procreate run in the MT19937-exact mode only (births change the "
                                       "id range the keyed order is sized by)")
        cfg.tags_length = int(rp["tags"]["tags_length"])                    # sugarscape.py:110
    cfg.foresight_lo, cfg.foresight_hi = 0, 0                               # Environment.foresightRange without the rule
    cfg.rule_spice = int(bool(rules.get("spice", False)))
    if cfg.rule_spice and not rules["moveEat"]:
        raise UnsupportedRuleError("spice is built with rule M (moveEat): examples/utilicalc_spice raises KeyError 'proximity' in the reference")
    return cfg


def init_params_from_settings(settings):
    """`ss_init_params` for `ss_init_world`: what the reference's `__main__` reads from settings.json to build the
    initial world (sugarscape.py:48-108, 138-146, 1494-1497, 1525-1547)."""
    from .capi import SsInitParams
    check_supported(settings)
    env, ag, rules = settings["env"], settings["agents"], settings["rules"]
    p = SsInitParams()
    sites = [env["sites"]["ne"], env["sites"]["sw"]]                    # northeast, southwest: the two sugar peaks
    p.n_sites = len(sites)
    for k, s in enumerate(sites):
        p.site_x[k], p.site_y[k], p.site_radius[k] = s["x"], s["y"], s["radius"]
    p.grow_to_capacity = 1 if rules["grow"] else 0
    groups = [env["total_agents"]["blue"], env["total_agents"]["red"]]  # `distributions`: blues first
    p.n_groups = len(groups)
    for k, c in enumerate(groups):
        p.group_count[k] = c
    p.max_metabolism, p.max_vision = ag["max_agent_metabolism"], ag["max_agent_vision"]
    p.endowment_min, p.endowment_max = ag["initial_endowments"]["min"], ag["initial_endowments"]["max"]
    p.age_min, p.age_max = ag["death_age_range"]["min"], ag["death_age_range"]["max"]
    fa = ag["fertility_age_ranges"]
    for sex, who in enumerate(("male", "female")):                      # fertility[0] = male tuple is indexed by sex 0
        r = fa[who]
        for k, key in enumerate(("min_start", "max_start", "min_end", "max_end")):
            p.childbearing[sex][k] = r[key]
    p.foresight_lo, p.foresight_hi = 0, 0                               # Environment.foresightRange without the rule
    tg = settings["rule_params"]["tags"]
    p.group_tags[0], p.group_tags[1] = int(tg["tags0"]), 2 ** int(tg["tags_length"]) - 1      # sugarscape.py:104-105
    if rules.get("spice"):                                              # sugarscape.py:1508-1513: southeast, northwest
        sp_sites = [env["sites"]["se"], env["sites"]["nw"]]
        p.n_spice_sites = len(sp_sites)
        for k, s in enumerate(sp_sites):
            p.spice_site_x[k], p.spice_site_y[k], p.spice_site_radius[k] = s["x"], s["y"], s["radius"]
    return p
