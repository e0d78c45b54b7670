import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import common
from sugarscape_utilicalc_b200 import config_from_settings
def check(n, vmax, seasons, steps=6, density=1/16):
    s, init = common.synthetic_case(n, False, vision=(1, vmax), seasons=seasons, density=density)
    if seasons: s["rule_params"]["seasons"]["period"] = 3
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options={"pass_impl": 3}), common.oracle_world(cfg)
    common.init_world(e, init, "keyed"); common.init_world(o, init, "keyed")
    try:
        common.compare_worlds(e, o, steps, what="x")
        print("ok   n=%d vmax=%d seasons=%s dens=%.2f" % (n, vmax, seasons, density), flush=True)
    except AssertionError as ex:
        print("FAIL n=%d vmax=%d seasons=%s dens=%.2f: %s %s" % (n, vmax, seasons, density, str(ex)[:150], e.health()), flush=True)
check(260, 6, True); check(50, 10, False); check(50, 10, False, density=0.16); check(64, 10, False, density=0.16); check(50, 6, False, density=0.16); check(50, 8, False, density=0.16); check(50, 9, False, density=0.16); check(100, 10, False, density=0.16); check(50, 10, True, density=0.16)
