import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import common
from sugarscape_utilicalc_b200 import config_from_settings, init_params_from_settings
s = common.load_golden("c1_default_moveeat_mt")["settings"]
s = {**s, "view": {**s["view"], "grid_size": {"x": 1000, "y": 1000}}, "env": {**s["env"], "total_agents": {"blue": 30000, "red": 20000}}}
cfg = config_from_settings(s, rng_mode="keyed")
e = common.engine_world(cfg, options={"trace": 1, "tma": int(sys.argv[1]) if len(sys.argv) > 1 else 1})
e.seed_mt19937(18)
e.init_world(init_params_from_settings(s))
print(e.counters(), flush=True)
for t in range(3):
    print(e.step(1), e.health(), flush=True)
