import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import common
from sugarscape_utilicalc_b200 import config_from_settings
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s, init = common.synthetic_case(n, False)
cfg = config_from_settings(s, rng_mode="keyed")
e = common.engine_world(cfg, options={"pass_impl": 3, "trace": 1})
print("created", flush=True)
common.init_world(e, init, "keyed")
print("uploaded", e.counters(), flush=True)
for t in range(3):
    st = e.step(1)
    print("step", t, st, e.health(), flush=True)
