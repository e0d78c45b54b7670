import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import common
from sugarscape_utilicalc_b200 import config_from_settings
n, vmax = int(sys.argv[1]), int(sys.argv[2])
s, init = common.synthetic_case(n, False, vision=(1, vmax))
cfg = config_from_settings(s, rng_mode="keyed")
e = common.engine_world(cfg, options={"pass_impl": 3})
common.init_world(e, init, "keyed")
try:
    st = e.step(3)
    print("n=%d vmax=%d ok" % (n, vmax), st["population"], e.health(), flush=True)
except Exception as ex:
    print("n=%d vmax=%d FAIL %s" % (n, vmax, str(ex)[-60:]), flush=True)
