import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
from sugarscape_utilicalc_b200.engine import Engine
s, init = common.synthetic_case(4096, False)
cfg = config_from_settings(s, rng_mode="keyed")
e = Engine(cfg); e.set_option("window", 128); common.init_world(e, init, "keyed")
e.set_option("trace", int(sys.argv[1]))
for i in range(12):
    t0=time.time(); st=e.step(1); print("STEP", i, st["population"], "%.3fs"%(time.time()-t0), e.counters()["passes"], flush=True)
