import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
from sugarscape_utilicalc_b200.sharded import VirtualRanks
from sugarscape_utilicalc_b200.engine import Engine
n=int(sys.argv[1])
s, init = common.synthetic_case(n, False)
cfg = config_from_settings(s, rng_mode="keyed")
for mode in ("plain", "shard1"):
    w = Engine(cfg) if mode == "plain" else VirtualRanks(cfg, 1)
    common.init_world(w, init, "keyed")
    w.step(3)
    t0=time.perf_counter(); w.step(5); dt=(time.perf_counter()-t0)/5
    print(mode, "ms/step %.2f" % (dt*1e3), flush=True)
    w.close()
