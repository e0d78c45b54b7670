import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
from sugarscape_utilicalc_b200.engine import Engine
n=int(sys.argv[1])
s, init = common.synthetic_case(n, False)
cfg = config_from_settings(s, rng_mode="keyed")
e = Engine(cfg); common.init_world(e, init, "keyed")
e.step(4)
e.set_option("trace",1); e.step(2)
