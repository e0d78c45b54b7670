import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
from sugarscape_utilicalc_b200.engine import Engine
n=int(sys.argv[1]); pol=bool(int(sys.argv[2])); win=int(sys.argv[3])
s, init = common.synthetic_case(n, pol)
if len(sys.argv)>4 and sys.argv[4]=="nostarve": s["rules"]["canStarve"]=False
cfg = config_from_settings(s, rng_mode="keyed")
e = Engine(cfg); e.set_option("window", win); common.init_world(e, init, "keyed")
st=e.step(5); print("pop", st["population"], "deaths", st["deaths"])
e.set_option("trace",1); e.step(2)
