"""Per-pass trace of the move passes (option "trace"): python tools/trace_c5.py <n> [pollution 0/1] [opt=value ...]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
n = int(sys.argv[1])
pol = len(sys.argv) > 2 and sys.argv[2] == "1"
opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[3:]}
s, init = common.synthetic_case(n, pol)
cfg = config_from_settings(s, rng_mode="keyed")
e = common.engine_world(cfg, options=opts)
common.init_world(e, init, "keyed")
e.step(4)
e.set_option("trace", 1)
e.step(2)
print(e.counters())
