"""Step timings of the dependency-round path: python tools/dr_time.py <n> <pollution 0/1> <steps> [util=1] [option=value ...]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import common
from sugarscape_utilicalc_b200 import config_from_settings
n, pol, steps = int(sys.argv[1]), int(sys.argv[2]) != 0, int(sys.argv[3])
opts = {"pass_impl": 3}
util = False
for kv in sys.argv[4:]:
    if kv.split("=")[0] == "util":
        util = int(kv.split("=")[1]) != 0
    elif kv.split("=")[0] == "lib":         # an A/B build of the library (tools/build_alt.sh)
        from sugarscape_utilicalc_b200 import capi
        capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "alt_%s.so" % kv.split("=")[1])
    else:
        opts[kv.split("=")[0]] = int(kv.split("=")[1])
s, init = common.synthetic_case(n, pol, utilicalc=util)
cfg = config_from_settings(s, rng_mode="keyed")
e = common.engine_world(cfg, options=opts)
common.init_world(e, init, "keyed")
e.step(5)
r0 = e.rounds()
e.set_option("profile", 1); e.set_option("reset_times", 0)
st = e.step(steps)
kt = e.kernel_times()
print("n=%d pol=%s %s: pop %d, rounds/step %.1f, kernel ms/step: %s health=%s" % (
    n, pol, opts, int(st["population"][-1]), (e.rounds() - r0) / steps, {k: round(v / steps, 3) for k, v in kt.items() if k not in ("last_call", "serial", "settle")}, e.health()), flush=True)
e.set_option("profile", 0)
e.step(steps)
print("   unprofiled: %.3f ms/step (device)" % (e.kernel_times()["last_call"] / steps), flush=True)
e.close()
