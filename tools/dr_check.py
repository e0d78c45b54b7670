"""Dependency-round kernels (pass_impl 3) against the CPU oracle on a few lattices, then step timings at scale."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import common
from sugarscape_utilicalc_b200 import config_from_settings

def check(n, pollution, steps, density=1 / 16, vision=(1, 6), opts=None):
    s, init = common.synthetic_case(n, pollution, density=density, vision=vision)
    cfg = config_from_settings(s, rng_mode="keyed")
    e, o = common.engine_world(cfg, options=opts or {"pass_impl": 3}), common.oracle_world(cfg)
    common.init_world(e, init, "keyed"); common.init_world(o, init, "keyed")
    assert e.counters()["pass_impl"] == 3, e.counters()
    try:
        common.compare_worlds(e, o, steps, what="dr %d" % n)
        print("ok   n=%d pol=%s dens=%.3f vis=%s tma=%s" % (n, pollution, density, vision, (opts or {}).get("tma", 1)), flush=True)
    except AssertionError as ex:
        print("FAIL n=%d pol=%s dens=%.3f vis=%s: %s health=%s" % (n, pollution, density, vision, str(ex)[:300], e.health()), flush=True)
    e.close(); o.close()

def timing(n, pollution, steps=10, impl=3):
    s, init = common.synthetic_case(n, pollution)
    cfg = config_from_settings(s, rng_mode="keyed")
    e = common.engine_world(cfg, options={"pass_impl": impl} if impl >= 0 else None)
    common.init_world(e, init, "keyed")
    e.step(3)
    e.set_option("profile", 1); e.set_option("reset_times", 0)
    t0 = time.perf_counter(); st = e.step(steps); dt = time.perf_counter() - t0
    kt = e.kernel_times()
    print("n=%d pol=%s impl=%d: %.3f ms/step wall(profiled), pop %d, kernel ms/step: %s health=%s" % (
        n, pollution, e.counters()["pass_impl"], dt / steps * 1e3, int(st["population"][-1]),
        {k: round(v / steps, 3) for k, v in kt.items() if k != "last_call"}, e.health()), flush=True)
    e.set_option("profile", 0)
    e.step(steps)
    print("   unprofiled: %.3f ms/step (device)" % (e.kernel_times()["last_call"] / steps), flush=True)
    e.close()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    extra = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[2:]}        # e.g. async_queue=1
    o3 = dict({"pass_impl": 3}, **extra)
    if what in ("all", "check"):
        check(64, False, 6, opts=o3); check(64, True, 6, opts=o3)
        check(200, False, 8, opts=o3); check(177, True, 8, opts=o3); check(257, True, 6, opts=o3); check(300, True, 6, vision=(1, 10), opts=o3)
        check(512, False, 6, opts=dict(o3, tma=0)); check(512, True, 6, opts=o3); check(1000, True, 4, opts=o3)
        check(150, True, 10, density=1 / 3, opts=o3); check(256, True, 4, density=0.9, vision=(1, 4), opts=o3)
    if what in ("all", "time"):
        timing(4096, True); timing(4096, False); timing(16384, False, steps=5)
