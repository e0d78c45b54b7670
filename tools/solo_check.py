"""Warp-per-window move passes (pass_impl=1 one agent at a time, 2 one lane per agent; argv[1]) against the CPU oracle on
synthetic worlds; prints per-case status and timing."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from sugarscape_utilicalc_b200 import config_from_settings  # noqa: E402

CASES = [(200, True, 8, (1, 6), 1 / 16, 64), (300, False, 8, (1, 6), 1 / 16, 96), (512, True, 8, (1, 6), 1 / 16, 96),
         (1000, True, 5, (1, 6), 1 / 16, 96), (1024, False, 6, (1, 6), 1 / 16, 128), (257, True, 6, (1, 6), 1 / 16, 96),
         (256, True, 5, (1, 4), 0.9, 64), (400, True, 6, (1, 10), 1 / 16, 128), (150, True, 12, (1, 6), 1 / 3, 72),
         (400, False, 8, (1, 6), 1 / 3, 72), (2048, True, 3, (1, 6), 1 / 16, 96), (2048, False, 3, (1, 6), 1 / 16, 128)]
IMPL = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bad = 0
for n, pol, steps, vis, dens, win in CASES:
    s, init = common.synthetic_case(n, pol, vision=vis, density=dens)
    cfg = config_from_settings(s, rng_mode="keyed")
    e = common.engine_world(cfg, options={"pass_impl": IMPL, "solo_window": win})
    o = common.oracle_world(cfg)
    common.init_world(e, init, "keyed")
    common.init_world(o, init, "keyed")
    c = e.counters()
    t0 = time.time()
    try:
        common.compare_worlds(e, o, steps, what="solo %d" % n)
        st = "ok"
    except AssertionError as ex:
        st = "MISMATCH " + str(ex)[:300]
        bad += 1
    c2 = e.counters()
    print("n=%d pol=%s vis=%s dens=%.2f win=%d: impl=%d window=%d passes/step=%.1f %s (%.1fs)" % (
        n, pol, vis, dens, win, c["pass_impl"], c["window"], c2["passes"] / steps, st, time.time() - t0), flush=True)
    e.close()
    o.close()
sys.exit(1 if bad else 0)
