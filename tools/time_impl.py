"""ms/step of the lattice path at size argv[1] (pollution = argv[2]) for the move-pass kernels argv[3:] (pass_impl values)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.engine import Engine  # noqa: E402

n, pol = int(sys.argv[1]), int(sys.argv[2]) != 0
w = synthetic.tiled_world(n)
cfg = config_from_settings(synthetic.synthetic_settings(n, pollution=pol), rng_mode="keyed")
ref = None
for spec in sys.argv[3:]:
    opts = dict(kv.split("=") for kv in spec.split(","))
    e = Engine(cfg)
    for k, v in opts.items():
        e.set_option(k, int(v))
    e.upload_cells(w["sugar"], w["cap"], w["pollution"])
    e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
    e.step(3)
    c0 = e.counters()
    t0 = time.perf_counter()
    st = e.step(10)
    dt = (time.perf_counter() - t0) / 10
    c1 = e.counters()
    kt = e.kernel_times() if hasattr(e, "kernel_times") else None
    dig = (int(st["population"][-1]), float(st["gini"][-1]))
    if ref is None:
        ref = dig
    print("n=%d pol=%d %s: %.3f ms/step, passes/step %.1f, impl %d window %d, digest %s %s" % (
        n, pol, spec, dt * 1e3, (c1["passes"] - c0["passes"]) / 10, c1["pass_impl"], c1["window"], dig,
        "same" if dig == ref else "DIFFERENT"), flush=True)
    e.close()
