"""Keyed utilicalc rule at size argv[1]: ms/step of the lattice path (warp-per-window kernel) and, with argv[2] = 1, of the
serial kernel on the same world."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1])
w = synthetic.tiled_world(n)
cfg = config_from_settings(synthetic.synthetic_settings(n, utilicalc=True), rng_mode="keyed")
for path in ([-1, 0] if len(sys.argv) > 2 and sys.argv[2] == "1" else [-1]):
    e = Engine(cfg, path=path)
    e.upload_cells(w["sugar"], w["cap"], w["pollution"])
    e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
    K = 5 if path == -1 else 1
    st = e.step(2 if path == -1 else 1)
    pop = int(st["population"][-1])
    c0 = e.counters()
    t0 = time.perf_counter()
    st = e.step(K)
    dt = (time.perf_counter() - t0) / K
    c1 = e.counters()
    print("n=%d agents %d path %d impl %d window %d: %.3f ms/step, %.1f M agent-steps/s, passes/step %.1f, pop %d gini %.6f" % (
        n, pop, c1["path"], c1["pass_impl"], c1["window"], dt * 1e3, pop / dt / 1e6, (c1["passes"] - c0["passes"]) / K,
        int(st["population"][-1]), float(st["gini"][-1])), flush=True)
    e.close()
