"""Runs argv[3] timesteps of the synthetic world of size argv[1] (pollution argv[2]) with engine options argv[4:] -- the
command ncu captures (tools/README in profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.engine import Engine  # noqa: E402

n, pol, steps = int(sys.argv[1]), int(sys.argv[2]) != 0, int(sys.argv[3])
w = synthetic.tiled_world(n)
cfg = config_from_settings(synthetic.synthetic_settings(n, pollution=pol), rng_mode="keyed")
e = Engine(cfg)
for kv in sys.argv[4:]:
    e.set_option(kv.split("=")[0], int(kv.split("=")[1]))
e.upload_cells(w["sugar"], w["cap"], w["pollution"])
e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
st = e.step(steps)
print("population", st["population"].astype(int).tolist(), e.counters(), "rounds", e.rounds(), e.health())
e.close()
