"""K steps of the synthetic world of size argv[1] under the utilicalc rule (keyed words) -- the command ncu wraps."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.engine import Engine  # noqa: E402
n, steps = int(sys.argv[1]), int(sys.argv[2])
w = synthetic.tiled_world(n)
e = Engine(config_from_settings(synthetic.synthetic_settings(n, utilicalc=True), rng_mode="keyed"))
e.upload_cells(w["sugar"], w["cap"], w["pollution"])
e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
print(e.step(steps)["population"].astype(int).tolist(), e.counters())
e.close()
