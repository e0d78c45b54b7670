import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from sugarscape_utilicalc_b200 import config_from_settings, synthetic
from sugarscape_utilicalc_b200.engine import Engine
from sugarscape_utilicalc_b200.strips import state_digest
n, steps = int(sys.argv[1]), int(sys.argv[2])
cfg = config_from_settings(synthetic.synthetic_settings(n), rng_mode="keyed")
w = synthetic.tiled_world(n)
e = Engine(cfg)
e.upload_cells(w["sugar"], w["cap"], None); e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
st = e.step(steps)
a = e.download_agents()
print("single pop", st["population"].astype(int).tolist(), "digest", *state_digest(a["id"], a["x"], a["y"], a["sugar"], n))
