"""Diffusion kernels against each other over many lattice sizes: impl 0 (column chains) and 2 (row bands) vs 1 (diagonal sweep)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import common
from sugarscape_utilicalc_b200 import config_from_settings
sizes = [int(a) for a in sys.argv[1:]] or [64, 100, 126, 127, 129, 130, 150, 160, 161, 177, 192, 193, 200, 224, 255, 256, 257, 300, 384, 385]
for n in sizes:
    s, init = common.synthetic_case(n, True)
    cfg = config_from_settings(s, rng_mode="keyed")
    ws = [common.engine_world(cfg, options={"diffuse_impl": k, "pass_impl": 0}) for k in (0, 1, 2)]
    for w in ws:
        common.init_world(w, init, "keyed")
        w.step(3)
    ref = ws[1].download_cells()["pollution"]
    out = []
    for k in (0, 2):
        got = ws[k].download_cells()["pollution"]
        bad = np.argwhere(got != ref)
        out.append("impl %d: %s" % (k, "ok" if len(bad) == 0 else "%d cells differ, first %s, rows %d..%d cols %d..%d" % (
            len(bad), bad[0], bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max())))
    print(n, "|", " | ".join(out), flush=True)
    for w in ws:
        w.close()
