"""Phase clocks of the warp-per-window move passes (instrumented build: tools/build_dbg.sh).
python tools/dbg_solo.py <n> <pollution 0/1> [opt=value ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sugarscape_utilicalc_b200 import capi  # noqa: E402
capi.LIB_PATH = os.path.join(ROOT, "sugarscape_utilicalc_b200", "csrc", "libsugarscape_b200_dbg.so")
from sugarscape_utilicalc_b200 import config_from_settings, synthetic  # noqa: E402
from sugarscape_utilicalc_b200.engine import Engine, _ptr  # noqa: E402

n, pol = int(sys.argv[1]), int(sys.argv[2]) != 0
opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[3:]}
w = synthetic.tiled_world(n)
cfg = config_from_settings(synthetic.synthetic_settings(n, pollution=pol), rng_mode="keyed")
e = Engine(cfg)
for k, v in opts.items():
    e.set_option(k, v)
e.upload_cells(w["sugar"], w["cap"], w["pollution"])
e.upload_agents(w["id"], w["x"], w["y"], w["vision"], w["metab"], w["sugar_agent"])
e.step(4)
e.set_option("debug", 1)
e.set_option("reset_counters", 0)
K = 4
e.step(K)
out = np.zeros(38, dtype=np.int64)
e._lib.ss_get_counters(e._h, _ptr(out), 38)
d = out[6:].astype(float)
us = 1e6 / 1.9e9
W = max(d[17], 1.0)
print("n=%d pol=%d %s: ms/step %.2f passes/step %.1f window %d" % (n, pol, opts, e.kernel_times()["last_call"] / K, out[1] / K, out[5] & 0xffff))
print("windows with work per step %.0f, listed/window %.1f, rounds/window %.1f, lanes/round %.1f, blocked %.0f per step" % (
    W / K, d[28] / W, d[25] / W, d[26] / max(d[25], 1), d[27] / K))
print("warp-time per window with work (us): scan %.1f sort %.1f | fill %.1f blocked-test %.1f batch-test %.1f q1 %.1f act %.1f" % (
    d[18] / W * us, d[19] / W * us, d[20] / W * us, d[21] / W * us, d[22] / W * us, d[23] / W * us, d[24] / W * us))
tot = sum(d[18:25])
print("share: scan %.2f sort %.2f fill %.2f blk %.2f dep %.2f q1 %.2f act %.2f; warp-seconds per step %.3f ms x warps" % (
    d[18] / tot, d[19] / tot, d[20] / tot, d[21] / tot, d[22] / tot, d[23] / tot, d[24] / tot, tot * us / 1e3 / K))
e.close()
