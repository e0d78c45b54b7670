"""Depth of a step's dependency graph (CPU study for the round-2 design): agents in keyed priority
order, an edge A -> B iff priority(A) < priority(B) and their vision crosses (start-of-step positions) share a cell --
the exact predicate of the move pass.  Prints the longest chain and how many agents become ready per round if every agent
acted as soon as all its predecessors had (the "dependency rounds" of the north star).
    python tools/dag_depth.py <n> [steps before measuring]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from sugarscape_utilicalc_b200 import config_from_settings, outputs  # noqa: E402

n = int(sys.argv[1])
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 5
s, init = common.synthetic_case(n, False)
cfg = config_from_settings(s, rng_mode="keyed")
o = common.oracle_world(cfg)
common.init_world(o, init, "keyed")
o.step(warm)
ag = o.download_agents()
ids, x, y, v = ag["id"], ag["x"].astype(np.int64), ag["y"].astype(np.int64), ag["vision"].astype(np.int64)
pr = outputs.keyed_priorities(s["env"]["random_seed"], warm, ids, int(init["id"].max()))
order = np.argsort(pr)
V = int(v.max())
grid = -np.ones((n, n), dtype=np.int64)
grid[x, y] = np.arange(len(ids))
rank = np.empty(len(ids), dtype=np.int64)
rank[order] = np.arange(len(ids))
depth = np.zeros(len(ids), dtype=np.int32)
offs = [(dx, dy) for dx in range(-2 * V, 2 * V + 1) for dy in range(-2 * V, 2 * V + 1) if (dx, dy) != (0, 0)
        and (abs(dx) <= V or abs(dy) <= V)]
n_edges = 0
for k in order:                                   # ascending priority: predecessors are final
    a, best = v[k], 0
    for dx, dy in offs:
        adx, ady = abs(dx), abs(dy)
        if not ((adx <= a and ady <= V) or (ady <= a and adx <= V) or (dy == 0 and adx <= a + V) or (dx == 0 and ady <= a + V)):
            continue
        j = grid[(x[k] + dx) % n, (y[k] + dy) % n]
        if j < 0 or rank[j] > rank[k]:
            continue
        b = v[j]
        if (adx <= a and ady <= b) or (ady <= a and adx <= b) or (dy == 0 and adx <= a + b) or (dx == 0 and ady <= a + b):
            n_edges += 1
            if depth[j] > best:
                best = depth[j]
    depth[k] = best + 1
hist = np.bincount(depth)
print("n=%d agents %d (after %d steps): predecessors per agent %.2f, longest chain %d" % (n, len(ids), warm, n_edges / len(ids), depth.max()))
cum = np.cumsum(hist[1:]) / len(ids)
print("agents ready per round (share of all): " + " ".join("%d:%.3f" % (r + 1, hist[r + 1] / len(ids)) for r in range(min(12, len(hist) - 1))))
print("rounds until 50 / 90 / 99 / 99.9 %% have acted: %s" % [int(np.searchsorted(cum, q) + 1) for q in (0.5, 0.9, 0.99, 0.999)])
