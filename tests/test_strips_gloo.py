"""CPU, world_size 2 over gloo: the host-side logic of the row-strip multi-GPU path (sugarscape_utilicalc_b200/strips.py) --
the partition of rows and agents, the per-rank uploads, the merge of the ranks' agent lists into the reference's list order
and the decomposition-independent state digest.  (The kernels of the path run on one GPU in tests/test_engine_gpu.py::
test_row_strips_*; the real multi-process run is bench.py --gpus N.)"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from sugarscape_utilicalc_b200 import config_from_settings, strips


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, steps = 96, 3
        s, init = common.synthetic_case(n, False)
        cfg = config_from_settings(s, rng_mode="keyed")
        o = common.oracle_world(cfg)                       # every rank steps the whole world on the CPU oracle ...
        common.init_world(o, init, "keyed")
        o.step(steps)
        ags, cells = o.download_agents(), o.download_cells()
        x0, x1 = strips.strip_rows(n, rank, world)
        # ... and keeps what a strip engine would hold: its rows and the agents standing on them, in priority order
        mine = (ags["x"] >= x0) & (ags["x"] < x1)
        part = {k: v[mine] for k, v in ags.items()}
        ok = bool((strips.strip_of_row(part["x"], n, world) == rank).all())
        parts = [None] * world
        dist.all_gather_object(parts, part)
        merged = strips.merge_agents(parts, int(cfg.keyed_seed), steps, int(init["id"].max()))
        ok &= all(np.array_equal(merged[k], ags[k]) for k in ("id", "x", "y", "vision", "metab", "sugar"))
        dg = [None] * world
        dist.all_gather_object(dg, strips.state_digest(part["id"], part["x"], part["y"], part["sugar"], n))
        xd, sd, cd = 0, 0, 0
        for d in dg:
            xd ^= d[0]; sd = (sd + d[1]) % (1 << 64); cd += d[2]
        whole = strips.state_digest(ags["id"], ags["x"], ags["y"], ags["sugar"], n)
        ok &= (xd, sd, cd) == (whole[0], whole[1] % (1 << 64), whole[2])
        # what the rank uploads: its rows, its agents, and the directory columns of the whole world
        w = dict(sugar=init["sugar"][x0:x1], cap=init["cap"][x0:x1])
        sp = strips.split_world(x0, x1, w, init, pinned=False)
        cnt = [None] * world
        dist.all_gather_object(cnt, (len(sp["local"][0]), sp["sugar"].shape, sp["max_id"], sp["vmax"]))
        ok &= sum(c[0] for c in cnt) == len(init["id"]) and len(sp["directory"][0]) == len(init["id"])
        ok &= all(c[2] == int(init["id"].max()) and c[3] == int(init["vision"].max()) for c in cnt)
        ok &= sp["sugar"].shape == (x1 - x0, n) and bool(((sp["local"][1] >= x0) & (sp["local"][1] < x1)).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_strip_host_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


@pytest.mark.parametrize("n,world", [(16384, 8), (300, 2), (512, 3), (130, 2), (1000, 7)])
def test_strip_rows_cover_the_lattice_once(n, world):
    rows = []
    for r in range(world):
        lo, hi = strips.strip_rows(n, r, world)
        rows.append((lo, hi))
        assert (strips.strip_of_row(np.arange(lo, hi), n, world) == r).all()
    assert rows[0][0] == 0 and rows[-1][1] == n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
