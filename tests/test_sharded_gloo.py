"""CPU, world_size 2 over gloo: the host-side collective helpers of the multi-GPU path
(variable-length log all-gather, window-row partition, termination agreement)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sugarscape_utilicalc_b200 import sharded


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
        def gather_counts(t):
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return torch.cat(out)

        def gather_padded(buf):
            out = [torch.zeros_like(buf) for _ in range(world)]
            dist.all_gather(out, buf)
            return torch.stack(out)

        ok = True
        for rnd, lens in enumerate([(3, 5), (0, 4), (7, 0), (0, 0)]):
            n = lens[rank]
            rng = np.random.default_rng(100 * rnd + rank)
            payload = torch.from_numpy(rng.integers(0, 256, size=max(n, 1) * sharded.ENTRY_BYTES, dtype=np.uint8))
            counts, parts = sharded.pad_and_gather(payload, n, world, gather_counts, gather_padded)
            ok &= counts == list(lens)
            for r in range(world):
                exp = np.random.default_rng(100 * rnd + r).integers(0, 256, size=max(lens[r], 1) * sharded.ENTRY_BYTES,
                                                                    dtype=np.uint8)[: lens[r] * sharded.ENTRY_BYTES]
                ok &= np.array_equal(parts[r].numpy(), exp)
            # both logs in one buffer: 48-byte entries and the 8-byte compact entries
            lens2 = (5 * rnd + 1, 2)
            n2 = lens2[rank]
            pay2 = torch.from_numpy(np.random.default_rng(7000 + 100 * rnd + rank).integers(
                0, 256, size=max(n2, 1) * sharded.COMPACT_BYTES, dtype=np.uint8))
            c1, p1, c2, p2 = sharded.pad_and_gather(payload, n, world, gather_counts, gather_padded, local2=pay2, n2=n2)
            ok &= c1 == list(lens) and c2 == list(lens2)
            for r in range(world):
                exp2 = np.random.default_rng(7000 + 100 * rnd + r).integers(
                    0, 256, size=max(lens2[r], 1) * sharded.COMPACT_BYTES, dtype=np.uint8)[: lens2[r] * sharded.COMPACT_BYTES]
                ok &= np.array_equal(p2[r].numpy(), exp2) and np.array_equal(p1[r].numpy(), parts[r].numpy())
            # the raw layout ss_shard_apply_gathered consumes: slice r at buf[r], compact entries at off2, counts (world, 2)
            g = sharded.gather_logs(payload, n, pay2, n2, world, gather_counts, gather_padded)
            ok &= g["counts_dev"].tolist() == [[lens[r], lens2[r]] for r in range(world)]
            ok &= g["mx"] == max(lens) and g["mx2"] == max(lens2) and g["off2"] % 16 == 0 and g["buf"].shape == (world, g["stride"])
            for r in range(world):
                ok &= np.array_equal(g["buf"][r, : lens[r] * sharded.ENTRY_BYTES].numpy(), p1[r].numpy())
                ok &= np.array_equal(g["buf"][r, g["off2"]: g["off2"] + lens2[r] * sharded.COMPACT_BYTES].numpy(), p2[r].numpy())
            # termination: every rank derives the same pending count from the gathered lengths
            pend = torch.tensor([100 - sum(counts)])
            dist.all_reduce(pend, op=dist.ReduceOp.MAX)
            ok &= int(pend) == 100 - sum(lens)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_varlen_log_allgather_world2_gloo():
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


@pytest.mark.parametrize("nwin,world", [(1, 2), (8, 2), (32, 8), (35, 4), (137, 8), (3, 8)])
def test_window_rows_partition_covers_every_row_once(nwin, world):
    rows = []
    for r in range(world):
        lo, hi = sharded.window_rows_for_rank(nwin, r, world)
        assert 0 <= lo <= hi <= nwin
        rows += list(range(lo, hi))
    assert rows == list(range(nwin))
