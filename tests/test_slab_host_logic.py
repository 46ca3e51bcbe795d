"""CPU coverage of the N > 1 path (world_size-2 gloo): the slab partition every rank computes, the shapes of the
per-rank host arrays, and the id-broadcast plumbing bench.py / tests/slab_worker.py rely on."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ny, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfd = importlib.import_module("macos-cfd_b200")
    row0, nrows = cfd.slab_rows(ny, rank, world)
    # the plumbing: rank 0 makes an opaque id, every rank receives the same bytes
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    mine = [None] * world
    dist.all_gather_object(mine, (row0, nrows, ids[0] == bytes(range(128))))
    # emulate one halo exchange of a row-numbered field over gloo: each rank sends its top 2 / bottom 2 owned rows
    import torch
    field = torch.arange(row0, row0 + nrows, dtype=torch.float64)  # value = global row index
    up, down = rank + 1, rank - 1
    got_up = torch.zeros(2, dtype=torch.float64)
    got_down = torch.zeros(2, dtype=torch.float64)
    ops = []
    if up < world:
        ops += [dist.P2POp(dist.isend, field[-2:].clone(), up), dist.P2POp(dist.irecv, got_up, up)]
    if down >= 0:
        ops += [dist.P2POp(dist.isend, field[:2].clone(), down), dist.P2POp(dist.irecv, got_down, down)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ok = True
    if up < world:
        ok &= got_up.tolist() == [row0 + nrows, row0 + nrows + 1]
    if down >= 0:
        ok &= got_down.tolist() == [row0 - 2, row0 - 1]
    if rank == 0:
        q.put((mine, ok))
    else:
        flag = [ok]
        q.put((None, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("ny", [256, 203, 8192])
def test_slab_partition_and_plumbing_world2(ny):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
    mine = next(m for m, _ in res if m is not None)
    rows = [(a, n) for a, n, same in mine]
    assert all(same for _, _, same in mine)
    assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
    for (a, n), (b, _) in zip(rows, rows[1:]):
        assert a + n == b  # contiguous, no overlap


def test_partition_formula_matches_library_for_many_sizes():
    cfd = importlib.import_module("macos-cfd_b200")
    for ny in (8, 9, 100, 203, 8192, 32768):
        for world in (1, 2, 3, 4, 8):
            rows = [cfd.slab_rows(ny, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][0] + rows[-1][1] == ny
            assert all(a + n == b for (a, n), (b, _) in zip(rows, rows[1:]))
            assert max(n for _, n in rows) - min(n for _, n in rows) <= 1
