"""The N > 1 protocol on CPU: two gloo ranks shard a candidate set the way the library does (contiguous
slices / whole grid rows), score their slices with the oracle, all-gather one (score, index) pair per rank
and merge; every rank must end with the single-process answer, ties and the accept rule included."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import room_map_cells, room_scan
    from oracle import binding as ob
    from slam_constructor_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # same inputs on every rank (replicated map and scan)
    if case == "ties":
        cells = np.zeros((200, 200, 2)); cells[..., 0] = 0.7; cells[..., 1] = 2
    else:
        cells = room_map_cells(rng, 120, 120, 0.1, ob.CELL_MEAN, passes=2)
    h = cells.shape[0]
    om = ob.OracleMap(h, h, 0.1, ob.CELL_MEAN)
    om.set_cells(cells)
    r, a = room_scan(rng, 61, np.deg2rad(270), pose=(0.2, -0.1, 0.3))
    sc = ob.OracleScan(r, a)
    nx, ny, nt = 13, 7, 5
    xs = 0.2 + 0.05 * (np.arange(nx) - nx // 2); ys = -0.1 + 0.05 * (np.arange(ny) - ny // 2); ts = 0.3 + 0.02 * (np.arange(nt) - 2)
    P = np.stack(np.meshgrid(ts, ys, xs, indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
    full = om.score(sc, ob.spe_params(), P)
    init = {"plain": float(np.median(full)), "ties": 0.1, "reject": float(full.max())}[case]
    b, e = sharding.grid_slice(nx, ny, nt, rank, world)
    assert (b % nx, e % nx) == (0, 0)
    mine = om.score(sc, ob.spe_params(), P[b:e]) if e > b else np.zeros(0)
    s, i = sharding.local_best(mine.tolist(), b)
    t = torch.tensor([s, float(i)], dtype=torch.float64)  # the library ships {f64 score, i64 index, i64 guard, pad}
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, t)
    best, idx = sharding.merge_best([(float(g[0]), int(g[1])) for g in gathered], init)
    # single-process reference: the oracle's sequential accept loop over the whole list
    import ctypes as C
    bs = C.c_double()
    want_idx = ob.orc.orc_argbest(ob.dptr(ob.f64(full)), len(full), init, C.byref(bs))
    out_q.put((rank, best, idx, bs.value, want_idx, b, e))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["plain", "ties", "reject"])
def test_two_rank_shard_and_merge(case):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][5] == 0 and res[0][6] == res[1][5] and res[1][6] == 13 * 7 * 5  # slices tile the set
    for rank, best, idx, want_best, want_idx, b, e in res:
        assert (best, idx) == (want_best, want_idx), (case, rank)
    if case == "ties":
        assert res[0][2] == 0
    if case == "reject":
        assert res[0][2] == -1


def test_slices_tile_any_set():
    from slam_constructor_b200 import sharding
    for total in (0, 1, 7, 100, 1020100, 10100):
        for n in (1, 2, 3, 4, 8):
            edges = [sharding.slice_of(total, r, n) for r in range(n)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[k][1] == edges[k + 1][0] for k in range(n - 1))
            assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1
