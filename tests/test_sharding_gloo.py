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


def _particle_worker(rank, world, port, out_q):
    """K6 on two gloo ranks: particles sharded by contiguous ranges, scan insertion and scoring on the owner,
    results all-gathered, resampling with maps cloned across ranks (send/recv), all with oracle maps"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import room_scan
    from oracle import binding as ob
    from slam_constructor_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)  # identical draws on every rank
    n, size = 7, 64
    lo, hi, chunk = sharding.particle_range(n, rank, world)
    mk = lambda: ob.OracleMap(size, size, 0.1, ob.CELL_MEAN, ob.GROW_TILED)
    mine = {i: mk() for i in range(lo, hi)}
    everything = [mk() for _ in range(n)]  # the single-process run, kept on every rank for comparison
    truth = np.array([0.1, 0.2, 0.1])
    for _ in range(2):
        r, a = room_scan(rng, 61, 2 * np.pi, half_w=2.5, half_h=2.0, pose=truth)
        poses = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 3))
        for i in range(n):
            everything[i].append_scan(ob.OracleScan(r, a), poses[i], 1.0, 0, ob.estimator())
            if i in mine:
                mine[i].append_scan(ob.OracleScan(r, a), poses[i], 1.0, 0, ob.estimator())
    # per-particle scores: local ones computed, padded chunk all-gathered
    cand = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 3, 3))
    sc = ob.OracleScan(r, a)
    buf = torch.full((chunk, 3), float("nan"), dtype=torch.float64)
    for i in range(lo, hi):
        buf[i - lo] = torch.from_numpy(mine[i].score(sc, ob.spe_params(), cand[i]))
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    scores = torch.cat(gathered)[:n].numpy()
    want = np.stack([everything[i].score(sc, ob.spe_params(), cand[i]) for i in range(n)])
    ok_scores = bool(np.array_equal(scores, want))
    # resampling across ranks
    src = [5, 1, 1, 0, 6, 6, 2]
    sends, recvs, local = sharding.resample_plan(src, rank, world)
    staged, reqs = {}, []
    for i, peer in sends:
        m = mine[src[i]]
        info = m.info()
        head = torch.tensor([info["w"], info["h"], info["ox"], info["oy"]], dtype=torch.int64)
        reqs.append(dist.isend(head, peer, tag=2 * i)); reqs.append(dist.isend(torch.from_numpy(m.cells().copy()), peer, tag=2 * i + 1))
    for i, peer in recvs:
        head = torch.zeros(4, dtype=torch.int64)
        dist.recv(head, peer, tag=2 * i)
        cells = torch.zeros((int(head[1]), int(head[0]), 2), dtype=torch.float64)
        dist.recv(cells, peer, tag=2 * i + 1)
        staged[i] = (head.tolist(), cells.numpy())
    for q in reqs:
        q.wait()
    nxt = {}
    for i, (kind, j) in local.items():
        if kind in ("keep", "move"):
            nxt[i] = (mine[j].info(), mine[j].cells().copy())
        elif kind == "copy":
            nxt[i] = (mine[j].info(), mine[j].cells().copy())
        else:
            (w, h, ox, oy), cells = staged[i]
            nxt[i] = (dict(mine[lo].info(), w=w, h=h, ox=ox, oy=oy), cells)
    moved = [j for kind, j in local.values() if kind in ("keep", "move")]
    ok_plan = len(moved) == len(set(moved))  # no map object is handed to two particles
    ok_maps = all(nxt[i][0] == everything[src[i]].info() and np.array_equal(nxt[i][1], everything[src[i]].cells(), equal_nan=True)
                  for i in range(lo, hi))
    out_q.put((rank, ok_scores, ok_plan, ok_maps, len(sends), len(recvs)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_particles_shard_and_resample():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_particle_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_scores, ok_plan, ok_maps, ns, nr in res:
        assert ok_scores and ok_plan and ok_maps, (rank, ok_scores, ok_plan, ok_maps)
    assert res[0][4] == res[1][5] and res[0][5] == res[1][4] and res[0][4] + res[0][5] > 0  # sends pair with receives


def test_window_chunks_cover_the_batch():
    from slam_constructor_b200 import sharding
    for M in (0, 1, 7, 8, 21, 101, 1000):
        for world in (1, 2, 3, 8):
            sharded, chunk = sharding.window_chunks(M, world)
            if not sharded:
                assert chunk == M and (world == 1 or M < 4 * world)
                continue
            spans = [(min(M, chunk * r), min(M, chunk * r + chunk)) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
