#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank computes each result twice on its own GPU -- through a distributed ctx (work sharded over the ranks,
NCCL exchange) and through a plain single-GPU ctx -- and the two must be identical:
  K1  candidate grid sharded by rows, 32-byte all-gather of the per-rank arg-max;
  K5  the root matches of BF-M3RSM sharded by index, all-gather of the bounds, identical host engine on every rank;
  K6  GMapping particles sharded by particle: batched insertion, per-particle scoring, lock-step hill climbing,
      all-gather of the results, and a resampling step that clones maps across ranks (ncclSend/ncclRecv).
Prints one JSON line on rank 0; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import slam_constructor_b200 as sg  # noqa: E402


def room(rng, n, fov, hw, hh, pose, noise=0.01):
    return bench.room_ranges(rng, n, fov, hw, hh, pose, noise)


def check_k1(dctx, lctx, out):
    rng = np.random.default_rng(11)
    size, scale = 400, 0.05
    pose = np.array([0.2, -0.1, 0.05])
    maps = []
    r, a = room(rng, 361, 2 * np.pi, 6.0, 5.0, pose)
    for ctx in (dctx, lctx):
        gm = sg.GridMap(ctx, size, size, scale, sg.CELL_MEAN, sg.GROW_NONE)
        sc = sg.Scan(ctx, r, a)
        ctx.append_scan(gm, sc, pose, 1.0, 0, sg.estimator(sg.EST_CONST))
        maps.append((gm, sc))
    xs = pose[0] + np.arange(-20, 21) * 0.02
    ys = pose[1] + np.arange(-20, 21) * 0.02
    ts = pose[2] + np.arange(-15, 16) * 0.01
    res = []
    for ctx, (gm, sc) in zip((dctx, lctx), maps):
        res.append(ctx.score_grid(gm, sc, sg.spe_params(), xs, ys, ts, want_scores=False)[1:])
    same = res[0] == res[1]
    out["k1_grid"] = {"best_idx": int(res[0][0]), "best": res[0][1], "identical": bool(same)}
    P = 5000
    poses = pose + rng.normal(0, [0.2, 0.2, 0.1], (P, 3))
    rl = []
    for ctx, (gm, sc) in zip((dctx, lctx), maps):
        rl.append(ctx.score_poses(gm, sc, sg.spe_params(), poses, want_scores=False)[1:])
    same2 = rl[0] == rl[1]
    out["k1_list"] = {"best_idx": int(rl[0][0]), "identical": bool(same2)}
    for gm, sc in maps:
        gm.close(); sc.close()
    return same and same2


def check_k5(dctx, lctx, out):
    rng = np.random.default_rng(12)
    size, scale = 512, 0.05
    truth = np.array([0.3, 0.2, 0.1])
    r, a = room(rng, 361, 2 * np.pi, 8.0, 6.0, truth)
    got = []
    for ctx in (dctx, lctx):
        gm = sg.GridMap(ctx, size, size, scale, sg.CELL_MEAN, sg.GROW_PLAIN)
        pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
        sc = sg.Scan(ctx, r, a)
        for _ in range(3):
            pyr.append_scan(sc, truth, 1.0, 0, sg.estimator(sg.EST_CONST))
        init = truth + [0.12, -0.08, 0.03]
        params = sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1)
        delta, prob, st = pyr.match_m3rsm(r, a, init, params, 0.5, 0.5, np.deg2rad(5), np.deg2rad(0.5), 0.05)
        got.append((tuple(np.asarray(delta).tolist()), float(prob)))
        pyr.close(); gm.close(); sc.close()
    same = got[0] == got[1]
    out["k5_m3rsm"] = {"delta": got[0][0], "prob": got[0][1], "identical": bool(same)}
    return same


def check_k6(dctx, lctx, out, n=10):
    rng = np.random.default_rng(13)
    size, scale = 128, 0.05
    est = sg.estimator(sg.EST_CONST)
    pd = sg.Particles(dctx, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
    pl = sg.Particles(lctx, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
    truth = np.array([0.2, -0.1, 0.3])
    ok = True
    world = dctx.nranks
    chunk = (n + world - 1) // world
    lo, hi = min(n, chunk * dctx.rank), min(n, chunk * (dctx.rank + 1))

    def maps_equal():
        for i in range(lo, hi):
            a_, b_ = pd.map(i), pl.map(i)
            if a_.info() != b_.info() or not np.array_equal(a_.download(), b_.download(), equal_nan=True):
                return False
        return True

    for k in range(3):
        r, a = room(rng, 181, 2 * np.pi, 5.0, 4.0, truth)
        poses = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 3))
        sd, sl = sg.Scan(dctx, r, a), sg.Scan(lctx, r, a)
        upd = np.ones(n, np.uint8)
        if k == 1:
            upd[::3] = 0
        cd = pd.append_scan(sd, poses, do_update=upd, est=est)
        cl = pl.append_scan(sl, poses, do_update=upd, est=est)
        ok &= bool(np.array_equal(cd, cl))
        sd.close(); sl.close()
        truth = truth + [0.3, 0.2, 0.05]  # walks out of the initial map: every map grows
    ok &= maps_equal()
    out["k6_append"] = {"identical": bool(ok)}
    r, a = room(rng, 181, 2 * np.pi, 5.0, 4.0, truth)
    sd, sl = sg.Scan(dctx, r, a), sg.Scan(lctx, r, a)
    params = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    cand = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 7, 3))
    s_ok = bool(np.array_equal(pd.score(sd, params, cand), pl.score(sl, params, cand), equal_nan=True))
    init = truth + rng.normal(0, [0.04, 0.04, 0.02], (n, 3))
    active = np.ones(n, np.uint8); active[n // 2] = 0
    hd = pd.match_hc(sd, params, init, 6, 0.1, 0.1, active=active)
    hl = pl.match_hc(sl, params, init, 6, 0.1, 0.1, active=active)
    h_ok = all(np.array_equal(x, y, equal_nan=True) for x, y in zip(hd, hl))
    # the estimator's carried cache (gm_cache = 2): per-particle state on the owning rank, two matches in a row
    params_c = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2)
    c_ok = True
    for _ in range(2):
        cd_, cl_ = pd.match_hc(sd, params_c, init, 6, 0.1, 0.1), pl.match_hc(sl, params_c, init, 6, 0.1, 0.1)
        c_ok &= all(np.array_equal(x, y, equal_nan=True) for x, y in zip(cd_, cl_))
    out["k6_hill_climb_carried_cache"] = {"identical": bool(c_ok)}
    h_ok &= c_ok
    out["k6_score"] = {"identical": s_ok}
    out["k6_hill_climb"] = {"identical": bool(h_ok), "tested_min": int(np.min(hd[2][active > 0]))}
    # resampling with sources on the other rank, kept-in-place particles, duplicates and dropped ones
    src = np.array([(i * 7 + 3) % n for i in range(n)], np.int32)
    src[1] = 1; src[2] = n - 1; src[n - 2] = 0; src[n - 1] = 0
    pd.resample(src); pl.resample(src)
    r_ok = maps_equal()
    cd = pd.append_scan(sd, np.tile(truth, (n, 1)), est=est)
    cl = pl.append_scan(sl, np.tile(truth, (n, 1)), est=est)
    r_ok &= bool(np.array_equal(cd, cl)) and maps_equal()
    out["k6_resample"] = {"identical": bool(r_ok), "src": src.tolist()}
    sd.close(); sl.close(); pd.close(); pl.close()
    return ok and s_ok and h_ok and r_ok


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [sg.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    dctx = sg.Context(local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    lctx = sg.Context(local_rank)
    out = {"n_gpus": world}
    ok = check_k1(dctx, lctx, out)
    ok &= check_k5(dctx, lctx, out)
    ok &= check_k6(dctx, lctx, out)
    flags = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["all_ranks_identical"] = bool(flags.item())
    if rank == 0:
        print(json.dumps(out))
    dctx.close(); lctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


if __name__ == "__main__":
    main()
