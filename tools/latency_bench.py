#!/usr/bin/env python
"""Per-call latency of the small-batch paths (BASELINE configs[0] and configs[1] shapes): one Monte-Carlo
batch (100 poses x 360 beams, 100x100 map @0.1) and one hill-climbing round (6 poses x 1081 beams,
800x800 @0.05), scan insertion (const / area) and the world-level loop; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import slam_constructor_b200 as sg  # noqa: E402


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def pyramid_numbers(ctx, size=4096, scale=0.025):
    """BASELINE configs[4] shape: from-scratch max-pyramid over a size x size map (MeanProbabilityCell)"""
    rng = np.random.default_rng(3)
    cells = np.zeros((size, size, 2))
    cells[..., 0] = rng.random((size, size)); cells[..., 1] = 1
    gm = sg.GridMap(ctx, size, size, scale, sg.CELL_MEAN, sg.GROW_PLAIN)
    gm.upload(cells)
    pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
    pyr.build()
    ts = []
    for _ in range(5):
        ctx.sync(); ctx.timer_begin(); pyr.build(); ts.append(ctx.timer_end())
    ms = float(np.median(ts))
    levels = pyr.levels()
    by = size * size * 2 * 8 * (1.0 + 1.0 / 3.0)  # compulsory traffic: the fine records read once, every coarser level written once
    pyr.close(); gm.close()
    # incremental maintenance on a map built from scans: one scan inserted into the fine map and folded up through
    # every level (K2/K3 + K4)
    gm = sg.GridMap(ctx, size, size, scale, sg.CELL_MEAN, sg.GROW_PLAIN)
    pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
    pose = np.array([0.317, -0.223, 0.1])
    r, a = bench.room_ranges(rng, 1081, 1.5 * np.pi, size * scale * 0.35, size * scale * 0.3, pose, 0.01)
    scan = sg.Scan(ctx, r, a)
    est = sg.estimator(sg.EST_CONST)
    for k in range(6):
        cells = pyr.append_scan(scan, pose + [0.01 * k, -0.01 * k, 0.002 * k], 0.9, 0, est, blur=0.3)
    l0 = ctx.launch_count()
    upd = timeit(lambda: pyr.append_scan(scan, pose, 0.9, 0, est, blur=0.3), n=20, warm=3)
    launches = (ctx.launch_count() - l0) / 23
    # BF-M3RSM match on it (13 levels, +-0.5 m, +-5 deg @0.5 deg = 21 rotations)
    params = sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1)
    init = pose + [0.12, -0.08, 0.03]
    _, _, st = pyr.match_m3rsm(r, a, init, params, 0.5, 0.5, np.deg2rad(5), np.deg2rad(0.5), 0.05)
    m3 = timeit(lambda: pyr.match_m3rsm(r, a, init, params, 0.5, 0.5, np.deg2rad(5), np.deg2rad(0.5), 0.05), n=10, warm=2)
    pyr.close(); gm.close(); scan.close()
    return {"grid": [size, size], "levels": levels, "build_ms": round(ms, 3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
            "append_scan_us": round(upd, 1), "append_scan_cells": int(cells), "append_scan_launches": round(launches, 1),
            "m3rsm_match_us": round(m3, 1), "m3rsm_stats": st}


def particle_numbers(ctx, n=256, size=2560, scale=0.05, beams=720):
    """BASELINE configs[3]: 256 particles x 720 beams, each particle with its own 2560 x 2560 map (copy-on-write tiles) in a
    35 x 30 m room; one batched scan insertion, the hill climbing of all particles (one launch without the OOPE cache; with the
    cache carried as upstream does), one resampling"""
    rng = np.random.default_rng(7)
    parts = sg.Particles(ctx, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
    est = sg.estimator(sg.EST_CONST)
    pose = np.array([0.317, -0.223, 0.1])
    r, a = bench.room_ranges(rng, beams, 2 * np.pi, 17.5, 15.0, pose, 0.01)
    scan = sg.Scan(ctx, r, a)
    poses = pose + rng.normal(0, [0.02, 0.02, 0.01], (n, 3))
    cells = parts.append_scan(scan, poses, est=est)
    upd = timeit(lambda: parts.append_scan(scan, poses, est=est), n=10, warm=2)
    params = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    hc = timeit(lambda: parts.match_hc(scan, params, poses), n=10, warm=2)
    params_c = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2)
    hc_c = timeit(lambda: parts.match_hc(scan, params_c, poses), n=5, warm=1)
    _, _, tested = parts.match_hc(scan, params_c, poses)
    st0 = parts.tile_stats()
    # a resampling that keeps a third of the particles (two copies of each), then the insertion that un-shares their tiles
    src = (np.arange(n) // 3 * 3).astype(np.int32)
    t0 = time.perf_counter(); parts.resample(src); res_us = (time.perf_counter() - t0) * 1e6
    st1 = parts.tile_stats()
    t0 = time.perf_counter(); parts.append_scan(scan, poses, est=est); ctx.sync(); first_us = (time.perf_counter() - t0) * 1e6
    st2 = parts.tile_stats()
    parts.close(); scan.close()
    tot = int(np.sum(cells))
    return {"particles": n, "grid": [size, size], "beams": beams, "append_scan_us": round(upd, 1), "cells_per_step": tot,
            "cell_updates_per_s": tot / (upd * 1e-6), "hill_climb_us": round(hc, 1), "hill_climb_carried_cache_us": round(hc_c, 1),
            "hill_climb_poses_tested": int(np.sum(tested)),
            "map_storage": {"kind": "copy-on-write 128x128 tiles" if st0["tiled"] else "dense", "pool_MB": st0["pool_bytes"] / 1e6,
                            "tiles_live": st0["tiles_live"], "dense_equivalent_MB": n * size * size * 5 * 8 / 1e6},
            "resample": {"e2e_us": round(res_us, 1), "device_us": st1["resample_us"], "bytes_copied": st1["resample_bytes"],
                         "tile_references_shared": st1["resample_tiles_shared"],
                         "first_insertion_after_us": round(first_us, 1), "tiles_cloned_by_it": st2["tiles_cloned"] - st1["tiles_cloned"]}}


def measure(ctx):
    rng = np.random.default_rng(1)
    out = {}
    for name, size, scale, beams, fov, P, model, est_kind in (
            ("tiny_mc_batch", 100, 0.1, 360, 2 * np.pi, 100, sg.CELL_MEAN, sg.EST_CONST),
            ("viny_hc_round", 800, 0.05, 1081, 1.5 * np.pi, 6, sg.CELL_TBM_CONSISTENT, sg.EST_AREA)):
        hw, hh = size * scale * 0.35, size * scale * 0.3
        gm = sg.GridMap(ctx, size, size, scale, model, sg.GROW_PLAIN)
        pose = np.array([0.317, -0.223, 0.1])
        r, a = bench.room_ranges(rng, beams, fov, hw, hh, pose, 0.01)
        scan = sg.Scan(ctx, r, a)
        tbm = model == sg.CELL_TBM_CONSISTENT
        est = sg.estimator(est_kind, occ=(0.95, 0.04) if tbm else (0.95, 1.0), empty=(0.01, 0.003) if tbm else (0.01, 1.0), shift=0.01 * scale)
        for _ in range(5):
            cells = ctx.append_scan(gm, scan, pose, 0.9, 0, est, blur=0.3)
        poses = pose + rng.normal(0, [0.1, 0.1, 0.05], (P, 3))
        params = sg.spe_params()
        e2e = timeit(lambda: ctx.score_poses(gm, scan, params, poses, want_scores=True))
        oneshot_kernel_us = ctx.last_kernel_ms() * 1e3  # phase-1 kernel of the one-shot path
        ctx.stage_poses(scan, params, poses)

        def dev():
            ctx.timer_begin(); ctx.score_launch(gm); return ctx.timer_end()
        for _ in range(20):
            dev()
        dev_us = float(np.median([dev() for _ in range(200)])) * 1e3
        shifts = rng.normal(0, [0.2, 0.2, 0.1], (100, 3))  # MC(0.2 / 0.1, 20 failed attempts, 100 poses): one dispersion segment
        mc_us = timeit(lambda: ctx.match_mc(gm, scan, params, pose + [0.06, -0.05, 0.03], shifts, 20, 100), n=100, warm=10)
        mc_out, _ = ctx.match_mc(gm, scan, params, pose + [0.06, -0.05, 0.03], shifts, 20, 100)
        mc_kernel_us = ctx.last_kernel_ms() * 1e3
        hc_us = timeit(lambda: ctx.match_hc(gm, scan, params, pose + [0.06, -0.05, 0.03]), n=100, warm=10)
        _, _, hc_tested, _ = ctx.match_hc(gm, scan, params, pose + [0.06, -0.05, 0.03])
        hc_kernel_us = ctx.last_kernel_ms() * 1e3
        upd = timeit(lambda: ctx.append_scan(gm, scan, pose, 0.9, 0, est, blur=0.3), n=50, warm=5)
        rec_bytes = {sg.CELL_MEAN: 2 * 16 + 8, sg.CELL_TBM_CONSISTENT: 2 * 48 + 8}[model]  # 2 x record + LUT entry
        out[name] = {"poses": P, "beams": beams, "score_call_e2e_us": round(e2e, 1), "score_device_us": round(dev_us, 1), "oneshot_terms_kernel_us": round(oneshot_kernel_us, 1),
                     "evals_per_s_e2e": P * beams / (e2e * 1e-6),
                     "monte_carlo_segment_e2e_us": round(mc_us, 1), "monte_carlo_segment_kernel_us": round(mc_kernel_us, 1),
                     "monte_carlo_poses_tested": int(mc_out["poses_nm"]),
                     "hill_climb_match_e2e_us": round(hc_us, 1), "hill_climb_match_kernel_us": round(hc_kernel_us, 1), "hill_climb_poses_tested": int(hc_tested),
                     "append_scan_e2e_us": round(upd, 1), "cells_per_scan": int(cells),
                     "cell_updates_per_s": cells / (upd * 1e-6),
                     "update_algorithmic_GBps": cells * rec_bytes / (upd * 1e-6) / 1e9}
        gm.close(); scan.close()
    out["pyramid_build"] = pyramid_numbers(ctx)
    out["gmapping_step"] = particle_numbers(ctx)
    return out


def main():
    ctx = sg.Context(0)
    out = measure(ctx)
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
