#!/usr/bin/env python
"""bench.py -- pose-beam likelihood evaluations per second of the brute-force scan matcher.

Workload (BASELINE.json configs[2], the one the metric is quoted on): one laser scan of 1081
beams scored under the candidate set of BruteForcePoseEnumerator(+-1 m @0.02, +-0.5 rad @0.01)
= 101 x 101 x 100 = 1 020 100 poses on a 2000 x 2000 grid at 0.05 m, obstacle OOPE, even
weights.  A "step" is one full pass: trig table -> cell-index tables -> scoring kernel ->
arg-max (-> 32-byte all-gather when N > 1).  Synthetic data (numpy, seeded).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); candidates are sharded by contiguous rows.
`value` is device-timed with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI
call (scan + axes H2D, best index/score D2H).  The cpu_baseline / --impl reference legs time
the UNMODIFIED reference (oracle/_ref/libslamref.so) on the host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP_SIZE, MAP_SCALE, N_BEAMS = 2000, 0.05, 1081
BF = dict(x=(-1.0, 1.0, 0.02), y=(-1.0, 1.0, 0.02), t=(-0.5, 0.5, 0.01))
BYTES_PER_EVAL = 32  # one 32-byte sector per map gather
L2_FLUSH = True


# ---------------------------------------------------------------- synthetic inputs (numpy only)
def room_ranges(rng, n, fov, half_w, half_h, pose, noise):
    ang = np.linspace(-fov / 2, fov / 2, n, endpoint=False)
    th = ang + pose[2]
    c, s = np.cos(th), np.sin(th)
    with np.errstate(divide="ignore"):
        tx = np.where(c > 0, (half_w - pose[0]) / c, np.where(c < 0, (-half_w - pose[0]) / c, np.inf))
        ty = np.where(s > 0, (half_h - pose[1]) / s, np.where(s < 0, (-half_h - pose[1]) / s, np.inf))
    return np.minimum(tx, ty) + rng.normal(0, noise, n), ang


def synth_map(rng, size, scale, half_w, half_h):
    """MeanProbabilityCell records {p, n}: a walled room with clutter, free interior, unknown outside"""
    cells = np.zeros((size, size, 2))
    cells[..., 0] = 0.5
    c = (np.arange(size) - size // 2 + 0.5) * scale
    X, Y = np.meshgrid(c, c)
    inside = (np.abs(X) < half_w) & (np.abs(Y) < half_h)
    wall = (np.abs(X) < half_w + 3 * scale) & (np.abs(Y) < half_h + 3 * scale) & ~inside
    n_obs = rng.integers(1, 30, (size, size)).astype(float)
    p_free = np.clip(rng.normal(0.03, 0.02, (size, size)), 0.005, 0.3)
    p_wall = np.clip(rng.normal(0.9, 0.05, (size, size)), 0.5, 0.99)
    cells[..., 0] = np.where(inside, p_free, np.where(wall, p_wall, 0.5))
    cells[..., 1] = np.where(inside | wall, n_obs, 0.0)
    for _ in range(60):  # clutter boxes
        bx, by = rng.uniform(-half_w, half_w), rng.uniform(-half_h, half_h)
        bw, bh = rng.uniform(0.2, 1.5, 2)
        box = (np.abs(X - bx) < bw) & (np.abs(Y - by) < bh) & inside
        cells[..., 0] = np.where(box, p_wall, cells[..., 0])
    return cells


def bf_axis(base, lo, hi, step, inclusive_le):
    """value list of BruteForcePoseEnumerator for one axis: FP accumulation with the reference's
    per-axis stop rule (brute_force_scan_matcher.h:23-54): x/y emit, then step while v < to;
    theta is emitted while t <= to"""
    vals, v = [], lo
    if inclusive_le:
        while v <= hi:
            vals.append(base + v)
            v += step
    else:
        while True:
            vals.append(base + v)
            if not (v < hi):
                break
            v += step
    return np.array(vals)


def make_workload(seed=42, theta_blocks=1):
    rng = np.random.default_rng(seed)
    half_w, half_h = 30.0, 22.0
    cells = synth_map(rng, MAP_SIZE, MAP_SCALE, half_w, half_h)
    true_pose = np.array([1.3, -2.1, 0.4])
    r, a = room_ranges(rng, N_BEAMS, np.deg2rad(270), half_w, half_h, true_pose, 0.01)
    base = true_pose + np.array([0.04, -0.03, 0.01])
    xs = bf_axis(base[0], *BF["x"], False)
    ys = bf_axis(base[1], *BF["y"], False)
    ts = bf_axis(base[2], *BF["t"], True)
    if theta_blocks > 1:  # weak scaling: the theta sweep goes on with the same step, 100 more values per extra GPU
        vals, v = [], BF["t"][0]
        while len(vals) < len(ts) * theta_blocks:
            vals.append(base[2] + v)
            v += BF["t"][2]
        ts = np.array(vals)
    return dict(cells=cells, r=r, a=a, base=base, xs=xs, ys=ys, ts=ts)


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """polls SM clock and throttle reasons of one GPU through NVML from a thread (about 1 kHz), so even a
    timed region of a few tens of milliseconds gets samples; falls back to `nvidia-smi -lms` if NVML is missing"""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index):
        self.rows, self.stop_flag, self.nv, self.p = [], False, None, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                           "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read_smi, daemon=True)
                self.t.start()
            except OSError:
                self.p = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), float(clk), int(rs)))
            except Exception:
                pass
            time.sleep(0.001)

    def _read_smi(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.max_mhz = float(f[1])
                self.rows.append((time.time(), float(f[0]), int(f[2], 16) if f[2].startswith("0x") else 0))
            except (ValueError, IndexError):
                pass

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.p is not None:
            time.sleep(0.05)
            self.p.terminate()
        if self.nv is None and self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        where = "timed region"
        if not inside:
            inside, where = self.rows, "whole run (timed region shorter than the sampling period)"
        mask = 0
        for r in inside:
            mask |= r[2]
        reasons = [n for n, bit in self.REASONS if mask & bit]
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(inside), "window": where}


# ---------------------------------------------------------------- reference (CPU) arm
class ReferenceScorer:
    """the unmodified reference (oracle/_ref, else the oracle port) scoring candidates of the workload
    on the host cores: the checker / baseline, never on the product path"""

    def __init__(self, wl):
        from oracle import binding as ob
        self.ob, self.wl = ob, wl
        self.kind = "reference" if ob.ref is not None else "port"
        self.P = np.stack(np.meshgrid(wl["ts"], wl["ys"], wl["xs"], indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
        self.r, self.a = ob.f64(wl["r"]), ob.f64(wl["a"])
        self.occ = np.ones(N_BEAMS, np.uint8)
        self.params = ob.spe_params()
        if self.kind == "reference":
            self.map = ob.RefMap(MAP_SIZE, MAP_SIZE, MAP_SCALE, ob.CELL_MEAN, ob.GROW_NONE)
            self.map.set_cells(wl["cells"])
        else:
            self.map = ob.OracleMap(MAP_SIZE, MAP_SIZE, MAP_SCALE, ob.CELL_MEAN)
            self.map.set_cells(wl["cells"])
            self.scan = ob.OracleScan(self.r, self.a)

    def sample(self, n_poses, seed=7):
        n_poses = min(n_poses, len(self.P) - 1)
        start = int(np.random.default_rng(seed).integers(0, len(self.P) - n_poses))
        return np.ascontiguousarray(self.P[start:start + n_poses])

    def score(self, poses, threads):
        """returns (seconds, scores, threads used)"""
        ob, wl = self.ob, self.wl
        out = np.empty(len(poses))
        t0 = time.perf_counter()
        if self.kind == "reference":
            ob.ref.ref_score_poses_mt(self.map.h_, N_BEAMS, ob.dptr(self.r), ob.dptr(self.a), ob.u8ptr(self.occ),
                                      ob.SPW_EVEN, self.params, wl["base"][0], wl["base"][1], wl["base"][2],
                                      ob.dptr(poses), len(poses), threads, ob.dptr(out))
        else:
            threads = 1
            out = self.map.score(self.scan, self.params, poses)
        return time.perf_counter() - t0, out, threads


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = make_workload()
    cores = host_cores()
    ref = ReferenceScorer(wl)
    poses = ref.sample(10000 * cores)  # bounded sample: roughly a second of work per step on all host threads
    for _ in range(args.warmup):
        ref.score(poses[:max(200, len(poses) // 10)], cores)
    total, thr = 0.0, cores
    for _ in range(args.steps):
        dt, _, thr = ref.score(poses, cores)
        total += dt
    value = args.steps * len(poses) * N_BEAMS / total
    sample = "%d consecutive candidates x %d beams per step of the same workload" % (len(poses), N_BEAMS)
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg["config"],
            "cpu_baseline": {"value": value, "unit": cfg["unit"], "cores": thr, "kind": ref.kind, "sample": sample},
            "e2e": {"value": value, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--grid-rows", type=int, default=0, help="experiment: rows per thread of the grid kernel (0 = automatic)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the fixed 1 020 100 candidates of configs[2] split N ways; weak = 100 more "
                         "theta values (1 020 100 more candidates) per extra GPU (also reported as a secondary key at N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = {"metric": "pose-beam likelihood evaluations/sec (brute-force scan matcher)", "unit": "evals/s",
           "config": {"workload": "configs[2]: brute-force matcher, +-1 m @0.02 / +-0.5 rad @0.01 = 101x101x100 = 1020100 "
                                  "candidate poses x 1081 beams, 2000x2000 grid @0.05 m, obstacle OOPE, even weights, "
                                  "MeanProbabilityCell map",
                      "candidates": 1020100, "beams": N_BEAMS, "grid": [MAP_SIZE, MAP_SIZE],
                      "sharding": "candidate rows (theta, y) split contiguously over ranks, map replicated; the per-rank arg-max "
                                  "(32 bytes) is exchanged inside the finalize kernel through NVLink peer mailboxes "
                                  "(ncclAllGather when peer mapping is unavailable)",
                      "scaling_mode": "weak: the theta sweep is extended by 100 values (1020100 candidates) per extra GPU"
                                      if args.scaling == "weak" else "strong: the fixed 1020100 candidates of configs[2] split over ranks",
                      "l2": "flushed (256 MB write) before every timed step" if L2_FLUSH else "not flushed"}}
    if args.impl == "reference":
        return run_reference_arm(args, cfg)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N > 1 with torchrun: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")

    import slam_constructor_b200 as sg
    dist = None
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [sg.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = sg.Context(local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    if args.grid_rows:
        ctx.set_option("grid_rows", args.grid_rows)

    wl = make_workload(theta_blocks=world if args.scaling == "weak" else 1)
    cfg["config"]["kernel_selection"] = ("k_score_grid4 (row de-duplication, register loads) when the GPU is full, k_score_grid5 (the same "
                                         "arithmetic behind a cp.async pipeline) when a rank's share leaves it partly empty")
    P = len(wl["xs"]) * len(wl["ys"]) * len(wl["ts"])
    cfg["config"]["candidates"] = P
    gmap = sg.GridMap(ctx, MAP_SIZE, MAP_SIZE, MAP_SCALE, sg.CELL_MEAN)
    gmap.upload(wl["cells"])
    scan = sg.Scan(ctx, wl["r"], wl["a"])
    params = sg.spe_params(sg.OOPE_OBSTACLE, sg.OIE_DISCREPANCY, trig=sg.TRIG_DEVICE)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    # ---- device-timed steps, inputs resident in HBM
    def timed_steps(workload, steps, sample_clocks):
        ctx.stage_grid(scan, params, workload["xs"], workload["ys"], workload["ts"])
        for _ in range(args.warmup):
            ctx.score_launch(gmap)
        _, i0, b0 = ctx.score_fetch()
        smp = sampler if sample_clocks else None
        barrier()
        l0 = ctx.launch_count()
        tw0 = time.time()
        tot, kern = 0.0, 0.0
        for _ in range(steps):
            if L2_FLUSH:
                ctx.flush_l2()
            ctx.timer_begin()
            ctx.score_launch(gmap)
            tot += ctx.timer_end()
            kern += ctx.last_kernel_ms()
        barrier()
        tw1 = time.time()
        n_launch = ctx.launch_count() - l0
        clk = smp.stop(tw0, tw1) if smp else None
        _, i1, b1 = ctx.score_fetch()
        assert (i1, b1) == (i0, b0), "result changed between launches"
        return dict(total_ms=tot, kern_ms=kern, launches=n_launch, clocks=clk, idx=i0, best=b0, stats=ctx.score_stats())

    sampler = ClockSampler(local_rank) if rank == 0 else None  # polls from here on; the timed region is cut out afterwards
    run = timed_steps(wl, args.steps, True)
    total_ms, kern_ms, launches, clocks, idx0, best0, st = (run[k] for k in ("total_ms", "kern_ms", "launches", "clocks", "idx", "best", "stats"))
    if world > 1:
        cfg["config"]["result_exchange"] = "peer memory (fused in k_exchange_finalize)" if st.get("peer_exchange") else "ncclAllGather"

    # ---- N > 1: the merged result against a single-rank recompute of the whole set on this GPU (outside the timed region)
    parity_vs_n1 = None
    if world > 1:
        solo = sg.Context(local_rank)
        smap = sg.GridMap(solo, MAP_SIZE, MAP_SIZE, MAP_SCALE, sg.CELL_MEAN)
        smap.upload(wl["cells"])
        sscan = sg.Scan(solo, wl["r"], wl["a"])
        _, sidx, sbest = solo.score_grid(smap, sscan, params, wl["xs"], wl["ys"], wl["ts"], want_scores=False)
        parity_vs_n1 = (sidx, sbest) == (idx0, best0)
        smap.close(); sscan.close(); solo.close()

    # ---- N > 1, secondary: weak scaling (the theta sweep grows with N)
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_secondary:
        wlw = make_workload(theta_blocks=world)
        rw = timed_steps(wlw, max(3, args.steps // 2), False)
        weak = dict(candidates=len(wlw["xs"]) * len(wlw["ys"]) * len(wlw["ts"]), steps=max(3, args.steps // 2), total_ms=rw["total_ms"])
        ctx.stage_grid(scan, params, wl["xs"], wl["ys"], wl["ts"])

    # ---- end to end through the host-buffer C-ABI call (what a GridScanMatcher adapter calls)
    e2e_steps = args.steps
    for _ in range(2):
        scan.upload(wl["r"], wl["a"])
        ctx.score_grid(gmap, scan, params, wl["xs"], wl["ys"], wl["ts"], want_scores=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scan.upload(wl["r"], wl["a"])
        _, idx2, best2 = ctx.score_grid(gmap, scan, params, wl["xs"], wl["ys"], wl["ts"], want_scores=False)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    assert (idx2, best2) == (idx0, best0)
    # per call: the scan (6 doubles + 1 byte per beam) and the three axes (the y-group / warp tables are re-used while the shape of
    # the candidate set stays the same)
    h2d = N_BEAMS * (6 * 8 + 1) + 8 * (len(wl["xs"]) + len(wl["ys"]) + len(wl["ts"]))
    d2h = 32

    if dist is not None:
        import torch
        t = torch.tensor([total_ms, kern_ms, e2e_s, weak["total_ms"] if weak else 0.0, 0.0 if parity_vs_n1 else 1.0], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kern_ms, e2e_s, weak_ms, parity_bad = (float(v) for v in t.tolist())
        parity_vs_n1 = parity_bad == 0.0
        if weak:
            weak["total_ms"] = weak_ms
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())

    if rank == 0:
        evals = P * N_BEAMS
        value = evals * args.steps / (total_ms * 1e-3)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (OSError, KeyError, ValueError):
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # dominant kernel: k_score_grid; algorithmic bytes = 32 B per pose-beam evaluation of this rank's slice
        k_evals = st["evals"]
        achieved = k_evals * BYTES_PER_EVAL / (kern_ms / args.steps * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("k_score_grid_dram_bytes_per_launch")
            except ValueError:
                traffic = None
        # context for frac > 1: the rate of random 8-byte loads that share nothing, over a table the size of the LUT
        gather_peak = ctx.probe_gather(MAP_SIZE * MAP_SIZE * 8, 1024)
        kernel_gathers = k_evals / (kern_ms / args.steps * 1e-3)
        line = {"metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg["config"],
                "e2e": {"value": evals * e2e_steps / e2e_s, "unit": cfg["unit"], "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / e2e_steps},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "k_score_grid", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "kernel_ms": kern_ms / args.steps,
                             "l2_random_gather": {"loads_per_s": gather_peak, "table_MB": MAP_SIZE * MAP_SIZE * 8 / 1e6,
                                                  "kernel_gathers_per_s": kernel_gathers, "ratio": kernel_gathers / gather_peak,
                                                  "what": "uniformly random 8-byte loads over an L2-resident table of the LUT's "
                                                          "size (slamgpu_probe_gather): gathers that share no sector"},
                             "note": "algorithmic bytes = 32 B (one sector) per pose-beam evaluation (SURVEY 8d); the map is L2/L1 "
                                     "resident and neighbouring candidates share sectors, so frac > 1 is sector reuse, not an "
                                     "efficiency: `binding` holds the counters that say what limits the kernel, `traffic` the "
                                     "DRAM bytes ncu measured for one launch"},
                "result": {"best_idx": idx0, "best_score": best0, "guard_hits": st["guard_hits"]}}
        line["roofline"]["kernel"] = {4: "k_score_grid4", 5: "k_score_grid5", 2: "k_score_grid2"}.get(st["variant"], "k_score_grid")
        bounds_file = os.path.join(ROOT, "profiles", "r02_k1_bounds.json")
        if os.path.exists(bounds_file):
            # what actually binds the kernel (ncu of this round, profiles/): issue slots, L1 wavefronts, measured DRAM traffic
            try:
                bd = json.load(open(bounds_file)).get(line["roofline"]["kernel"])
                if bd:
                    line["roofline"]["traffic"] = bd.get("dram_bytes_per_launch", traffic)
                    line["roofline"]["binding"] = dict(bd)
                    if "floors_ms" in bd and world == 1:
                        # this run's kernel time against the floors the ncu counters give for the same launch: the largest
                        # fraction is how close the kernel is to ANY pipe's limit (the 32 B/evaluation figure above is not one)
                        k_ms = kern_ms / args.steps
                        line["roofline"]["binding"]["frac_of_floor_this_run"] = {k: v / k_ms for k, v in bd["floors_ms"].items()}
            except ValueError:
                pass
        if world > 1:
            line["parity_vs_n1"] = parity_vs_n1
            if weak:
                line["weak_scaling"] = {"value": weak["candidates"] * N_BEAMS * weak["steps"] / (weak["total_ms"] * 1e-3), "unit": cfg["unit"],
                                        "candidates": weak["candidates"], "ms_per_step": weak["total_ms"] / weak["steps"],
                                        "what": "the theta sweep extended by 100 values per extra GPU (round 1's headline mode)"}
        if args.gpus == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            ref = ReferenceScorer(wl)
            ref_poses = ref.sample(30000 * cores)
            best_dt, ref_scores, thr = min((ref.score(ref_poses, cores) for _ in range(3)), key=lambda r: r[0])
            got, _, _ = ctx.score_poses(gmap, scan, sg.spe_params(trig=sg.TRIG_DEVICE), ref_poses)
            # the reference as shipped has no threads: the same code on ONE core, on a smaller sample
            one_dt, _, _ = min((ref.score(ref_poses[:40000], 1) for _ in range(2)), key=lambda r: r[0])
            line["cpu_baseline"] = {"value": len(ref_poses) * N_BEAMS / best_dt, "unit": cfg["unit"], "cores": thr,
                                    "kind": ref.kind,
                                    "single_thread": {"value": 40000 * N_BEAMS / one_dt, "unit": cfg["unit"], "cores": 1,
                                                      "sample": "40000 consecutive candidates, best of 2"},
                                    "sample": "%d consecutive candidates x %d beams of the same workload, best of 3; GPU "
                                              "scores on the sample bit-equal to the reference: %s"
                                              % (len(ref_poses), N_BEAMS, bool(np.array_equal(got, ref_scores)))}
        if args.gpus == 1 and not args.no_secondary:
            # the other rows of the path at the shapes of configs[0], [1] and [4] (per-call latency, scan
            # insertion, pyramid build); a second or two in total
            try:
                from tools import latency_bench
                line["secondary"] = latency_bench.measure(ctx)
            except Exception as e:  # the headline line must survive a secondary failure
                line["secondary"] = {"error": repr(e)}
        print(json.dumps(line))
    gmap.close(); scan.close(); ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
