"""ctypes binding of libslamgpu.so -- one Python method per C-ABI entry point.

Fails loudly: a missing library is an ImportError-like RuntimeError on first use, a
missing GPU is SlamGpuError(SLAMGPU_E_NODEVICE) from Context().  Nothing here computes.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CELL_LWW, CELL_AFFINE, CELL_MEAN, CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, CELL_GMAPPING, CELL_CREDIBILIST = range(7)
STRIDE = {0: 3, 1: 2, 2: 2, 3: 6, 4: 6, 5: 5, 6: 6}
OIE_DISCREPANCY, OIE_OCCUPANCY = 0, 1
OOPE_OBSTACLE, OOPE_MAX, OOPE_MEAN, OOPE_OVERLAP, OOPE_GMAPPING = range(5)
GROW_NONE, GROW_PLAIN, GROW_TILED = range(3)
EST_CONST, EST_AREA = 0, 1
TRIG_DEVICE, TRIG_HOST = 0, 1
SPW_EVEN, SPW_VINY, SPW_AHR = range(3)
OMQE_IDLE, OMQE_AHR = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)


class SpeParams(C.Structure):
    _fields_ = [("oope", C.c_int32), ("oie", C.c_int32), ("win_v", C.c_double), ("win_h", C.c_double),
                ("prerotated", C.c_int32), ("trig_mode", C.c_int32), ("gm_fullness_th", C.c_double),
                ("gm_window", C.c_int32), ("gm_cache", C.c_int32)]


class Estimator(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("occ_p", C.c_double), ("occ_q", C.c_double),
                ("empty_p", C.c_double), ("empty_q", C.c_double), ("low_qual", C.c_double), ("unknown_qual", C.c_double),
                ("shift_amount", C.c_double)]


class GmCache(C.Structure):
    """slamgpu_gm_cache: the GMapping OOPE's one-entry cache (cell, probability; prob == -1: empty)"""
    _fields_ = [("cx", C.c_int32), ("cy", C.c_int32), ("prob", C.c_double)]


def spe_params(oope=OOPE_OBSTACLE, oie=OIE_DISCREPANCY, win_v=0.0, win_h=0.0, prerotated=0, trig=TRIG_DEVICE, gm_th=0.1,
               gm_window=1, gm_cache=0):
    return SpeParams(oope, oie, win_v, win_h, prerotated, trig, gm_th, gm_window, gm_cache)


def estimator(type=EST_CONST, occ=(0.95, 1.0), empty=(0.01, 1.0), low_qual=0.01, unknown_qual=0.5, shift=-1.0):
    return Estimator(type, 0, occ[0], occ[1], empty[0], empty[1], low_qual, unknown_qual, shift)


class SlamGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("slamgpu error %d: %s" % (code, msg))
        self.code = code


def point_weights(kind, r, a):
    """ScanPointWeighting of the filtered points (host, libm)"""
    r, a = _f64(r), _f64(a)
    w = np.empty(len(r))
    rc = lib().slamgpu_point_weights(kind, len(r), _dp(r), _dp(a), _dp(w))
    if rc != 0:
        raise SlamGpuError(rc, "slamgpu_point_weights")
    return w


def mapping_quality(kind, r, a):
    """ObservationMappingQualityEstimator of the raw points (host, libm)"""
    r, a = _f64(r), _f64(a)
    q = np.empty(len(r))
    rc = lib().slamgpu_mapping_quality(kind, len(r), _dp(r), _dp(a), _dp(q))
    if rc != 0:
        raise SlamGpuError(rc, "slamgpu_mapping_quality")
    return q


def library_path():
    return os.path.join(HERE, "lib", "libslamgpu.so")


_lib = None

# every symbol include/slamgpu.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "slamgpu_abi_version", "slamgpu_device_count", "slamgpu_ctx_create", "slamgpu_nccl_unique_id", "slamgpu_ctx_create_dist",
    "slamgpu_ctx_destroy", "slamgpu_last_error", "slamgpu_sync", "slamgpu_ctx_set_option", "slamgpu_timer_begin", "slamgpu_timer_end",
    "slamgpu_last_kernel_ms", "slamgpu_launch_count", "slamgpu_flush_l2", "slamgpu_model_stride", "slamgpu_default_unknown", "slamgpu_map_create",
    "slamgpu_map_destroy", "slamgpu_map_info", "slamgpu_map_upload", "slamgpu_map_download", "slamgpu_map_read_cell",
    "slamgpu_map_reset_cell", "slamgpu_map_update_cell", "slamgpu_map_lut_download", "slamgpu_map_upload_lut", "slamgpu_scan_create",
    "slamgpu_scan_destroy", "slamgpu_scan_upload", "slamgpu_scan_filter", "slamgpu_point_weights", "slamgpu_mapping_quality", "slamgpu_score_poses", "slamgpu_score_grid", "slamgpu_score_poses_chained", "slamgpu_match_hc", "slamgpu_match_mc", "slamgpu_debug_div", "slamgpu_debug_hill_climb", "slamgpu_debug_m3rsm", "slamgpu_probe_gather", "slamgpu_stage_poses",
    "slamgpu_stage_grid", "slamgpu_score_launch", "slamgpu_score_fetch", "slamgpu_score_stats", "slamgpu_raycast", "slamgpu_raycast_segments", "slamgpu_estimate_occupancy",
    "slamgpu_append_scan", "slamgpu_append_beams", "slamgpu_pyramid_create", "slamgpu_pyramid_destroy", "slamgpu_pyramid_levels",
    "slamgpu_pyramid_level_info", "slamgpu_pyramid_build", "slamgpu_pyramid_level_download", "slamgpu_pyramid_rescale",
    "slamgpu_pyramid_append_scan", "slamgpu_pyramid_append_beams", "slamgpu_score_windows", "slamgpu_match_m3rsm", "slamgpu_particles_create", "slamgpu_particles_destroy",
    "slamgpu_particles_count", "slamgpu_particles_tile_stats", "slamgpu_particles_map", "slamgpu_particles_score", "slamgpu_particles_match_hc",
    "slamgpu_particles_append_scan", "slamgpu_particles_resample",
]


def lib():
    """the loaded library; raises if it was never built (python -m slam_constructor_b200.build)"""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: build it with `python -m slam_constructor_b200.build` "
                           "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pvp = C.POINTER(C.c_void_p)
    L.slamgpu_last_error.restype = C.c_char_p
    L.slamgpu_last_error.argtypes = [vp]
    L.slamgpu_ctx_create.argtypes = [C.c_int, pvp]
    L.slamgpu_nccl_unique_id.argtypes = [vp]
    L.slamgpu_ctx_create_dist.argtypes = [C.c_int, C.c_int, C.c_int, vp, pvp]
    L.slamgpu_ctx_destroy.argtypes = [vp]
    L.slamgpu_ctx_destroy.restype = None
    L.slamgpu_sync.argtypes = [vp]
    L.slamgpu_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
    L.slamgpu_timer_begin.argtypes = [vp]
    L.slamgpu_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.slamgpu_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.slamgpu_launch_count.restype = i64
    L.slamgpu_launch_count.argtypes = [vp]
    L.slamgpu_flush_l2.argtypes = [vp]
    L.slamgpu_model_stride.argtypes = [C.c_int]
    L.slamgpu_default_unknown.argtypes = [C.c_int, c_dp]
    L.slamgpu_default_unknown.restype = None
    L.slamgpu_map_create.argtypes = [vp, i32, i32, dbl, i32, i32, c_dp, pvp]
    L.slamgpu_map_destroy.argtypes = [vp]
    L.slamgpu_map_destroy.restype = None
    L.slamgpu_map_info.argtypes = [vp, c_ip, c_ip, c_dp, c_ip, c_ip, c_ip]
    L.slamgpu_map_upload.argtypes = [vp, c_dp, i32, i32, i32, i32]
    L.slamgpu_map_download.argtypes = [vp, c_dp]
    L.slamgpu_map_read_cell.argtypes = [vp, i32, i32, c_dp]
    L.slamgpu_map_reset_cell.argtypes = [vp, i32, i32, c_dp]
    L.slamgpu_map_update_cell.argtypes = [vp, i32, i32, i32, dbl, dbl, dbl, dbl, dbl]
    L.slamgpu_map_lut_download.argtypes = [vp, i32, c_dp, c_dp]
    L.slamgpu_map_upload_lut.argtypes = [vp, i32, c_dp, dbl, i32, i32, i32, i32]
    L.slamgpu_scan_create.argtypes = [vp, pvp]
    L.slamgpu_scan_destroy.argtypes = [vp]
    L.slamgpu_scan_destroy.restype = None
    L.slamgpu_scan_upload.argtypes = [vp, i32, i32, c_dp, c_dp, c_u8p, c_dp, c_dp]
    L.slamgpu_scan_filter.argtypes = [vp, i32, c_dp, c_dp, c_u8p, c_dp, C.c_uint32, dbl, c_ip]
    L.slamgpu_point_weights.argtypes = [i32, i32, c_dp, c_dp, c_dp]
    L.slamgpu_mapping_quality.argtypes = [i32, i32, c_dp, c_dp, c_dp]
    sp = C.POINTER(SpeParams)
    L.slamgpu_score_poses.argtypes = [vp, vp, vp, sp, c_dp, i64, dbl, c_dp, c_lp, c_dp]
    L.slamgpu_score_grid.argtypes = [vp, vp, vp, sp, c_dp, i32, c_dp, i32, c_dp, i32, dbl, c_dp, c_lp, c_dp]
    L.slamgpu_score_poses_chained.argtypes = [vp, vp, vp, sp, c_dp, i64, C.POINTER(GmCache), c_dp, C.POINTER(GmCache)]
    L.slamgpu_probe_gather.argtypes = [vp, i64, i32, c_dp]
    L.slamgpu_match_mc.argtypes = [vp, vp, vp, sp, c_dp, dbl, i32, c_dp, i32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_dp, c_dp, i32, c_ip]
    L.slamgpu_debug_div.argtypes = [vp, i32, c_dp, c_dp, c_dp]
    L.slamgpu_match_hc.argtypes = [vp, vp, vp, sp, c_dp, C.c_uint32, dbl, dbl, c_dp, c_dp, c_lp, c_dp, i32, c_ip, C.POINTER(GmCache)]
    L.slamgpu_stage_poses.argtypes = [vp, vp, sp, c_dp, i64]
    L.slamgpu_stage_grid.argtypes = [vp, vp, sp, c_dp, i32, c_dp, i32, c_dp, i32]
    L.slamgpu_score_launch.argtypes = [vp, vp, dbl]
    L.slamgpu_score_fetch.argtypes = [vp, c_dp, c_lp, c_dp]
    L.slamgpu_score_stats.argtypes = [vp, c_lp]
    L.slamgpu_raycast.argtypes = [vp, vp, vp, c_dp, c_lp, c_ip, i64, c_lp]
    ep = C.POINTER(Estimator)
    L.slamgpu_raycast_segments.argtypes = [vp, dbl, c_dp, i32, c_lp, c_ip, i64, c_lp]
    L.slamgpu_estimate_occupancy.argtypes = [vp, ep, i32, c_dp, c_dp, c_u8p, c_dp]
    L.slamgpu_append_scan.argtypes = [vp, vp, vp, c_dp, dbl, i32, ep, dbl, dbl, c_dp, c_lp]
    L.slamgpu_append_beams.argtypes = [vp, vp, i32, c_dp, c_u8p, c_dp, ep, dbl, dbl, c_lp]
    L.slamgpu_pyramid_create.argtypes = [vp, vp, i32, pvp]
    L.slamgpu_pyramid_destroy.argtypes = [vp]
    L.slamgpu_pyramid_destroy.restype = None
    L.slamgpu_pyramid_levels.argtypes = [vp]
    L.slamgpu_pyramid_level_info.argtypes = [vp, i32, c_ip, c_ip, c_dp, c_ip, c_ip]
    L.slamgpu_pyramid_build.argtypes = [vp]
    L.slamgpu_pyramid_level_download.argtypes = [vp, i32, c_dp, c_dp]
    L.slamgpu_pyramid_rescale.argtypes = [vp, dbl]
    L.slamgpu_pyramid_append_scan.argtypes = [vp, vp, c_dp, dbl, i32, ep, dbl, dbl, c_dp, c_lp]
    L.slamgpu_pyramid_append_beams.argtypes = [vp, i32, c_dp, c_u8p, c_dp, ep, dbl, dbl, c_lp]
    L.slamgpu_score_windows.argtypes = [vp, pvp, i32, c_ip, c_dp, i64, c_dp, sp, c_dp]
    L.slamgpu_match_m3rsm.argtypes = [vp, i32, c_dp, c_dp, c_dp, c_dp, sp, dbl, dbl, dbl, dbl, dbl, dbl, c_dp, c_dp, c_lp]
    L.slamgpu_particles_create.argtypes = [vp, i32, i32, i32, dbl, i32, i32, c_dp, pvp]
    L.slamgpu_particles_destroy.argtypes = [vp]
    L.slamgpu_particles_destroy.restype = None
    L.slamgpu_particles_count.argtypes = [vp]
    L.slamgpu_particles_tile_stats.argtypes = [vp, c_lp]
    L.slamgpu_particles_map.argtypes = [vp, i32]
    L.slamgpu_particles_map.restype = vp
    L.slamgpu_particles_score.argtypes = [vp, vp, sp, c_dp, i32, c_dp]
    L.slamgpu_particles_match_hc.argtypes = [vp, vp, sp, c_dp, c_u8p, C.c_uint32, dbl, dbl, c_dp, c_dp, c_lp]
    L.slamgpu_particles_append_scan.argtypes = [vp, vp, c_dp, c_u8p, dbl, i32, ep, dbl, dbl, c_dp, c_lp]
    L.slamgpu_particles_resample.argtypes = [vp, c_ip]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


class Context:
    """slamgpu_ctx: one per world / per GPU rank"""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None):
        self.L = lib()
        h = C.c_void_p()
        if nranks > 1:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            r = self.L.slamgpu_ctx_create_dist(device, rank, nranks, buf, C.byref(h))
        else:
            r = self.L.slamgpu_ctx_create(device, C.byref(h))
        if r != 0:
            raise SlamGpuError(r, self.L.slamgpu_last_error(None).decode())
        self.h = h
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        r = lib().slamgpu_nccl_unique_id(buf)
        if r != 0:
            raise SlamGpuError(r, lib().slamgpu_last_error(None).decode())
        return bytes(buf.raw)

    def check(self, r):
        if r != 0:
            raise SlamGpuError(r, self.L.slamgpu_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.slamgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        pass  # explicit close(); maps/scans hold the ctx alive by reference

    def set_option(self, name, value):
        self.check(self.L.slamgpu_ctx_set_option(self.h, name.encode(), int(value)))

    def sync(self):
        self.check(self.L.slamgpu_sync(self.h))

    def timer_begin(self):
        self.check(self.L.slamgpu_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float()
        self.check(self.L.slamgpu_timer_end(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self.check(self.L.slamgpu_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.L.slamgpu_launch_count(self.h)

    def flush_l2(self):
        self.check(self.L.slamgpu_flush_l2(self.h))

    # ---- K1 ----
    def score_poses(self, gmap, scan, params, poses, init_score=-np.inf, want_scores=True):
        poses = _f64(poses).reshape(-1, 3)
        P = len(poses)
        out = np.full(P, np.nan) if want_scores else None
        idx, best = C.c_int64(), C.c_double()
        self.check(self.L.slamgpu_score_poses(self.h, gmap.h, scan.h, C.byref(params), _dp(poses), P, init_score, _dp(out),
                                              C.byref(idx), C.byref(best)))
        return out, idx.value, best.value

    def score_grid(self, gmap, scan, params, xs, ys, ts, init_score=-np.inf, want_scores=True):
        xs, ys, ts = _f64(xs), _f64(ys), _f64(ts)
        P = len(xs) * len(ys) * len(ts)
        out = np.full(P, np.nan) if want_scores else None
        idx, best = C.c_int64(), C.c_double()
        self.check(self.L.slamgpu_score_grid(self.h, gmap.h, scan.h, C.byref(params), _dp(xs), len(xs), _dp(ys), len(ys),
                                             _dp(ts), len(ts), init_score, _dp(out), C.byref(idx), C.byref(best)))
        return out, idx.value, best.value

    def score_poses_chained(self, gmap, scan, params, poses, state=None):
        """GMapping OOPE with its cache carried from pose to pose; returns (scores, states after each pose)"""
        poses = _f64(poses).reshape(-1, 3)
        P = len(poses)
        out = np.full(P, np.nan)
        st_in = state if state is not None else GmCache(0, 0, -1.0)
        states = (GmCache * max(P, 1))()
        self.check(self.L.slamgpu_score_poses_chained(self.h, gmap.h, scan.h, C.byref(params), _dp(poses), P, C.byref(st_in), _dp(out),
                                                      states))
        return out, [GmCache(s.cx, s.cy, s.prob) for s in states[:P]]

    def probe_gather(self, table_bytes=32 << 20, loads_per_thread=512):
        out = C.c_double()
        self.check(self.L.slamgpu_probe_gather(self.h, table_bytes, loads_per_thread, C.byref(out)))
        return out.value

    def match_mc(self, gmap, scan, params, best_pose, noise, max_failed, max_poses, best_prob=None, failed=0, poses_nm=0, log_cap=0):
        """a segment of MonteCarloScanMatcher::process_scan over the given pose shifts; returns (out dict, log) or None when the
        device kernel declines the request"""
        best, noise = _f64(best_pose), _f64(noise).reshape(-1, 3)
        out, served = np.zeros(10), C.c_int32()
        log = np.zeros((max(log_cap, 1), 4))
        self.check(self.L.slamgpu_match_mc(self.h, gmap.h, scan.h, C.byref(params), _dp(best), best_prob if best_prob is not None else np.nan,
                                           0 if best_prob is None else 1, _dp(noise), len(noise), failed, poses_nm, max_failed, max_poses,
                                           _dp(out), _dp(log) if log_cap else None, log_cap, C.byref(served)))
        if not served.value:
            return None
        keys = ("x", "y", "theta", "prob", "consumed", "failed", "poses_nm", "reset", "guard", "logged")
        return dict(zip(keys, out.tolist())), log[:min(int(out[9]), log_cap)]

    def debug_div(self, a, b):
        a, b = _f64(a), _f64(b)
        out = np.empty(len(a))
        self.check(self.L.slamgpu_debug_div(self.h, len(a), _dp(a), _dp(b), _dp(out)))
        return out

    def match_hc(self, gmap, scan, params, init_pose, max_failed_rounds=6, tr=0.1, rot=0.1, log_cap=0, gm_state=None):
        """HillClimbingScanMatcher::process_scan as one call; returns (pose, prob, tested, log or None)"""
        init, out = _f64(init_pose), np.zeros(3)
        prob, tested, count = C.c_double(), C.c_int64(), C.c_int32()
        log = np.zeros((max(log_cap, 1), 4))
        self.check(self.L.slamgpu_match_hc(self.h, gmap.h, scan.h, C.byref(params), _dp(init), max_failed_rounds, tr, rot, _dp(out),
                                           C.byref(prob), C.byref(tested), _dp(log) if log_cap else None, log_cap, C.byref(count),
                                           C.byref(gm_state) if gm_state is not None else None))
        return out, prob.value, tested.value, (log[:min(count.value, log_cap)] if log_cap and count.value >= 0 else None)

    def stage_poses(self, scan, params, poses):
        poses = _f64(poses).reshape(-1, 3)
        self.check(self.L.slamgpu_stage_poses(self.h, scan.h, C.byref(params), _dp(poses), len(poses)))
        self._staged_P = len(poses)

    def stage_grid(self, scan, params, xs, ys, ts):
        xs, ys, ts = _f64(xs), _f64(ys), _f64(ts)
        self.check(self.L.slamgpu_stage_grid(self.h, scan.h, C.byref(params), _dp(xs), len(xs), _dp(ys), len(ys), _dp(ts),
                                             len(ts)))
        self._staged_P = len(xs) * len(ys) * len(ts)

    def score_launch(self, gmap, init_score=-np.inf):
        self.check(self.L.slamgpu_score_launch(self.h, gmap.h, init_score))

    def score_fetch(self, want_scores=False):
        out = np.full(self._staged_P, np.nan) if want_scores else None
        idx, best = C.c_int64(), C.c_double()
        self.check(self.L.slamgpu_score_fetch(self.h, _dp(out), C.byref(idx), C.byref(best)))
        return out, idx.value, best.value

    def score_stats(self):
        st = np.zeros(8, dtype=np.int64)
        self.check(self.L.slamgpu_score_stats(self.h, st.ctypes.data_as(c_lp)))
        return dict(guard_hits=int(st[0]), variant=int(st[1]), evals=int(st[2]), slice_begin=int(st[3]), slice_len=int(st[4]), rows_per_thread=int(st[5]), peer_exchange=int(st[6]))

    # ---- K2 / K3 ----
    def raycast(self, gmap, scan, pose, want_cells=True):
        pose = _f64(pose)
        offs = np.zeros(scan.n + 1, dtype=np.int64)
        total = C.c_int64()
        self.check(self.L.slamgpu_raycast(self.h, gmap.h, scan.h, _dp(pose), offs.ctypes.data_as(c_lp), None, 0,
                                          C.byref(total)))
        if not want_cells:
            return offs, None
        cells = np.zeros((max(total.value, 1), 2), dtype=np.int32)
        self.check(self.L.slamgpu_raycast(self.h, gmap.h, scan.h, _dp(pose), offs.ctypes.data_as(c_lp),
                                          cells.ctypes.data_as(c_ip), total.value, C.byref(total)))
        return offs, cells[:total.value]

    def raycast_segments(self, scale, segments):
        seg = _f64(segments).reshape(-1, 4)
        n = len(seg)
        offs = np.zeros(n + 1, dtype=np.int64)
        total = C.c_int64()
        self.check(self.L.slamgpu_raycast_segments(self.h, scale, _dp(seg), n, offs.ctypes.data_as(c_lp), None, 0,
                                                   C.byref(total)))
        cells = np.zeros((max(total.value, 1), 2), dtype=np.int32)
        self.check(self.L.slamgpu_raycast_segments(self.h, scale, _dp(seg), n, offs.ctypes.data_as(c_lp),
                                                   cells.ctypes.data_as(c_ip), total.value, C.byref(total)))
        return offs, cells[:total.value]

    def estimate_occupancy(self, est, beams, bounds, is_occ):
        beams, bounds = _f64(beams).reshape(-1, 4), _f64(bounds).reshape(-1, 4)
        occ = np.ascontiguousarray(is_occ, dtype=np.uint8)
        out = np.zeros((len(beams), 2))
        self.check(self.L.slamgpu_estimate_occupancy(self.h, C.byref(est), len(beams), _dp(beams), _dp(bounds),
                                                     occ.ctypes.data_as(c_u8p), _dp(out)))
        return out

    def append_scan(self, gmap, scan, pose, quality=1.0, margin=0, est=None, blur=0.0, max_range=np.inf,
                    point_quality=None):
        est = est or estimator()
        pose = _f64(pose)
        pq = _f64(point_quality) if point_quality is not None else None
        n = C.c_int64()
        self.check(self.L.slamgpu_append_scan(self.h, gmap.h, scan.h, _dp(pose), quality, margin, C.byref(est), blur,
                                              max_range, _dp(pq), C.byref(n)))
        return n.value


class GridMap:
    """slamgpu_map: the device-resident dense grid map"""

    def __init__(self, ctx, w, h, scale, model=CELL_LWW, grow=GROW_NONE, unknown=None):
        self.ctx, self.model = ctx, model
        u = _f64(unknown) if unknown is not None else None
        hnd = C.c_void_p()
        ctx.check(ctx.L.slamgpu_map_create(ctx.h, w, h, scale, model, grow, _dp(u), C.byref(hnd)))
        self.h = hnd

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.slamgpu_map_destroy(self.h)
            self.h = None

    def info(self):
        w, h, ox, oy, st = (C.c_int32() for _ in range(5))
        sc = C.c_double()
        self.ctx.check(self.ctx.L.slamgpu_map_info(self.h, w, h, sc, ox, oy, st))
        return dict(w=w.value, h=h.value, scale=sc.value, ox=ox.value, oy=oy.value, stride=st.value)

    def upload(self, cells, ox=None, oy=None):
        cells = _f64(cells)
        h, w = cells.shape[:2]
        ox = w // 2 if ox is None else ox
        oy = h // 2 if oy is None else oy
        self.ctx.check(self.ctx.L.slamgpu_map_upload(self.h, _dp(cells), w, h, ox, oy))

    def download(self):
        i = self.info()
        out = np.empty((i["h"], i["w"], i["stride"]))
        self.ctx.check(self.ctx.L.slamgpu_map_download(self.h, _dp(out)))
        return out

    def lut(self, oie=OIE_DISCREPANCY):
        i = self.info()
        out = np.empty((i["h"], i["w"]))
        unk = C.c_double()
        self.ctx.check(self.ctx.L.slamgpu_map_lut_download(self.h, oie, _dp(out), C.byref(unk)))
        return out, unk.value

    def upload_lut(self, lut, unknown_value, oie=OIE_DISCREPANCY, ox=None, oy=None):
        lut = _f64(lut)
        h, w = lut.shape
        ox = w // 2 if ox is None else ox
        oy = h // 2 if oy is None else oy
        self.ctx.check(self.ctx.L.slamgpu_map_upload_lut(self.h, oie, _dp(lut), unknown_value, w, h, ox, oy))

    def filter_scan(self, r, a, pose, occ=None, skip_rate=0, max_range=-1.0):
        """indices of the points WeightedMeanPointProbabilitySPE::filter_scan keeps"""
        r, a, pose = _f64(r), _f64(a), _f64(pose)
        o = np.ascontiguousarray(occ, dtype=np.uint8) if occ is not None else None
        keep = np.zeros(len(r), dtype=np.int32)
        k = self.ctx.L.slamgpu_scan_filter(self.h, len(r), _dp(r), _dp(a), o.ctypes.data_as(c_u8p) if o is not None else None,
                                           _dp(pose), skip_rate, max_range, keep.ctypes.data_as(c_ip))
        if k < 0:
            self.ctx.check(k)
        return keep[:k]

    def read_cell(self, x, y):
        rec = np.zeros(8)
        self.ctx.check(self.ctx.L.slamgpu_map_read_cell(self.h, x, y, _dp(rec)))
        return rec[:STRIDE[self.model]]

    def reset_cell(self, x, y, rec):
        r = np.zeros(8)
        r[:len(rec)] = rec
        self.ctx.check(self.ctx.L.slamgpu_map_reset_cell(self.h, x, y, _dp(r)))

    def update_cell(self, x, y, is_occ, p, q, obst=(0.0, 0.0), quality=1.0):
        self.ctx.check(self.ctx.L.slamgpu_map_update_cell(self.h, x, y, int(is_occ), p, q, obst[0], obst[1], quality))


class Scan:
    """slamgpu_scan: filtered scan points + pose-independent weights on the device"""

    def __init__(self, ctx, a=None, b=None, occ=None, factor=None, weight=None, cartesian=False):
        self.ctx = ctx
        hnd = C.c_void_p()
        ctx.check(ctx.L.slamgpu_scan_create(ctx.h, C.byref(hnd)))
        self.h = hnd
        self.n = 0
        if a is not None:
            self.upload(a, b, occ, factor, weight, cartesian)

    def upload(self, a, b, occ=None, factor=None, weight=None, cartesian=False):
        a, b = _f64(a), _f64(b)
        n = len(a)
        occ = np.ascontiguousarray(occ, dtype=np.uint8) if occ is not None else None
        factor = _f64(factor) if factor is not None else None
        weight = _f64(weight) if weight is not None else None
        self.ctx.check(self.ctx.L.slamgpu_scan_upload(self.h, n, int(cartesian), _dp(a), _dp(b),
                                                      occ.ctypes.data_as(c_u8p) if occ is not None else None, _dp(factor),
                                                      _dp(weight)))
        self.n = n

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.slamgpu_scan_destroy(self.h)
            self.h = None


class Pyramid:
    """slamgpu_pyramid: the max-pyramid over a fine GridMap (level 0)"""

    def __init__(self, ctx, fine, oie=OIE_DISCREPANCY):
        self.ctx, self.fine, self.oie = ctx, fine, oie
        hnd = C.c_void_p()
        ctx.check(ctx.L.slamgpu_pyramid_create(ctx.h, fine.h, oie, C.byref(hnd)))
        self.h = hnd

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.slamgpu_pyramid_destroy(self.h)
            self.h = None

    def levels(self):
        n = self.ctx.L.slamgpu_pyramid_levels(self.h)
        if n < 0:
            self.ctx.check(n)
        return n

    def level_info(self, level):
        w, h, ox, oy = (C.c_int32() for _ in range(4))
        sc = C.c_double()
        self.ctx.check(self.ctx.L.slamgpu_pyramid_level_info(self.h, level, w, h, sc, ox, oy))
        return dict(w=w.value, h=h.value, scale=sc.value, ox=ox.value, oy=oy.value, stride=STRIDE[self.fine.model])

    def level(self, level, want_impact=False):
        i = self.level_info(level)
        cells = np.empty((i["h"], i["w"], i["stride"]))
        imp = np.empty((i["h"], i["w"])) if want_impact else None
        self.ctx.check(self.ctx.L.slamgpu_pyramid_level_download(self.h, level, _dp(cells), _dp(imp)))
        return (cells, imp) if want_impact else cells

    def build(self):
        self.ctx.check(self.ctx.L.slamgpu_pyramid_build(self.h))

    def rescale(self, target):
        r = self.ctx.L.slamgpu_pyramid_rescale(self.h, target)
        if r < 0:
            self.ctx.check(r)
        return r

    def append_scan(self, scan, pose, quality=1.0, margin=0, est=None, blur=0.0, max_range=np.inf, point_quality=None):
        est = est or estimator()
        pose = _f64(pose)
        pq = _f64(point_quality) if point_quality is not None else None
        n = C.c_int64()
        self.ctx.check(self.ctx.L.slamgpu_pyramid_append_scan(self.h, scan.h, _dp(pose), quality, margin, C.byref(est), blur,
                                                              max_range, _dp(pq), C.byref(n)))
        return n.value

    def match_m3rsm(self, r, a, pose, params, x_limit=1.0, y_limit=1.0, rot_limit=np.deg2rad(5), ang_step=np.deg2rad(0.1),
                    transl_step=0.05, max_finest_prob_diff=0.0, weight=None):
        r, a, pose = _f64(r), _f64(a), _f64(pose)
        w = _f64(weight) if weight is not None else None
        delta, prob, st = np.zeros(3), C.c_double(), np.zeros(4, dtype=np.int64)
        self.ctx.check(self.ctx.L.slamgpu_match_m3rsm(self.h, len(r), _dp(r), _dp(a), _dp(w), _dp(pose), C.byref(params), x_limit,
                                                      y_limit, rot_limit, ang_step, transl_step, max_finest_prob_diff, _dp(delta),
                                                      C.byref(prob), st.ctypes.data_as(c_lp)))
        return delta, prob.value, dict(scored=int(st[0]), calls=int(st[1]), branches=int(st[2]), rotations=int(st[3]))

    def score_windows(self, scans, scan_id, windows, pose, params):
        arr = (C.c_void_p * len(scans))(*[s.h for s in scans])
        sid = np.ascontiguousarray(scan_id, dtype=np.int32)
        win = _f64(windows).reshape(-1, 4)
        pose = _f64(pose)
        out = np.full(len(win), np.nan)
        self.ctx.check(self.ctx.L.slamgpu_score_windows(self.h, arr, len(scans), sid.ctypes.data_as(c_ip), _dp(win), len(win),
                                                        _dp(pose), C.byref(params), _dp(out)))
        return out


class _BorrowedMap(GridMap):
    """a particle's map: same methods as GridMap, owned by the Particles object"""

    def __init__(self, ctx, handle, model):
        self.ctx, self.model, self.h = ctx, model, C.c_void_p(handle)

    def close(self):
        self.h = None


class Particles:
    """slamgpu_particles: n GMapping particles, each with its own device map"""

    def __init__(self, ctx, n, w, h, scale, model=CELL_GMAPPING, grow=GROW_TILED):
        self.ctx, self.n, self.model = ctx, n, model
        hnd = C.c_void_p()
        ctx.check(ctx.L.slamgpu_particles_create(ctx.h, n, w, h, scale, model, grow, None, C.byref(hnd)))
        self.h = hnd

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.slamgpu_particles_destroy(self.h)
            self.h = None

    def map(self, i):
        return _BorrowedMap(self.ctx, self.ctx.L.slamgpu_particles_map(self.h, i), self.model)

    def tile_stats(self):
        st = np.zeros(8, dtype=np.int64)
        self.ctx.check(self.ctx.L.slamgpu_particles_tile_stats(self.h, st.ctypes.data_as(c_lp)))
        keys = ("tiled", "tiles_live", "tiles_cloned", "tile_bytes", "pool_bytes", "resample_bytes", "resample_tiles_shared", "resample_us")
        return dict(zip(keys, (int(v) for v in st)))

    def score(self, scan, params, poses):
        poses = _f64(poses).reshape(self.n, -1, 3)
        c = poses.shape[1]
        out = np.full((self.n, c), np.nan)
        self.ctx.check(self.ctx.L.slamgpu_particles_score(self.h, scan.h, C.byref(params), _dp(poses), c, _dp(out)))
        return out

    def match_hc(self, scan, params, init_poses, max_failed_rounds=6, tr=0.1, rot=0.1, active=None):
        init = _f64(init_poses).reshape(self.n, 3)
        act = np.ascontiguousarray(active, dtype=np.uint8) if active is not None else None
        poses, probs, tested = np.zeros((self.n, 3)), np.zeros(self.n), np.zeros(self.n, dtype=np.int64)
        self.ctx.check(self.ctx.L.slamgpu_particles_match_hc(self.h, scan.h, C.byref(params), _dp(init),
                                                             act.ctypes.data_as(c_u8p) if act is not None else None,
                                                             max_failed_rounds, tr, rot, _dp(poses), _dp(probs),
                                                             tested.ctypes.data_as(c_lp)))
        return poses, probs, tested

    def append_scan(self, scan, poses, do_update=None, quality=1.0, margin=0, est=None, blur=0.0, max_range=np.inf):
        est = est or estimator()
        poses = _f64(poses).reshape(self.n, 3)
        upd = np.ascontiguousarray(do_update, dtype=np.uint8) if do_update is not None else None
        cells = np.zeros(self.n, dtype=np.int64)
        self.ctx.check(self.ctx.L.slamgpu_particles_append_scan(self.h, scan.h, _dp(poses),
                                                                upd.ctypes.data_as(c_u8p) if upd is not None else None, quality,
                                                                margin, C.byref(est), blur, max_range, None,
                                                                cells.ctypes.data_as(c_lp)))
        return cells

    def resample(self, src):
        src = np.ascontiguousarray(src, dtype=np.int32)
        self.ctx.check(self.ctx.L.slamgpu_particles_resample(self.h, src.ctypes.data_as(c_ip)))
