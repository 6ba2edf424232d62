"""ctypes bindings of the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``orc``  -- oracle/_build/liboracle.so, the plain-C restatement (slam_oracle.h)
* ``ref``  -- oracle/_ref/libslamref.so, the unmodified reference compiled in place
              (None when it was never built, e.g. on a clean checkout without
              /root/reference)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libslamref.so")

CELL_LWW, CELL_AFFINE, CELL_MEAN, CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, CELL_GMAPPING, CELL_CREDIBILIST = range(7)
STRIDE = {0: 3, 1: 2, 2: 2, 3: 6, 4: 6, 5: 5, 6: 6}
OIE_DISCREPANCY, OIE_OCCUPANCY = 0, 1
OOPE_OBSTACLE, OOPE_MAX, OOPE_MEAN, OOPE_OVERLAP, OOPE_GMAPPING = range(5)
GROW_NONE, GROW_PLAIN, GROW_TILED = range(3)
EST_CONST, EST_AREA = 0, 1
SPW_EVEN, SPW_VINY, SPW_AHR = range(3)
OMQE_IDLE, OMQE_AHR = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class Scan(C.Structure):
    _fields_ = [("n", C.c_int32), ("cartesian", C.c_int32), ("a", c_dp), ("b", c_dp), ("occ", c_u8p),
                ("factor", c_dp), ("weight", c_dp)]


class SpeParams(C.Structure):
    _fields_ = [("oope", C.c_int32), ("oie", C.c_int32), ("win_v", C.c_double), ("win_h", C.c_double),
                ("prerotated", C.c_int32), ("gm_fullness_th", C.c_double), ("gm_window", C.c_int32)]


class GmCache(C.Structure):
    _fields_ = [("cx", C.c_int32), ("cy", C.c_int32), ("prob", C.c_double)]


class Estimator(C.Structure):
    _fields_ = [("type", C.c_int32), ("occ_p", C.c_double), ("occ_q", C.c_double), ("empty_p", C.c_double),
                ("empty_q", C.c_double), ("low_qual", C.c_double), ("unknown_qual", C.c_double),
                ("shift_amount", C.c_double)]


class MatchResult(C.Structure):
    _fields_ = [("best_prob", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("dth", C.c_double),
                ("poses_tested", C.c_int64)]


def build(ref=True):
    """(re)build the checkers; building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE, "_build/liboracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def dptr(a):
    return a.ctypes.data_as(c_dp)


def iptr(a):
    return a.ctypes.data_as(c_ip)


def u8ptr(a):
    return a.ctypes.data_as(c_u8p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def spe_params(oope=OOPE_OBSTACLE, oie=OIE_DISCREPANCY, win_v=0.0, win_h=0.0, prerotated=0, gm_th=0.1, gm_window=1):
    return SpeParams(oope, oie, win_v, win_h, prerotated, gm_th, gm_window)


def estimator(type=EST_CONST, occ=(0.95, 1.0), empty=(0.01, 1.0), low_qual=0.01, unknown_qual=0.5, shift=-1.0):
    return Estimator(type, occ[0], occ[1], empty[0], empty[1], low_qual, unknown_qual, shift)


class Mt19937(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]


class Normal(C.Structure):
    _fields_ = [("mean", C.c_double), ("stddev", C.c_double), ("saved", C.c_double), ("saved_available", C.c_int)]


def _load_orc():
    if not os.path.exists(ORC_SO):
        build(ref=False)
    L = C.CDLL(ORC_SO)
    L.orc_map_create.restype = C.c_void_p
    L.orc_map_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_dp]
    L.orc_default_unknown.argtypes = [C.c_int, c_dp]
    L.orc_default_unknown.restype = None
    L.orc_map_clone.restype = C.c_void_p
    L.orc_map_clone.argtypes = [C.c_void_p]
    L.orc_map_destroy.argtypes = [C.c_void_p]
    L.orc_map_cells.restype = c_dp
    L.orc_map_cells.argtypes = [C.c_void_p]
    L.orc_map_info.argtypes = [C.c_void_p, c_ip, c_ip, c_dp, c_ip, c_ip, c_ip]
    L.orc_map_at.restype = c_dp
    L.orc_map_at.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_map_ensure_inside.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_map_reset_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
    L.orc_world_to_cell.argtypes = [C.c_double, C.c_double]
    L.orc_are_equal.argtypes = [C.c_double, C.c_double]
    L.orc_raycast.argtypes = [C.c_double] * 5 + [c_ip, C.c_int]
    L.orc_bresenham.argtypes = [C.c_int] * 4 + [c_ip, C.c_int]
    L.orc_rasterize_rect.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int, c_ip]
    L.orc_rect_overlap.restype = C.c_double
    L.orc_rect_overlap.argtypes = [C.c_double] * 8
    L.orc_estimate_occupancy.argtypes = [C.POINTER(Estimator)] + [C.c_double] * 8 + [C.c_int, c_dp]
    L.orc_cell_update.argtypes = [C.c_int, c_dp, C.c_int] + [C.c_double] * 5
    L.orc_cell_impact.restype = C.c_double
    L.orc_cell_impact.argtypes = [C.c_int, C.c_int, c_dp, C.c_double, C.c_double]
    L.orc_build_lut.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
    L.orc_filter_scan.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p, C.c_double, C.c_double, C.c_double,
                                  C.c_uint, C.c_double, c_ip]
    L.orc_point_weights.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.orc_point_probability.restype = C.c_double
    L.orc_point_probability.argtypes = [C.c_void_p, C.POINTER(SpeParams), C.c_double, C.c_double, C.POINTER(GmCache)]
    L.orc_scan_probability.restype = C.c_double
    L.orc_scan_probability.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams), C.c_double, C.c_double,
                                       C.c_double, C.POINTER(GmCache)]
    L.orc_score_poses.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams), c_dp, C.c_int64, c_dp,
                                  C.POINTER(GmCache)]
    L.orc_argbest.restype = C.c_int64
    L.orc_argbest.argtypes = [c_dp, C.c_int64, C.c_double, c_dp]
    L.orc_append_scan.restype = C.c_int64
    L.orc_append_scan.argtypes = [C.c_void_p, C.POINTER(Scan)] + [C.c_double] * 5 + [C.POINTER(Estimator), C.c_double,
                                                                                   C.c_double, C.c_int, c_ip, C.c_int64]
    L.orc_pyramid_create.restype = C.c_void_p
    L.orc_pyramid_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_dp, C.c_int]
    L.orc_pyramid_destroy.argtypes = [C.c_void_p]
    L.orc_pyramid_levels.argtypes = [C.c_void_p]
    L.orc_pyramid_level.restype = C.c_void_p
    L.orc_pyramid_level.argtypes = [C.c_void_p, C.c_int]
    L.orc_pyramid_rescale.argtypes = [C.c_void_p, C.c_double]
    L.orc_pyramid_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5
    L.orc_pyramid_append_scan.restype = C.c_int64
    L.orc_pyramid_append_scan.argtypes = [C.c_void_p, C.POINTER(Scan)] + [C.c_double] * 5 + [C.POINTER(Estimator),
                                                                                           C.c_double, C.c_double, C.c_int]
    L.orc_match_bound.restype = C.c_double
    L.orc_match_bound.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams)] + [C.c_double] * 8
    L.orc_bf_enumerate.restype = C.c_int64
    L.orc_bf_enumerate.argtypes = [C.c_double] * 12 + [c_dp, C.c_int64, c_dp, c_ip, c_dp, c_ip, c_dp, c_ip]
    L.orc_match_list.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams)] + [C.c_double] * 3 + [
        c_dp, C.c_int64, C.POINTER(MatchResult)]
    L.orc_match_hill_climbing.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams)] + [C.c_double] * 3 + [
        C.c_uint, C.c_double, C.c_double, C.POINTER(MatchResult), C.POINTER(GmCache)]
    L.orc_match_monte_carlo.argtypes = [C.c_void_p, C.POINTER(Scan), C.POINTER(SpeParams)] + [C.c_double] * 3 + [
        C.c_uint, C.c_double, C.c_double, C.c_uint, C.c_uint, C.POINTER(MatchResult)]
    L.orc_angle_histogram_values.argtypes = [C.c_int, c_dp, c_dp, C.POINTER(C.c_uint32)]
    L.orc_mt_seed.argtypes = [C.POINTER(Mt19937), C.c_uint32]
    L.orc_normal_sample.restype = C.c_double
    L.orc_normal_sample.argtypes = [C.POINTER(Normal), C.POINTER(Mt19937)]
    return L


def _load_ref():
    if not os.path.exists(REF_SO):
        return None
    L = C.CDLL(REF_SO)
    L.ref_init_area_shift.argtypes = [C.c_double, C.c_double]
    L.ref_map_create.restype = C.c_void_p
    L.ref_map_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_dp, C.c_int]
    L.ref_map_destroy.argtypes = [C.c_void_p]
    L.ref_map_levels.argtypes = [C.c_void_p]
    L.ref_map_info.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip, c_dp, c_ip, c_ip]
    L.ref_map_export.argtypes = [C.c_void_p, C.c_int, c_dp]
    L.ref_map_import.argtypes = [C.c_void_p, c_dp]
    L.ref_map_at.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
    L.ref_map_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5
    L.ref_raycast.argtypes = [C.c_double] * 5 + [c_ip, C.c_int]
    L.ref_bresenham.argtypes = [C.c_int] * 4 + [c_ip, C.c_int]
    L.ref_rasterize_rect.argtypes = [C.c_double, C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int, c_ip, C.c_int]
    L.ref_rect_overlap.restype = C.c_double
    L.ref_rect_overlap.argtypes = [C.c_double] * 8
    L.ref_estimate_occupancy.argtypes = [C.POINTER(Estimator)] + [C.c_double] * 8 + [C.c_int, c_dp]
    L.ref_cell_update.argtypes = [C.c_int, c_dp, C.c_int] + [C.c_double] * 5
    L.ref_cell_impact.restype = C.c_double
    L.ref_cell_impact.argtypes = [C.c_int, C.c_int, c_dp, C.c_double, C.c_double]
    L.ref_filter_scan.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p, C.c_double, C.c_double, C.c_double,
                                  C.c_uint, C.c_double, c_ip]
    L.ref_point_weights.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.ref_score_poses.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p, c_dp, C.c_int, C.c_int, C.c_uint, C.c_double,
                                  C.POINTER(SpeParams), C.c_double, C.c_double, C.c_double, c_dp, C.c_int64, c_dp, c_ip]
    L.ref_bf_enumerate.restype = C.c_int64
    L.ref_bf_enumerate.argtypes = [C.c_double] * 12 + [c_dp, C.c_int64]
    common = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p, C.c_int, C.POINTER(SpeParams), C.c_double, C.c_double, C.c_double]
    L.ref_match_bf.argtypes = common + [C.c_double] * 9 + [C.POINTER(MatchResult), c_dp, C.c_int64]
    L.ref_match_hc.argtypes = common + [C.c_uint, C.c_double, C.c_double, C.POINTER(MatchResult), c_dp, C.c_int64]
    L.ref_match_mc.argtypes = common + [C.c_uint, C.c_double, C.c_double, C.c_uint, C.c_uint, C.POINTER(MatchResult),
                                        c_dp, C.c_int64]
    L.ref_match_bf_m3rsm.argtypes = common + [C.c_double] * 5 + [C.POINTER(MatchResult)]
    L.ref_append_scan.restype = C.c_int64
    L.ref_append_scan.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p] + [C.c_double] * 5 + [
        C.POINTER(Estimator), C.c_double, C.c_double, C.c_int]
    L.ref_pyramid_rescale.argtypes = [C.c_void_p, C.c_double]
    L.ref_match_bound.restype = C.c_double
    L.ref_match_bound.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.POINTER(SpeParams)] + [C.c_double] * 8
    L.ref_gmapping_point_probs.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.ref_normal_samples.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_int, c_dp]
    L.ref_score_poses_mt.restype = C.c_int64
    L.ref_score_poses_mt.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_u8p, C.c_int, C.POINTER(SpeParams), C.c_double,
                                     C.c_double, C.c_double, c_dp, C.c_int64, C.c_int, c_dp]
    return L


orc = _load_orc()
ref = _load_ref()


class OracleScan:
    """keeps the numpy buffers alive behind an orc_scan struct"""

    def __init__(self, a, b, occ=None, weight=None, factor=None, cartesian=False):
        self.a, self.b = f64(a), f64(b)
        n = len(self.a)
        self.occ = np.ascontiguousarray(occ if occ is not None else np.ones(n), dtype=np.uint8)
        self.weight = f64(weight if weight is not None else np.full(n, 1.0 / max(n, 1)))
        self.factor = f64(factor) if factor is not None else None
        self.s = Scan(n, int(cartesian), dptr(self.a), dptr(self.b), u8ptr(self.occ),
                      dptr(self.factor) if self.factor is not None else None, dptr(self.weight))

    @property
    def n(self):
        return self.s.n


class OracleMap:
    def __init__(self, w=None, h=None, scale=None, model=CELL_LWW, grow=GROW_NONE, unknown=None, handle=None, owner=True):
        self.model = model
        self.owner = owner
        if handle is not None:
            self.h_ = handle
        else:
            u = dptr(f64(unknown)) if unknown is not None else None
            self.h_ = orc.orc_map_create(w, h, scale, model, grow, u)

    def __del__(self):
        if getattr(self, "owner", False) and getattr(self, "h_", None):
            orc.orc_map_destroy(self.h_)
            self.h_ = None

    def info(self):
        w, h, ox, oy, st = (C.c_int32() for _ in range(5))
        sc = C.c_double()
        orc.orc_map_info(self.h_, w, h, sc, ox, oy, st)
        return dict(w=w.value, h=h.value, scale=sc.value, ox=ox.value, oy=oy.value, stride=st.value)

    def cells(self):
        """numpy view [h][w][stride] on the oracle's storage (invalid after growth)"""
        i = self.info()
        p = orc.orc_map_cells(self.h_)
        return np.ctypeslib.as_array(p, shape=(i["h"], i["w"], i["stride"]))

    def set_cells(self, arr):
        self.cells()[...] = arr

    def lut(self, oie=OIE_DISCREPANCY):
        i = self.info()
        out = np.empty((i["h"], i["w"]))
        unk = C.c_double()
        orc.orc_build_lut(self.h_, oie, dptr(out), C.byref(unk))
        return out, unk.value

    def score(self, scan, params, poses, cache=None):
        poses = f64(poses).reshape(-1, 3)
        out = np.empty(len(poses))
        orc.orc_score_poses(self.h_, C.byref(scan.s), C.byref(params), dptr(poses), len(poses), dptr(out),
                            C.byref(cache) if cache is not None else None)
        return out

    def append_scan(self, scan, pose, quality=1.0, margin=0.0, est=None, blur=0.0, max_range=np.inf, omqe=OMQE_IDLE,
                    log_cap=0):
        est = est or estimator()
        log = np.zeros((max(log_cap, 1), 2), dtype=np.int32)
        n = orc.orc_append_scan(self.h_, C.byref(scan.s), pose[0], pose[1], pose[2], quality, margin, C.byref(est), blur,
                                max_range, omqe, iptr(log) if log_cap else None, log_cap)
        return n, log[:min(n, log_cap)]


def raycast(fn, bx, by, ex, ey, scale):
    cap = 4096
    buf = np.zeros((cap, 2), dtype=np.int32)
    n = fn(bx, by, ex, ey, scale, iptr(buf), cap)
    if n > cap:
        buf = np.zeros((n, 2), dtype=np.int32)
        n = fn(bx, by, ex, ey, scale, iptr(buf), n)
    return buf[:n].copy()


class RefMap:
    def __init__(self, w, h, scale, model=CELL_LWW, grow=GROW_NONE, unknown=None, pyramid_oie=-1):
        assert ref is not None, "oracle/_ref/libslamref.so is not built"
        self.model = model
        u = dptr(f64(unknown)) if unknown is not None else None
        self.h_ = ref.ref_map_create(w, h, scale, model, grow, u, pyramid_oie)

    def __del__(self):
        if getattr(self, "h_", None):
            ref.ref_map_destroy(self.h_)
            self.h_ = None

    def info(self, level=0):
        w, h, ox, oy = (C.c_int32() for _ in range(4))
        sc = C.c_double()
        ref.ref_map_info(self.h_, level, w, h, sc, ox, oy)
        return dict(w=w.value, h=h.value, scale=sc.value, ox=ox.value, oy=oy.value, stride=STRIDE[self.model])

    def levels(self):
        return ref.ref_map_levels(self.h_)

    def export(self, level=0):
        i = self.info(level)
        out = np.empty((i["h"], i["w"], i["stride"]))
        ref.ref_map_export(self.h_, level, dptr(out))
        return out

    def set_cells(self, arr):
        arr = f64(arr)
        ref.ref_map_import(self.h_, dptr(arr))
