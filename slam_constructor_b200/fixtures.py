"""The on-disk state triple of slam-constructor: what `lslam2D_bag_runner` dumps
(src/ros/lslam2D_bag_runner.cpp:96-117) and `sm_runner` replays (src/utils/sm_runner.cpp:36-62).

  <name>.pose2D   text  "x y theta"
  <name>.scan2D   text  "n" then n lines "range angle is_occupied"        (LaserScan2D operator<< / >>,
                                                                           src/core/states/sensor_data.h:178-201)
  <name>.map      binary GridMap::save_state of a (Unbounded)PlainGridMap (src/core/maps/plain_grid_map.h:79-129):
                  int32 height, int32 width, f64 scale, int32 origin.x, int32 origin.y, then h*w cells row major,
                  each cell = GridCell::serialize (f64 prob, f64 quality, u8 is_unknown; grid_cell.h:37-47) and,
                  for TBM cells, four more f64 (unknown, empty, occupied, conflict; tbm_grid_cells.h:37-51)

Readers return / writers take the dense record arrays of include/slamgpu.h, so a dumped state goes straight
onto the device.  Cell classes that do not serialise their counters upstream (MeanProbabilityCell::_n,
GmappingBaseCell hits/tries/obstacle) come back with those counters as the reference's load_state leaves them:
n = 0 for known cells is represented here as n = 1 so the cell stays "known"."""
import struct

import numpy as np

from .capi import (CELL_AFFINE, CELL_CREDIBILIST, CELL_GMAPPING, CELL_LWW, CELL_MEAN, CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, STRIDE)

_HEADER = struct.Struct("<iidii")
_BASE = np.dtype([("p", "<f8"), ("q", "<f8"), ("unknown", "u1")])
_TBM = np.dtype([("p", "<f8"), ("q", "<f8"), ("unknown", "u1"), ("u", "<f8"), ("e", "<f8"), ("o", "<f8"), ("c", "<f8")])


def read_pose2d(path):
    return np.array([float(v) for v in open(path).read().split()[:3]])


def write_pose2d(path, pose):
    with open(path, "w") as f:
        f.write("%.17g %.17g %.17g\n" % tuple(pose))


def read_scan2d(path):
    """-> (range[n], angle[n], occupied[n] uint8)"""
    tok = open(path).read().split()
    n = int(tok[0])
    v = np.array(tok[1:1 + 3 * n], dtype=np.float64).reshape(n, 3)
    return v[:, 0].copy(), v[:, 1].copy(), v[:, 2].astype(np.uint8)


def write_scan2d(path, r, a, occ=None):
    occ = np.ones(len(r), np.uint8) if occ is None else occ
    with open(path, "w") as f:
        f.write("%d\n" % len(r))
        for ri, ai, oi in zip(r, a, occ):
            f.write("%.17g %.17g %d\n" % (ri, ai, int(oi)))


def _cell_dtype(model):
    return _TBM if model in (CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, CELL_CREDIBILIST) else _BASE


def read_map(path, model):
    """-> dict(cells[h][w][stride], w, h, scale, ox, oy) for the given cell model"""
    data = open(path, "rb").read()
    h, w, scale, ox, oy = _HEADER.unpack_from(data, 0)
    dt = _cell_dtype(model)
    if len(data) != _HEADER.size + h * w * dt.itemsize:
        raise ValueError("%s: %d bytes do not hold a %dx%d map of this cell class" % (path, len(data), w, h))
    raw = np.frombuffer(data, dtype=dt, count=h * w, offset=_HEADER.size).reshape(h, w)
    known = (raw["unknown"] == 0).astype(np.float64)
    cells = np.zeros((h, w, STRIDE[model]))
    cells[..., 0] = raw["p"]
    if model == CELL_LWW:
        cells[..., 1] = raw["q"]; cells[..., 2] = known
    elif model == CELL_AFFINE:
        cells[..., 1] = known
    elif model == CELL_MEAN:
        cells[..., 1] = known  # _n is not serialised upstream
    elif model in (CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, CELL_CREDIBILIST):
        cells[..., 1] = raw["q"]; cells[..., 2] = raw["u"]; cells[..., 3] = raw["e"]; cells[..., 4] = raw["o"]
        cells[..., 5] = known
    elif model == CELL_GMAPPING:
        cells[..., 4] = known  # hits / tries / obstacle are not serialised upstream
    return dict(cells=cells, w=w, h=h, scale=scale, ox=ox, oy=oy)


def write_map(path, cells, model, scale, ox, oy):
    """the inverse of read_map (the fields the reference serialises)"""
    cells = np.asarray(cells, dtype=np.float64)
    h, w = cells.shape[:2]
    dt = _cell_dtype(model)
    raw = np.zeros((h, w), dtype=dt)
    raw["p"] = cells[..., 0]
    known_col = {CELL_LWW: 2, CELL_AFFINE: 1, CELL_MEAN: 1, CELL_TBM_CONSISTENT: 5, CELL_TBM_UNKNOWN_EVEN: 5, CELL_CREDIBILIST: 5, CELL_GMAPPING: 4}[model]
    raw["unknown"] = (cells[..., known_col] == 0).astype(np.uint8)
    raw["q"] = cells[..., 1] if model in (CELL_LWW, CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN) else 1.0
    if dt is _TBM:
        raw["u"] = cells[..., 2]; raw["e"] = cells[..., 3]; raw["o"] = cells[..., 4]; raw["c"] = 0.0
    with open(path, "wb") as f:
        f.write(_HEADER.pack(h, w, scale, ox, oy))
        f.write(raw.tobytes())


def write_pgm(path, cells):
    """GridMapToPgmDumber::dump_map (src/utils/map_dumpers.h:65-89): binary PGM of the occupancy, top row
    first, intensity = uint8(255 * (1 - occupancy)), occupancy -1 (a never-observed GMapping cell) as 0.5"""
    occ = np.asarray(cells, dtype=np.float64)[..., 0]
    h, w = occ.shape
    occ = np.where(occ == -1, 0.5, np.clip(occ, 0.0, 1.0))
    img = (255 * (1.0 - occ)).astype(np.uint8)[::-1]
    with open(path, "wb") as f:
        f.write(b"P5\n%d\n%d\n255\n" % (w, h))
        f.write(img.tobytes())
