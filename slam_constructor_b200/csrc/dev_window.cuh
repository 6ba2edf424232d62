// dev_window.cuh -- the device view of a grid map level and the window OOPEs shared by K1 and K5.
#pragma once
#include "dev_math.cuh"

#ifndef SG_LUT_PAD
#define SG_LUT_PAD 1
#endif

namespace {

struct MapView {
  const double *lut;    // padded score LUT (pitch doubles per row, ring of unknown)
  const double *cells;  // dense records (GMapping OOPE gathers these); NULL for a tiled map
  double *const *tiles; // copy-on-write particle maps: tile pointers [th][tw] (128 x 128 cells each)
  int tw;
  int w, h, ox, oy, pitch, stride, model;
  double scale;
  double unknown_lut;
  double unknown_rec[SLAMGPU_MAX_STRIDE];
};

SG_DEV int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
SG_DEV int cell_of(double q_floor) {  // floor(q) -> int without overflow
  q_floor = q_floor < -1e9 ? -1e9 : (q_floor > 1e9 ? 1e9 : q_floor);
  return (int)q_floor;
}

SG_DEV const double *view_cell(const MapView &m, int ix, int iy) {  // record of an internal cell (inside the map)
  if (m.tiles) {
    const double *t = m.tiles[(iy >> 7) * m.tw + (ix >> 7)];
    return t + ((size_t)(iy & 127) * 128 + (ix & 127)) * m.stride;
  }
  return m.cells + ((size_t)iy * m.w + ix) * m.stride;
}

SG_DEV double lut_at(const MapView &m, int cx, int cy) {  // external cell -> impact (unknown outside)
  int ix = clampi(cx + m.ox, -1, m.w) + SG_LUT_PAD, iy = clampi(cy + m.oy, -1, m.h) + SG_LUT_PAD;
  return __ldg(m.lut + (size_t)iy * m.pitch + ix);
}

// ---- window OOPEs: occupancy_observation_probability.h:29-99 over GridRasterizedRectangle
// (src/core/maps/grid_rasterization.h:26-64) and LightWeightRectangle::overlap
// (src/core/geometry_primitives.h:252-318)
struct Lwr { double b, t, l, r; };
SG_DEV bool lwr_contains(const Lwr &a, double x, double y) { return sg::are_ordered(a.l, x, a.r) && sg::are_ordered(a.b, y, a.t); }
SG_DEV double lwr_area(const Lwr &a) { return sg::mul(sg::sub(a.t, a.b), sg::sub(a.r, a.l)); }
SG_DEV Lwr lwr_intersect(const Lwr &a0, const Lwr &b0) {
  for (int pass = 0; pass < 2; ++pass) {
    const Lwr &a = pass ? b0 : a0;
    const Lwr &that = pass ? a0 : b0;
    unsigned nm = 0;
    double cl = a.l, cr = a.r, ct = a.t, cb = a.b;
    if (lwr_contains(a, that.l, that.b)) { ++nm; cl = that.l; cb = that.b; }
    if (lwr_contains(a, that.r, that.b)) { ++nm; cr = that.r; cb = that.b; }
    if (lwr_contains(a, that.l, that.t)) { ++nm; cl = that.l; ct = that.t; }
    if (lwr_contains(a, that.r, that.t)) { ++nm; cr = that.r; ct = that.t; }
    if (nm) return Lwr{cb, ct, cl, cr};
  }
  return Lwr{0, 0, 0, 0};
}
SG_DEV double lwr_overlap(const Lwr &a, const Lwr &b) {
  if (lwr_area(a) != 0) {
    Lwr i = lwr_intersect(a, b);
    return sg::div(lwr_area(i), lwr_area(a));
  }
  if (lwr_area(b) != 0) return lwr_contains(b, a.l, a.b) ? 1.0 : 0.0;
  return (sg::are_equal(a.t, b.t) && sg::are_equal(a.b, b.b) && sg::are_equal(a.l, b.l) && sg::are_equal(a.r, b.r)) ? 1.0 : 0.0;
}

template <int MODE>
SG_DEV double window_probability(const MapView &m, double X, double Y, double win_v, double win_h) {
  const double s = m.scale;
  double half_v = sg::div(win_v, 2.0), half_h = sg::div(win_h, 2.0);
  Lwr win{sg::sub(Y, half_v), sg::add(Y, half_v), sg::sub(X, half_h), sg::add(X, half_h)};
  double area = lwr_area(win);
  int lx, ly, rx, ry;
  if (area == INFINITY) {
    lx = -m.ox; ly = -m.oy; rx = m.w - 1 - m.ox; ry = m.h - 1 - m.oy;
  } else if (area == 0) {
    lx = rx = cell_of(floor(sg::div(win.l, s)));
    ly = ry = cell_of(floor(sg::div(win.b, s)));
  } else {
    lx = cell_of(floor(sg::div(win.l, s))); ly = cell_of(floor(sg::div(win.b, s)));
    rx = cell_of(floor(sg::div(win.r, s))); ry = cell_of(floor(sg::div(win.t, s)));
  }
  double acc = 0, wsum = 0;
  unsigned nm = 0;
  for (int x = lx; x <= rx; ++x)
    for (int y = ly; y <= ry; ++y) {
      double impact = lut_at(m, x, y);
      if (MODE == SLAMGPU_OOPE_MAX) {
        acc = sg::maxd(impact, acc);
      } else if (MODE == SLAMGPU_OOPE_MEAN) {
        acc = sg::add(acc, impact); nm += 1;
      } else {
        Lwr cell;  // world_cell_bounds, regular_squares_grid.h:108-118
        if (s == INFINITY) cell = Lwr{-INFINITY, INFINITY, -INFINITY, INFINITY};
        else cell = Lwr{sg::mul(s, (double)y), sg::mul(s, (double)(y + 1)), sg::mul(s, (double)x), sg::mul(s, (double)(x + 1))};
        double wgt = lwr_overlap(win, cell);
        acc = sg::add(acc, sg::mul(impact, wgt));
        wsum = sg::add(wsum, wgt);
      }
    }
  if (MODE == SLAMGPU_OOPE_MAX) return acc;
  if (MODE == SLAMGPU_OOPE_MEAN) return nm ? sg::div(acc, (double)nm) : 0.5;
  return wsum != 0 ? sg::div(acc, wsum) : 0.5;
}


}  // namespace
