// dev_geometry.cuh -- segment ray casting and the const / area cell occupancy estimators.
//
// Device restatement of RegularSquaresGrid::world_to_cells
// (src/core/maps/regular_squares_grid.h:56-101), DiscreteSegment2D
// (src/core/geometry_discrete_primitives.h:55-104), ConstOccupancyEstimator
// (src/core/maps/const_occupancy_estimator.h:11-15) and AreaOccupancyEstimator
// (src/core/maps/area_occupancy_estimator.h:27-240 over Segment2D / Ray / Rectangle of
// src/core/geometry_primitives.h).  IEEE double, reference operation order, no FMA.
#pragma once
#include "dev_math.cuh"

namespace sg {

// ---------------------------------------------------------------- ray casting
// A resumable walker: next() yields the cells of the segment in the reference's order.
struct RayWalker {
  int px, py, sx, sy, endx, endy, inc_x, inc_y;
  int cells_nm, n;  // (a ray spans at most 2^26 cells per axis: the callers refuse longer ones)
  double e, e_x_inc, e_y_inc;
  bool done;

  SG_DEV void init(double bx, double by, double ex, double ey, double scale) {
    double d_x = sub(ex, bx), d_y = sub(ey, by);
    inc_x = 0 < d_x ? 1 : -1; inc_y = 0 < d_y ? 1 : -1;
    px = sx = world_to_cell(bx, scale); py = sy = world_to_cell(by, scale);
    endx = world_to_cell(ex, scale); endy = world_to_cell(ey, scale);
    long long ax = (long long)endx - px, ay = (long long)endy - py;
    const long long span = (ax < 0 ? -ax : ax) + (ay < 0 ? -ay : ay) + 1;
    cells_nm = span > 0x7fffffffll ? 0x7fffffff : (int)span;
    double midx = mul(add((double)px, 0.5), scale), midy = mul(add((double)py, 0.5), scale);
    double mid_seg_y = add(mul(d_x, by), mul(sub(midx, bx), d_y));
    e = sub(mid_seg_y, mul(midy, d_x));
    e_x_inc = mul(mul((double)inc_x, scale), d_y);
    e_y_inc = mul(mul((double)(-inc_y), scale), d_x);
    n = 0;
    done = false;
  }
  // current cell is (px, py); returns false when the walk must fail over to Bresenham
  SG_DEV bool advance() {
    ++n;
    if (px == endx && py == endy) { done = true; return true; }
    if (cells_nm < n) return false;  // fail-over on fp rounding errors
    double e_x = add(e, e_x_inc), e_y = add(e, e_y_inc);
    double diff = sub(fabs(e_y), fabs(e_x));
    // are_equal(diff, 0.0) (math_utils.h:10-16) is |diff - 0| <= 1e-7 * max(1, |diff|, 0): for |diff| <= 1 the scale is
    // 1, above 1 the test fails either way -- so it is exactly |diff| <= 1e-7, without the chain of max / multiply.
    // The three-way choice of the reference (diagonal step on a tie -- only along the axis that is left when the other
    // has arrived --, x step if the error favours it, else y step) is spelled with selects: the lanes of a warp walk
    // different rays, and as branches the three arms run one after the other for every step.
    const bool tie = fabs(diff) <= 1e-7;
    const bool x_wins = 0 < diff;
    const bool step_x = tie ? px != endx : x_wins;
    const bool step_y = tie ? (px == endx || py != endy) : !x_wins;
    e = tie ? 0.0 : (x_wins ? e_x : e_y);
    px += step_x ? inc_x : 0;
    py += step_y ? inc_y : 0;
    return true;
  }
};

// integer Bresenham of the fail-over path, resumable
struct Bresenham {
  int primary, secondary, limit, d_primary, d_secondary, inc_p, inc_s, error;
  bool y_primary, done;
  SG_DEV void init(int bx, int by, int ex, int ey) {
    int dx = ex - bx, dy = ey - by;
    y_primary = abs(dx) < abs(dy);
    limit = y_primary ? ey : ex;
    primary = y_primary ? by : bx; d_primary = y_primary ? dy : dx;
    secondary = y_primary ? bx : by; d_secondary = y_primary ? dx : dy;
    inc_p = 0 < d_primary ? 1 : -1; inc_s = 0 < d_secondary ? 1 : -1;
    error = 0; done = false;
  }
  SG_DEV int x() const { return y_primary ? secondary : primary; }
  SG_DEV int y() const { return y_primary ? primary : secondary; }
  SG_DEV void advance() {
    if (primary == limit) { done = true; return; }
    int e_p = error + inc_p * d_secondary;
    int e_b = e_p - inc_s * d_primary;
    primary += inc_p;
    if (abs(e_p) < abs(e_b)) error = e_p;
    else { secondary += inc_s; error = e_b; }
  }
};

// walks the whole segment, calling emit(k, x, y) per cell; returns the number of cells
template <class Emit>
SG_DEV int raycast(double bx, double by, double ex, double ey, double scale, Emit emit) {
  RayWalker w;
  w.init(bx, by, ex, ey, scale);
  for (;;) {
    emit((int)w.n, w.px, w.py);
    if (!w.advance()) break;
    if (w.done) return (int)w.n;
  }
  Bresenham b;
  b.init(w.sx, w.sy, w.endx, w.endy);
  int n = 0;
  for (;;) {
    emit(n, b.x(), b.y());
    ++n;
    b.advance();
    if (b.done) break;
  }
  return n;
}

// ---------------------------------------------------------------- area occupancy estimator
struct Seg { double bx, by, ex, ey; bool horiz, vert; };
SG_DEV Seg mkseg(double bx, double by, double ex, double ey) {
  Seg s{bx, by, ex, ey, false, false};
  s.horiz = are_equal(by, ey);
  s.vert = are_equal(bx, ex);
  return s;
}
enum { LOC_BOT = 0, LOC_LEFT = 1, LOC_TOP = 2, LOC_RIGHT = 3 };
struct Isect { double x, y; int loc; };
SG_DEV bool isect_horiz(const Isect &i) { return i.loc == LOC_BOT || i.loc == LOC_TOP; }
struct Isects { Isect v[4]; int n; };
struct Rect { double bot, top, left, right; };

SG_DEV Seg rect_edge(const Rect &c, int loc) {
  switch (loc) {
    case LOC_BOT: return mkseg(c.left, c.bot, c.right, c.bot);
    case LOC_TOP: return mkseg(c.left, c.top, c.right, c.top);
    case LOC_LEFT: return mkseg(c.left, c.bot, c.left, c.top);
    default: return mkseg(c.right, c.bot, c.right, c.top);
  }
}
// Segment2D::contains (axis aligned only)
SG_DEV bool seg_contains(const Seg &s, double x, double y) {
  if (s.horiz) return are_equal(y, s.by) && are_ordered(s.bx, x, s.ex);
  if (s.vert) return are_equal(x, s.bx) && are_ordered(s.by, y, s.ey);
  return false;
}
// Segment2D::contains_intersection
SG_DEV bool seg_contains_isect(const Seg &s, double x, double y) {
  bool xp = are_ordered(s.bx, x, s.ex) || are_ordered(s.ex, x, s.bx);
  bool yp = are_ordered(s.by, y, s.ey) || are_ordered(s.ey, y, s.by);
  return xp && yp;
}
// Ray::intersect
SG_DEV void ray_isect(double rbx, double rby, double rdx, double rdy, const Seg &s, int loc, Isects &out) {
  if (s.horiz) {
    if (are_equal(rdy, 0.0)) return;
    double alpha = div(sub(s.by, rby), rdy);
    double ix = add(rbx, mul(alpha, rdx));
    if (ix < s.bx || s.ex < ix) return;
    out.v[out.n++] = Isect{ix, s.by, loc};
    return;
  }
  if (s.vert) {
    if (are_equal(rdx, 0.0)) return;
    double alpha = div(sub(s.bx, rbx), rdx);
    double iy = add(rby, mul(alpha, rdy));
    if (iy < s.by || s.ey < iy) return;
    out.v[out.n++] = Isect{s.bx, iy, loc};
  }
}
SG_DEV bool pt_eq(const Isect &a, const Isect &b) { return are_equal(a.x, b.x) && are_equal(a.y, b.y); }
// Rectangle::find_intersections(Ray): edges in the order top, left, bot, right; duplicates dropped
SG_DEV Isects rect_isect_ray(const Rect &c, double rbx, double rby, double rdx, double rdy) {
  Isects r;
  r.n = 0;
  ray_isect(rbx, rby, rdx, rdy, rect_edge(c, LOC_TOP), LOC_TOP, r);
  ray_isect(rbx, rby, rdx, rdy, rect_edge(c, LOC_LEFT), LOC_LEFT, r);
  ray_isect(rbx, rby, rdx, rdy, rect_edge(c, LOC_BOT), LOC_BOT, r);
  ray_isect(rbx, rby, rdx, rdy, rect_edge(c, LOC_RIGHT), LOC_RIGHT, r);
  if (1 < r.n && pt_eq(r.v[0], r.v[r.n - 1])) --r.n;
  int m = 0;
  for (int i = 0; i < r.n; ++i)
    if (m == 0 || !pt_eq(r.v[m - 1], r.v[i])) r.v[m++] = r.v[i];
  r.n = m;
  return r;
}
SG_DEV Isects rect_isect_seg(const Rect &c, const Seg &s) {
  Isects all = rect_isect_ray(c, s.bx, s.by, sub(s.ex, s.bx), sub(s.ey, s.by)), r;
  r.n = 0;
  for (int i = 0; i < all.n; ++i)
    if (seg_contains_isect(s, all.v[i].x, all.v[i].y)) r.v[r.n++] = all.v[i];
  return r;
}
SG_DEV bool has_on_edge_line(const Rect &c, const Seg &s) {
  if (s.vert) return are_equal(s.bx, c.left) || are_equal(s.bx, c.right);
  if (s.horiz) return are_equal(s.by, c.bot) || are_equal(s.by, c.top);
  return false;
}
SG_DEV bool on_some_edge(const Rect &c, double x, double y) {
  for (int loc = 0; loc < 4; ++loc)
    if (seg_contains(rect_edge(c, loc), x, y)) return true;
  return false;
}
SG_DEV bool rect_contains(const Rect &c, double x, double y) { return are_ordered(c.left, x, c.right) && are_ordered(c.bot, y, c.top); }

enum { POS_UNRELATED = 0, POS_LIES_INSIDE, POS_STOPS_INSIDE, POS_STARTS_INSIDE, POS_PIERCES, POS_TOUCHES };

SG_DEV int classify(const Seg &s, const Rect &c) {
  bool beg_in, end_in;
  bool beg_edge = on_some_edge(c, s.bx, s.by), end_edge = on_some_edge(c, s.ex, s.ey);
  if (beg_edge && end_edge) {
    beg_in = end_in = false;
  } else {
    bool bc = rect_contains(c, s.bx, s.by), ec = rect_contains(c, s.ex, s.ey);
    if (!beg_edge && !end_edge) { beg_in = bc; end_in = ec; }
    else if (beg_edge) { beg_in = false; end_in = ec; }
    else { beg_in = bc; end_in = !bc; }
  }
  if (beg_in != end_in) return beg_in ? POS_STARTS_INSIDE : POS_STOPS_INSIDE;
  if (beg_in) return POS_LIES_INSIDE;
  Isects is = rect_isect_seg(c, s);
  if (is.n == 1) return POS_TOUCHES;
  if (is.n == 2) return POS_PIERCES;
  return POS_UNRELATED;
}

SG_DEV void final_estimate(const slamgpu_estimator &e, double chunk, double total, bool is_occ, double *p, double *q) {
  double rate = div(chunk, total);
  if (is_occ) {
    *p = maxd(rate, e.empty_p);
    *q = e.occ_q;
  } else {
    if (0.5 < rate) rate = sub(1.0, rate);
    *p = e.empty_p;
    *q = mul(e.empty_q, rate);
  }
}

// are_on_the_same_side (left-to-right evaluation as written upstream)
SG_DEV bool same_side(double l1x, double l1y, double l2x, double l2y, double p1x, double p1y, double p2x, double p2y) {
  double dx = sub(l2x, l1x), dy = sub(l2y, l1y);
  double a = sub(add(sub(mul(dy, p1y), mul(dx, p1x)), mul(dy, p1x)), mul(dx, p1y));
  double b = sub(add(sub(mul(dy, p2y), mul(dx, p2x)), mul(dy, p2x)), mul(dx, p2y));
  return 0 < mul(a, b);
}

SG_DEV double chunk_area(const Seg &beam, const Rect &c, bool is_occ, const Isects &in) {
  double cell_area = mul(sub(c.top, c.bot), sub(c.right, c.left));
  if (in.n == 0) return div(cell_area, 2.0);
  double corner_x = 0, corner_y = 0, area = 0;
  bool tri = isect_horiz(in.v[0]) != isect_horiz(in.v[1]);
  if (tri) {
    for (int i = 0; i < 2; ++i) {
      switch (in.v[i].loc) {
        case LOC_BOT: corner_y = c.bot; break;
        case LOC_TOP: corner_y = c.top; break;
        case LOC_LEFT: corner_x = c.left; break;
        case LOC_RIGHT: corner_x = c.right; break;
      }
    }
    area = 0.5;
    for (int i = 0; i < 2; ++i) {
      if (isect_horiz(in.v[i])) area = mul(area, fabs(sub(in.v[i].x, corner_x)));
      else area = mul(area, fabs(sub(in.v[i].y, corner_y)));
    }
  } else {
    corner_x = c.left; corner_y = c.bot;
    double base_sum = 0;
    for (int i = 0; i < 2; ++i) {
      if (isect_horiz(in.v[i])) base_sum = add(base_sum, fabs(sub(in.v[i].x, corner_x)));
      else base_sum = add(base_sum, fabs(sub(in.v[i].y, corner_y)));
    }
    area = mul(mul(0.5, sub(c.top, c.bot)), base_sum);
  }
  if (is_occ && same_side(in.v[0].x, in.v[0].y, in.v[1].x, in.v[1].y, beam.bx, beam.by, corner_x, corner_y))
    area = sub(cell_area, area);
  return area;
}

// CellOccupancyEstimator::estimate_occupancy(beam, cell_bounds, is_occ) -> Occupancy{p, q}
// (NaN, NaN) is the reference's "invalid" occupancy: the cell update is skipped.
SG_DEV void estimate_occupancy(const slamgpu_estimator &e, double shift_amount, double bx, double by, double ex, double ey,
                               double cbot, double ctop, double cleft, double cright, bool is_occ, double *p, double *q) {
  if (e.type == SLAMGPU_EST_CONST) {
    *p = is_occ ? e.occ_p : e.empty_p;
    *q = is_occ ? e.occ_q : e.empty_q;
    return;
  }
  Rect c{cbot, ctop, cleft, cright};
  Seg beam = mkseg(bx, by, ex, ey);
  if (has_on_edge_line(c, beam)) {  // ensure_segment_not_on_edge, :68-86
    double shx = 0, shy = 0;
    if (beam.horiz) shy = mul(are_equal(beam.by, c.top) ? -1.0 : 1.0, shift_amount);
    else if (beam.vert) shx = mul(are_equal(beam.bx, c.right) ? -1.0 : 1.0, shift_amount);
    beam = mkseg(add(beam.bx, shx), add(beam.by, shy), add(beam.ex, shx), add(beam.ey, shy));
  }
  const double nan = NAN;
  switch (classify(beam, c)) {
    case POS_UNRELATED:
    case POS_TOUCHES: *p = *q = nan; return;
    case POS_PIERCES:
    case POS_STARTS_INSIDE:
      if (is_occ) { *p = *q = nan; return; }
      break;
    case POS_LIES_INSIDE:
      if (is_occ) { *p = *q = nan; return; }
      *p = e.empty_p; *q = e.unknown_qual; return;
    default: break;
  }
  Isects in;
  if (is_occ) {  // a ray perpendicular to the beam through its end point (:142-146)
    in = rect_isect_ray(c, beam.ex, beam.ey, sub(beam.by, beam.ey), sub(beam.ex, beam.bx));
  } else {
    Isects all = rect_isect_ray(c, beam.bx, beam.by, sub(beam.ex, beam.bx), sub(beam.ey, beam.by));
    in.n = 0;
    for (int i = 0; i < all.n; ++i)
      if (seg_contains_isect(beam, all.v[i].x, all.v[i].y)) in.v[in.n++] = all.v[i];
  }
  double cell_area = mul(sub(c.top, c.bot), sub(c.right, c.left));
  if (in.n == 1) {
    if (!is_occ) { *p = e.empty_p; *q = e.unknown_qual; return; }
    Isects raw = rect_isect_seg(c, beam);
    if (raw.n <= 1) {  // stops at the front vertex: the whole cell is occupied
      final_estimate(e, cell_area, cell_area, is_occ, p, q);
      return;
    }
    in = raw;  // rear vertex: treat the cell as an empty pierce
    is_occ = false;
  }
  double chunk = chunk_area(beam, c, is_occ, in);
  final_estimate(e, chunk, cell_area, is_occ, p, q);
}

}  // namespace sg
