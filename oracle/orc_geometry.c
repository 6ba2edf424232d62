/*
 * orc_geometry.c -- fuzzy compares, world->cell, segment ray cast, Bresenham
 * fail-over, rectangle rasterisation/overlap and the const/area cell occupancy
 * estimators.  TEST INFRASTRUCTURE (see slam_oracle.h).  Parity: pinned.
 */
#include "slam_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>

#define MAXD(a, b) (((a) < (b)) ? (b) : (a)) /* std::max(a,b) */
#define MIND(a, b) (((b) < (a)) ? (b) : (a)) /* std::min(a,b) */

/* ---- src/core/math_utils.h:10-51 ---- */
static int eq_eps(double a, double b, double eps) { return fabs(a - b) <= eps; }
int orc_are_equal(double a, double b) {
  double sc = MAXD(1.0, MAXD(fabs(a), fabs(b)));
  return eq_eps(a, b, 1e-7 * sc);
}
int orc_less(double a, double b) { return a < b + DBL_EPSILON; }
int orc_less_or_equal(double a, double b) { return orc_are_equal(a, b) || orc_less(a, b); }
static int are_ordered(double a, double b, double c) { return orc_less_or_equal(a, b) && orc_less_or_equal(b, c); }

/* ---- src/core/maps/regular_squares_grid.h:40-46 ---- */
int orc_world_to_cell(double v, double scale) { return (int)floor(v / scale); }

/* ---- src/core/geometry_discrete_primitives.h:67-100 (integer Bresenham) ---- */
int orc_bresenham(int bx, int by, int ex, int ey, int32_t *xy, int cap) {
  int dx = ex - bx, dy = ey - by;
  int y_primary = abs(dx) < abs(dy);
  int limit = y_primary ? ey : ex;
  int primary = y_primary ? by : bx, d_primary = y_primary ? dy : dx;
  int secondary = y_primary ? bx : by, d_secondary = y_primary ? dx : dy;
  int inc_p = 0 < d_primary ? 1 : -1, inc_s = 0 < d_secondary ? 1 : -1;
  int error = 0, n = 0;
  for (;;) {
    if (n < cap) {
      xy[2 * n] = y_primary ? secondary : primary;
      xy[2 * n + 1] = y_primary ? primary : secondary;
    }
    ++n;
    if (primary == limit) break;
    int e_p = error + inc_p * d_secondary;
    int e_b = e_p - inc_s * d_primary;
    primary += inc_p;
    if (abs(e_p) < abs(e_b)) {
      error = e_p;
    } else {
      secondary += inc_s;
      error = e_b;
    }
  }
  return n;
}

/* ---- src/core/maps/regular_squares_grid.h:56-101 (modified 4-connected line) ---- */
int orc_raycast(double bx, double by, double ex, double ey, double scale, int32_t *xy, int cap) {
  double d_x = ex - bx, d_y = ey - by;
  int inc_x = 0 < d_x ? 1 : -1, inc_y = 0 < d_y ? 1 : -1;
  int px = orc_world_to_cell(bx, scale), py = orc_world_to_cell(by, scale);
  const int sx = px, sy = py;
  const int endx = orc_world_to_cell(ex, scale), endy = orc_world_to_cell(ey, scale);
  long cells_nm = labs((long)endx - px) + labs((long)endy - py) + 1;
  double midx = (px + 0.5) * scale, midy = (py + 0.5) * scale;
  double mid_seg_y = d_x * by + (midx - bx) * d_y;
  double e = mid_seg_y - midy * d_x;
  double e_x_inc = inc_x * scale * d_y;
  double e_y_inc = -inc_y * scale * d_x;
  long n = 0;
  for (;;) {
    if (n < cap) { xy[2 * n] = px; xy[2 * n + 1] = py; }
    ++n;
    if (px == endx && py == endy) break;
    if (cells_nm < n) /* fail-over on fp rounding errors */
      return orc_bresenham(sx, sy, endx, endy, xy, cap);
    double e_x = e + e_x_inc, e_y = e + e_y_inc;
    double diff = fabs(e_y) - fabs(e_x);
    if (orc_are_equal(diff, 0)) {
      if (px == endx) py += inc_y;
      else if (py == endy) px += inc_x;
      else { px += inc_x; py += inc_y; }
      e = 0;
    } else if (0 < diff) {
      px += inc_x; e = e_x;
    } else {
      py += inc_y; e = e_y;
    }
  }
  return (int)n;
}

/* ---- src/core/maps/grid_rasterization.h:26-47 ----
 * lbrt = {lb.x, lb.y, rt.x, rt.y}; iteration is x outer / y inner (:56-64).
 * returns number of cells (0 if rt < lb). */
int orc_rasterize_rect(double scale, int w, int h, int ox, int oy, double bot, double top, double left, double right,
                       int include_border, int32_t *lbrt) {
  double offset = include_border ? 0 : 1e-9;
  double area = (top - bot) * (right - left); /* LightWeightRectangle::area geometry_primitives.h:186-190 */
  int inf = area == INFINITY, empty = area == 0;
  if (!empty && !inf) {
    lbrt[0] = orc_world_to_cell(left + offset, scale);
    lbrt[1] = orc_world_to_cell(bot + offset, scale);
    lbrt[2] = orc_world_to_cell(right - offset, scale);
    lbrt[3] = orc_world_to_cell(top - offset, scale);
  } else if (empty) {
    lbrt[0] = lbrt[2] = orc_world_to_cell(left, scale);
    lbrt[1] = lbrt[3] = orc_world_to_cell(bot, scale);
  } else {
    lbrt[0] = 0 - ox; lbrt[1] = 0 - oy;
    lbrt[2] = w - 1 - ox; lbrt[3] = h - 1 - oy;
  }
  if (lbrt[2] < lbrt[0] || lbrt[3] < lbrt[1]) return 0;
  return (lbrt[2] - lbrt[0] + 1) * (lbrt[3] - lbrt[1] + 1);
}

/* ---- LightWeightRectangle::overlap / intersect_internal, geometry_primitives.h:252-318 ---- */
typedef struct { double b, t, l, r; } lwr;
static int lwr_contains(const lwr *a, double x, double y) { return are_ordered(a->l, x, a->r) && are_ordered(a->b, y, a->t); }
static double lwr_area(const lwr *a) { return (a->t - a->b) * (a->r - a->l); }
static lwr lwr_intersect(const lwr *a, const lwr *that, int reversed) {
  unsigned nm = 0;
  double cl = a->l, cr = a->r, ct = a->t, cb = a->b;
  if (lwr_contains(a, that->l, that->b)) { ++nm; cl = that->l; cb = that->b; }
  if (lwr_contains(a, that->r, that->b)) { ++nm; cr = that->r; cb = that->b; }
  if (lwr_contains(a, that->l, that->t)) { ++nm; cl = that->l; ct = that->t; }
  if (lwr_contains(a, that->r, that->t)) { ++nm; cr = that->r; ct = that->t; }
  if (nm == 0) {
    lwr z = {0, 0, 0, 0};
    return reversed ? z : lwr_intersect(that, a, 1);
  }
  /* nm == 3 asserts in the reference */
  lwr res = {cb, ct, cl, cr};
  return res;
}
double orc_rect_overlap(double ab, double at, double al, double ar, double bb, double bt, double bl, double br) {
  lwr a = {ab, at, al, ar}, b = {bb, bt, bl, br};
  if (lwr_area(&a)) {
    lwr i = lwr_intersect(&a, &b, 0);
    return lwr_area(&i) / lwr_area(&a);
  }
  if (lwr_area(&b)) return lwr_contains(&b, a.l, a.b) ? 1.0 : 0.0;
  return (orc_are_equal(a.t, b.t) && orc_are_equal(a.b, b.b) && orc_are_equal(a.l, b.l) && orc_are_equal(a.r, b.r)) ? 1.0 : 0.0;
}

/* =========================================================================
 * AreaOccupancyEstimator, src/core/maps/area_occupancy_estimator.h:27-240,
 * over Segment2D / Ray / Rectangle of src/core/geometry_primitives.h.
 * ========================================================================= */
typedef struct { double bx, by, ex, ey; int horiz, vert; } seg; /* Segment2D :37-99 */
static seg mkseg(double bx, double by, double ex, double ey) {
  seg s = {bx, by, ex, ey, 0, 0};
  s.horiz = orc_are_equal(by, ey);
  s.vert = orc_are_equal(bx, ex);
  return s;
}
enum { LOC_BOT = 0, LOC_LEFT = 1, LOC_TOP = 2, LOC_RIGHT = 3 };
typedef struct { double x, y; int loc; } isect;
static int isect_horiz(const isect *i) { return i->loc == LOC_BOT || i->loc == LOC_TOP; }
typedef struct { isect v[4]; int n; } isects;
typedef struct { double bot, top, left, right; seg e[4]; /* bot, top, left, right */ } rect;
static rect mkrect(double b, double t, double l, double r) {
  rect c = {b, t, l, r, {{0}}};
  c.e[0] = mkseg(l, b, r, b);
  c.e[1] = mkseg(l, t, r, t);
  c.e[2] = mkseg(l, b, l, t);
  c.e[3] = mkseg(r, b, r, t);
  return c;
}
/* Segment2D::contains :59-68 (axis aligned only; anything else asserts) */
static int seg_contains(const seg *s, double x, double y) {
  if (s->horiz) return orc_are_equal(y, s->by) && are_ordered(s->bx, x, s->ex);
  if (s->vert) return orc_are_equal(x, s->bx) && are_ordered(s->by, y, s->ey);
  return 0;
}
/* Segment2D::contains_intersection :70-88 */
static int seg_contains_isect(const seg *s, double x, double y) {
  int xp = are_ordered(s->bx, x, s->ex) || are_ordered(s->ex, x, s->bx);
  int yp = are_ordered(s->by, y, s->ey) || are_ordered(s->ey, y, s->by);
  return xp && yp;
}
/* Ray::intersect :126-166 */
static void ray_isect(double rbx, double rby, double rdx, double rdy, const seg *s, int loc, isects *out) {
  if (s->horiz) {
    if (orc_are_equal(rdy, 0)) return;
    double alpha = (s->by - rby) / rdy;
    double ix = rbx + alpha * rdx;
    if (ix < s->bx || s->ex < ix) return;
    isect i = {ix, s->by, loc};
    out->v[out->n++] = i;
    return;
  }
  if (s->vert) {
    if (orc_are_equal(rdx, 0)) return;
    double alpha = (s->bx - rbx) / rdx;
    double iy = rby + alpha * rdy;
    if (iy < s->by || s->ey < iy) return;
    isect i = {s->bx, iy, loc};
    out->v[out->n++] = i;
  }
}
static int pt_eq(const isect *a, const isect *b) { return orc_are_equal(a->x, b->x) && orc_are_equal(a->y, b->y); }
/* Rectangle::find_intersections(Ray) :370-388 -- order top, left, bot, right */
static isects rect_isect_ray(const rect *c, double rbx, double rby, double rdx, double rdy) {
  isects r; r.n = 0;
  ray_isect(rbx, rby, rdx, rdy, &c->e[1], LOC_TOP, &r);
  ray_isect(rbx, rby, rdx, rdy, &c->e[2], LOC_LEFT, &r);
  ray_isect(rbx, rby, rdx, rdy, &c->e[0], LOC_BOT, &r);
  ray_isect(rbx, rby, rdx, rdy, &c->e[3], LOC_RIGHT, &r);
  if (1 < r.n && pt_eq(&r.v[0], &r.v[r.n - 1])) --r.n;
  /* std::unique */
  int m = 0;
  for (int i = 0; i < r.n; ++i)
    if (m == 0 || !pt_eq(&r.v[m - 1], &r.v[i])) r.v[m++] = r.v[i];
  r.n = m;
  return r;
}
/* Rectangle::find_intersections(Segment2D) :362-368 */
static isects rect_isect_seg(const rect *c, const seg *s) {
  isects all = rect_isect_ray(c, s->bx, s->by, s->ex - s->bx, s->ey - s->by), r;
  r.n = 0;
  for (int i = 0; i < all.n; ++i)
    if (seg_contains_isect(s, all.v[i].x, all.v[i].y)) r.v[r.n++] = all.v[i];
  return r;
}
/* Rectangle::has_on_edge_line :342-350 */
static int has_on_edge_line(const rect *c, const seg *s) {
  if (s->vert) return orc_are_equal(s->bx, c->left) || orc_are_equal(s->bx, c->right);
  if (s->horiz) return orc_are_equal(s->by, c->bot) || orc_are_equal(s->by, c->top);
  return 0;
}
/* Rectangle::find_containing_edge :352-358 (as a bool) */
static int on_some_edge(const rect *c, double x, double y) {
  for (int i = 0; i < 4; ++i)
    if (seg_contains(&c->e[i], x, y)) return 1;
  return 0;
}
static int rect_contains(const rect *c, double x, double y) { return are_ordered(c->left, x, c->right) && are_ordered(c->bot, y, c->top); }

enum { POS_UNRELATED = 0, POS_LIES_INSIDE, POS_STOPS_INSIDE, POS_STARTS_INSIDE, POS_PIERCES, POS_TOUCHES };

/* classify_segment :88-109 + modified_is_inside :112-137 */
static int classify(const seg *s, const rect *c) {
  int beg_in, end_in;
  int beg_edge = on_some_edge(c, s->bx, s->by), end_edge = on_some_edge(c, s->ex, s->ey);
  if (beg_edge && end_edge) {
    beg_in = end_in = 0;
  } else {
    int bc = rect_contains(c, s->bx, s->by), ec = rect_contains(c, s->ex, s->ey);
    if (!beg_edge && !end_edge) { beg_in = bc; end_in = ec; }
    else if (beg_edge) { beg_in = 0; end_in = ec; }
    else { beg_in = bc; end_in = !bc; }
  }
  if (beg_in ^ end_in) return beg_in ? POS_STARTS_INSIDE : POS_STOPS_INSIDE;
  if (beg_in) return POS_LIES_INSIDE;
  isects is = rect_isect_seg(c, s);
  switch (is.n) {
  case 0: return POS_UNRELATED;
  case 1: return POS_TOUCHES;
  case 2: return POS_PIERCES;
  }
  return POS_UNRELATED;
}

/* estimate_occupancy(double, double, bool) :225-240 */
static void final_estimate(const orc_estimator *e, double chunk, double total, int is_occ, double *pq) {
  double rate = chunk / total;
  if (is_occ) {
    pq[0] = MAXD(rate, e->empty_p);
    pq[1] = e->occ_q;
  } else {
    if (0.5 < rate) rate = 1 - rate;
    pq[0] = e->empty_p;
    pq[1] = e->empty_q * rate;
  }
}

/* are_on_the_same_side :218-223 (restated verbatim, left-to-right evaluation) */
static int same_side(double l1x, double l1y, double l2x, double l2y, double p1x, double p1y, double p2x, double p2y) {
  double dx = l2x - l1x, dy = l2y - l1y;
  double a = dy * p1y - dx * p1x + dy * p1x - dx * p1y;
  double b = dy * p2y - dx * p2x + dy * p2x - dx * p2y;
  return 0 < a * b;
}

/* compute_chunk_area :162-216 */
static double chunk_area(const seg *beam, const rect *c, int is_occ, const isects *in) {
  double cell_area = (c->top - c->bot) * (c->right - c->left);
  if (in->n == 0) return cell_area / 2;
  double corner_x = 0, corner_y = 0, area = 0;
  int tri = isect_horiz(&in->v[0]) ^ isect_horiz(&in->v[1]);
  if (tri) {
    for (int i = 0; i < 2; ++i) {
      switch (in->v[i].loc) {
      case LOC_BOT: corner_y = c->bot; break;
      case LOC_TOP: corner_y = c->top; break;
      case LOC_LEFT: corner_x = c->left; break;
      case LOC_RIGHT: corner_x = c->right; break;
      }
    }
    area = 0.5;
    for (int i = 0; i < 2; ++i) {
      if (isect_horiz(&in->v[i])) area *= fabs(in->v[i].x - corner_x);
      else area *= fabs(in->v[i].y - corner_y);
    }
  } else {
    corner_x = c->left; corner_y = c->bot;
    double base_sum = 0;
    for (int i = 0; i < 2; ++i) {
      if (isect_horiz(&in->v[i])) base_sum += fabs(in->v[i].x - corner_x);
      else base_sum += fabs(in->v[i].y - corner_y);
    }
    area = 0.5 * (c->top - c->bot) * base_sum;
  }
  if (is_occ && same_side(in->v[0].x, in->v[0].y, in->v[1].x, in->v[1].y, beam->bx, beam->by, corner_x, corner_y))
    area = cell_area - area;
  return area;
}

void orc_estimate_occupancy(const orc_estimator *e, double bx, double by, double ex, double ey, double cbot, double ctop,
                            double cleft, double cright, int is_occ, double *pq) {
  /* ConstOccupancyEstimator, const_occupancy_estimator.h:11-15 */
  if (e->type == ORC_EST_CONST) {
    pq[0] = is_occ ? e->occ_p : e->empty_p;
    pq[1] = is_occ ? e->occ_q : e->empty_q;
    return;
  }
  rect c = mkrect(cbot, ctop, cleft, cright);
  seg beam = mkseg(bx, by, ex, ey);
  /* ensure_segment_not_on_edge :68-86; Shift_Amount is function-static (Q10) */
  double shift_amount = e->shift_amount >= 0 ? e->shift_amount : e->low_qual * (ctop - cbot);
  if (has_on_edge_line(&c, &beam)) {
    double shx = 0, shy = 0;
    if (beam.horiz) shy = (orc_are_equal(beam.by, c.top) ? -1 : 1) * shift_amount;
    else if (beam.vert) shx = (orc_are_equal(beam.bx, c.right) ? -1 : 1) * shift_amount;
    beam = mkseg(beam.bx + shx, beam.by + shy, beam.ex + shx, beam.ey + shy);
  }
  const double nan = NAN;
  switch (classify(&beam, &c)) {
  case POS_UNRELATED:
  case POS_TOUCHES: pq[0] = pq[1] = nan; return;
  case POS_PIERCES:
  case POS_STARTS_INSIDE:
    if (is_occ) { pq[0] = pq[1] = nan; return; }
    break;
  case POS_LIES_INSIDE:
    if (is_occ) { pq[0] = pq[1] = nan; return; }
    pq[0] = e->empty_p; pq[1] = e->unknown_qual; return;
  case POS_STOPS_INSIDE: break;
  }
  /* find_intersections(beam, cell, is_occ) :139-160 */
  isects all, in;
  if (is_occ) {
    in = rect_isect_ray(&c, beam.ex, beam.ey, beam.by - beam.ey, beam.ex - beam.bx);
  } else {
    all = rect_isect_ray(&c, beam.bx, beam.by, beam.ex - beam.bx, beam.ey - beam.by);
    in.n = 0;
    for (int i = 0; i < all.n; ++i)
      if (seg_contains_isect(&beam, all.v[i].x, all.v[i].y)) in.v[in.n++] = all.v[i];
  }
  double cell_area = (c.top - c.bot) * (c.right - c.left);
  if (in.n == 1) {
    if (!is_occ) { pq[0] = e->empty_p; pq[1] = e->unknown_qual; return; }
    isects raw = rect_isect_seg(&c, &beam);
    if (raw.n <= 1) { /* stops at the front vertex: whole cell occupied (case 0 asserts upstream) */
      final_estimate(e, cell_area, cell_area, is_occ, pq);
      return;
    }
    in = raw; /* rear vertex: treat the cell as empty */
    is_occ = 0;
  }
  double chunk = chunk_area(&beam, &c, is_occ, &in);
  final_estimate(e, chunk, cell_area, is_occ, pq);
}
