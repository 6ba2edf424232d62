/*
 * orc_cells.c -- dense grid map storage (with the reference's growth rules),
 * cell update rules and the per-cell observation impact (the "score LUT").
 * TEST INFRASTRUCTURE (see slam_oracle.h).  Parity: pinned.
 */
#include "slam_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD(a, b) (((a) < (b)) ? (b) : (a))

int orc_model_stride(int model) {
  static const int s[ORC_CELL_MODELS] = {3, 2, 2, 6, 6, 5, 6};
  return (model >= 0 && model < ORC_CELL_MODELS) ? s[model] : 0;
}

/* prototypes: MockGridCell{0.5,0} test/core/mock_grid_cell.h:10-11; naive_grid_cells.h:8,27;
 * tbm_grid_cells.h:10 + TBM() transferable_belief_model.h:63-68; gmapping_grid_cell.h:14 */
void orc_default_unknown(int model, double *r) {
  memset(r, 0, sizeof(double) * ORC_MAX_STRIDE);
  switch (model) {
  case ORC_CELL_LWW: r[0] = 0.5; r[1] = 0; r[2] = 0; break;
  case ORC_CELL_AFFINE: r[0] = 0.5; r[1] = 0; break;
  case ORC_CELL_MEAN: r[0] = 0.5; r[1] = 0; break;
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN:
  case ORC_CELL_CREDIBILIST: r[0] = 0.5; r[1] = 1; r[2] = 1; r[3] = 0; r[4] = 0; r[5] = 0; break;
  case ORC_CELL_GMAPPING: r[0] = -1; break;
  }
}

orc_map *orc_map_create(int w, int h, double scale, int model, int grow, const double *unknown_rec) {
  orc_map *m = (orc_map *)calloc(1, sizeof(orc_map));
  m->w = w; m->h = h; m->scale = scale;
  m->ox = w / 2; m->oy = h / 2; /* regular_squares_grid.h:120-122 */
  m->model = model; m->stride = orc_model_stride(model); m->grow = grow;
  if (unknown_rec) memcpy(m->unknown, unknown_rec, sizeof(double) * m->stride);
  else orc_default_unknown(model, m->unknown);
  size_t n = (size_t)w * h;
  m->cells = (double *)malloc(sizeof(double) * m->stride * (n ? n : 1));
  for (size_t i = 0; i < n; ++i) memcpy(m->cells + i * m->stride, m->unknown, sizeof(double) * m->stride);
  return m;
}
void orc_map_destroy(orc_map *m) { if (m) { free(m->cells); free(m); } }
orc_map *orc_map_clone(const orc_map *s) {
  orc_map *m = (orc_map *)malloc(sizeof(orc_map));
  *m = *s;
  size_t n = (size_t)s->w * s->h * s->stride;
  m->cells = (double *)malloc(sizeof(double) * (n ? n : 1));
  memcpy(m->cells, s->cells, sizeof(double) * n);
  return m;
}
void orc_map_info(const orc_map *m, int32_t *w, int32_t *h, double *scale, int32_t *ox, int32_t *oy, int32_t *stride) {
  *w = m->w; *h = m->h; *scale = m->scale; *ox = m->ox; *oy = m->oy; *stride = m->stride;
}
double *orc_map_cells(orc_map *m) { return m->cells; }

static int inside(const orc_map *m, int ix, int iy) { return 0 <= ix && ix < m->w && 0 <= iy && iy < m->h; }
/* plain_grid_map.h:69-73 / lazy_tiled_grid_map.h:140-147: unknown cell outside */
const double *orc_map_at(const orc_map *m, int x, int y) {
  int ix = x + m->ox, iy = y + m->oy;
  if (!inside(m, ix, iy)) return m->unknown;
  return m->cells + ((size_t)iy * m->w + ix) * m->stride;
}
/* regular_squares_grid.h:124-126; unbounded maps: plain_grid_map.h:77 */
int orc_map_has_cell(const orc_map *m, int x, int y) {
  if (m->grow != ORC_GROW_NONE) return 1;
  return inside(m, x + m->ox, y + m->oy);
}

static void regrow(orc_map *m, unsigned prep_x, unsigned prep_y, unsigned new_w, unsigned new_h) {
  double *nc = (double *)malloc(sizeof(double) * m->stride * (size_t)new_w * new_h);
  for (size_t i = 0; i < (size_t)new_w * new_h; ++i) memcpy(nc + i * m->stride, m->unknown, sizeof(double) * m->stride);
  for (int y = 0; y < m->h; ++y)
    memcpy(nc + (((size_t)y + prep_y) * new_w + prep_x) * m->stride, m->cells + (size_t)y * m->w * m->stride,
           sizeof(double) * m->stride * m->w);
  free(m->cells);
  m->cells = nc;
  m->w = (int)new_w; m->h = (int)new_h;
  m->ox += (int)prep_x; m->oy += (int)prep_y;
}

/* UnboundedPlainGridMap::ensure_inside plain_grid_map.h:133-173 (incl. its
 * unsigned/double conversions) and UnboundedLazyTiledGridMap::ensure_inside
 * lazy_tiled_grid_map.h:151-181 */
int orc_map_ensure_inside(orc_map *m, int x, int y) {
  int cx = x + m->ox, cy = y + m->oy;
  if (inside(m, cx, cy)) return 0;
  if (m->grow == ORC_GROW_PLAIN) {
    unsigned w = m->w, h = m->h, prep_x = 0, app_x = 0, prep_y = 0, app_y = 0;
    if (cx < 0) prep_x = 0 - cx; else if ((int)w <= cx) app_x = cx - w + 1;
    if (cy < 0) prep_y = 0 - cy; else if ((int)h <= cy) app_y = cy - h + 1;
    unsigned new_w = prep_x + w + app_x, new_h = prep_y + h + app_y;
    const double rate = 1.2;
    if (w < new_w && new_w < rate * w) {
      double sc = prep_x / (new_w - w); /* unsigned division, as written upstream */
      prep_x += (rate * w - new_w) * sc;
      new_w = rate * w;
      app_x = new_w - (prep_x + w);
    }
    if (h < new_h && new_h < rate * h) {
      double sc = prep_y / (new_h - h);
      prep_y += (rate * h - new_h) * sc;
      new_h = rate * h;
      app_y = new_h - (prep_y + h);
    }
    (void)app_x; (void)app_y;
    regrow(m, prep_x, prep_y, new_w, new_h);
    return 1;
  }
  if (m->grow == ORC_GROW_TILED) {
    const unsigned bits = 7, tile = 1u << bits;
    unsigned tx = (m->w + tile - 1) / tile, ty = (m->h + tile - 1) / tile;
    unsigned prep_x = 0, app_x = 0, prep_y = 0, app_y = 0;
    if (cx < 0) prep_x = 1 + ((0 - cx) >> bits); else if (m->w <= cx) app_x = 1 + ((cx - m->w) >> bits);
    if (cy < 0) prep_y = 1 + ((0 - cy) >> bits); else if (m->h <= cy) app_y = 1 + ((cy - m->h) >> bits);
    unsigned ntx = prep_x + tx + app_x, nty = prep_y + ty + app_y;
    regrow(m, prep_x * tile, prep_y * tile, ntx * tile, nty * tile);
    return 1;
  }
  return 0;
}

void orc_map_reset_cell(orc_map *m, int x, int y, const double *rec) {
  if (m->grow != ORC_GROW_NONE) orc_map_ensure_inside(m, x, y);
  int ix = x + m->ox, iy = y + m->oy;
  if (!inside(m, ix, iy)) { ++m->oob_updates; return; }
  memcpy(m->cells + ((size_t)iy * m->w + ix) * m->stride, rec, sizeof(double) * m->stride);
}

/* ---- transferable_belief_model.h:70-143 ---- */
typedef struct { double b[4]; } tbm; /* unknown, empty, occupied, conflict */
static void tbm_default(tbm *t) { t->b[0] = 1.0; t->b[1] = t->b[2] = t->b[3] = 0.0; }
static tbm tbm_conj(const tbm *l, const tbm *r) {
  tbm t = {{0.0, 0.0, 0.0, 0.0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) t.b[i | j] += l->b[i] * r->b[j];
  double tot = t.b[0] + t.b[1] + t.b[2] + t.b[3];
  if (tot == 0.0) tbm_default(&t);
  else { t.b[0] /= tot; t.b[1] /= tot; t.b[2] /= tot; t.b[3] /= tot; }
  return t;
}
static void tbm_norm_conflict(tbm *t) {
  double w = t->b[0] + t->b[1] + t->b[2];
  if (w == 0.0) tbm_default(t);
  else { t->b[0] /= w; t->b[1] /= w; t->b[2] /= w; t->b[3] = 0.0; }
}
/* aoo2tbm tbm_grid_cells.h:57-66 */
static tbm aoo2tbm(double p, double q, double quality) {
  tbm t;
  if (isnan(p) || isnan(q)) { tbm_default(&t); return t; }
  double est = q * quality;
  double occupied = p * est, empty = (1 - p) * est;
  t.b[0] = 1.0 - occupied - empty; t.b[1] = empty; t.b[2] = occupied; t.b[3] = 0.0;
  return t;
}

/* operator+= of each cell model (grid_cell.h:27-30, naive_grid_cells.h:14-20,33-40,
 * tbm_grid_cells.h:12-19,93-106, gmapping_grid_cell.h:20-33) */
void orc_cell_update(int model, double *r, int aoo_is_occ, double p, double q, double ox, double oy, double quality) {
  (void)aoo_is_occ;
  int valid = !isnan(p) && !isnan(q);
  switch (model) {
  case ORC_CELL_LWW: r[0] = p; r[1] = q; r[2] = 1; break;
  case ORC_CELL_AFFINE:
    if (!valid) return;
    r[0] = (1.0 - quality) * r[0] + quality * p;
    r[1] = 1;
    break;
  case ORC_CELL_MEAN: {
    if (!valid) return;
    r[1] += 1;
    double that_p = 0.5 + (p - 0.5) * quality;
    r[0] = (r[0] * (r[1] - 1) + that_p) / r[1];
    break;
  }
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN:
  case ORC_CELL_CREDIBILIST: { /* credibilist/grid_cell.h:23-29 + TBM_prob_conversion.h:8-21: the unknown-even arithmetic */
    if (!valid) return;
    tbm b = {{r[2], r[3], r[4], 0.0}}, m = aoo2tbm(p, q, quality);
    b = tbm_conj(&b, &m);
    tbm_norm_conflict(&b);
    r[2] = b.b[0]; r[3] = b.b[1]; r[4] = b.b[2];
    if (model == ORC_CELL_TBM_CONSISTENT) {
      double qual = b.b[2] + b.b[1];
      r[0] = b.b[2] / qual; r[1] = qual;
    } else {
      r[0] = b.b[2] + 0.5 * b.b[0]; r[1] = 1.0;
    }
    r[5] = 1;
    break;
  }
  case ORC_CELL_GMAPPING: {
    if (!valid) return;
    r[4] += 1; /* tries */
    int free_ = p <= 0.5;
    double aoo_p = free_ ? 0.0 : p;
    r[0] = (r[0] * (r[4] - 1) + aoo_p) / r[4];
    if (free_) return;
    r[3] += 1; /* hits */
    r[1] = (r[1] * (r[3] - 1) + ox) / r[3];
    r[2] = (r[2] * (r[3] - 1) + oy) / r[3];
    break;
  }
  }
}

/* GridCell::discrepancy grid_cell.h:33-35; TbmBaseCell tbm_grid_cells.h:21-35; Gmapping :35-38 */
double orc_cell_discrepancy(int model, const double *r, double p, double q, double ox, double oy, double quality) {
  switch (model) {
  case ORC_CELL_LWW:
  case ORC_CELL_AFFINE:
  case ORC_CELL_MEAN: return fabs(r[0] - p);
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN: {
    tbm that = aoo2tbm(p, q, quality), b = {{r[2], r[3], r[4], 0.0}};
    double total_unknown = that.b[0] + b.b[0];
    double d_occ = fabs(that.b[2] - b.b[2]);
    tbm comb = tbm_conj(&that, &b);
    double unknown = total_unknown / 2.0;
    double known = 1 - unknown;
    double known_disc = known * (comb.b[3] + d_occ) / 2.0;
    return unknown / 2 + known_disc;
  }
  case ORC_CELL_CREDIBILIST: { /* 1 - score: credibilist/grid_cell.h:31-40, disjunctive transferable_belief_model.h:145-162 */
    tbm that = aoo2tbm(p, q, quality), b = {{r[2], r[3], r[4], 0.0}}, t = {{0.0, 0.0, 0.0, 0.0}};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) t.b[i & j] += that.b[i] * b.b[j];
    double tot = t.b[0] + t.b[1] + t.b[2] + t.b[3];
    double occ = tot == 0.0 ? 0.0 : t.b[2] / tot; /* TBM::normalize: all-zero -> default (occupied 0) */
    return 1.0 - occ;
  }
  case ORC_CELL_GMAPPING: {
    double d = pow(r[1] - ox, 2) + pow(r[2] - oy, 2);
    double sim = exp(-d / 0.05);
    return 1.0 - sim;
  }
  }
  return 0;
}

/* observation_impact_estimators.h:14-28 with obstacle_AOO {true,{1,1},obst,1} grid_scan_matcher.h:44-46 */
double orc_cell_impact(int model, int oie, const double *r, double ox, double oy) {
  if (oie == ORC_OIE_OCCUPANCY) return r[0];
  return 1.0 - orc_cell_discrepancy(model, r, 1.0, 1.0, ox, oy, 1.0);
}

void orc_build_lut(const orc_map *m, int oie, double *lut, double *unknown_value) {
  size_t n = (size_t)m->w * m->h;
  for (size_t i = 0; i < n; ++i) lut[i] = orc_cell_impact(m->model, oie, m->cells + i * m->stride, 0, 0);
  if (unknown_value) *unknown_value = orc_cell_impact(m->model, oie, m->unknown, 0, 0);
}
