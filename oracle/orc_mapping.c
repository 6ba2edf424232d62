/*
 * orc_mapping.c -- scan insertion (ray cast -> per-cell occupancy estimate ->
 * cell update with wall blur) and the M3RSM max-pyramid with its incremental
 * update rule and Match upper bound.
 * TEST INFRASTRUCTURE (see slam_oracle.h).  Parity: pinned.
 */
#include "slam_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define MAXD(a, b) (((a) < (b)) ? (b) : (a))

typedef void (*update_fn)(void *ctx, int x, int y, int is_occ, double p, double q, double ox, double oy, double quality);

static void map_update(void *ctx, int x, int y, int is_occ, double p, double q, double ox, double oy, double quality) {
  orc_map *m = (orc_map *)ctx;
  if (m->grow != ORC_GROW_NONE) orc_map_ensure_inside(m, x, y);
  int ix = x + m->ox, iy = y + m->oy;
  if (ix < 0 || ix >= m->w || iy < 0 || iy >= m->h) { ++m->oob_updates; return; } /* reference asserts */
  orc_cell_update(m->model, m->cells + ((size_t)iy * m->w + ix) * m->stride, is_occ, p, q, ox, oy, quality);
}

static double dist_sq_i(int ax, int ay, int bx, int by) { return pow(ax - bx, 2) + pow(ay - by, 2); }

/* GridMapScanAdder::append_scan grid_map_scan_adders.h:54-75 +
 * WallDistanceBlurringScanAdder::handle_scan_point :138-172, blur_cell_dist :176-189 */
static int64_t append_scan_impl(update_fn upd, void *ctx, double scale, const orc_scan *s, double px, double py,
                                double pth, double scan_quality, double scan_margin, const orc_estimator *est_in,
                                double blur, double max_range, int omqe, int32_t *log_xy, int64_t log_cap) {
  if (s->n == 0) return 0;
  orc_estimator est = *est_in;
  uint32_t *hv = NULL;
  if (omqe == ORC_OMQE_AHR) {
    hv = (uint32_t *)malloc(sizeof(uint32_t) * s->n);
    orc_angle_histogram_values(s->n, s->a, s->b, hv);
  }
  const double max_range_sq = pow(max_range, 2);
  int64_t total = 0;
  int cap = 64;
  int32_t *cells = (int32_t *)malloc(sizeof(int32_t) * 2 * cap);
  size_t last_pt_i = s->n - scan_margin - 1;
  for (size_t i = scan_margin; i <= last_pt_i; ++i) {
    double r = s->cartesian ? sqrt(pow(s->a[i], 2) + pow(s->b[i], 2)) : s->a[i];
    double a = s->cartesian ? atan2(s->b[i], s->a[i]) : s->b[i];
    double wx = px + r * cos(pth + a), wy = py + r * sin(pth + a);
    double quality = scan_quality * (hv ? 1.0 / hv[i] : 1.0);
    int is_occ = s->occ[i];
    double len_sq = pow(wx - px, 2) + pow(wy - py, 2);
    if (max_range_sq < len_sq) continue;
    int rx = orc_world_to_cell(px, scale), ry = orc_world_to_cell(py, scale);
    int obx = orc_world_to_cell(wx, scale), oby = orc_world_to_cell(wy, scale);
    double obst_dist_sq = dist_sq_i(rx, ry, obx, oby);
    double blur_dist = 0;
    if (is_occ) {
      blur_dist = blur / scale;
      if (blur_dist < 0) blur_dist *= -len_sq;
    }
    double hole_dist_sq = pow(blur_dist, 2);
    int n = orc_raycast(px, py, wx, wy, scale, cells, cap);
    if (n > cap) {
      cap = n;
      cells = (int32_t *)realloc(cells, sizeof(int32_t) * 2 * cap);
      n = orc_raycast(px, py, wx, wy, scale, cells, cap);
    }
    for (int k = 0; k < n; ++k) {
      if (log_xy && total + k < log_cap) { log_xy[2 * (total + k)] = cells[2 * k]; log_xy[2 * (total + k) + 1] = cells[2 * k + 1]; }
    }
    total += n;
    /* the obstacle cell first */
    int lx = cells[2 * (n - 1)], ly = cells[2 * (n - 1) + 1];
    double base[2];
    if (est.type == ORC_EST_AREA && est.shift_amount < 0) est.shift_amount = est.low_qual * (scale * (ly + 1) - scale * ly);
    orc_estimate_occupancy(&est, px, py, wx, wy, scale * ly, scale * (ly + 1), scale * lx, scale * (lx + 1), is_occ, base);
    upd(ctx, lx, ly, is_occ, base[0], base[1], wx, wy, quality);
    for (int k = 0; k < n - 1; ++k) {
      int cx = cells[2 * k], cy = cells[2 * k + 1];
      double d_sq = dist_sq_i(cx, cy, obx, oby);
      double occ[2];
      orc_estimate_occupancy(&est, px, py, wx, wy, scale * cy, scale * (cy + 1), scale * cx, scale * (cx + 1), 0, occ);
      int aoo_occ = 0;
      if (d_sq < hole_dist_sq && hole_dist_sq < obst_dist_sq) {
        aoo_occ = 1;
        double prob_scale = 1.0 - d_sq / hole_dist_sq;
        occ[0] = base[0] * prob_scale;
      }
      upd(ctx, cx, cy, aoo_occ, occ[0], occ[1], wx, wy, quality);
    }
  }
  free(cells);
  free(hv);
  return total;
}

int64_t orc_append_scan(orc_map *m, const orc_scan *s, double px, double py, double pth, double scan_quality,
                        double scan_margin, const orc_estimator *est, double blur, double max_range, int omqe,
                        int32_t *log_xy, int64_t log_cap) {
  return append_scan_impl(map_update, m, m->scale, s, px, py, pth, scan_quality, scan_margin, est, blur, max_range,
                          omqe, log_xy, log_cap);
}

/* =========================================================================
 * RescalableCachingGridMap (rescalable_caching_grid_map.h:27-42,79-105,171-194)
 * + M3RSMRescalableGridMap (scan_matchers/m3rsm_engine.h:17-131)
 * ========================================================================= */
struct orc_pyramid {
  orc_map **lv;
  int n, cap, oie, model, grow;
  double unknown[ORC_MAX_STRIDE];
};

static int ge_pow2(int i) { int p = 1; while (p < i) p *= 2; return p; }

static void insert_before_last(orc_pyramid *p, orc_map *m) {
  if (p->n == p->cap) { p->cap *= 2; p->lv = (orc_map **)realloc(p->lv, sizeof(orc_map *) * p->cap); }
  p->lv[p->n] = p->lv[p->n - 1];
  p->lv[p->n - 1] = m;
  ++p->n;
}

/* ensure_map_cache_is_continuous :171-194 */
static void ensure_continuous(orc_pyramid *p) {
  if (p->n < 2) return;
  const orc_map *pc = p->lv[p->n - 2];
  int pc_w = pc->w, pc_h = pc->h;
  double pc_scale = pc->scale;
  if (pc_w <= 2 && pc_h <= 2) return;
  pc_w = ge_pow2(pc_w); pc_h = ge_pow2(pc_h);
  while (2 < pc_w || 2 < pc_h) {
    /* std::ceil(pc_w / Map_Scale_Factor) with an unsigned divisor: integer division */
    int hw = (int)ceil((double)((unsigned)pc_w / 2u)), hh = (int)ceil((double)((unsigned)pc_h / 2u));
    pc_w = MAXD(2, hw); pc_h = MAXD(2, hh);
    pc_scale *= 2;
    insert_before_last(p, orc_map_create(pc_w, pc_h, pc_scale, p->model, p->grow, p->unknown));
  }
}

orc_pyramid *orc_pyramid_create(int w, int h, double scale, int model, int grow, const double *unknown_rec, int oie) {
  orc_pyramid *p = (orc_pyramid *)calloc(1, sizeof(orc_pyramid));
  p->cap = 8; p->lv = (orc_map **)malloc(sizeof(orc_map *) * p->cap);
  p->oie = oie; p->model = model; p->grow = grow;
  if (unknown_rec) memcpy(p->unknown, unknown_rec, sizeof(double) * orc_model_stride(model));
  else orc_default_unknown(model, p->unknown);
  p->lv[0] = orc_map_create(w, h, scale, model, grow, p->unknown);
  p->lv[1] = orc_map_create(1, 1, INFINITY, model, grow, p->unknown);
  p->n = 2;
  ensure_continuous(p);
  return p;
}
void orc_pyramid_destroy(orc_pyramid *p) {
  if (!p) return;
  for (int i = 0; i < p->n; ++i) orc_map_destroy(p->lv[i]);
  free(p->lv); free(p);
}
int orc_pyramid_levels(orc_pyramid *p) { ensure_continuous(p); return p->n; }
orc_map *orc_pyramid_level(orc_pyramid *p, int level) { return p->lv[level]; }
/* rescale :79-91 */
int orc_pyramid_rescale(orc_pyramid *p, double target_scale) {
  ensure_continuous(p);
  int id = 0;
  while (!(target_scale <= p->lv[id]->scale)) ++id;
  return id;
}

/* update :100-105 -> post_area_update / update_coarser_maps m3rsm_engine.h:89-126 */
void orc_pyramid_update(orc_pyramid *p, int x, int y, int aoo_is_occ, double ap, double aq, double obx, double oby,
                        double quality) {
  orc_map *fine = p->lv[0];
  map_update(fine, x, y, aoo_is_occ, ap, aq, obx, oby, quality);
  double rec[ORC_MAX_STRIDE];
  memcpy(rec, orc_map_at(fine, x, y), sizeof(double) * fine->stride);
  double impact = orc_cell_impact(p->model, p->oie, rec, 0, 0);
  double s = fine->scale;
  double bot = s * y, top = s * (y + 1), left = s * x, right = s * (x + 1);
  for (int id = 1; id <= orc_pyramid_levels(p) - 1; ++id) {
    orc_map *cm = p->lv[id];
    int updated = 0;
    int32_t lbrt[4];
    int cnt = orc_rasterize_rect(cm->scale, cm->w, cm->h, cm->ox, cm->oy, bot, top, left, right, 0, lbrt);
    if (cnt > 0)
      for (int cx = lbrt[0]; cx <= lbrt[2]; ++cx)
        for (int cy = lbrt[1]; cy <= lbrt[3]; ++cy) {
          const double *cr = orc_map_at(cm, cx, cy);
          /* is_unknown of a record */
          int unknown;
          switch (p->model) {
          case ORC_CELL_LWW: unknown = cr[2] == 0; break;
          case ORC_CELL_AFFINE: unknown = cr[1] == 0; break;
          case ORC_CELL_MEAN: unknown = cr[1] == 0; break;
          case ORC_CELL_GMAPPING: unknown = cr[4] == 0; break;
          default: unknown = cr[5] == 0; break;
          }
          double c_impact = orc_cell_impact(p->model, p->oie, cr, 0, 0);
          if (!unknown && orc_less_or_equal(impact, c_impact)) continue;
          orc_map_reset_cell(cm, cx, cy, rec);
          updated = 1;
        }
    if (!updated) break;
  }
}

static void pyr_update(void *ctx, int x, int y, int is_occ, double p, double q, double ox, double oy, double quality) {
  orc_pyramid_update((orc_pyramid *)ctx, x, y, is_occ, p, q, ox, oy, quality);
}
int64_t orc_pyramid_append_scan(orc_pyramid *p, const orc_scan *s, double px, double py, double pth, double scan_quality,
                                double scan_margin, const orc_estimator *est, double blur, double max_range, int omqe) {
  return append_scan_impl(pyr_update, p, p->lv[0]->scale, s, px, py, pth, scan_quality, scan_margin, est, blur,
                          max_range, omqe, NULL, 0);
}

/* Match::Match m3rsm_engine.h:156-180 -- upper bound of the scan probability for
 * (rotation, translation window); `s` is the scan the Match holds (pre-rotated
 * Cartesian when spe->prerotated) */
double orc_match_bound(orc_pyramid *p, const orc_scan *s, const orc_spe_params *spe, double px, double py, double pth,
                       double rotation, double wbot, double wtop, double wleft, double wright) {
  double vside = wtop - wbot, hside = wright - wleft;
  double cx = wleft + hside / 2, cy = wbot + vside / 2; /* LightWeightRectangle::center :191-193 */
  double dx, dy, dth;
  if (spe->prerotated) { dx = px + cx; dy = py + cy; dth = 0; }
  else { dx = px + cx; dy = py + cy; dth = pth + rotation; }
  double target = MAXD(vside, hside);
  int id = orc_pyramid_rescale(p, target);
  orc_spe_params q = *spe;
  q.win_v = vside; q.win_h = hside;
  return orc_scan_probability(p->lv[id], s, &q, dx, dy, dth, NULL);
}
