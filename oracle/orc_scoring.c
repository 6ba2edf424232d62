/*
 * orc_scoring.c -- scan filtering, point weights, occupancy-observation
 * probability estimators and the weighted-mean scan probability.
 * TEST INFRASTRUCTURE (see slam_oracle.h).  Parity: pinned.
 */
#include "slam_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD(a, b) (((a) < (b)) ? (b) : (a))

/* ScanPoint2D::range/angle/x/y, sensor_data.h:47-70 */
static double sp_range(const orc_scan *s, int i) {
  return s->cartesian ? sqrt(pow(s->a[i], 2) + pow(s->b[i], 2)) : s->a[i];
}
static double sp_angle(const orc_scan *s, int i) { return s->cartesian ? atan2(s->b[i], s->a[i]) : s->b[i]; }
static double sp_x(const orc_scan *s, int i) { return s->cartesian ? s->a[i] : s->a[i] * cos(s->b[i]); }
static double sp_y(const orc_scan *s, int i) { return s->cartesian ? s->b[i] : s->a[i] * sin(s->b[i]); }

/* AngleHistogram, src/core/features/angle_histogram.h:9-102 (20 bins) */
static double ox_angle(double bx, double by, double x, double y) {
  double d_x = x - bx, d_y = y - by;
  if (d_y == 0) return 0;
  double d_d = sqrt(d_x * d_x + d_y * d_y);
  double angle = acos(d_x / d_d);
  if (d_y < 0 && d_x != 0) angle = M_PI - angle;
  return angle;
}
void orc_angle_histogram_values(int n, const double *range, const double *angle, uint32_t *values) {
  enum { NB = 20 };
  unsigned hist[NB] = {0};
  const double step = (180 * M_PI / 180) / NB; /* deg2rad(180) / _n */
  int *bin = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int i = 1; i < n; ++i) {
    double a = ox_angle(range[i - 1] * cos(angle[i - 1]), range[i - 1] * sin(angle[i - 1]), range[i] * cos(angle[i]),
                        range[i] * sin(angle[i]));
    bin[i] = (int)(size_t)floor(a / step);
    hist[bin[i]]++;
  }
  for (int i = 0; i < n; ++i) values[i] = i == 0 ? (uint32_t)n : hist[bin[i]];
  free(bin);
}

/* EvenSPW / VinySlamSPW / AngleHistogramReciprocalSPW, weighted_mean_point_probability_spe.h:21-60 */
void orc_point_weights(int spw, int n, const double *range, const double *angle, double *w) {
  if (spw == ORC_SPW_EVEN) {
    double c = 1.0 / n;
    for (int i = 0; i < n; ++i) w[i] = c;
  } else if (spw == ORC_SPW_VINY) {
    for (int i = 0; i < n; ++i) {
      double a = angle[i];
      double wt = fabs(sin(a)) + fabs(cos(a));
      if (0.9 < fabs(cos(a))) wt = 3;
      else if (0.8 < fabs(cos(a))) wt = 2;
      w[i] = wt * sqrt(range[i]);
    }
  } else {
    uint32_t *v = (uint32_t *)malloc(sizeof(uint32_t) * (n > 0 ? n : 1));
    orc_angle_histogram_values(n, range, angle, v);
    for (int i = 0; i < n; ++i) w[i] = 1.0 / v[i];
    free(v);
  }
}

/* WeightedMeanPointProbabilitySPE::filter_scan :75-95 + should_skip_point :136-141 (polar raw scans) */
int orc_filter_scan(const orc_map *m, int n, const double *range, const double *angle, const uint8_t *occ, double px,
                    double py, double pth, unsigned skip_rate, double max_range, int32_t *keep) {
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (skip_rate && (unsigned)i % skip_rate) continue;
    double wx = px + range[i] * cos(pth + angle[i]);
    double wy = py + range[i] * sin(pth + angle[i]);
    int cx = orc_world_to_cell(wx, m->scale), cy = orc_world_to_cell(wy, m->scale);
    int skip = !occ[i] || !orc_map_has_cell(m, cx, cy) || (orc_less(0.0, max_range) && orc_less(max_range, range[i]));
    if (skip) continue;
    keep[k++] = i;
  }
  return k;
}

/* OOPEs: occupancy_observation_probability.h:12-99; GMapping: slams/gmapping/gmapping_occupancy_observation_pe.h:17-38 */
double orc_point_probability(const orc_map *m, const orc_spe_params *p, double X, double Y, orc_gm_cache *cache) {
  const double s = m->scale;
  if (p->oope == ORC_OOPE_OBSTACLE) {
    int cx = orc_world_to_cell(X, s), cy = orc_world_to_cell(Y, s);
    return orc_cell_impact(m->model, p->oie, orc_map_at(m, cx, cy), X, Y);
  }
  if (p->oope == ORC_OOPE_GMAPPING) {
    int cx = orc_world_to_cell(X, s), cy = orc_world_to_cell(Y, s);
    if (cache && cx == cache->cx && cy == cache->cy && cache->prob != -1) return cache->prob;
    double best = 0;
    for (int dx = -p->gm_window; dx <= p->gm_window; ++dx)
      for (int dy = -p->gm_window; dy <= p->gm_window; ++dy) {
        const double *r = orc_map_at(m, cx + dx, cy + dy);
        if (r[0] < p->gm_fullness_th) continue;
        double v = 1.0 - orc_cell_discrepancy(m->model, r, 1.0, 1.0, X, Y, 1.0);
        best = MAXD(best, v);
      }
    if (cache) { cache->cx = cx; cache->cy = cy; cache->prob = best; }
    return best;
  }
  /* window = sp_analysis_area.move_center(obstacle), geometry_primitives.h:205-209 */
  double half_v = p->win_v / 2, half_h = p->win_h / 2;
  double bot = Y - half_v, top = Y + half_v, left = X - half_h, right = X + half_h;
  int32_t lbrt[4];
  int cnt = orc_rasterize_rect(s, m->w, m->h, m->ox, m->oy, bot, top, left, right, 1, lbrt);
  double acc = 0, wsum = 0;
  unsigned nm = 0;
  if (cnt > 0)
    for (int x = lbrt[0]; x <= lbrt[2]; ++x)
      for (int y = lbrt[1]; y <= lbrt[3]; ++y) {
        double impact = orc_cell_impact(m->model, p->oie, orc_map_at(m, x, y), X, Y);
        if (p->oope == ORC_OOPE_MAX) {
          acc = MAXD(impact, acc);
        } else if (p->oope == ORC_OOPE_MEAN) {
          acc += impact; nm += 1;
        } else { /* overlap; world_cell_bounds regular_squares_grid.h:108-118 */
          double cb, ct, cl, cr;
          if (s == INFINITY) { cb = cl = -INFINITY; ct = cr = INFINITY; }
          else { cb = s * y; ct = s * (y + 1); cl = s * x; cr = s * (x + 1); }
          double wgt = orc_rect_overlap(bot, top, left, right, cb, ct, cl, cr);
          acc += impact * wgt; wsum += wgt;
        }
      }
  if (p->oope == ORC_OOPE_MAX) return acc;
  if (p->oope == ORC_OOPE_MEAN) return nm ? acc / nm : 0.5;
  return wsum ? acc / wsum : 0.5;
}

/* WeightedMeanPointProbabilitySPE::estimate_scan_probability :97-133;
 * ScanPoint2D::move_origin sensor_data.h:83-87,103-105; RawTrigonometryProvider trigonometry_utils.h:21-27 */
double orc_scan_probability(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double px, double py,
                            double pth, orc_gm_cache *cache) {
  double total_weight = 0, total_probability = 0;
  for (int i = 0; i < s->n; ++i) {
    double X, Y;
    if (p->prerotated) {
      X = sp_x(s, i) + px; Y = sp_y(s, i) + py;
    } else {
      double r = sp_range(s, i), a = sp_angle(s, i);
      X = px + r * cos(pth + a); Y = py + r * sin(pth + a);
    }
    double prob = orc_point_probability(m, p, X, Y, cache);
    double w = s->weight[i];
    double f = s->factor ? s->factor[i] : 1.0;
    total_probability += prob * w * f;
    total_weight += w;
  }
  if (total_weight == 0) return NAN;
  return total_probability / total_weight;
}

void orc_score_poses(const orc_map *m, const orc_scan *s, const orc_spe_params *p, const double *poses, int64_t P,
                     double *scores, orc_gm_cache *cache) {
  for (int64_t k = 0; k < P; ++k)
    scores[k] = orc_scan_probability(m, s, p, poses[3 * k], poses[3 * k + 1], poses[3 * k + 2], cache);
}

/* the accept loop of PoseEnumerationScanMatcher::process_scan :48-65 over a fixed
 * candidate list: strict '<', so the lowest index wins ties; -1 = keep the initial pose */
int64_t orc_argbest(const double *scores, int64_t P, double init_score, double *best_score) {
  double best = init_score;
  int64_t idx = -1;
  for (int64_t k = 0; k < P; ++k)
    if (best < scores[k]) { best = scores[k]; idx = k; }
  if (best_score) *best_score = best;
  return idx;
}
