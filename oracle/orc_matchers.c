/*
 * orc_matchers.c -- pose enumerators and the sequential accept loop of the
 * pose-enumeration scan matchers (brute force, hill climbing, Monte Carlo),
 * including std::mt19937 and libstdc++'s std::normal_distribution.
 * TEST INFRASTRUCTURE (see slam_oracle.h).  Parity: pinned.
 */
#include "slam_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* BruteForcePoseEnumerator, scan_matchers/brute_force_scan_matcher.h:10-64.
 * FP accumulation with per-axis stop rules (Q2): x/y step while v < to, theta
 * is emitted while t <= to.  Also returns the per-axis value lists. */
int64_t orc_bf_enumerate(double bx, double by, double bth, double fx, double tx, double sx, double fy, double ty,
                         double sy, double ft, double tt, double st, double *poses, int64_t cap, double *xs,
                         int32_t *nx, double *ys, int32_t *ny, double *ts, int32_t *nt) {
  double x = fx, y = fy, t = ft;
  int64_t n = 0;
  int32_t cx = 0, cy = 0, ct = 0;
  int first_row = 1, first_plane = 1;
  while (t <= tt) {
    if (poses && n < cap) { poses[3 * n] = bx + x; poses[3 * n + 1] = by + y; poses[3 * n + 2] = bth + t; }
    if (first_row && xs) xs[cx] = bx + x;
    if (first_row) ++cx;
    if (first_plane && x == fx) { if (ys) ys[cy] = by + y; ++cy; }
    if (x == fx && y == fy) { if (ts) ts[ct] = bth + t; ++ct; }
    ++n;
    if (x < tx) { x += sx; continue; }
    x = fx; first_row = 0;
    if (y < ty) { y += sy; continue; }
    y = fy; first_plane = 0;
    t += st;
  }
  if (nx) *nx = cx;
  if (ny) *ny = cy;
  if (nt) *nt = ct;
  return n;
}

/* PoseEnumerationScanMatcher::process_scan pose_enumeration_scan_matcher.h:31-77 over a fixed list */
void orc_match_list(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy, double ith,
                    const double *poses, int64_t P, orc_match_result *out) {
  double bx = ix, by = iy, bt = ith;
  double best = orc_scan_probability(m, s, p, ix, iy, ith, NULL);
  for (int64_t k = 0; k < P; ++k) {
    double v = orc_scan_probability(m, s, p, poses[3 * k], poses[3 * k + 1], poses[3 * k + 2], NULL);
    if (best < v) { best = v; bx = poses[3 * k]; by = poses[3 * k + 1]; bt = poses[3 * k + 2]; }
  }
  out->best_prob = best; out->dx = bx - ix; out->dy = by - iy; out->dth = bt - ith; out->poses_tested = P + 1;
}

/* HillClimbingScanMatcher: FailedRoundsLimitedPoseEnumerator<Distorsion1DPoseEnumerator>,
 * hill_climbing_scan_matcher.h:10-126 (frame_rotation is dropped upstream, Q5) */
void orc_match_hill_climbing(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy,
                             double ith, unsigned max_failed_rounds, double tr0, double rot0, orc_match_result *out,
                             orc_gm_cache *cache) {
  double bx = ix, by = iy, bt = ith;
  double best = orc_scan_probability(m, s, p, ix, iy, ith, cache);
  unsigned failed_rounds = 0;
  double tr = tr0, rot = rot0;
  unsigned action = 0;
  int base_set = 0, round_failed = 1;
  double rbx = 0, rby = 0, rbt = 0;
  const double fsin = sin(0.0), fcos = cos(0.0);
  int64_t tested = 1;
  while (failed_rounds < max_failed_rounds) {
    if (!(action < 6)) {
      if (round_failed) { tr *= 0.5; rot *= 0.5; ++failed_rounds; }
      action = 0; base_set = 0; round_failed = 1;
    }
    if (!base_set) { rbx = bx; rby = by; rbt = bt; base_set = 1; }
    double x = rbx, y = rby, t = rbt;
    double dir = action % 2 ? -1 : 1;
    switch (action % 3) {
    case 0: x += fcos * dir * tr; y += fsin * dir * tr; break;
    case 1: x += -fsin * dir * tr; y += fcos * dir * tr; break;
    case 2: t += dir * rot; break;
    }
    ++action;
    double v = orc_scan_probability(m, s, p, x, y, t, cache);
    ++tested;
    int ok = best < v;
    round_failed &= !ok;
    if (!ok) continue;
    best = v; bx = x; by = y; bt = t;
  }
  out->best_prob = best; out->dx = bx - ix; out->dy = by - iy; out->dth = bt - ith; out->poses_tested = tested;
}

/* ---- std::mt19937 (ISO C++ [rand.eng.mers]) ---- */
void orc_mt_seed(orc_mt19937 *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
uint32_t orc_mt_next(orc_mt19937 *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
/* libstdc++ std::generate_canonical<double, 53>(mt19937): two 32-bit draws */
static double canonical(orc_mt19937 *g) {
  double sum = 0, tmp = 1;
  for (int k = 2; k != 0; --k) { sum += (double)orc_mt_next(g) * tmp; tmp *= 4294967296.0; }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}
/* libstdc++ std::normal_distribution<double>::operator() (Marsaglia polar, cached 2nd variate) */
double orc_normal_sample(orc_normal *d, orc_mt19937 *g) {
  double ret;
  if (d->saved_available) {
    d->saved_available = 0;
    ret = d->saved;
  } else {
    double x, y, r2;
    do {
      x = 2.0 * canonical(g) - 1.0;
      y = 2.0 * canonical(g) - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    double mult = sqrt(-2 * log(r2) / r2);
    d->saved = x * mult;
    d->saved_available = 1;
    ret = y * mult;
  }
  return ret * d->stddev + d->mean;
}

/* MonteCarloScanMatcher: GaussianPoseEnumerator, monte_carlo_scan_matcher.h:10-82 (Q4).
 * One call = a freshly constructed matcher (engine seeded with `seed`). */
void orc_match_monte_carlo(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy,
                           double ith, unsigned seed, double tr0, double rot0, unsigned fal, unsigned attempts,
                           orc_match_result *out) {
  orc_mt19937 g;
  orc_mt_seed(&g, seed);
  double bx = ix, by = iy, bt = ith;
  double best = orc_scan_probability(m, s, p, ix, iy, ith, NULL);
  unsigned failed = 0, poses_nm = 0;
  double tr = tr0, rot = rot0;
  orc_normal nx = {0, tr, 0, 0}, ny = {0, tr, 0, 0}, nt = {0, rot, 0, 0};
  int64_t tested = 1;
  while (failed < fal && poses_nm < attempts) {
    double sx = orc_normal_sample(&nx, &g);
    double sy = orc_normal_sample(&ny, &g);
    double st = orc_normal_sample(&nt, &g);
    double x = bx + sx, y = by + sy, t = bt + st;
    double v = orc_scan_probability(m, s, p, x, y, t, NULL);
    ++tested;
    int ok = best < v;
    ++poses_nm;
    if (!ok) {
      ++failed;
    } else if (!(failed <= fal / 3)) {
      failed = 0; tr = tr * 0.5; rot = rot * 0.5;
      orc_normal fx = {0, tr, 0, 0}, fr = {0, rot, 0, 0};
      nx = fx; ny = fx; nt = fr;
    }
    if (!ok) continue;
    best = v; bx = x; by = y; bt = t;
  }
  out->best_prob = best; out->dx = bx - ix; out->dy = by - iy; out->dth = bt - ith; out->poses_tested = tested;
}
