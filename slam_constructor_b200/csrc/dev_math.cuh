// dev_math.cuh -- device-side arithmetic shared by the kernels.
//
// Everything here is IEEE double in the reference's operation order.  The library is
// compiled with -fmad=false, and the parity-critical expressions additionally use the
// explicit round-to-nearest intrinsics so no FMA contraction can ever be introduced.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/slamgpu.h"

#define SG_DEV __device__ __forceinline__
#define SG_HD __host__ __device__ __forceinline__

namespace sg {

SG_DEV double add(double a, double b) { return __dadd_rn(a, b); }
SG_DEV double sub(double a, double b) { return __dsub_rn(a, b); }
SG_DEV double mul(double a, double b) { return __dmul_rn(a, b); }
SG_DEV double div(double a, double b) { return __ddiv_rn(a, b); }

// Division inside the TBM update chain.  A belief mass that is observed "empty" thousands of times decays to the
// smallest subnormals and stays there for good, and __ddiv_rn leaves its fast path for such numerators (measured on
// B200: 441 cycles against 112, and the slow path is a call, so the chain's independent divisions queue up).  The
// absorbing state has an exact shortcut: a subnormal a is k units of 2^-1074 (k integer), and a / b rounds back to
// k units whenever |k/b - k| < 1/2, which k*|1 - b| < 1/4 with b >= 1/2 guarantees.  There b is a sum of masses
// within rounding of 1 and k is 1 or 2, so the shortcut always fires; everything else takes the real division.
SG_DEV unsigned biased_exponent(double a) { return ((unsigned)__double2hiint(a) >> 20) & 0x7ffu; }
SG_DEV bool is_tiny(double a) { return a != 0.0 && biased_exponent(a) < 123u; }  // 0 < |a| < 2^-900, subnormals included
SG_DEV double div_chain(double a, double b) {
  const unsigned ea = biased_exponent(a);
  if (ea < 123u) {
    if (ea == 0u && b >= 0.5) {                          // a is zero or subnormal: k units of 2^-1074
      const double k = (fabs(a) * 0x1p537) * 0x1p537;    // exact
      if (fabs(b - 1.0) * k < 0.25) return a;
    }
    // a tiny but normal quotient: scale the numerator up (exact), divide on the fast path, scale back (exact, because
    // the result is a normal number)
    const double aa = fabs(a), ab = fabs(b);
    if (ab >= 0x1p-60 && ab <= 0x1p60) {
      const double a600 = aa * 0x1p600;
      if (a600 >= ab * 0x1p-422) {
        const double q = __ddiv_rn(a600, ab) * 0x1p-600;
        return ((a < 0) != (b < 0)) ? -q : q;
      }
    }
  }
  return __ddiv_rn(a, b);
}
SG_DEV double maxd(double a, double b) { return (a < b) ? b : a; }  // std::max(a, b)
SG_DEV double mind(double a, double b) { return (b < a) ? b : a; }  // std::min(a, b)

// src/core/math_utils.h:10-51
SG_DEV bool are_equal(double a, double b) {
  double sc = maxd(1.0, maxd(fabs(a), fabs(b)));
  return fabs(sub(a, b)) <= mul(1e-7, sc);
}
SG_DEV bool less(double a, double b) { return a < add(b, 2.220446049250313e-16); }
SG_DEV bool less_or_equal(double a, double b) { return are_equal(a, b) || less(a, b); }
SG_DEV bool are_ordered(double a, double b, double c) { return less_or_equal(a, b) && less_or_equal(b, c); }

// RegularSquaresGrid::world_to_cell, src/core/maps/regular_squares_grid.h:40-46
SG_DEV int world_to_cell(double v, double scale) { return (int)floor(div(v, scale)); }

// floor(v / scale) with the reference's rounding (a correctly rounded division, then floor), without
// paying for the division: floor(v * (1/scale)) is checked against both cell borders with exact FMA
// residuals, and only a point within a few ulps of a border takes the real division.
SG_DEV double floor_div(double v, double scale, double inv_scale) {
  double f = floor(v * inv_scale);
  const double lo = fma(-f, scale, v);        // v - f*scale, one rounding
  const double hi = fma(f + 1.0, scale, -v);  // (f+1)*scale - v
  const double margin = fabs(v) * 8.9e-16 + 1e-300;  // 4 ulp(v): RN(v/scale) cannot cross a border farther than this
  if (!(lo >= margin && hi >= margin)) f = floor(div(v, scale));
  return f;
}

// world_to_cell plus the trig guard: `slack` is an upper bound of |v - v_reference|
// when v was built from device trigonometry; returns true if the cell could differ.
SG_DEV int world_to_cell_guard(double v, double scale, double slack, bool *unsafe) {
  double q = div(v, scale);
  double f = floor(q);
  double lo = q - f;           // distance to the lower border, in cells
  double hi = (f + 1.0) - q;   // distance to the upper border
  double tol = slack / scale + 4.0 * 2.220446049250313e-16 * fabs(q);
  *unsafe = (lo <= tol) || (hi <= tol);
  f = f < -1e9 ? -1e9 : (f > 1e9 ? 1e9 : f);
  return (int)f;
}

// ---------------------------------------------------------------- cell models
SG_HD int model_stride(int model) {
  switch (model) {
    case SLAMGPU_CELL_LWW: return 3;
    case SLAMGPU_CELL_AFFINE: return 2;
    case SLAMGPU_CELL_MEAN: return 2;
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN:
    case SLAMGPU_CELL_CREDIBILIST: return 6;
    case SLAMGPU_CELL_GMAPPING: return 5;
  }
  return 0;
}

SG_DEV bool rec_is_unknown(int model, const double *r) {
  switch (model) {
    case SLAMGPU_CELL_LWW: return r[2] == 0;
    case SLAMGPU_CELL_AFFINE: return r[1] == 0;
    case SLAMGPU_CELL_MEAN: return r[1] == 0;
    case SLAMGPU_CELL_GMAPPING: return r[4] == 0;
    default: return r[5] == 0;
  }
}

// TBM belief {unknown, empty, occupied, conflict}: src/core/maps/transferable_belief_model.h:63-143
struct Tbm { double u, e, o, c; };
// CAREFUL selects the divisions with the exact shortcuts for zero / subnormal numerators; the cell update decides once
// per update whether any mass is in that range, so ordinary cells pay nothing for it.  With a divisor within 2^-5 of
// one (a sum of masses), a numerator of at most 8 units of 2^-1074 divides to itself: |k/b - k| = k|1 - b|/b < 1/2.
// That is one predicate per divisor and an integer compare per numerator; anything else goes to div_chain.
SG_DEV bool near_one(double b) { return fabs(b - 1.0) < 0x1p-5; }
SG_DEV bool few_units(double a) { return (unsigned long long)__double_as_longlong(a) <= 8ull; }  // +0 .. 8 * 2^-1074
template <bool CAREFUL>
SG_DEV double tbm_div(double a, double b, bool b_near_one) {
  if (!CAREFUL) return __ddiv_rn(a, b);
  if (b_near_one && few_units(a)) return a;
  return div_chain(a, b);
}
template <bool CAREFUL = false>
SG_DEV Tbm tbm_conj(const Tbm &l, const Tbm &r) {
  // t[i|j] += l[i]*r[j] over i, j in {0:u, 1:e, 2:o, 3:c}, accumulated in (i, j) order
  double lb[4] = {l.u, l.e, l.o, l.c}, rb[4] = {r.u, r.e, r.o, r.c};
  double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) t[i | j] = add(t[i | j], mul(lb[i], rb[j]));
  double tot = add(add(add(t[0], t[1]), t[2]), t[3]);
  Tbm out;
  if (tot == 0.0) { out.u = 1.0; out.e = out.o = out.c = 0.0; return out; }
  const bool n1 = CAREFUL && near_one(tot);
  out.u = tbm_div<CAREFUL>(t[0], tot, n1); out.e = tbm_div<CAREFUL>(t[1], tot, n1);
  out.o = tbm_div<CAREFUL>(t[2], tot, n1); out.c = tbm_div<CAREFUL>(t[3], tot, n1);
  return out;
}
template <bool CAREFUL = false>
SG_DEV void tbm_norm_conflict(Tbm &t) {
  double w = add(add(t.u, t.e), t.o);
  if (w == 0.0) { t.u = 1.0; t.e = t.o = t.c = 0.0; return; }
  const bool n1 = CAREFUL && near_one(w);
  t.u = tbm_div<CAREFUL>(t.u, w, n1); t.e = tbm_div<CAREFUL>(t.e, w, n1); t.o = tbm_div<CAREFUL>(t.o, w, n1); t.c = 0.0;
}
// aoo2tbm, src/core/maps/tbm_grid_cells.h:57-66
SG_DEV Tbm aoo2tbm(double p, double q, double quality) {
  Tbm t;
  if (isnan(p) || isnan(q)) { t.u = 1.0; t.e = t.o = t.c = 0.0; return t; }
  double est = mul(q, quality);
  double occupied = mul(p, est), empty = mul(sub(1.0, p), est);
  t.u = sub(sub(1.0, occupied), empty); t.e = empty; t.o = occupied; t.c = 0.0;
  return t;
}

// operator+= of each cell model: grid_cell.h:27-30, naive_grid_cells.h:14-20,33-40,
// tbm_grid_cells.h:12-19,93-106, slams/gmapping/gmapping_grid_cell.h:20-33
SG_DEV void cell_update(int model, double *r, double p, double q, double obx, double oby, double quality) {
  bool valid = !isnan(p) && !isnan(q);
  switch (model) {
    case SLAMGPU_CELL_LWW: r[0] = p; r[1] = q; r[2] = 1; break;
    case SLAMGPU_CELL_AFFINE:
      if (!valid) return;
      r[0] = add(mul(sub(1.0, quality), r[0]), mul(quality, p));
      r[1] = 1;
      break;
    case SLAMGPU_CELL_MEAN: {
      if (!valid) return;
      r[1] = add(r[1], 1.0);
      double that_p = add(0.5, mul(sub(p, 0.5), quality));
      r[0] = div(add(mul(r[0], sub(r[1], 1.0)), that_p), r[1]);
      break;
    }
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN:
    case SLAMGPU_CELL_CREDIBILIST: {  // CredibilistCell::operator+= is TbmUnknownEvenOccCell's arithmetic (grid_cell.h:23-29, TBM_prob_conversion.h:8-21)
      if (!valid) return;
      Tbm b = {r[2], r[3], r[4], 0.0}, m = aoo2tbm(p, q, quality);
      // a tiny or subnormal mass anywhere: take the divisions with the exact shortcuts (see div_chain)
      const bool careful = is_tiny(b.u) || is_tiny(b.e) || is_tiny(b.o);
      if (careful) { b = tbm_conj<true>(b, m); tbm_norm_conflict<true>(b); }
      else { b = tbm_conj<false>(b, m); tbm_norm_conflict<false>(b); }
      r[2] = b.u; r[3] = b.e; r[4] = b.o;
      if (model == SLAMGPU_CELL_TBM_CONSISTENT) {
        double qual = add(b.o, b.e);
        r[0] = careful ? tbm_div<true>(b.o, qual, near_one(qual)) : div(b.o, qual); r[1] = qual;
      } else {
        r[0] = add(b.o, mul(0.5, b.u)); r[1] = 1.0;
      }
      r[5] = 1;
      break;
    }
    case SLAMGPU_CELL_GMAPPING: {
      if (!valid) return;
      r[4] = add(r[4], 1.0);  // tries
      bool free_ = p <= 0.5;
      double aoo_p = free_ ? 0.0 : p;
      r[0] = div(add(mul(r[0], sub(r[4], 1.0)), aoo_p), r[4]);
      if (free_) return;
      r[3] = add(r[3], 1.0);  // hits
      r[1] = div(add(mul(r[1], sub(r[3], 1.0)), obx), r[3]);
      r[2] = div(add(mul(r[2], sub(r[3], 1.0)), oby), r[3]);
      break;
    }
  }
}

// discrepancy against the obstacle AOO {true, {1, 1}, obst, 1}:
// grid_cell.h:33-35, tbm_grid_cells.h:21-35, gmapping_grid_cell.h:35-38
SG_DEV double cell_discrepancy_obstacle(int model, const double *r, double obx, double oby) {
  switch (model) {
    case SLAMGPU_CELL_LWW:
    case SLAMGPU_CELL_AFFINE:
    case SLAMGPU_CELL_MEAN: return fabs(sub(r[0], 1.0));
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN: {
      Tbm that = aoo2tbm(1.0, 1.0, 1.0), b = {r[2], r[3], r[4], 0.0};
      double total_unknown = add(that.u, b.u);
      double d_occ = fabs(sub(that.o, b.o));
      Tbm comb = tbm_conj(that, b);
      double unknown = div(total_unknown, 2.0);
      double known = sub(1.0, unknown);
      double known_disc = div(mul(known, add(comb.c, d_occ)), 2.0);
      return add(div(unknown, 2.0), known_disc);
    }
    case SLAMGPU_CELL_CREDIBILIST: {
      // 1 - score(aoo): disjunctive(AOO_to_TBM(aoo), belief).occupied(), src/slams/credibilist/grid_cell.h:31-40,
      // transferable_belief_model.h:145-162 -- products accumulated in (this, that) order into t[this & that], then normalised
      const Tbm that = aoo2tbm(1.0, 1.0, 1.0);
      const double lb[4] = {that.u, that.e, that.o, that.c}, rb[4] = {r[2], r[3], r[4], 0.0};
      double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) t[i & j] = add(t[i & j], mul(lb[i], rb[j]));
      const double tot = add(add(add(t[0], t[1]), t[2]), t[3]);
      const double occ = tot == 0.0 ? 0.0 : div(t[2], tot);
      return sub(1.0, occ);
    }
    case SLAMGPU_CELL_GMAPPING: {
      double dx = sub(r[1], obx), dy = sub(r[2], oby);
      double d = add(mul(dx, dx), mul(dy, dy));
      double sim = exp(div(-d, 0.05));
      return sub(1.0, sim);
    }
  }
  return 0;
}

// observation_impact_estimators.h:14-28
SG_DEV double cell_impact(int model, int oie, const double *r, double obx, double oby) {
  if (oie == SLAMGPU_OIE_OCCUPANCY) return r[0];
  return sub(1.0, cell_discrepancy_obstacle(model, r, obx, oby));
}

}  // namespace sg
