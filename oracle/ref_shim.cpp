/*
 * ref_shim.cpp -- C entry points over the UNMODIFIED reference headers, included
 * in place from /root/reference (never copied).  Built by oracle/Makefile into
 * oracle/_ref/libslamref.so (git-ignored; it does travel to the GPU box).
 *
 * TEST INFRASTRUCTURE ONLY: used to validate the plain-C restatement
 * (slam_oracle.h), to generate the golden fixtures under tests/golden/ and as the
 * "reference" CPU baseline of bench.py.  The product never loads it.
 *
 * The exported functions mirror slam_oracle.h one to one (prefix ref_), using
 * the same dense record layout, so a test can call both with identical inputs.
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <ostream>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

/* test-only access to cell internals (Mean::_n, Gmapping hits/tries/obst, map levels) */
#define private public
#define protected public
#include "src/core/maps/plain_grid_map.h"
#include "src/core/maps/lazy_tiled_grid_map.h"
#include "src/core/maps/naive_grid_cells.h"
#include "src/core/maps/tbm_grid_cells.h"
#include "src/core/maps/const_occupancy_estimator.h"
#include "src/core/maps/area_occupancy_estimator.h"
#include "src/core/maps/grid_map_scan_adders.h"
#include "src/core/maps/rescalable_caching_grid_map.h"
#include "src/core/scan_matchers/observation_impact_estimators.h"
#include "src/core/scan_matchers/occupancy_observation_probability.h"
#include "src/core/scan_matchers/weighted_mean_point_probability_spe.h"
#include "src/core/scan_matchers/brute_force_scan_matcher.h"
#include "src/core/scan_matchers/monte_carlo_scan_matcher.h"
#include "src/core/scan_matchers/hill_climbing_scan_matcher.h"
#include "src/core/scan_matchers/m3rsm_engine.h"
#include "src/core/scan_matchers/bf_multi_res_scan_matcher.h"
#include "src/slams/gmapping/gmapping_grid_cell.h"
#include "src/slams/gmapping/gmapping_occupancy_observation_pe.h"
#include "src/slams/credibilist/grid_cell.h"
#undef private
#undef protected

#include "slam_oracle.h"

#include <csetjmp>
#include <csignal>
/* the reference keeps its asserts on (CMakeLists.txt:4, no NDEBUG); a failed assert
 * would abort the test process, so guarded entry points turn it into a status */
static sigjmp_buf g_abort_jmp;
static volatile int g_guard_active = 0;
static void on_abort(int) { if (g_guard_active) siglongjmp(g_abort_jmp, 1); }
template <class F, class G> static void ref_guarded(F body, G on_assert) {
  auto prev = std::signal(SIGABRT, on_abort);
  g_guard_active = 1;
  if (sigsetjmp(g_abort_jmp, 1) == 0) body(); else on_assert();
  g_guard_active = 0;
  std::signal(SIGABRT, prev);
}

namespace {

int shim_stride(int model) {
  static const int s[ORC_CELL_MODELS] = {3, 2, 2, 6, 6, 5, 6};
  return s[model];
}
void shim_default_unknown(int model, double *r) {
  std::memset(r, 0, sizeof(double) * ORC_MAX_STRIDE);
  switch (model) {
  case ORC_CELL_LWW: r[0] = 0.5; break;
  case ORC_CELL_AFFINE: r[0] = 0.5; break;
  case ORC_CELL_MEAN: r[0] = 0.5; break;
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN:
  case ORC_CELL_CREDIBILIST: r[0] = 0.5; r[1] = 1; r[2] = 1; break;
  case ORC_CELL_GMAPPING: r[0] = -1; break;
  }
}

std::shared_ptr<GridCell> make_cell(int model, const double *r) {
  switch (model) {
  case ORC_CELL_LWW: {
    auto c = std::make_shared<GridCell>(Occupancy{r[0], r[1]});
    c->_is_unknown = r[2] == 0;
    return c;
  }
  case ORC_CELL_AFFINE: {
    auto c = std::make_shared<AffineQualityMergeCell>();
    c->_occupancy.prob_occ = r[0];
    c->_is_unknown = r[1] == 0;
    return c;
  }
  case ORC_CELL_MEAN: {
    auto c = std::make_shared<MeanProbabilityCell>();
    c->_occupancy.prob_occ = r[0];
    c->_n = r[1];
    c->_is_unknown = r[1] == 0;
    return c;
  }
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN: {
    std::shared_ptr<TbmBaseCell> c;
    if (model == ORC_CELL_TBM_CONSISTENT) c = std::make_shared<TbmOccConsistentCell>();
    else c = std::make_shared<TbmUnknownEvenOccCell>();
    c->_occupancy = Occupancy{r[0], r[1]};
    c->_belief = TBM(r[2], r[3], r[4], 0.0);
    c->_is_unknown = r[5] == 0;
    return c;
  }
  case ORC_CELL_CREDIBILIST: {
    auto c = std::make_shared<CredibilistCell>();
    c->_occupancy = Occupancy{r[0], r[1]};
    c->_belief = TBM(r[2], r[3], r[4], 0.0);
    c->_is_unknown = r[5] == 0;
    return c;
  }
  case ORC_CELL_GMAPPING: {
    auto c = std::make_shared<GmappingBaseCell>();
    c->_occupancy.prob_occ = r[0];
    c->obst = Point2D{r[1], r[2]};
    c->_hits = int(r[3]);
    c->_tries = int(r[4]);
    c->_is_unknown = r[4] == 0;
    return c;
  }
  }
  return nullptr;
}

void export_cell(int model, const GridCell &c, double *r) {
  switch (model) {
  case ORC_CELL_LWW:
    r[0] = c.occupancy().prob_occ; r[1] = c.occupancy().estimation_quality; r[2] = c.is_unknown() ? 0 : 1; break;
  case ORC_CELL_AFFINE: r[0] = c.occupancy().prob_occ; r[1] = c.is_unknown() ? 0 : 1; break;
  case ORC_CELL_MEAN: r[0] = c.occupancy().prob_occ; r[1] = static_cast<const MeanProbabilityCell &>(c)._n; break;
  case ORC_CELL_TBM_CONSISTENT:
  case ORC_CELL_TBM_UNKNOWN_EVEN: {
    auto &t = static_cast<const TbmBaseCell &>(c);
    r[0] = c.occupancy().prob_occ; r[1] = c.occupancy().estimation_quality;
    r[2] = t._belief.unknown(); r[3] = t._belief.empty(); r[4] = t._belief.occupied();
    r[5] = c.is_unknown() ? 0 : 1;
    break;
  }
  case ORC_CELL_CREDIBILIST: {
    auto &t = static_cast<const CredibilistCell &>(c);
    r[0] = c.occupancy().prob_occ; r[1] = c.occupancy().estimation_quality;
    r[2] = t._belief.unknown(); r[3] = t._belief.empty(); r[4] = t._belief.occupied();
    r[5] = c.is_unknown() ? 0 : 1;
    break;
  }
  case ORC_CELL_GMAPPING: {
    auto &g = static_cast<const GmappingBaseCell &>(c);
    r[0] = c.occupancy().prob_occ; r[1] = g.obst.x; r[2] = g.obst.y; r[3] = g._hits; r[4] = g._tries;
    break;
  }
  }
}

struct RefMap {
  std::shared_ptr<GridMap> map;
  int model = 0, grow = 0, oie = -1; /* oie >= 0: M3RSM pyramid */
};

std::shared_ptr<GridMap> make_plain(int grow, std::shared_ptr<GridCell> proto, const GridMapParams &gp) {
  switch (grow) {
  case ORC_GROW_PLAIN: return std::make_shared<UnboundedPlainGridMap>(proto, gp);
  case ORC_GROW_TILED: return std::make_shared<UnboundedLazyTiledGridMap>(proto, gp);
  default: return std::make_shared<PlainGridMap>(proto, gp);
  }
}

std::shared_ptr<ObservationImpactEstimator> make_oie(int oie) {
  if (oie == ORC_OIE_OCCUPANCY) return std::make_shared<OccupancyOIE>();
  return std::make_shared<DiscrepancyOIE>();
}

std::shared_ptr<OccupancyObservationProbabilityEstimator> make_oope(const orc_spe_params *p) {
  auto oie = make_oie(p->oie);
  switch (p->oope) {
  case ORC_OOPE_MAX: return std::make_shared<MaxOccupancyObservationPE>(oie);
  case ORC_OOPE_MEAN: return std::make_shared<MeanOccupancyObservationPE>(oie);
  case ORC_OOPE_OVERLAP: return std::make_shared<OverlapWeightedOccupancyObservationPE>(oie);
  case ORC_OOPE_GMAPPING: return std::make_shared<GmappingOccupancyObservationPE>(p->gm_fullness_th, p->gm_window);
  default: return std::make_shared<ObstacleBasedOccupancyObservationPE>(oie);
  }
}

std::shared_ptr<ScanPointWeighting> make_spw(int spw) {
  if (spw == ORC_SPW_VINY) return std::make_shared<VinySlamSPW>();
  if (spw == ORC_SPW_AHR) return std::make_shared<AngleHistogramReciprocalSPW>();
  return std::make_shared<EvenSPW>();
}

LaserScan2D make_scan(int n, const double *range, const double *angle, const uint8_t *occ) {
  LaserScan2D s;
  s.points().reserve(n);
  for (int i = 0; i < n; ++i) s.points().push_back(ScanPoint2D::make_polar(range[i], angle[i], occ ? occ[i] != 0 : true));
  return s;
}

std::shared_ptr<CellOccupancyEstimator> make_est(const orc_estimator *e) {
  Occupancy occ{e->occ_p, e->occ_q}, empty{e->empty_p, e->empty_q};
  if (e->type == ORC_EST_AREA) return std::make_shared<AreaOccupancyEstimator>(occ, empty, e->low_qual, e->unknown_qual);
  return std::make_shared<ConstOccupancyEstimator>(occ, empty);
}

std::shared_ptr<WallDistanceBlurringScanAdder> make_adder(const orc_estimator *e, double blur, double max_range, int omqe) {
  std::shared_ptr<ObservationMappingQualityEstimator> q;
  if (omqe == ORC_OMQE_AHR) q = std::make_shared<AngleHistogramResiprocalOMQE>();
  else q = std::make_shared<IdleOMQE>();
  return WallDistanceBlurringScanAdder::builder()
      .set_occupancy_estimator(make_est(e))
      .set_observation_quality_estimator(q)
      .set_blur_distance(blur)
      .set_max_usable_range(max_range)
      .build();
}

using SPE = WeightedMeanPointProbabilitySPE;
std::shared_ptr<SPE> make_spe(const orc_spe_params *p, int spw, unsigned skip_rate, double max_range) {
  return std::make_shared<SPE>(make_oope(p), make_spw(spw), skip_rate, max_range);
}

struct CountingObserver : GridScanMatcherObserver {
  int64_t tests = 0;
  std::vector<double> *scores = nullptr;
  void on_scan_test(const RobotPose &, const LaserScan2D &, double s) override {
    ++tests;
    if (scores) scores->push_back(s);
  }
};

} // namespace

extern "C" {

/* fixes the function-static Shift_Amount of AreaOccupancyEstimator (Q10,
 * area_occupancy_estimator.h:71) to low_qual * side for this process */
void ref_init_area_shift(double low_qual, double side) {
  AreaOccupancyEstimator aoe{Occupancy{0.95, 1.0}, Occupancy{0.01, 1.0}, low_qual, 0.5};
  aoe.estimate_occupancy(Segment2D{{-10 * side, 0.3 * side}, {10 * side, 0.6 * side}}, Rectangle{0, side, 0, side}, false);
}

void *ref_map_create(int w, int h, double scale, int model, int grow, const double *unknown_rec, int pyramid_oie) {
  double rec[ORC_MAX_STRIDE];
  if (unknown_rec) std::memcpy(rec, unknown_rec, sizeof(double) * shim_stride(model));
  else shim_default_unknown(model, rec);
  auto proto = make_cell(model, rec);
  auto m = new RefMap;
  m->model = model; m->grow = grow; m->oie = pyramid_oie;
  GridMapParams gp{w, h, scale};
  if (pyramid_oie >= 0) {
    auto oie = make_oie(pyramid_oie);
    switch (grow) {
    case ORC_GROW_PLAIN: m->map = std::make_shared<M3RSMRescalableGridMap<UnboundedPlainGridMap>>(oie, proto, gp); break;
    case ORC_GROW_TILED: m->map = std::make_shared<M3RSMRescalableGridMap<UnboundedLazyTiledGridMap>>(oie, proto, gp); break;
    default: m->map = std::make_shared<M3RSMRescalableGridMap<PlainGridMap>>(oie, proto, gp); break;
    }
  } else {
    m->map = make_plain(grow, proto, gp);
  }
  return m;
}
void ref_map_destroy(void *h) { delete static_cast<RefMap *>(h); }

static GridMap &level_map(RefMap *m, int level) {
  if (m->oie < 0) return *m->map;
  switch (m->grow) {
  case ORC_GROW_PLAIN: return static_cast<M3RSMRescalableGridMap<UnboundedPlainGridMap> &>(*m->map).map(level);
  case ORC_GROW_TILED: return static_cast<M3RSMRescalableGridMap<UnboundedLazyTiledGridMap> &>(*m->map).map(level);
  default: return static_cast<M3RSMRescalableGridMap<PlainGridMap> &>(*m->map).map(level);
  }
}
int ref_map_levels(void *h) {
  auto m = static_cast<RefMap *>(h);
  if (m->oie < 0) return 1;
  switch (m->grow) {
  case ORC_GROW_PLAIN: return static_cast<M3RSMRescalableGridMap<UnboundedPlainGridMap> &>(*m->map).scales_nm();
  case ORC_GROW_TILED: return static_cast<M3RSMRescalableGridMap<UnboundedLazyTiledGridMap> &>(*m->map).scales_nm();
  default: return static_cast<M3RSMRescalableGridMap<PlainGridMap> &>(*m->map).scales_nm();
  }
}
void ref_map_info(void *h, int level, int32_t *w, int32_t *hh, double *scale, int32_t *ox, int32_t *oy) {
  auto &g = level_map(static_cast<RefMap *>(h), level);
  *w = g.width(); *hh = g.height(); *scale = g.scale(); *ox = g.origin().x; *oy = g.origin().y;
}
/* dense export, internal row-major [h][w][stride] */
void ref_map_export(void *h, int level, double *cells) {
  auto m = static_cast<RefMap *>(h);
  auto &g = level_map(m, level);
  int st = shim_stride(m->model);
  for (int y = 0; y < g.height(); ++y)
    for (int x = 0; x < g.width(); ++x)
      export_cell(m->model, g[g.internal2external({x, y})], cells + ((size_t)y * g.width() + x) * st);
}
void ref_map_import(void *h, const double *cells) {
  auto m = static_cast<RefMap *>(h);
  auto &g = *m->map;
  int st = shim_stride(m->model), w = g.width(), hh = g.height();
  auto org = g.origin();
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < w; ++x) g.reset(DiscretePoint2D{x, y} - org, *make_cell(m->model, cells + ((size_t)y * w + x) * st));
}
void ref_map_at(void *h, int x, int y, double *rec) {
  auto m = static_cast<RefMap *>(h);
  export_cell(m->model, (*m->map)[{x, y}], rec);
}
void ref_map_update(void *h, int x, int y, int is_occ, double p, double q, double obx, double oby, double quality) {
  auto m = static_cast<RefMap *>(h);
  m->map->update({x, y}, AreaOccupancyObservation{is_occ != 0, Occupancy{p, q}, Point2D{obx, oby}, quality});
}

int ref_raycast(double bx, double by, double ex, double ey, double scale, int32_t *xy, int cap) {
  RegularSquaresGrid g{100, 100, scale};
  auto cells = g.world_to_cells(Segment2D{{bx, by}, {ex, ey}});
  for (size_t i = 0; i < cells.size() && (int)i < cap; ++i) { xy[2 * i] = cells[i].x; xy[2 * i + 1] = cells[i].y; }
  return (int)cells.size();
}
int ref_bresenham(int bx, int by, int ex, int ey, int32_t *xy, int cap) {
  std::vector<DiscretePoint2D> pts = DiscreteSegment2D{{bx, by}, {ex, ey}};
  for (size_t i = 0; i < pts.size() && (int)i < cap; ++i) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  return (int)pts.size();
}
int ref_rasterize_rect(double scale, int w, int h, double bot, double top, double left, double right, int include_border,
                       int32_t *xy, int cap) {
  RegularSquaresGrid g{w, h, scale};
  auto v = GridRasterizedRectangle{g, LightWeightRectangle{bot, top, left, right}, include_border != 0}.to_vector();
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; }
  return (int)v.size();
}
double ref_rect_overlap(double ab, double at, double al, double ar, double bb, double bt, double bl, double br) {
  double r = -1; /* -1: the reference asserted on this input */
  ref_guarded([&] { r = LightWeightRectangle{ab, at, al, ar}.overlap(LightWeightRectangle{bb, bt, bl, br}); }, [&] { r = -1; });
  return r;
}
void ref_estimate_occupancy(const orc_estimator *e, double bx, double by, double ex, double ey, double cbot, double ctop,
                            double cleft, double cright, int is_occ, double *pq) {
  auto est = make_est(e);
  /* {inf, inf}: the reference asserted on this input */
  ref_guarded([&] {
    auto o = est->estimate_occupancy(Segment2D{{bx, by}, {ex, ey}}, Rectangle{cbot, ctop, cleft, cright}, is_occ != 0);
    pq[0] = o.prob_occ; pq[1] = o.estimation_quality;
  }, [&] { pq[0] = pq[1] = std::numeric_limits<double>::infinity(); });
}

void ref_cell_update(int model, double *rec, int is_occ, double p, double q, double obx, double oby, double quality) {
  auto c = make_cell(model, rec);
  *c += AreaOccupancyObservation{is_occ != 0, Occupancy{p, q}, Point2D{obx, oby}, quality};
  export_cell(model, *c, rec);
}
double ref_cell_impact(int model, int oie, const double *rec, double obx, double oby) {
  auto c = make_cell(model, rec);
  auto aoo = AreaOccupancyObservation{true, {1.0, 1.0}, {obx, oby}, 1.0};
  return make_oie(oie)->estimate_impact(*c, aoo);
}

int ref_filter_scan(void *h, int n, const double *range, const double *angle, const uint8_t *occ, double px, double py,
                    double pth, unsigned skip_rate, double max_range, int32_t *keep) {
  auto m = static_cast<RefMap *>(h);
  orc_spe_params p{};
  auto spe = make_spe(&p, ORC_SPW_EVEN, skip_rate, max_range);
  auto raw = make_scan(n, range, angle, occ);
  auto f = spe->filter_scan(raw, RobotPose{px, py, pth}, *m->map);
  int k = 0, j = 0;
  for (auto &sp : f.points()) {
    while (j < n && !(range[j] == sp.range() && angle[j] == sp.angle() && (occ ? occ[j] != 0 : true) == sp.is_occupied() &&
                      !(skip_rate && j % skip_rate)))
      ++j;
    keep[k++] = j++;
  }
  return k;
}
void ref_point_weights(int spw, int n, const double *range, const double *angle, double *w) {
  auto s = make_scan(n, range, angle, nullptr);
  auto sw = make_spw(spw);
  sw->reset(s);
  for (int i = 0; i < n; ++i) w[i] = sw->weight(s.points(), i);
}

/* filter_scan at (fx, fy, fth), then estimate_scan_probability for each pose.
 * cartesian != 0: the given scan is Cartesian (x, y) and is scored as pre-rotated
 * when p->prerotated (no filtering, weights from `spw` over the scan as given). */
void ref_score_poses(void *h, int n, const double *a, const double *b, const uint8_t *occ, const double *factor,
                     int cartesian, int spw, unsigned skip_rate, double max_range, const orc_spe_params *p, double fx,
                     double fy, double fth, const double *poses, int64_t P, double *scores, int32_t *n_filtered) {
  auto m = static_cast<RefMap *>(h);
  auto spe = make_spe(p, spw, skip_rate, max_range);
  LaserScan2D scan;
  if (cartesian) {
    for (int i = 0; i < n; ++i) scan.points().push_back(ScanPoint2D::make_cartesian({a[i], b[i]}, occ ? occ[i] != 0 : true));
    spe->_spw->reset(scan);
  } else {
    scan = spe->filter_scan(make_scan(n, a, b, occ), RobotPose{fx, fy, fth}, *m->map);
  }
  if (factor && (int)scan.points().size() == n)
    for (int i = 0; i < n; ++i) scan.points()[i].set_factor(factor[i]);
  if (n_filtered) *n_filtered = (int32_t)scan.points().size();
  ScanProbabilityEstimator::SPEParams sp;
  sp.sp_analysis_area = LightWeightRectangle{-p->win_v / 2, p->win_v / 2, -p->win_h / 2, p->win_h / 2};
  sp.scan_is_prerotated = p->prerotated != 0;
  for (int64_t k = 0; k < P; ++k)
    scores[k] = spe->estimate_scan_probability(scan, RobotPose{poses[3 * k], poses[3 * k + 1], poses[3 * k + 2]}, *m->map, sp);
}

int64_t ref_bf_enumerate(double bx, double by, double bth, double fx, double tx, double sx, double fy, double ty, double sy,
                         double ft, double tt, double st, double *poses, int64_t cap) {
  BruteForcePoseEnumerator e{fx, tx, sx, fy, ty, sy, ft, tt, st};
  RobotPose base{bx, by, bth};
  int64_t n = 0;
  while (e.has_next()) {
    auto p = e.next(base);
    if (poses && n < cap) { poses[3 * n] = p.x; poses[3 * n + 1] = p.y; poses[3 * n + 2] = p.theta; }
    ++n;
    e.feedback(false);
  }
  return n;
}

static void run_matcher(GridScanMatcher &sm, RefMap *m, int n, const double *range, const double *angle,
                        const uint8_t *occ, double ix, double iy, double ith, orc_match_result *out, double *scores,
                        int64_t scores_cap) {
  TransformedLaserScan ts;
  ts.scan = make_scan(n, range, angle, occ);
  ts.quality = 1.0;
  auto obs = std::make_shared<CountingObserver>();
  std::vector<double> sc;
  if (scores) obs->scores = &sc;
  sm.subscribe(obs);
  RobotPoseDelta d;
  out->best_prob = sm.process_scan(ts, RobotPose{ix, iy, ith}, *m->map, d);
  out->dx = d.x; out->dy = d.y; out->dth = d.theta; out->poses_tested = obs->tests;
  for (size_t i = 0; scores && i < sc.size() && (int64_t)i < scores_cap; ++i) scores[i] = sc[i];
}

void ref_match_bf(void *h, int n, const double *range, const double *angle, const uint8_t *occ, int spw,
                  const orc_spe_params *p, double ix, double iy, double ith, double fx, double tx, double sx, double fy,
                  double ty, double sy, double ft, double tt, double st, orc_match_result *out, double *scores,
                  int64_t scores_cap) {
  BruteForceScanMatcher sm{make_spe(p, spw, 0, -1), fx, tx, sx, fy, ty, sy, ft, tt, st};
  run_matcher(sm, static_cast<RefMap *>(h), n, range, angle, occ, ix, iy, ith, out, scores, scores_cap);
}
void ref_match_hc(void *h, int n, const double *range, const double *angle, const uint8_t *occ, int spw,
                  const orc_spe_params *p, double ix, double iy, double ith, unsigned fal, double tr, double rot,
                  orc_match_result *out, double *scores, int64_t scores_cap) {
  HillClimbingScanMatcher sm{make_spe(p, spw, 0, -1), fal, tr, rot};
  run_matcher(sm, static_cast<RefMap *>(h), n, range, angle, occ, ix, iy, ith, out, scores, scores_cap);
}
void ref_match_mc(void *h, int n, const double *range, const double *angle, const uint8_t *occ, int spw,
                  const orc_spe_params *p, double ix, double iy, double ith, unsigned seed, double tr, double rot,
                  unsigned fal, unsigned attempts, orc_match_result *out, double *scores, int64_t scores_cap) {
  MonteCarloScanMatcher sm{make_spe(p, spw, 0, -1), seed, tr, rot, fal, attempts};
  run_matcher(sm, static_cast<RefMap *>(h), n, range, angle, occ, ix, iy, ith, out, scores, scores_cap);
}
void ref_match_bf_m3rsm(void *h, int n, const double *range, const double *angle, const uint8_t *occ, int spw,
                        const orc_spe_params *p, double ix, double iy, double ith, double xlim, double ylim,
                        double rotlim, double ang_step, double tr_step, orc_match_result *out) {
  BruteForceMultiResolutionScanMatcher sm{make_spe(p, spw, 0, -1), xlim, ylim, rotlim, ang_step, tr_step};
  run_matcher(sm, static_cast<RefMap *>(h), n, range, angle, occ, ix, iy, ith, out, nullptr, 0);
}

int64_t ref_append_scan(void *h, int n, const double *range, const double *angle, const uint8_t *occ, double px, double py,
                        double pth, double scan_quality, double scan_margin, const orc_estimator *est, double blur,
                        double max_range, int omqe) {
  auto m = static_cast<RefMap *>(h);
  auto adder = make_adder(est, blur, max_range, omqe);
  auto scan = make_scan(n, range, angle, occ);
  adder->append_scan(*m->map, RobotPose{px, py, pth}, scan, scan_quality, scan_margin);
  return 0;
}

int ref_pyramid_rescale(void *h, double target) {
  auto m = static_cast<RefMap *>(h);
  m->map->rescale(target);
  int id = 0;
  switch (m->grow) {
  case ORC_GROW_PLAIN: id = static_cast<M3RSMRescalableGridMap<UnboundedPlainGridMap> &>(*m->map).scale_id(); break;
  case ORC_GROW_TILED: id = static_cast<M3RSMRescalableGridMap<UnboundedLazyTiledGridMap> &>(*m->map).scale_id(); break;
  default: id = static_cast<M3RSMRescalableGridMap<PlainGridMap> &>(*m->map).scale_id(); break;
  }
  m->map->rescale(0);
  return id;
}

/* Match::Match (m3rsm_engine.h:156-180) on a Cartesian pre-rotated scan (or a polar one) */
double ref_match_bound(void *h, int n, const double *a, const double *b, int cartesian, int spw, const orc_spe_params *p,
                       double px, double py, double pth, double rotation, double wbot, double wtop, double wleft,
                       double wright) {
  auto m = static_cast<RefMap *>(h);
  auto spe = make_spe(p, spw, 0, -1);
  auto scan = std::make_shared<LaserScan2D>();
  for (int i = 0; i < n; ++i)
    scan->points().push_back(cartesian ? ScanPoint2D::make_cartesian({a[i], b[i]}, true) : ScanPoint2D::make_polar(a[i], b[i], true));
  spe->_spw->reset(*scan);
  RobotPose pose{px, py, pth};
  Match mt{rotation, LightWeightRectangle{wbot, wtop, wleft, wright}, spe, scan, p->prerotated != 0, pose, *m->map};
  m->map->rescale(0);
  return mt.prob_upper_bound;
}

/* GMapping OOPE with its 1-entry cache (gmapping_occupancy_observation_pe.h:17-44) over a point list */
void ref_gmapping_point_probs(void *h, double fullness_th, int window, int n, const double *X, const double *Y, double *out) {
  auto m = static_cast<RefMap *>(h);
  GmappingOccupancyObservationPE pe{fullness_th, (unsigned)window};
  for (int i = 0; i < n; ++i)
    out[i] = pe.probability(AreaOccupancyObservation{true, {1.0, 1.0}, {X[i], Y[i]}, 1.0}, LightWeightRectangle{}, *m->map);
}

/* std::normal_distribution / mt19937 of this libstdc++ (for the RNG restatement pin) */
void ref_normal_samples(unsigned seed, double mean, double stddev, int n, double *out) {
  std::mt19937 g{seed};
  std::normal_distribution<> d{mean, stddev};
  for (int i = 0; i < n; ++i) out[i] = d(g);
}

/* brute-force candidate scoring on T std::threads: the
 * candidate list is split into contiguous chunks, each thread with private SPE,
 * scan and trig provider; the map is shared read-only. returns best index */
int64_t ref_score_poses_mt(void *h, int n, const double *range, const double *angle, const uint8_t *occ, int spw,
                           const orc_spe_params *p, double fx, double fy, double fth, const double *poses, int64_t P,
                           int threads, double *scores) {
  auto m = static_cast<RefMap *>(h);
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    th.emplace_back([=]() {
      int64_t lo = P * t / threads, hi = P * (t + 1) / threads;
      auto spe = make_spe(p, spw, 0, -1);
      auto scan = spe->filter_scan(make_scan(n, range, angle, occ), RobotPose{fx, fy, fth}, *m->map);
      scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
      for (int64_t k = lo; k < hi; ++k)
        scores[k] = spe->estimate_scan_probability(scan, RobotPose{poses[3 * k], poses[3 * k + 1], poses[3 * k + 2]}, *m->map,
                                                       ScanProbabilityEstimator::SPEParams{});
    });
  }
  for (auto &t : th) t.join();
  double best = -1;
  int64_t bi = -1;
  for (int64_t k = 0; k < P; ++k)
    if (best < scores[k]) { best = scores[k]; bi = k; }
  return bi;
}

} // extern "C"
