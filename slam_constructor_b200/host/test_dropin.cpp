// test_dropin.cpp -- the reference's own world classes with the CUDA plug-ins swapped in, run side
// by side with the unmodified CPU classes on the same synthetic scans; every pose and every map cell
// must be identical.  Built here against the reference headers (host/Makefile), run on the GPU box by
// tests/test_gpu_dropin.py.
#include <chrono>
#include <cstdio>
#include <functional>
#include <random>
#include <unordered_set>

#include "slamgpu_init.h"
#include "src/utils/init_slam.h"
#include "src/core/maps/plain_grid_map.h"
#include "src/core/maps/rescalable_caching_grid_map.h"
#include "src/core/scan_matchers/m3rsm_engine.h"
#include "src/core/states/single_state_hypothesis_laser_scan_grid_world.h"

namespace {

int g_failed = 0;
#define CHECK(cond, ...)                                      \
  do {                                                        \
    if (!(cond)) {                                            \
      ++g_failed;                                             \
      std::printf("FAIL %s:%d %s | ", __FILE__, __LINE__, #cond); \
      std::printf(__VA_ARGS__);                               \
      std::printf("\n");                                      \
    }                                                         \
  } while (0)

// a laser scan of a rectangular room (|x| < hw, |y| < hh) taken from `pose`
LaserScan2D room_scan(const RobotPose &pose, int n, double fov, double hw, double hh, std::mt19937 &rng, double noise) {
  LaserScan2D scan;
  std::normal_distribution<double> nd(0.0, noise);
  for (int i = 0; i < n; ++i) {
    double a = -fov / 2 + fov * i / n;
    double c = std::cos(a + pose.theta), s = std::sin(a + pose.theta);
    double tx = c > 0 ? (hw - pose.x) / c : (c < 0 ? (-hw - pose.x) / c : 1e300);
    double ty = s > 0 ? (hh - pose.y) / s : (s < 0 ? (-hh - pose.y) / s : 1e300);
    double r = std::min(tx, ty) + nd(rng);
    bool occ = (i % 37) != 5;  // a few "no return" beams
    scan.points().push_back(ScanPoint2D::make_polar(r, a, occ));
  }
  return scan;
}

struct Config {
  const char *name;
  int map_cells; double scale;
  std::shared_ptr<GridCell> proto;
  int estimator; Occupancy occ, empty;
  double blur;
  int matcher;  // 0 MC, 1 HC, 2 BF
  int weighting; // 0 even, 1 viny
  int beams; double fov;
  int steps;
  double loc_q, raw_q;
};

std::shared_ptr<ScanPointWeighting> make_spw(int kind) {
  if (kind == 1) return std::make_shared<VinySlamSPW>();
  return std::make_shared<EvenSPW>();
}

std::shared_ptr<ScanProbabilityEstimator> make_spe(std::shared_ptr<ScanPointWeighting> spw) {
  auto oope = std::make_shared<ObstacleBasedOccupancyObservationPE>(std::make_shared<DiscrepancyOIE>());
  return std::make_shared<WeightedMeanPointProbabilitySPE>(oope, spw);
}

bool same_cells(const GridMap &a, const GridMap &b, const char *what) {
  bool ok = a.width() == b.width() && a.height() == b.height() && a.origin().x == b.origin().x && a.origin().y == b.origin().y;
  CHECK(ok, "%s: geometry %dx%d@(%d,%d) vs %dx%d@(%d,%d)", what, a.width(), a.height(), a.origin().x, a.origin().y, b.width(),
        b.height(), b.origin().x, b.origin().y);
  if (!ok) return false;
  long bad = 0, known = 0;
  auto org = a.origin();
  for (int y = 0; y < a.height(); ++y)
    for (int x = 0; x < a.width(); ++x) {
      GridMap::Coord c{x - org.x, y - org.y};
      const auto &ca = a[c];
      Occupancy oa = ca.occupancy();
      bool ua = ca.is_unknown();
      const auto &cb = b[c];
      Occupancy obb = cb.occupancy();
      known += !ua;
      if (ua != cb.is_unknown() || oa.prob_occ != obb.prob_occ || oa.estimation_quality != obb.estimation_quality) {
        if (bad++ < 3) std::printf("  cell (%d,%d): %.17g/%.17g/%d vs %.17g/%.17g/%d\n", c.x, c.y, oa.prob_occ, oa.estimation_quality,
                                   (int)ua, obb.prob_occ, obb.estimation_quality, (int)cb.is_unknown());
      }
    }
  CHECK(bad == 0, "%s: %ld cells differ", what, bad);
  CHECK(known > 100, "%s: only %ld known cells", what, known);
  return bad == 0;
}

struct CountingObserver : GridScanMatcherObserver {
  long tests = 0, updates = 0;
  double last = 0;
  void on_scan_test(const RobotPose &, const LaserScan2D &, double s) override { ++tests; last = s; }
  void on_pose_update(const RobotPose &, const LaserScan2D &, double) override { ++updates; }
};

void run_world_pair(const Config &cfg_in, std::shared_ptr<slamgpu::Context> ctx) {
  Config cfg = cfg_in;
  if (const char *soak = std::getenv("SLAMGPU_SOAK_STEPS")) cfg.steps = std::max(cfg.steps, std::atoi(soak));  // long sequences on request
  std::printf("== %s\n", cfg.name);
  GridMapParams gmp{cfg.map_cells, cfg.map_cells, cfg.scale};
  // ---- the reference, unmodified
  auto spw_ref = make_spw(cfg.weighting);
  auto spe_ref = make_spe(spw_ref);
  SingleStateHypothesisLSGWProperties pr;
  pr.localized_scan_quality = cfg.loc_q; pr.raw_scan_quality = cfg.raw_q;
  pr.grid_map = std::make_shared<UnboundedPlainGridMap>(cfg.proto, gmp);
  if (cfg.matcher == 0) pr.gsm = std::make_shared<MonteCarloScanMatcher>(spe_ref, 42, 0.2, 0.1, 20, 100);
  else if (cfg.matcher == 1) pr.gsm = std::make_shared<HillClimbingScanMatcher>(spe_ref, 6, 0.1, 0.1);
  else pr.gsm = std::make_shared<BruteForceScanMatcher>(spe_ref, -0.3, 0.3, 0.05, -0.3, 0.3, 0.05, -0.1, 0.1, 0.02);
  std::shared_ptr<CellOccupancyEstimator> est;
  if (cfg.estimator == SLAMGPU_EST_AREA) est = std::make_shared<AreaOccupancyEstimator>(cfg.occ, cfg.empty);
  else est = std::make_shared<ConstOccupancyEstimator>(cfg.occ, cfg.empty);
  pr.gmsa = WallDistanceBlurringScanAdder::builder().set_blur_distance(cfg.blur).set_occupancy_estimator(est)
              .set_observation_quality_estimator(std::make_shared<IdleOMQE>()).set_max_usable_range(25.0).build();
  SingleStateHypothesisLaserScanGridWorld ref_world{pr};
  // ---- the same world class with the CUDA plug-ins
  auto spw_gpu = make_spw(cfg.weighting);
  auto spe_gpu = make_spe(spw_gpu);
  SingleStateHypothesisLSGWProperties pg;
  pg.localized_scan_quality = cfg.loc_q; pg.raw_scan_quality = cfg.raw_q;
  pg.grid_map = std::make_shared<slamgpu::CudaGridMap>(ctx, cfg.proto, gmp, SLAMGPU_GROW_PLAIN);
  if (cfg.matcher == 0) pg.gsm = std::make_shared<slamgpu::CudaMonteCarloScanMatcher>(ctx, spe_gpu, spw_gpu, 42, 0.2, 0.1, 20, 100);
  else if (cfg.matcher == 1) pg.gsm = std::make_shared<slamgpu::CudaHillClimbingScanMatcher>(ctx, spe_gpu, spw_gpu, 6, 0.1, 0.1);
  else pg.gsm = std::make_shared<slamgpu::CudaBruteForceScanMatcher>(ctx, spe_gpu, spw_gpu, -0.3, 0.3, 0.05, -0.3, 0.3, 0.05, -0.1, 0.1, 0.02);
  slamgpu::CudaScanAdder::Properties ap;
  ap.estimator = cfg.estimator; ap.base_occupied = cfg.occ; ap.base_empty = cfg.empty; ap.blur_distance = cfg.blur;
  ap.max_usable_range = 25.0;
  pg.gmsa = std::make_shared<slamgpu::CudaScanAdder>(ap);
  SingleStateHypothesisLaserScanGridWorld gpu_world{pg};

  auto obs_ref = std::make_shared<CountingObserver>(), obs_gpu = std::make_shared<CountingObserver>();
  ref_world.add_sm_observer(obs_ref);
  gpu_world.add_sm_observer(obs_gpu);

  std::mt19937 rng(7);
  std::normal_distribution<double> odo(0.0, 0.02), odo_t(0.0, 0.01);
  RobotPose truth{0.3, -0.2, 0.1};
  double ref_ms = 0, gpu_ms = 0;
  // both worlds start at the origin; the first scan is matched against an empty map
  RobotPose prev = truth;
  for (int step = 0; step < cfg.steps; ++step) {
    RobotPoseDelta motion = step == 0 ? RobotPoseDelta{truth.x, truth.y, truth.theta}
                                      : RobotPoseDelta{0.08 * std::cos(0.3 * step), 0.06 * std::sin(0.2 * step), 0.03};
    if (step > 0) { truth += motion; }
    RobotPoseDelta odom = step == 0 ? motion : RobotPoseDelta{motion.x + odo(rng), motion.y + odo(rng), motion.theta + odo_t(rng)};
    auto scan = room_scan(truth, cfg.beams, cfg.fov, 4.0, 3.0, rng, 0.01);
    TransformedLaserScan a{odom, scan, 1.0}, b{odom, scan, 1.0};
    b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();  // no shared mutable state between the two worlds
    auto t0 = std::chrono::steady_clock::now();
    ref_world.handle_sensor_data(a);
    auto t1 = std::chrono::steady_clock::now();
    gpu_world.handle_sensor_data(b);
    static_cast<const slamgpu::CudaGridMap &>(gpu_world.map()).flush();  // the queued scan insertion belongs to this scan
    auto t2 = std::chrono::steady_clock::now();
    if (step > 0) { ref_ms += std::chrono::duration<double, std::milli>(t1 - t0).count(); gpu_ms += std::chrono::duration<double, std::milli>(t2 - t1).count(); }
    const RobotPose &p1 = ref_world.pose(), &p2 = gpu_world.pose();
    CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "%s step %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)", cfg.name,
          step, p1.x, p1.y, p1.theta, p2.x, p2.y, p2.theta);
    CHECK(obs_ref->tests == obs_gpu->tests && obs_ref->updates == obs_gpu->updates && obs_ref->last == obs_gpu->last,
          "%s step %d: observers saw %ld/%ld tests, %ld/%ld updates", cfg.name, step, obs_ref->tests, obs_gpu->tests, obs_ref->updates,
          obs_gpu->updates);
    if (step == 1 || step + 1 == cfg.steps) same_cells(ref_world.map(), gpu_world.map(), cfg.name);
    if (g_failed > 5) return;
    prev = truth;
  }
  std::printf("   %d scans, %ld candidate poses scored on the device, final pose %.6f %.6f %.6f\n", cfg.steps, obs_gpu->tests,
              gpu_world.pose().x, gpu_world.pose().y, gpu_world.pose().theta);
  std::printf("   TIMING per scan (handle_sensor_data: match + insert, wall clock, scans 1..%d): reference CPU %.3f ms, CUDA plug-ins %.3f ms\n",
              cfg.steps - 1, ref_ms / (cfg.steps - 1), gpu_ms / (cfg.steps - 1));
}

// a CUDA matcher handed a plain host map of the reference (score-LUT snapshot path)
void run_host_map(std::shared_ptr<slamgpu::Context> ctx) {
  std::printf("== CUDA matchers on a reference UnboundedPlainGridMap\n");
  GridMapParams gmp{200, 200, 0.05};
  auto map = std::make_shared<UnboundedPlainGridMap>(std::make_shared<TbmUnknownEvenOccCell>(), gmp);
  auto est = std::make_shared<ConstOccupancyEstimator>(Occupancy{0.95, 0.3}, Occupancy{0.01, 0.1});
  auto adder = WallDistanceBlurringScanAdder::builder().set_blur_distance(0.2).set_occupancy_estimator(est)
                 .set_observation_quality_estimator(std::make_shared<IdleOMQE>()).build();
  std::mt19937 rng(11);
  RobotPose truth{0.2, 0.1, -0.2};
  for (int k = 0; k < 3; ++k) {
    auto s = room_scan(truth, 360, 2 * M_PI, 4.0, 3.0, rng, 0.01);
    adder->append_scan(*map, truth, s, 1.0);
  }
  auto scan = room_scan(truth, 181, 1.5 * M_PI, 4.0, 3.0, rng, 0.005);
  RobotPose init{truth.x + 0.07, truth.y - 0.05, truth.theta + 0.03};
  for (int which = 0; which < 3; ++which) {
    auto spw1 = make_spw(1), spw2 = make_spw(1);
    auto spe1 = make_spe(spw1), spe2 = make_spe(spw2);
    std::shared_ptr<GridScanMatcher> m1, m2;
    if (which == 0) { m1 = std::make_shared<MonteCarloScanMatcher>(spe1, 3, 0.2, 0.1, 20, 100); m2 = std::make_shared<slamgpu::CudaMonteCarloScanMatcher>(ctx, spe2, spw2, 3, 0.2, 0.1, 20, 100); }
    if (which == 1) { m1 = std::make_shared<HillClimbingScanMatcher>(spe1, 6, 0.1, 0.1); m2 = std::make_shared<slamgpu::CudaHillClimbingScanMatcher>(ctx, spe2, spw2, 6, 0.1, 0.1); }
    if (which == 2) { m1 = std::make_shared<BruteForceScanMatcher>(spe1, -0.2, 0.2, 0.02, -0.2, 0.2, 0.02, -0.06, 0.06, 0.01); m2 = std::make_shared<slamgpu::CudaBruteForceScanMatcher>(ctx, spe2, spw2, -0.2, 0.2, 0.02, -0.2, 0.2, 0.02, -0.06, 0.06, 0.01); }
    TransformedLaserScan a{RobotPoseDelta{}, scan, 1.0}, b{RobotPoseDelta{}, scan, 1.0};
    b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
    RobotPoseDelta d1, d2;
    double s1 = m1->process_scan(a, init, *map, d1), s2 = m2->process_scan(b, init, *map, d2);
    CHECK(s1 == s2 && d1.x == d2.x && d1.y == d2.y && d1.theta == d2.theta, "matcher %d: %.17g (%.17g %.17g %.17g) vs %.17g (%.17g %.17g %.17g)",
          which, s1, d1.x, d1.y, d1.theta, s2, d2.x, d2.y, d2.theta);
    std::printf("   matcher %d: best %.9f, delta %.4f %.4f %.4f\n", which, s2, d2.x, d2.y, d2.theta);
  }
}

// BF-M3RSM: the reference's M3RSMRescalableGridMap + BruteForceMultiResolutionScanMatcher world vs the CUDA pair
void run_m3rsm(std::shared_ptr<slamgpu::Context> ctx) {
  std::printf("== BF-M3RSM world: pyramid map + multi-resolution matcher\n");
  GridMapParams gmp{160, 160, 0.05};
  auto mk_spe = [](std::shared_ptr<ScanPointWeighting> spw) {
    auto oope = std::make_shared<MaxOccupancyObservationPE>(std::make_shared<DiscrepancyOIE>());
    return std::make_shared<WeightedMeanPointProbabilitySPE>(oope, spw);
  };
  auto spw1 = make_spw(0), spw2 = make_spw(0);
  SingleStateHypothesisLSGWProperties pr, pg;
  pr.localized_scan_quality = pg.localized_scan_quality = 0.9;
  pr.raw_scan_quality = pg.raw_scan_quality = 0.6;
  auto proto = std::make_shared<MeanProbabilityCell>();
  pr.grid_map = std::make_shared<M3RSMRescalableGridMap<UnboundedPlainGridMap>>(std::make_shared<DiscrepancyOIE>(), proto, gmp);
  pr.gsm = std::make_shared<BruteForceMultiResolutionScanMatcher>(mk_spe(spw1), 0.4, 0.4, deg2rad(3), deg2rad(0.5), 0.05);
  auto est = std::make_shared<ConstOccupancyEstimator>(Occupancy{0.95, 1.0}, Occupancy{0.01, 1.0});
  pr.gmsa = WallDistanceBlurringScanAdder::builder().set_blur_distance(0.3).set_occupancy_estimator(est)
              .set_observation_quality_estimator(std::make_shared<IdleOMQE>()).build();
  pg.grid_map = std::make_shared<slamgpu::CudaPyramidGridMap>(ctx, std::make_shared<DiscrepancyOIE>(), proto, gmp, SLAMGPU_GROW_PLAIN);
  auto gsm = std::make_shared<slamgpu::CudaBfMultiResScanMatcher>(ctx, mk_spe(spw2), spw2, 0.4, 0.4, deg2rad(3), deg2rad(0.5), 0.05);
  pg.gsm = gsm;
  slamgpu::CudaScanAdder::Properties ap;
  ap.blur_distance = 0.3;
  pg.gmsa = std::make_shared<slamgpu::CudaScanAdder>(ap);
  SingleStateHypothesisLaserScanGridWorld ref_world{pr}, gpu_world{pg};
  std::mt19937 rng(21);
  std::normal_distribution<double> odo(0.0, 0.03), odo_t(0.0, 0.01);
  RobotPose truth{0.1, 0.2, -0.1};
  double m3_ref = 0, m3_gpu = 0;
  for (int step = 0; step < 6; ++step) {
    RobotPoseDelta motion = step == 0 ? RobotPoseDelta{truth.x, truth.y, truth.theta} : RobotPoseDelta{0.07, -0.04, 0.02};
    if (step > 0) { truth += motion; }
    RobotPoseDelta odom = step == 0 ? motion : RobotPoseDelta{motion.x + odo(rng), motion.y + odo(rng), motion.theta + odo_t(rng)};
    auto scan = room_scan(truth, 181, 2 * M_PI, 3.0, 2.5, rng, 0.01);
    TransformedLaserScan a{odom, scan, 1.0}, b{odom, scan, 1.0};
    b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
    auto t0 = std::chrono::steady_clock::now();
    ref_world.handle_sensor_data(a);
    auto t1 = std::chrono::steady_clock::now();
    gpu_world.handle_sensor_data(b);
    auto t2 = std::chrono::steady_clock::now();
    if (step > 0) { m3_ref += std::chrono::duration<double, std::milli>(t1 - t0).count(); m3_gpu += std::chrono::duration<double, std::milli>(t2 - t1).count(); }
    const RobotPose &p1 = ref_world.pose(), &p2 = gpu_world.pose();
    CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "m3rsm step %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)", step, p1.x,
          p1.y, p1.theta, p2.x, p2.y, p2.theta);
    if (g_failed > 5) return;
  }
  ref_world.map().rescale(0);
  same_cells(ref_world.map(), gpu_world.map(), "m3rsm fine level");
  std::printf("   6 scans; last scan: %ld matches scored in %ld K5 calls, %ld branches, %ld rotations; final pose %.6f %.6f %.6f\n",
              (long)gsm->stats()[0], (long)gsm->stats()[1], (long)gsm->stats()[2], (long)gsm->stats()[3], gpu_world.pose().x,
              gpu_world.pose().y, gpu_world.pose().theta);
  std::printf("   TIMING per scan (handle_sensor_data, wall clock, scans 1..5): reference CPU %.3f ms, CUDA plug-ins %.3f ms\n", m3_ref / 5, m3_gpu / 5);
}

// the shipped presets (config/slams/tiny_slam_base.properties, viny_slam_base.properties with their includes),
// keys spelled out here because the config files do not travel to the GPU box; both factories read them
void run_presets(std::shared_ptr<slamgpu::Context> ctx) {
  struct Preset { const char *name; std::vector<std::pair<const char *, const char *>> kv; int beams; double fov; };
  std::vector<Preset> presets = {
    {"tiny_slam_base.properties",
     {{"slam/mapping/blur", "0.5"}, {"slam/occupancy_estimator/type", "const"}, {"slam/occupancy_estimator/base_occupied/prob", "0.95"},
      {"slam/occupancy_estimator/base_empty/prob", "0.01"}, {"slam/mapping/grid/area/type", "mean_probability"},
      {"slam/mapping/corrected_pose_quality", "0.9"}, {"slam/mapping/raw_pose_quality", "0.6"},
      {"slam/mapping/grid/type", "unbounded_plain"}, {"slam/map/height_in_meters", "10"}, {"slam/map/width_in_meters", "10"},
      {"slam/map/meters_per_cell", "0.1"}, {"slam/scmtch/type", "MC"}, {"slam/scmtch/MC/dispersion/translation", "0.2"},
      {"slam/scmtch/MC/dispersion/rotation", "0.1"}, {"slam/scmtch/MC/dispersion/failed_attempts_limit", "20"},
      {"slam/scmtch/MC/attempts_limit", "100"}, {"slam/scmtch/MC/seed", "42"}, {"slam/scmtch/spe/type", "wmpp"},
      {"slam/scmtch/spe/wmpp/weighting/type", "even"}}, 360, 2 * M_PI},
    {"viny_slam_base.properties",
     {{"slam/mapping/blur", "0.3"}, {"slam/occupancy_estimator/type", "const"}, {"slam/occupancy_estimator/base_occupied/prob", "0.95"},
      {"slam/occupancy_estimator/base_occupied/qual", "0.04"}, {"slam/occupancy_estimator/base_empty/prob", "0.01"},
      {"slam/occupancy_estimator/base_empty/qual", "0.003"}, {"slam/mapping/grid/area/type", "tbm_consistent"},
      {"slam/mapping/corrected_pose_quality", "0.9"}, {"slam/mapping/raw_pose_quality", "0.6"},
      {"slam/mapping/grid/type", "unbounded_plain"}, {"slam/map/height_in_meters", "10"}, {"slam/map/width_in_meters", "10"},
      {"slam/map/meters_per_cell", "0.1"}, {"slam/scmtch/type", "MC"}, {"slam/scmtch/MC/dispersion/translation", "0.2"},
      {"slam/scmtch/MC/dispersion/rotation", "0.1"}, {"slam/scmtch/MC/dispersion/failed_attempts_limit", "20"},
      {"slam/scmtch/MC/attempts_limit", "100"}, {"slam/scmtch/MC/seed", "7"}, {"slam/scmtch/spe/type", "wmpp"},
      {"slam/scmtch/spe/wmpp/weighting/type", "viny"}}, 541, 1.5 * M_PI},
    {"hill climbing + area estimator + AHR mapping quality (factory keys)",
     {{"slam/mapping/blur", "0.2"}, {"slam/occupancy_estimator/type", "area"}, {"slam/mapping/grid/area/type", "tbm_unknown_even_occ"},
      {"slam/occupancy_estimator/base_occupied/qual", "0.3"}, {"slam/occupancy_estimator/base_empty/qual", "0.1"},
      {"slam/mapping/grid/type", "unbounded_plain"}, {"slam/map/height_in_meters", "8"}, {"slam/map/width_in_meters", "8"},
      {"slam/map/meters_per_cell", "0.05"}, {"slam/scmtch/type", "HC"}, {"slam/scmtch/spe/type", "wmpp"},
      {"slam/scmtch/spe/wmpp/weighting/type", "ahr"}, {"slam/mapping/observation_quality_estimator/typetype", "ahr"},
      {"slam/mapping/max_range", "20"}}, 361, 1.5 * M_PI},
  };
  for (auto &ps : presets) {
    std::printf("== factories on %s\n", ps.name);
    MapPropertiesProvider props;
    for (auto &kv : ps.kv) props.set_property(kv.first, kv.second);
    auto ref_world = init_1h_slam(props);
    auto gpu_world = slamgpu::init_cuda_1h_slam(props, ctx);
    std::mt19937 rng(31);
    std::normal_distribution<double> odo(0.0, 0.02), odo_t(0.0, 0.01);
    RobotPose truth{0.2, -0.1, 0.15};
    for (int step = 0; step < 12; ++step) {
      RobotPoseDelta motion = step == 0 ? RobotPoseDelta{truth.x, truth.y, truth.theta} : RobotPoseDelta{0.06, 0.03 * std::cos(0.5 * step), 0.02};
      if (step > 0) { truth += motion; }
      RobotPoseDelta odom = step == 0 ? motion : RobotPoseDelta{motion.x + odo(rng), motion.y + odo(rng), motion.theta + odo_t(rng)};
      auto scan = room_scan(truth, ps.beams, ps.fov, 3.5, 3.0, rng, 0.01);
      TransformedLaserScan a{odom, scan, 1.0}, b{odom, scan, 1.0};
      b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
      ref_world->handle_sensor_data(a);
      gpu_world->handle_sensor_data(b);
      const RobotPose &p1 = ref_world->pose(), &p2 = gpu_world->pose();
      CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "%s step %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)", ps.name,
            step, p1.x, p1.y, p1.theta, p2.x, p2.y, p2.theta);
      if (g_failed > 5) return;
    }
    same_cells(ref_world->map(), gpu_world->map(), ps.name);
    std::printf("   12 scans, final pose %.6f %.6f %.6f\n", gpu_world->pose().x, gpu_world->pose().y, gpu_world->pose().theta);
  }
}

// GmappingParticleFilter with the CUDA plug-ins.  Upstream seeds every particle from std::random_device
// (gmapping_world.h:48), so two runs never agree bit for bit: both filters must simply track the truth.
void run_gmapping(std::shared_ptr<slamgpu::Context> ctx) {
  std::printf("== GMapping particle filter (reference classes, CUDA plug-ins), 8 particles\n");
  MapPropertiesProvider props;
  for (auto kv : std::vector<std::pair<const char *, const char *>>{
         {"slam/particles/number", "8"}, {"slam/map/height_in_meters", "12"}, {"slam/map/width_in_meters", "12"},
         {"slam/map/meters_per_cell", "0.05"}, {"slam/scmtch/spe/type", "wmpp"}, {"slam/scmtch/spe/wmpp/weighting/type", "even"},
         {"slam/particles/sm_delta_lim/xy/min", "0.05"}, {"slam/particles/sm_delta_lim/xy/max", "0.06"},
         {"slam/particles/sm_delta_lim/theta/min", "0.02"}, {"slam/particles/sm_delta_lim/theta/max", "0.03"},
         {"slam/particles/sample/xy/sigma", "0.02"}, {"slam/particles/sample/theta/sigma", "0.01"}})
    props.set_property(kv.first, kv.second);
  auto ref = init_gmapping(props);
  auto gpu = slamgpu::init_cuda_gmapping(props, ctx);
  std::mt19937 rng(41);
  RobotPose truth{0.0, 0.0, 0.0};
  for (int step = 0; step < 10; ++step) {
    RobotPoseDelta motion = step == 0 ? RobotPoseDelta{0, 0, 0} : RobotPoseDelta{0.10, 0.04, 0.03};
    truth += motion;
    auto scan = room_scan(truth, 360, 2 * M_PI, 3.0, 2.5, rng, 0.005);
    TransformedLaserScan a{motion, scan, 1.0}, b{motion, scan, 1.0};
    b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
    ref->handle_sensor_data(a);
    gpu->handle_sensor_data(b);
  }
  const RobotPose &p1 = ref->pose(), &p2 = gpu->pose();
  double e1 = std::hypot(p1.x - truth.x, p1.y - truth.y), e2 = std::hypot(p2.x - truth.x, p2.y - truth.y);
  std::printf("   truth %.3f %.3f %.3f | reference %.3f %.3f %.3f (err %.3f) | cuda %.3f %.3f %.3f (err %.3f)\n", truth.x, truth.y,
              truth.theta, p1.x, p1.y, p1.theta, e1, p2.x, p2.y, p2.theta, e2);
  CHECK(std::isfinite(p2.x) && std::isfinite(p2.y) && std::isfinite(p2.theta), "gmapping: pose is not finite");
  CHECK(e2 < 0.25 && std::fabs(p2.theta - truth.theta) < 0.15, "gmapping: the CUDA filter lost track (err %.3f)", e2);
  long known = 0;
  const GridMap &m = gpu->map();
  auto org = m.origin();
  for (int y = 0; y < m.height(); ++y)
    for (int x = 0; x < m.width(); ++x) known += !m[GridMap::Coord{x - org.x, y - org.y}].is_unknown();
  CHECK(known > 2000, "gmapping: only %ld known cells in the heaviest particle's map", known);
}

// ---------------------------------------------------------------------------------------------------------------
// The batched filter (slamgpu_gmapping.h): per-particle maps, lock-step hill climbing, batched insertion.
MapPropertiesProvider gmapping_props(const char *particles, bool noiseless, const char *sigma_xy = "0.03", const char *sigma_th = "0.015") {
  MapPropertiesProvider props;
  for (auto kv : std::vector<std::pair<const char *, const char *>>{
         {"slam/particles/number", particles}, {"slam/map/height_in_meters", "12"}, {"slam/map/width_in_meters", "12"},
         {"slam/map/meters_per_cell", "0.05"}, {"slam/scmtch/spe/type", "wmpp"}, {"slam/scmtch/spe/wmpp/weighting/type", "even"},
         {"slam/particles/sm_delta_lim/xy/min", "0.15"}, {"slam/particles/sm_delta_lim/xy/max", noiseless ? "0.15" : "0.25"},
         {"slam/particles/sm_delta_lim/theta/min", "0.05"}, {"slam/particles/sm_delta_lim/theta/max", noiseless ? "0.05" : "0.08"},
         {"slam/particles/sample/xy/sigma", noiseless ? "0" : sigma_xy}, {"slam/particles/sample/theta/sigma", noiseless ? "0" : sigma_th}})
    props.set_property(kv.first, kv.second);
  return props;
}

// a GmappingWorld-shaped particle built ONLY from reference components (matcher, adder, lazy tiled map), with its own
// map and an injectable seed: the sequential statement of what the batched filter must compute
// (gmapping_world.h:59-119; the real class seeds itself from std::random_device, :48)
struct SeqParticle {
  using Engine = std::mt19937;
  SeqParticle(const PropertiesProvider &props, const GMappingParams &g, std::uint32_t seed)
    : map{std::make_shared<UnboundedLazyTiledGridMap>(std::make_shared<GmappingBaseCell>(), init_grid_map_params(props))}
    , matcher{std::make_shared<HillClimbingScanMatcher>(init_gmapping_prob_estimator(props), 6, 0.1, 0.1)}
    , adder{init_scan_adder(props)}, engine(seed), guess_rv{g.pose_guess_rv}, next_rv{g.next_sm_delta_rv} {
    const PropertiesProvider *pp = &props;
    fresh_matcher = [pp] { return std::make_shared<HillClimbingScanMatcher>(init_gmapping_prob_estimator(*pp), 6, 0.1, 0.1); };
    rearm();
  }
  // `*new_particle = *sampled`: the lazy tiled map forks copy-on-write.  The copy gets its own matcher (upstream's copies
  // go on sharing the source's matcher object, and through it the OOPE's cell cache, with the particle they came from)
  SeqParticle(const SeqParticle &o)
    : map{std::make_shared<UnboundedLazyTiledGridMap>(*o.map)}, matcher{o.fresh_matcher()}, adder{o.adder}, pose{o.pose}, odom{o.odom}
    , weight{o.weight}, master{o.master}, first{o.first}, engine{o.engine}, guess_rv{o.guess_rv}, next_rv{o.next_rv}
    , since{o.since}, next{o.next} { fresh_matcher = o.fresh_matcher; }
  std::function<std::shared_ptr<GridScanMatcher>()> fresh_matcher;
  void rearm() { since.reset(); next = next_rv.sample(engine); }
  void move(const RobotPoseDelta &d) {
    double dth = (pose - odom).theta, sn = std::sin(dth), cs = std::cos(dth);
    RobotPoseDelta corrected{cs * d.x - sn * d.y, sn * d.x + cs * d.y, d.theta};
    odom += d; since += corrected.abs(); pose += corrected;
  }
  void observe(TransformedLaserScan &scan) {
    if (since.sq_dist() < next.sq_dist() && std::fabs(since.theta) < next.theta) return;
    if (!first) pose += guess_rv.sample(engine);
    RobotPoseDelta delta;
    double prob = matcher->process_scan(scan, pose, *map, delta);
    pose += delta;
    const bool was_first = first;
    if (0.0 < prob || first) { adder->append_scan(*map, pose, scan.scan, scan.quality, 0); first = false; }
    if (!was_first) weight = prob * weight;  // Properties::first_scan_keeps_weight (see slamgpu_gmapping.h)
    rearm();
  }
  void make_master() {
    using G = GaussianRV1D<Engine>;
    master = true;
    guess_rv = RobotPoseDeltaRV<Engine>{G{0, 0}, G{0, 0}, G{0, 0}};
    next_rv = RobotPoseDeltaRV<Engine>{G{0, 0}, G{0, 0}, G{0, 0}};
  }
  std::shared_ptr<UnboundedLazyTiledGridMap> map;
  std::shared_ptr<GridScanMatcher> matcher;
  std::shared_ptr<GridMapScanAdder> adder;
  RobotPose pose{0, 0, 0}, odom{0, 0, 0};
  double weight = 1.0;
  bool master = false, first = true;
  Engine engine;
  RobotPoseDeltaRV<Engine> guess_rv, next_rv;
  RobotPoseDelta since, next;
};

// GmappingParticleFilter + ParticleFilter + UniformResamling over SeqParticles, draws from `seeds`
struct SeqFilter {
  SeqFilter(const PropertiesProvider &props, unsigned n, std::function<std::uint32_t()> seeds) : seeds{seeds} {
    auto g = init_gmapping_params(props);
    for (unsigned i = 0; i < n; ++i) { ps.push_back(std::make_shared<SeqParticle>(props, g, seeds())); ps.back()->weight = 1.0 / n; }
    ps[heaviest()]->make_master();
  }
  std::size_t heaviest() const {
    std::size_t b = 0;
    for (std::size_t i = 1; i < ps.size(); ++i) if (!(ps[i]->weight < ps[b]->weight)) b = i;
    return b;
  }
  void normalize() { double t = 0; for (auto &p : ps) t += p->weight; for (auto &p : ps) p->weight = p->weight / t; }
  void step(TransformedLaserScan &scan) {
    for (auto &p : ps) p->move(scan.pose_delta);
    traversed += scan.pose_delta.abs();
    for (auto &p : ps) p->observe(scan);
    normalize();
    if (traversed.sq_dist() <= 0.5 && std::fabs(traversed.theta <= 0.2)) return;
    double sq = 0; for (auto &p : ps) sq += p->weight * p->weight;
    if (!((1.0 / sq) * 2 < ps.size())) return;
    std::vector<unsigned> inds(ps.size());
    std::mt19937 engine(seeds());
    std::uniform_real_distribution<> u(0, 1);
    for (std::size_t i = 0; i < ps.size(); ++i) {
      double sample = u(engine), tw = 0;
      for (std::size_t j = 0; j < ps.size(); ++j) { tw += ps[j]->weight; if (sample < tw) { inds[i] = (unsigned)j; break; } }
    }
    std::vector<std::shared_ptr<SeqParticle>> nxt;
    std::unordered_set<unsigned> seen;
    for (unsigned i : inds) {
      auto sp = ps[i];
      if (seen.count(i)) { sp = std::make_shared<SeqParticle>(*sp); sp->master = false; } else { seen.insert(i); }
      nxt.push_back(sp);
    }
    ps = std::move(nxt);
    normalize();
    ++resamplings;
    traversed.reset();
    bool has_master = false; for (auto &p : ps) has_master |= p->master;
    if (!has_master) ps[heaviest()]->make_master();
  }
  std::function<std::uint32_t()> seeds;
  std::vector<std::shared_ptr<SeqParticle>> ps;
  RobotPoseDelta traversed;
  std::size_t resamplings = 0;
};

bool same_as_device_map(const GridMap &ref, std::shared_ptr<slamgpu::Context> ctx, slamgpu_map *dev, const GridMapParams &gmp, const char *what) {
  slamgpu::CudaGridMap view(ctx, std::make_shared<GmappingBaseCell>(), gmp, SLAMGPU_GROW_TILED, dev);
  return same_cells(ref, view, what);
}

void run_gmapping_batched(std::shared_ptr<slamgpu::Context> ctx) {
  // ---- A: against REAL GmappingWorld objects (fresh map each).  Their engines are seeded from std::random_device, so the
  // random variables are made degenerate (sigma 0, min == max): every draw is then a constant, whatever the engine state.
  {
    std::printf("== batched GMapping filter vs GmappingWorld particles with their own maps (degenerate noise), 3 particles\n");
    auto props = gmapping_props("3", true);
    props.set_property("slam/particles/first_scan_keeps_weight", "false");  // upstream's literal weight product
    auto gp = init_gmapping_params(props);
    std::vector<std::shared_ptr<GmappingWorld>> ref;
    for (int i = 0; i < 3; ++i) {
      auto map = std::make_shared<UnboundedLazyTiledGridMap>(std::make_shared<GmappingBaseCell>(), init_grid_map_params(props));
      auto shw = SingleStateHypothesisLSGWProperties{
        1.0, 1.0, 0, map, std::make_shared<HillClimbingScanMatcher>(init_gmapping_prob_estimator(props), 6, 0.1, 0.1), init_scan_adder(props)};
      ref.push_back(std::make_shared<GmappingWorld>(shw, gp));
      ref.back()->set_weight(1.0 / 3);
    }
    ref[2]->mark_master();  // the heaviest of equal weights is the last one (particle_filter.h:114-121)
    auto gpu = slamgpu::init_cuda_gmapping_batched(props, ctx);
    CHECK(gpu->particle_is_master(2) && !gpu->particle_is_master(0), "batched filter: wrong initial master");
    std::mt19937 rng(43);
    RobotPose truth{0, 0, 0};
    long matched_total = 0;
    for (int step = 0; step < 9; ++step) {
      RobotPoseDelta motion = step == 0 ? RobotPoseDelta{0, 0, 0} : RobotPoseDelta{0.06, 0.03, 0.02};
      truth += motion;
      auto scan = room_scan(truth, 360, 2 * M_PI, 3.0, 2.5, rng, 0.005);
      TransformedLaserScan a{motion, scan, 1.0}, b{motion, scan, 1.0};
      b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
      for (auto &w : ref) { w->update_robot_pose(motion); w->handle_observation(a); }
      double tot = 0;
      for (auto &w : ref) tot += w->weight();
      for (auto &w : ref) w->set_weight(w->weight() / tot);
      gpu->handle_sensor_data(b);
      matched_total += (long)gpu->last_step().matched;
      for (int i = 0; i < 3; ++i) {
        const RobotPose &p1 = ref[i]->pose(), &p2 = gpu->particle_pose(i);
        CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "step %d particle %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)",
              step, i, p1.x, p1.y, p1.theta, p2.x, p2.y, p2.theta);
        // (every particle's first match scores 0 on its empty map: upstream's weights end as 0/0, and so must ours)
        const double w1 = ref[i]->weight(), w2 = gpu->particle_weight(i);
        CHECK((std::isnan(w1) && std::isnan(w2)) || std::fabs(w1 - w2) <= 1e-5 * w1, "step %d particle %d: weight %.17g vs %.17g", step, i, w1, w2);
      }
      if (g_failed > 5) return;
    }
    // the master matches every scan, the others only past the (constant) gate: both paths were exercised
    CHECK(matched_total > 9 && matched_total < 27, "gate: %ld particle-steps matched out of 27", matched_total);
    for (int i = 0; i < 3; ++i) same_as_device_map(ref[i]->map(), ctx, gpu->particle_map(i), init_grid_map_params(props), "batched GMapping particle map");
    std::printf("   9 scans, %ld particle-steps matched, heaviest pose %.6f %.6f %.6f\n", matched_total, gpu->pose().x, gpu->pose().y, gpu->pose().theta);
  }
  // ---- B: real noise, resampling, diverging particles: against the sequential statement above, same seeds
  {
    std::printf("== batched GMapping filter vs sequential reference components, seeded, 6 particles\n");
    // pose noise far beyond what six halving rounds of hill climbing recover: some particles match badly, the weights
    // spread and the filter resamples
    auto props = gmapping_props("6", false, "1.5", "0.6");
    std::uint32_t c1 = 1000, c2 = 1000;
    SeqFilter ref(props, 6, [&c1] { return c1 += 7; });
    const std::string OOPE_Pfx = "slam/scmtch/oope/";
    slamgpu::CudaGmappingParticleFilter::Properties p;
    p.setup.oope = SLAMGPU_OOPE_GMAPPING; p.setup.gm_cache = 2;
    p.spw = init_swp(props);
    p.spe = std::make_shared<WeightedMeanPointProbabilitySPE>(std::make_shared<GmappingOccupancyObservationPE>(0.1, 1), p.spw);
    p.map_params = init_grid_map_params(props);
    p.adder = slamgpu::init_cuda_scan_adder_properties(props);
    p.particles = 6;
    p.seed_source = [&c2] { return c2 += 7; };
    slamgpu::CudaGmappingParticleFilter gpu(ctx, p, init_gmapping_params(props));
    std::mt19937 rng(47);
    RobotPose truth{0, 0, 0};
    double t_ref = 0, t_gpu = 0;
    std::vector<double> per_ref, per_gpu;  // steps in which at least one particle scan-matched
    for (int step = 0; step < 16; ++step) {
      RobotPoseDelta motion = step == 0 ? RobotPoseDelta{0, 0, 0} : RobotPoseDelta{0.12, 0.05, 0.04};
      truth += motion;
      auto scan = room_scan(truth, 360, 2 * M_PI, 3.0, 2.5, rng, 0.005);
      TransformedLaserScan a{motion, scan, 1.0}, b{motion, scan, 1.0};
      b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
      auto t0 = std::chrono::steady_clock::now();
      ref.step(a);
      auto t1 = std::chrono::steady_clock::now();
      gpu.handle_sensor_data(b);
      auto t2 = std::chrono::steady_clock::now();
      if (step > 0) {
        t_ref += std::chrono::duration<double, std::milli>(t1 - t0).count(); t_gpu += std::chrono::duration<double, std::milli>(t2 - t1).count();
        if (gpu.last_step().matched) { per_ref.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count()); per_gpu.push_back(std::chrono::duration<double, std::milli>(t2 - t1).count()); }
      }
      CHECK(ref.resamplings == gpu.resamplings(), "step %d: %zu resamplings vs %zu", step, ref.resamplings, gpu.resamplings());
      if (std::getenv("SLAMGPU_TEST_VERBOSE")) {
        int64_t st[8] = {0};
        slamgpu_score_stats(ctx->handle(), st);
        std::printf("   step %2d matched %zu tested %ld variant %ld ref %.2f ms gpu %.2f ms (match %.2f insert %.2f resample %.2f) weights", step,
                    gpu.last_step().matched, (long)gpu.last_step().poses_tested, (long)st[1],
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
                    gpu.last_step().match_ms, gpu.last_step().insert_ms, gpu.last_step().resample_ms);
        for (int i = 0; i < 6; ++i) std::printf(" %.4f", gpu.particle_weight(i));
        std::printf("\n");
      }
      for (int i = 0; i < 6; ++i) {
        const RobotPose &p1 = ref.ps[i]->pose, &p2 = gpu.particle_pose(i);
        CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "step %d particle %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)",
              step, i, p1.x, p1.y, p1.theta, p2.x, p2.y, p2.theta);
        CHECK(std::fabs(ref.ps[i]->weight - gpu.particle_weight(i)) <= 1e-5 * ref.ps[i]->weight, "step %d particle %d: weight %.17g vs %.17g",
              step, i, ref.ps[i]->weight, gpu.particle_weight(i));
        CHECK(ref.ps[i]->master == gpu.particle_is_master(i), "step %d particle %d: master flag", step, i);
      }
      if (g_failed > 5) return;
    }
    CHECK(gpu.resamplings() > 0, "the sequence never resampled: the test does not cover it");
    CHECK(ref.heaviest() == gpu.heaviest_particle(), "heaviest particle %zu vs %zu", ref.heaviest(), gpu.heaviest_particle());
    for (int i = 0; i < 6; ++i) same_as_device_map(*ref.ps[i]->map, ctx, gpu.particle_map(i), p.map_params, "seeded GMapping particle map");
    same_cells(*ref.ps[ref.heaviest()]->map, gpu.map(), "published map");
    double err = std::hypot(gpu.pose().x - truth.x, gpu.pose().y - truth.y);
    CHECK(err < 0.6, "batched filter lost track: err %.3f", err);
    std::sort(per_ref.begin(), per_ref.end()); std::sort(per_gpu.begin(), per_gpu.end());
    std::printf("   16 scans, %zu resamplings, err %.3f m; per scan: reference components %.2f ms, batched CUDA filter %.2f ms "
                "(median of the %zu matching steps: %.2f vs %.2f ms)\n",
                gpu.resamplings(), err, t_ref / 15, t_gpu / 15, per_gpu.size(), per_ref[per_ref.size() / 2], per_gpu[per_gpu.size() / 2]);
  }
}

}  // namespace

int main() {
  slamgpu::error_mode() = slamgpu::ErrorMode::Throw;  // this binary reports failures itself
  std::shared_ptr<slamgpu::Context> ctx;
  try {
    ctx = std::make_shared<slamgpu::Context>(0);
  } catch (const slamgpu::Error &e) {
    std::printf("NO-DEVICE %s\n", e.what());
    return e.code == SLAMGPU_E_NODEVICE ? 77 : 1;
  }
  try {
    Config tiny{"tinySLAM: mean cell, const estimator, blur 0.5, MC(42, 0.2/0.1, 20, 100), even weights, 360 beams", 100, 0.1,
                std::make_shared<MeanProbabilityCell>(), SLAMGPU_EST_CONST, {0.95, 1.0}, {0.01, 1.0}, 0.5, 0, 0, 360, 2 * M_PI, 25, 0.9, 0.6};
    Config viny{"vinySLAM: tbm_consistent cell, const 0.95/0.04 0.01/0.003, blur 0.3, MC, viny weights, 361 beams over 270 deg", 100, 0.1,
                std::make_shared<TbmOccConsistentCell>(), SLAMGPU_EST_CONST, {0.95, 0.04}, {0.01, 0.003}, 0.3, 0, 1, 361, 1.5 * M_PI, 25, 0.9, 0.6};
    Config hc{"hill climbing + area estimator: tbm_unknown_even cell, HC(6, 0.1, 0.1), 0.05 m cells, 541 beams", 200, 0.05,
              std::make_shared<TbmUnknownEvenOccCell>(), SLAMGPU_EST_AREA, {0.95, 0.3}, {0.01, 0.1}, 0.3, 1, 1, 541, 1.5 * M_PI, 15, 0.9, 0.6};
    Config bf{"brute force: affine cell, area estimator, BF(+-0.3 @0.05, +-0.1 rad @0.02)", 160, 0.05,
              std::make_shared<AffineQualityMergeCell>(), SLAMGPU_EST_AREA, {0.95, 1.0}, {0.01, 1.0}, 0.2, 2, 0, 181, 1.5 * M_PI, 6, 0.9, 0.6};
    run_world_pair(tiny, ctx);
    run_world_pair(viny, ctx);
    run_world_pair(hc, ctx);
    run_world_pair(bf, ctx);
    run_host_map(ctx);
    run_m3rsm(ctx);
    run_presets(ctx);
    run_gmapping(ctx);
    run_gmapping_batched(ctx);
  } catch (const std::exception &e) {
    std::printf("FAIL exception: %s\n", e.what());
    return 1;
  }
  std::printf(g_failed ? "RESULT: %d FAILED\n" : "RESULT: ALL PASSED\n", g_failed);
  return g_failed ? 1 : 0;
}
