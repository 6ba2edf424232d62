// test_backend_key.cpp -- compiled against the reference headers WITH integration/slam_backend_key.patch applied (the
// Makefile patches a throw-away copy): the reference's own init_1h_slam(props) builds the CPU world from a preset and the
// CUDA world from the same preset plus `slam/backend=cuda`; both must produce identical poses and maps.
#include <cmath>
#include <cstdio>
#include <random>

#include "slamgpu_init.h"
#include "src/utils/init_slam.h"

static int g_failed = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++g_failed; std::printf("FAIL "); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static LaserScan2D room_scan(const RobotPose &pose, int n, double fov, double hw, double hh, std::mt19937 &rng, double noise) {
  LaserScan2D scan;
  scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
  std::normal_distribution<double> nd(0.0, noise);
  for (int i = 0; i < n; ++i) {
    double a = -fov / 2 + fov * i / n, th = a + pose.theta, c = std::cos(th), s = std::sin(th);
    double tx = c > 0 ? (hw - pose.x) / c : (c < 0 ? (-hw - pose.x) / c : INFINITY);
    double ty = s > 0 ? (hh - pose.y) / s : (s < 0 ? (-hh - pose.y) / s : INFINITY);
    scan.points().emplace_back(std::min(tx, ty) + nd(rng), a);
  }
  return scan;
}

int main() {
  slamgpu::error_mode() = slamgpu::ErrorMode::Throw;  // this binary reports failures itself
  std::shared_ptr<slamgpu::Context> ctx;
  try {
    ctx = std::make_shared<slamgpu::Context>(0);
  } catch (const slamgpu::Error &e) {
    std::printf("NO-DEVICE %s\n", e.what());
    return e.code == SLAMGPU_E_NODEVICE ? 77 : 1;
  }
  slamgpu::install_backend(ctx);
  const std::vector<std::pair<const char *, const char *>> tiny = {
      {"slam/mapping/blur", "0.5"}, {"slam/occupancy_estimator/type", "const"}, {"slam/occupancy_estimator/base_occupied/prob", "0.95"},
      {"slam/occupancy_estimator/base_empty/prob", "0.01"}, {"slam/mapping/grid/area/type", "mean_probability"},
      {"slam/mapping/corrected_pose_quality", "0.9"}, {"slam/mapping/raw_pose_quality", "0.6"},
      {"slam/mapping/grid/type", "unbounded_plain"}, {"slam/map/height_in_meters", "10"}, {"slam/map/width_in_meters", "10"},
      {"slam/map/meters_per_cell", "0.1"}, {"slam/scmtch/type", "MC"}, {"slam/scmtch/MC/dispersion/translation", "0.2"},
      {"slam/scmtch/MC/dispersion/rotation", "0.1"}, {"slam/scmtch/MC/dispersion/failed_attempts_limit", "20"},
      {"slam/scmtch/MC/attempts_limit", "100"}, {"slam/scmtch/MC/seed", "42"}, {"slam/scmtch/spe/type", "wmpp"},
      {"slam/scmtch/spe/wmpp/weighting/type", "even"}};
  MapPropertiesProvider cpu_props, gpu_props;
  for (auto &kv : tiny) { cpu_props.set_property(kv.first, kv.second); gpu_props.set_property(kv.first, kv.second); }
  gpu_props.set_property("slam/backend", "cuda");
  auto ref_world = init_1h_slam(cpu_props);   // the reference's factory, reference classes
  auto gpu_world = init_1h_slam(gpu_props);   // the SAME factory, CUDA plug-ins
  CHECK(dynamic_cast<const slamgpu::CudaGridMap *>(&gpu_world->map()) != nullptr, "slam/backend=cuda did not select the CUDA map");
  CHECK(dynamic_cast<const slamgpu::CudaGridMap *>(&ref_world->map()) == nullptr, "the default back end is not the reference's");
  std::mt19937 rng(31);
  std::normal_distribution<double> odo(0.0, 0.02), odo_t(0.0, 0.01);
  RobotPose truth{0.2, -0.1, 0.15};
  for (int step = 0; step < 40; ++step) {
    RobotPoseDelta motion = step == 0 ? RobotPoseDelta{truth.x, truth.y, truth.theta} : RobotPoseDelta{0.06, 0.03 * std::cos(0.5 * step), 0.02};
    if (step > 0) truth += motion;
    RobotPoseDelta odom = step == 0 ? motion : RobotPoseDelta{motion.x + odo(rng), motion.y + odo(rng), motion.theta + odo_t(rng)};
    auto scan = room_scan(truth, 360, 2 * M_PI, 3.5, 3.0, rng, 0.01);
    TransformedLaserScan a{odom, scan, 1.0}, b{odom, scan, 1.0};
    b.scan.trig_provider = std::make_shared<RawTrigonometryProvider>();
    ref_world->handle_sensor_data(a);
    gpu_world->handle_sensor_data(b);
    const RobotPose &p1 = ref_world->pose(), &p2 = gpu_world->pose();
    CHECK(p1.x == p2.x && p1.y == p2.y && p1.theta == p2.theta, "step %d: pose (%.17g %.17g %.17g) vs (%.17g %.17g %.17g)", step, p1.x, p1.y,
          p1.theta, p2.x, p2.y, p2.theta);
  }
  const GridMap &m1 = ref_world->map(), &m2 = gpu_world->map();
  CHECK(m1.width() == m2.width() && m1.height() == m2.height(), "map size %dx%d vs %dx%d", m1.width(), m1.height(), m2.width(), m2.height());
  long bad = 0, known = 0;
  for (int y = 0; y < m1.height() && y < m2.height(); ++y)
    for (int x = 0; x < m1.width() && x < m2.width(); ++x) {
      auto c = m1.internal2external({x, y});
      const GridCell &c1 = m1[c], &c2 = m2[c];
      bad += !(c1.is_unknown() == c2.is_unknown() && double(c1) == double(c2));
      known += !c1.is_unknown();
    }
  CHECK(bad == 0 && known > 100, "%ld cells differ (%ld known)", bad, known);
  std::printf(g_failed ? "RESULT: %d FAILED\n" : "RESULT: ALL PASSED\n", g_failed);
  return g_failed ? 1 : 0;
}
