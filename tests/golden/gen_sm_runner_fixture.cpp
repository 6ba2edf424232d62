// gen_sm_runner_fixture.cpp -- writes the on-disk triple sm_runner consumes (src/utils/sm_runner.cpp:36-62:
// <name>.pose2D text, <name>.scan2D text, <name>.map binary = GridMap::save_state) with the UNMODIFIED reference
// classes, plus what the reference's own matchers answer on it.  Run once here:
//   g++ -std=c++14 -O2 -w -I/root/reference tests/golden/gen_sm_runner_fixture.cpp -o /tmp/gen && /tmp/gen tests/golden/sm_runner
#include <cmath>
#include <cstdio>
#include <fstream>
#include <random>

#include "src/core/maps/const_occupancy_estimator.h"
#include "src/core/maps/grid_map_scan_adders.h"
#include "src/core/maps/plain_grid_map.h"
#include "src/core/maps/naive_grid_cells.h"
#include "src/core/maps/tbm_grid_cells.h"
#include "src/core/scan_matchers/hill_climbing_scan_matcher.h"
#include "src/core/scan_matchers/monte_carlo_scan_matcher.h"
#include "src/core/scan_matchers/observation_impact_estimators.h"
#include "src/core/scan_matchers/occupancy_observation_probability.h"
#include "src/core/scan_matchers/weighted_mean_point_probability_spe.h"
#include "src/utils/map_dumpers.h"

static LaserScan2D room_scan(const RobotPose &pose, int n, double fov, double hw, double hh, std::mt19937 &rng) {
  LaserScan2D scan;
  std::normal_distribution<double> nd(0.0, 0.01);
  for (int i = 0; i < n; ++i) {
    double a = -fov / 2 + fov * i / n;
    double c = std::cos(a + pose.theta), s = std::sin(a + pose.theta);
    double tx = c > 0 ? (hw - pose.x) / c : (c < 0 ? (-hw - pose.x) / c : 1e300);
    double ty = s > 0 ? (hh - pose.y) / s : (s < 0 ? (-hh - pose.y) / s : 1e300);
    scan.points().push_back(ScanPoint2D::make_polar(std::min(tx, ty) + nd(rng), a, i % 29 != 3));
  }
  return scan;
}

template <class Cell>
static void one(const std::string &dir, const std::string &name, std::ofstream &json, bool first) {
  std::mt19937 rng(5);
  GridMapParams gmp{80, 80, 0.1};
  auto map = std::make_shared<UnboundedPlainGridMap>(std::make_shared<Cell>(), gmp);
  auto est = std::make_shared<ConstOccupancyEstimator>(Occupancy{0.95, 0.5}, Occupancy{0.01, 0.2});
  auto adder = WallDistanceBlurringScanAdder::builder().set_blur_distance(0.3).set_occupancy_estimator(est)
                 .set_observation_quality_estimator(std::make_shared<IdleOMQE>()).build();
  RobotPose truth{0.25, -0.15, 0.2};
  for (int k = 0; k < 3; ++k) {
    auto s = room_scan(truth, 180, 2 * M_PI, 3.0, 2.5, rng);
    adder->append_scan(*map, truth, s, 1.0);
  }
  auto scan = room_scan(truth, 91, 1.5 * M_PI, 3.0, 2.5, rng);
  RobotPose init{truth.x + 0.06, truth.y - 0.05, truth.theta + 0.03};
  { std::ofstream f(dir + "/" + name + ".pose2D"); f.precision(17); f << init.x << " " << init.y << " " << init.theta << "\n"; }
  { std::ofstream f(dir + "/" + name + ".scan2D"); f.precision(17); f << scan; }
  { auto buf = map->save_state(); std::ofstream f(dir + "/" + name + ".map", std::ios::binary); f.write(buf.data(), buf.size()); }
  { std::ofstream f(dir + "/" + name + ".pgm", std::ios::binary); GridMapToPgmDumber::dump_map(f, *map); }
  auto spw = std::make_shared<EvenSPW>();
  auto oope = std::make_shared<ObstacleBasedOccupancyObservationPE>(std::make_shared<DiscrepancyOIE>());
  auto spe = std::make_shared<WeightedMeanPointProbabilitySPE>(oope, spw);
  HillClimbingScanMatcher hc{spe, 6, 0.1, 0.1};
  TransformedLaserScan tls{RobotPoseDelta{}, scan, 1.0};
  RobotPoseDelta d;
  double p = hc.process_scan(tls, init, *map, d);
  char buf[512];
  std::snprintf(buf, sizeof buf, "%s\"%s\": {\"width\": %d, \"height\": %d, \"origin\": [%d, %d], \"hc\": {\"prob\": %.17g, \"delta\": [%.17g, %.17g, %.17g]}}",
                first ? "" : ",\n ", name.c_str(), map->width(), map->height(), map->origin().x, map->origin().y, p, d.x, d.y, d.theta);
  json << buf;
}

int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  std::ofstream json(dir + "/expected.json");
  json << "{";
  one<AffineQualityMergeCell>(dir, "affine_cell", json, true);
  one<TbmUnknownEvenOccCell>(dir, "tbm_cell", json, false);
  json << "}\n";
  return 0;
}
