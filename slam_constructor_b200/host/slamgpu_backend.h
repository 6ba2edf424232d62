// slamgpu_backend.h -- the B200 back end as plug-ins for slam-constructor's own interfaces.
//
// Header-only C++14, compiled against the UNMODIFIED reference headers (add the reference's
// root to the include path: -I<slam-constructor>); it talks to the GPU only through the C ABI
// of include/slamgpu.h.  Nothing in here computes a score or a cell value on the CPU.
//
//   slamgpu::CudaGridMap            : GridMap            device-resident map (+ lazy host mirror)
//   slamgpu::CudaScanAdder          : GridMapScanAdder   queues handle_scan_point, one K2/K3 launch per scan
//   slamgpu::CudaBruteForceScanMatcher          : GridScanMatcher   BruteForceScanMatcher's candidate set on K1 (grid kernel)
//   slamgpu::CudaPoseEnumerationScanMatcher<PE> : GridScanMatcher   any copyable PoseEnumerator (Monte-Carlo, hill climbing,
//                                                 polar brute force) by speculative batches on K1 (list kernel)
//   slamgpu::CudaMonteCarloScanMatcher, slamgpu::CudaHillClimbingScanMatcher   the two stock instances
//   slamgpu::CudaPyramidGridMap     : CudaGridMap        fine map + max-pyramid (M3RSMRescalableGridMap)
//   slamgpu::CudaBfMultiResScanMatcher : GridScanMatcher BruteForceMultiResolutionScanMatcher on K5
//
// They drop into SingleStateHypothesisLSGWProperties{grid_map, gsm, gmsa}
// (src/core/states/single_state_hypothesis_laser_scan_grid_world.h:13-21) and hence into the
// tinySLAM / vinySLAM world and into GmappingWorld / GmappingParticleFilter
// (src/slams/gmapping/gmapping_world.h:36-128) unchanged.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "src/core/maps/area_occupancy_estimator.h"
#include "src/core/maps/const_occupancy_estimator.h"
#include "src/core/maps/grid_map.h"
#include "src/core/maps/grid_map_scan_adders.h"
#include "src/core/maps/naive_grid_cells.h"
#include "src/core/maps/tbm_grid_cells.h"
#include "src/core/scan_matchers/bf_multi_res_scan_matcher.h"
#include "src/core/scan_matchers/brute_force_scan_matcher.h"
#include "src/core/scan_matchers/grid_scan_matcher.h"
#include "src/core/scan_matchers/hill_climbing_scan_matcher.h"
#include "src/core/scan_matchers/monte_carlo_scan_matcher.h"
#include "src/core/scan_matchers/observation_impact_estimators.h"
#include "src/core/scan_matchers/occupancy_observation_probability.h"
#include "src/core/scan_matchers/weighted_mean_point_probability_spe.h"
#include "src/slams/gmapping/gmapping_grid_cell.h"
#include "src/slams/gmapping/gmapping_occupancy_observation_pe.h"

#include "slamgpu.h"

namespace slamgpu {

class Error : public std::runtime_error {
public:
  Error(int code, const std::string &msg) : std::runtime_error("slamgpu error " + std::to_string(code) + ": " + msg), code(code) {}
  int code;
};

// How the plug-ins report a failure.  The reference is exception-free by convention: a broken contract is an assert
// (abort), a bad configuration prints to std::cerr and calls std::exit(-1) (init_scan_matching.h:212-215), "unknown" is a
// NaN.  The default mode follows it, so nothing is ever thrown through the reference's world classes (or a ROS node built
// on them); a host that prefers exceptions -- the test binaries do -- opts in with error_mode() = ErrorMode::Throw.
enum class ErrorMode { Fatal, Throw };
inline ErrorMode &error_mode() { static ErrorMode m = ErrorMode::Fatal; return m; }
[[noreturn]] inline void fail(int code, const std::string &msg) {  // a C-ABI call returned an error (device lost, out of memory, ...)
  if (error_mode() == ErrorMode::Throw) throw Error(code, msg);
  std::cerr << "[slamgpu] error " << code << ": " << msg << std::endl;
  std::exit(-1);
}
[[noreturn]] inline void contract_violation(const char *msg) {  // the plug-in is used against its contract (upstream: assert)
  if (error_mode() == ErrorMode::Throw) throw std::logic_error(msg);
  std::cerr << "[slamgpu] contract violation: " << msg << std::endl;
  std::abort();
}

// one CUDA context (stream, scratch) per world; shared by the map, the adder and the matcher
class Context {
public:
  explicit Context(int device = 0) {
    int r = slamgpu_ctx_create(device, &_h);
    if (r != SLAMGPU_OK) fail(r, slamgpu_last_error(nullptr));
  }
  ~Context() { slamgpu_ctx_destroy(_h); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  slamgpu_ctx *handle() const { return _h; }
  void check(int r) const {
    if (r != SLAMGPU_OK) fail(r, slamgpu_last_error(_h));
  }
private:
  slamgpu_ctx *_h = nullptr;
};

// the cell model enum of a reference cell class
inline int cell_model_of(const GridCell &proto) {
  if (dynamic_cast<const MeanProbabilityCell *>(&proto)) return SLAMGPU_CELL_MEAN;
  if (dynamic_cast<const AffineQualityMergeCell *>(&proto)) return SLAMGPU_CELL_AFFINE;
  if (dynamic_cast<const TbmOccConsistentCell *>(&proto)) return SLAMGPU_CELL_TBM_CONSISTENT;
  if (dynamic_cast<const TbmUnknownEvenOccCell *>(&proto)) return SLAMGPU_CELL_TBM_UNKNOWN_EVEN;
  if (dynamic_cast<const GmappingBaseCell *>(&proto)) return SLAMGPU_CELL_GMAPPING;
#ifdef SLAM_CTOR_SLAM_CREDIBILIST_GRID_CELL_H  // src/slams/credibilist/grid_cell.h included before this header
  if (dynamic_cast<const CredibilistCell *>(&proto)) return SLAMGPU_CELL_CREDIBILIST;
#endif
  return SLAMGPU_CELL_LWW;  // GridCell itself (last write wins) and test doubles built on it
}

// Host view of one device cell record, handed out by CudaGridMap::operator[] (publishers, PGM dumps,
// tests read occupancy()/is_unknown() from it; nothing on the hot path does).
class MirrorCell : public GridCell {
public:
  MirrorCell() : GridCell{Occupancy{0.5, 1}}, _model{SLAMGPU_CELL_LWW} { std::memset(_rec, 0, sizeof _rec); }
  MirrorCell(int model, const double *rec) : GridCell{occupancy_of(model, rec)}, _model{model} {
    std::memcpy(_rec, rec, sizeof _rec);
    if (!unknown_of(model, rec)) { on_update(); }
  }
  std::unique_ptr<GridCell> clone() const override { return std::make_unique<MirrorCell>(*this); }
  void operator+=(const AreaOccupancyObservation &) override {
    contract_violation("MirrorCell is read-only: update the CudaGridMap, not the cell");
  }
  double discrepancy(const AreaOccupancyObservation &aoo) const override {
    switch (_model) {
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN: {
      TbmOccConsistentCell probe;  // the formula only needs the belief, restored through the public deserializer
      Serializer s(GridCell::serialize());
      s << _rec[2] << _rec[3] << _rec[4] << 0.0;
      probe.deserialize(s.result());
      return probe.discrepancy(aoo);
    }
    case SLAMGPU_CELL_GMAPPING: {
      auto d = std::pow(_rec[1] - aoo.obstacle.x, 2) + std::pow(_rec[2] - aoo.obstacle.y, 2);
      return 1.0 - std::exp(-d / 0.05);
    }
    default: return GridCell::discrepancy(aoo);
    }
  }
  int model() const { return _model; }
  const double *record() const { return _rec; }

  static Occupancy occupancy_of(int model, const double *r) {
    switch (model) {
    case SLAMGPU_CELL_LWW: return Occupancy{r[0], r[1]};
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN:
    case SLAMGPU_CELL_CREDIBILIST: return Occupancy{r[0], r[1]};
    default: return Occupancy{r[0], 1};
    }
  }
  static bool unknown_of(int model, const double *r) {
    switch (model) {
    case SLAMGPU_CELL_LWW: return r[2] == 0;
    case SLAMGPU_CELL_AFFINE: return r[1] == 0;
    case SLAMGPU_CELL_MEAN: return r[1] == 0;
    case SLAMGPU_CELL_GMAPPING: return r[4] == 0;
    default: return r[5] == 0;
    }
  }
private:
  int _model;
  double _rec[SLAMGPU_MAX_STRIDE];
};

// what a scan adder needs to hand to the device with each beam
struct AdderParams {
  slamgpu_estimator est;
  double blur = 0;
  double max_range = std::numeric_limits<double>::infinity();
  bool operator==(const AdderParams &o) const {
    return std::memcmp(&est, &o.est, sizeof est) == 0 && blur == o.blur && max_range == o.max_range;
  }
};

//============================================================================//
// CudaGridMap: GridMap whose cells live in HBM.  Growth follows the reference's unbounded maps
// (SLAMGPU_GROW_PLAIN = UnboundedPlainGridMap, SLAMGPU_GROW_TILED = UnboundedLazyTiledGridMap,
// SLAMGPU_GROW_NONE = PlainGridMap).
class CudaGridMap : public GridMap {
public:
  CudaGridMap(std::shared_ptr<Context> ctx, std::shared_ptr<GridCell> prototype, const GridMapParams &params = MapValues::gmp,
              int grow = SLAMGPU_GROW_PLAIN)
    : GridMap{prototype, params}, _ctx{ctx}, _model{cell_model_of(*prototype)}, _grow{grow} {
    _ctx->check(slamgpu_map_create(_ctx->handle(), params.width_cells, params.height_cells, params.meters_per_cell, _model, grow,
                                   nullptr, &_map));
    _stride = slamgpu_model_stride(_model);
    slamgpu_default_unknown(_model, _unknown_rec);
    _unknown_cell = MirrorCell(_model, _unknown_rec);
    refresh_info();
  }
  // a view of a device map owned by somebody else (a particle's map, slamgpu_particles_map): same interface,
  // never destroyed here; rebind() points the view at another map of the same model
  CudaGridMap(std::shared_ptr<Context> ctx, std::shared_ptr<GridCell> prototype, const GridMapParams &params, int grow,
              slamgpu_map *borrowed)
    : GridMap{prototype, params}, _ctx{ctx}, _map{borrowed}, _owned{false}, _model{cell_model_of(*prototype)}, _grow{grow} {
    _stride = slamgpu_model_stride(_model);
    slamgpu_default_unknown(_model, _unknown_rec);
    _unknown_cell = MirrorCell(_model, _unknown_rec);
    refresh_info();
  }
  void rebind(slamgpu_map *m) {
    if (_owned) { contract_violation("CudaGridMap::rebind: this map owns its device map"); }
    flush();
    _map = m;
    touched();
  }
  ~CudaGridMap() override { if (_owned) slamgpu_map_destroy(_map); }
  CudaGridMap(const CudaGridMap &) = delete;
  CudaGridMap &operator=(const CudaGridMap &) = delete;

  // ---- RegularSquaresGrid
  int width() const override { flush(); return _w; }
  int height() const override { flush(); return _h; }
  DiscretePoint2D origin() const override { flush(); return DiscretePoint2D{_ox, _oy}; }
  bool has_cell(const Coord &c) const override {
    if (_grow != SLAMGPU_GROW_NONE) { return true; }  // unbounded maps: plain_grid_map.h:77
    return GridMap::has_cell(c);
  }

  // ---- GridMap
  const GridCell &operator[](const Coord &c) const override {
    flush();
    int ix = c.x + _ox, iy = c.y + _oy;
    if (ix < 0 || ix >= _w || iy < 0 || iy >= _h) { return _unknown_cell; }
    ensure_mirror();
    MirrorCell &slot = _ring[_ring_next++ % _ring.size()];  // valid until kRing more lookups
    slot = MirrorCell(_model, _mirror.data() + ((size_t)iy * _w + ix) * _stride);
    return slot;
  }
  void update(const Coord &c, const AreaOccupancyObservation &aoo) override {
    flush();
    _ctx->check(slamgpu_map_update_cell(_map, c.x, c.y, aoo.is_occupied, aoo.occupancy.prob_occ, aoo.occupancy.estimation_quality,
                                        aoo.obstacle.x, aoo.obstacle.y, aoo.quality));
    touched();
  }
  void reset(const Coord &c, const GridCell &cell) override {
    flush();
    double rec[SLAMGPU_MAX_STRIDE];
    record_of(cell, rec);
    _ctx->check(slamgpu_map_reset_cell(_map, c.x, c.y, rec));
    touched();
  }

  // ---- device side
  std::shared_ptr<Context> context() const { return _ctx; }
  slamgpu_map *device() const { flush(); return _map; }
  int cell_model() const { return _model; }
  // dense copy of the records, [h][w][stride]
  const std::vector<double> &records() const { flush(); ensure_mirror(); return _mirror; }

  // GridMapScanAdder::handle_scan_point arguments, queued until the map is next looked at
  void queue_beam(bool is_occ, double quality, const Segment2D &beam, const AdderParams &p) {
    if (!_q_occ.empty() && (!(p == _q_params) || beam.beg().x != _q_beams[0] || beam.beg().y != _q_beams[1])) { flush(); }
    _q_params = p;
    _q_beams.push_back(beam.beg().x); _q_beams.push_back(beam.beg().y);
    _q_beams.push_back(beam.end().x); _q_beams.push_back(beam.end().y);
    _q_occ.push_back(is_occ ? 1 : 0);
    _q_quality.push_back(quality);
  }
  void flush() const {
    if (_q_occ.empty()) { return; }
    std::vector<double> beams;
    std::vector<uint8_t> occ;
    std::vector<double> quality;
    beams.swap(_q_beams); occ.swap(_q_occ); quality.swap(_q_quality);  // re-entrancy: the queue is empty from here on
    _cells_updated += append_beams_on_device((int32_t)occ.size(), beams.data(), occ.data(), quality.data(), _q_params);
    const_cast<CudaGridMap *>(this)->touched();
  }
  int64_t cells_updated() const { return _cells_updated; }

protected:
  virtual int64_t append_beams_on_device(int32_t n, const double *beams, const uint8_t *occ, const double *quality,
                                         const AdderParams &p) const {
    int64_t cells = 0;
    _ctx->check(slamgpu_append_beams(_ctx->handle(), _map, n, beams, occ, quality, &p.est, p.blur, p.max_range, &cells));
    return cells;
  }
  slamgpu_map *raw_device_map() const { return _map; }

private:
public:
  // the device map changed behind this object's back (e.g. a batched particle insertion)
  void touched() { _mirror_valid = false; refresh_info(); }
private:
  void refresh_info() const {
    int32_t st;
    double sc;
    _ctx->check(slamgpu_map_info(_map, &_w, &_h, &sc, &_ox, &_oy, &st));
  }
  void ensure_mirror() const {
    if (_mirror_valid) { return; }
    _mirror.resize((size_t)_w * _h * _stride);
    _ctx->check(slamgpu_map_download(_map, _mirror.data()));
    _mirror_valid = true;
  }
  void record_of(const GridCell &cell, double *rec) const {
    std::memcpy(rec, _unknown_rec, sizeof _unknown_rec);
    if (auto m = dynamic_cast<const MirrorCell *>(&cell)) { std::memcpy(rec, m->record(), sizeof _unknown_rec); return; }
    if (cell.is_unknown()) { return; }
    const Occupancy &o = cell.occupancy();
    switch (_model) {
    case SLAMGPU_CELL_LWW: rec[0] = o.prob_occ; rec[1] = o.estimation_quality; rec[2] = 1; return;
    case SLAMGPU_CELL_AFFINE: rec[0] = o.prob_occ; rec[1] = 1; return;
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN:
      if (auto t = dynamic_cast<const TbmBaseCell *>(&cell)) {
        rec[0] = o.prob_occ; rec[1] = o.estimation_quality;
        rec[2] = t->belief().unknown(); rec[3] = t->belief().empty(); rec[4] = t->belief().occupied(); rec[5] = 1;
        return;
      }
      break;
    default: break;
    }
    // MeanProbabilityCell / GmappingBaseCell keep their counters private: a written cell of those
    // classes cannot be moved onto the device through the public API
    contract_violation("CudaGridMap::reset: cannot take over a known cell of this class");
  }

private:
  static constexpr std::size_t kRing = 4096;
  std::shared_ptr<Context> _ctx;
  slamgpu_map *_map = nullptr;
  bool _owned = true;
  int _model, _grow, _stride = 0;
  double _unknown_rec[SLAMGPU_MAX_STRIDE];
  MirrorCell _unknown_cell;
  mutable int32_t _w = 0, _h = 0, _ox = 0, _oy = 0;
  mutable std::vector<double> _mirror;
  mutable bool _mirror_valid = false;
  mutable std::vector<MirrorCell> _ring = std::vector<MirrorCell>(kRing);
  mutable std::size_t _ring_next = 0;
  mutable std::vector<double> _q_beams, _q_quality;
  mutable std::vector<uint8_t> _q_occ;
  mutable AdderParams _q_params;
  mutable int64_t _cells_updated = 0;
};

//============================================================================//
// CudaScanAdder: WallDistanceBlurringScanAdder on the device.  GridMapScanAdder::append_scan is not
// virtual (grid_map_scan_adders.h:54); it walks the scan on the host (libm trig, mapping quality) and
// calls handle_scan_point per beam, which here only queues the beam on the CudaGridMap; the map
// inserts the whole scan with one slamgpu_append_beams call the next time anybody looks at it.
class CudaScanAdder : public GridMapScanAdder {
public:
  struct Properties {
    int estimator = SLAMGPU_EST_CONST;  // SLAMGPU_EST_CONST | SLAMGPU_EST_AREA
    Occupancy base_occupied{0.95, 1.0}, base_empty{0.01, 1.0};
    double low_qual = 0.01, unknown_qual = 0.5;  // area estimator only
    double blur_distance = 0;
    double max_usable_range = std::numeric_limits<double>::infinity();
    std::shared_ptr<ObservationMappingQualityEstimator> observation_quality_estimator = std::make_shared<IdleOMQE>();
  };
  explicit CudaScanAdder(const Properties &p) : GridMapScanAdder{make_estimator(p), p.observation_quality_estimator}, _props{p} {
    std::memset(&_params.est, 0, sizeof _params.est);
    _params.est.type = p.estimator;
    _params.est.occ_p = p.base_occupied.prob_occ; _params.est.occ_q = p.base_occupied.estimation_quality;
    _params.est.empty_p = p.base_empty.prob_occ; _params.est.empty_q = p.base_empty.estimation_quality;
    _params.est.low_qual = p.low_qual; _params.est.unknown_qual = p.unknown_qual;
    _params.est.shift_amount = -1;
    _params.blur = p.blur_distance;
    _params.max_range = p.max_usable_range;
  }
protected:
  void handle_scan_point(GridMap &map, bool is_occ, double scan_quality, const Segment2D &beam) const override {
    auto *cm = dynamic_cast<CudaGridMap *>(&map);
    if (!cm) { contract_violation("CudaScanAdder updates a CudaGridMap only (there is no CPU fallback)"); }
    if (_params.est.type == SLAMGPU_EST_AREA && _params.est.shift_amount < 0) {
      // AreaOccupancyEstimator::ensure_segment_not_on_edge keeps a function-static shift computed from
      // the first cell it ever sees (area_occupancy_estimator.h:71): the obstacle cell of the first beam
      auto c = map.world_to_cell(beam.end());
      auto b = map.world_cell_bounds(c);
      _params.est.shift_amount = area_shift_amount(_props.low_qual * b.side());
    }
    cm->queue_beam(is_occ, scan_quality, beam, _params);
  }
private:
  static std::shared_ptr<CellOccupancyEstimator> make_estimator(const Properties &p) {
    if (p.estimator == SLAMGPU_EST_AREA) {
      return std::make_shared<AreaOccupancyEstimator>(p.base_occupied, p.base_empty, p.low_qual, p.unknown_qual);
    }
    return std::make_shared<ConstOccupancyEstimator>(p.base_occupied, p.base_empty);
  }
  static double area_shift_amount(double first) {
    static double shift = first;  // process-wide and fixed by its first use, like the reference's static
    return shift;
  }
  Properties _props;
  mutable AdderParams _params;
};

//============================================================================//
// scan scoring set-up shared by the matchers

struct ScoreSetup {
  int oope = SLAMGPU_OOPE_OBSTACLE, oie = SLAMGPU_OIE_DISCREPANCY;
  double gm_fullness_th = 0.1;
  int gm_window = 1;
  int gm_cache = 0;  // 2: carry the GMapping OOPE's one-entry cache from candidate to candidate, as the estimator object does
  bool generic_oie = false;  // an OIE class this header does not know: score through a host-built LUT
};

// which kernels reproduce this estimator chain (the GMapping OOPE keeps its parameters private:
// pass a ScoreSetup explicitly for it)
inline ScoreSetup detect_score_setup(const ScanProbabilityEstimator &spe) {
  ScoreSetup s;
  auto oope = spe.occupancy_observation_probability_estimator();
  if (dynamic_cast<const ObstacleBasedOccupancyObservationPE *>(oope.get())) s.oope = SLAMGPU_OOPE_OBSTACLE;
  else if (dynamic_cast<const MaxOccupancyObservationPE *>(oope.get())) s.oope = SLAMGPU_OOPE_MAX;
  else if (dynamic_cast<const MeanOccupancyObservationPE *>(oope.get())) s.oope = SLAMGPU_OOPE_MEAN;
  else if (dynamic_cast<const OverlapWeightedOccupancyObservationPE *>(oope.get())) s.oope = SLAMGPU_OOPE_OVERLAP;
  else if (dynamic_cast<const GmappingOccupancyObservationPE *>(oope.get())) { s.oope = SLAMGPU_OOPE_GMAPPING; s.gm_cache = 2; }
  else contract_violation("slamgpu: unknown OccupancyObservationProbabilityEstimator class");
  auto oie = oope->impact_estimator();
  if (dynamic_cast<const DiscrepancyOIE *>(oie.get())) s.oie = SLAMGPU_OIE_DISCREPANCY;
  else if (dynamic_cast<const OccupancyOIE *>(oie.get())) s.oie = SLAMGPU_OIE_OCCUPANCY;
  else if (s.oope != SLAMGPU_OOPE_GMAPPING) s.generic_oie = true;
  return s;
}

// the device view of the map a matcher was given: a CudaGridMap is used in place; any other GridMap
// is snapshotted as a score LUT evaluated with the reference's own OIE (slow: a host pass per call)
class MapBinding {
public:
  explicit MapBinding(std::shared_ptr<Context> ctx) : _ctx{ctx} {}
  ~MapBinding() { if (_snapshot) slamgpu_map_destroy(_snapshot); }
  slamgpu_map *bind(const GridMap &map, const ScanProbabilityEstimator &spe, const ScoreSetup &setup) {
    if (auto cm = dynamic_cast<const CudaGridMap *>(&map)) {
      if (setup.generic_oie) contract_violation("slamgpu: a custom ObservationImpactEstimator needs a host map");
      return cm->device();
    }
    if (setup.oope == SLAMGPU_OOPE_GMAPPING) contract_violation("slamgpu: the GMapping OOPE gathers cell records: use a CudaGridMap");
    const int w = map.width(), h = map.height();
    const auto org = map.origin();
    if (!_snapshot) {
      _ctx->check(slamgpu_map_create(_ctx->handle(), w, h, map.scale(), SLAMGPU_CELL_LWW, SLAMGPU_GROW_NONE, nullptr, &_snapshot));
    }
    auto oie = spe.occupancy_observation_probability_estimator()->impact_estimator();
    _lut.resize((size_t)w * h);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        _lut[(size_t)y * w + x] = oie->estimate_obstacle_impact(map[GridMap::Coord{x - org.x, y - org.y}]);
    double unknown = oie->estimate_obstacle_impact(*map.new_cell());
    _ctx->check(slamgpu_map_upload_lut(_snapshot, setup.generic_oie ? 0 : setup.oie, _lut.data(), unknown, w, h, org.x, org.y));
    return _snapshot;
  }
private:
  std::shared_ptr<Context> _ctx;
  slamgpu_map *_snapshot = nullptr;
  std::vector<double> _lut;
};

// filtered scan -> device (weights are the SPE's own ScanPointWeighting, evaluated once per scan)
class ScanBinding {
public:
  explicit ScanBinding(std::shared_ptr<Context> ctx) : _ctx{ctx} { _ctx->check(slamgpu_scan_create(_ctx->handle(), &_scan)); }
  ~ScanBinding() { slamgpu_scan_destroy(_scan); }
  slamgpu_scan *upload(const LaserScan2D &scan, const ScanPointWeighting &spw) {
    const auto &pts = scan.points();
    const std::size_t n = pts.size();
    _a.resize(n); _b.resize(n); _w.resize(n); _f.resize(n); _occ.resize(n);
    bool factor = false;
    for (std::size_t i = 0; i < n; ++i) {
      _a[i] = pts[i].range(); _b[i] = pts[i].angle();
      _w[i] = spw.weight(pts, i);
      _f[i] = pts[i].factor(); factor |= _f[i] != 1.0;
      _occ[i] = pts[i].is_occupied() ? 1 : 0;
    }
    _ctx->check(slamgpu_scan_upload(_scan, (int32_t)n, 0, _a.data(), _b.data(), _occ.data(), factor ? _f.data() : nullptr, _w.data()));
    return _scan;
  }
private:
  std::shared_ptr<Context> _ctx;
  slamgpu_scan *_scan = nullptr;
  std::vector<double> _a, _b, _w, _f;
  std::vector<uint8_t> _occ;
};

inline slamgpu_spe_params make_spe_params(const ScoreSetup &s, int trig_mode) {
  slamgpu_spe_params p;
  std::memset(&p, 0, sizeof p);
  p.oope = s.oope; p.oie = s.generic_oie ? 0 : s.oie;
  p.trig_mode = trig_mode;
  p.gm_fullness_th = s.gm_fullness_th; p.gm_window = s.gm_window;
  p.gm_cache = s.oope == SLAMGPU_OOPE_GMAPPING ? s.gm_cache : 0;
  return p;
}

//============================================================================//
// Base of the CUDA matchers: PoseEnumerationScanMatcher::process_scan
// (pose_enumeration_scan_matcher.h:31-77) with the candidate loop on the device.
class CudaScanMatcherBase : public GridScanMatcher {
public:
  CudaScanMatcherBase(std::shared_ptr<Context> ctx, std::shared_ptr<ScanProbabilityEstimator> spe,
                      std::shared_ptr<ScanPointWeighting> spw, int trig_mode = SLAMGPU_TRIG_DEVICE)
    : GridScanMatcher{spe}, _ctx{ctx}, _spw{spw}, _setup{detect_score_setup(*spe)}, _trig_mode{trig_mode}
    , _map_binding{ctx}, _scan_binding{ctx} {}

  void set_score_setup(const ScoreSetup &s) { _setup = s; }
  // candidates scored on the device by the last process_scan (the initial pose included)
  std::size_t poses_tested() const { return _poses_tested; }

protected:
  struct Prepared { LaserScan2D scan; slamgpu_map *map; slamgpu_scan *dscan; slamgpu_spe_params params; };

  Prepared prepare(const TransformedLaserScan &raw_scan, const RobotPose &init_pose, const GridMap &map) {
    Prepared p;
    // WeightedMeanPointProbabilitySPE::filter_scan (host): point filter + resets the shared SPW
    p.scan = filter_scan(raw_scan.scan, init_pose, map);
    p.map = _map_binding.bind(map, *scan_probability_estimator(), _setup);
    p.dscan = _scan_binding.upload(p.scan, *_spw);
    p.params = make_spe_params(_setup, _trig_mode);
    return p;
  }
  // scores of a pose list, in order
  std::vector<double> score(const Prepared &p, const std::vector<RobotPose> &poses) {
    std::vector<double> flat(poses.size() * 3), out(poses.size());
    for (std::size_t i = 0; i < poses.size(); ++i) { flat[3 * i] = poses[i].x; flat[3 * i + 1] = poses[i].y; flat[3 * i + 2] = poses[i].theta; }
    int64_t idx; double best;
    _gm_states.clear();
    if (p.params.oope == SLAMGPU_OOPE_GMAPPING && p.params.gm_cache == 2) {
      // one estimator object, one cache: the list is scored as the sequence it would be evaluated in, starting from
      // the cache left by the last candidate really evaluated (see consumed())
      _gm_states.resize(poses.size());
      _ctx->check(slamgpu_score_poses_chained(_ctx->handle(), p.map, p.dscan, &p.params, flat.data(), (int64_t)poses.size(), &_gm_state,
                                              out.data(), _gm_states.data()));
    } else {
      _ctx->check(slamgpu_score_poses(_ctx->handle(), p.map, p.dscan, &p.params, flat.data(), (int64_t)poses.size(),
                                      -std::numeric_limits<double>::infinity(), out.data(), &idx, &best));
    }
    _poses_tested += poses.size();
    return out;
  }
  // the reference evaluated the candidates of the last score() call up to and including index k
  void consumed(std::size_t k) {
    if (k < _gm_states.size()) { _gm_state = _gm_states[k]; }
  }
  bool has_observers() {
    bool any = false;
    do_for_each_observer([&any](ObsPtr) { any = true; });
    return any;
  }

  std::shared_ptr<Context> _ctx;
  std::shared_ptr<ScanPointWeighting> _spw;
  ScoreSetup _setup;
  int _trig_mode;
  MapBinding _map_binding;
  ScanBinding _scan_binding;
  std::size_t _poses_tested = 0;
  slamgpu_gm_cache _gm_state{0, 0, -1.0};
  std::vector<slamgpu_gm_cache> _gm_states;
};

//============================================================================//
// Any copyable PoseEnumerator, by speculation.  The enumerator decides the next candidate from the
// current best pose and the accept/reject feedback, so candidates cannot be listed up front in
// general.  A copy of the enumerator is run ahead as if every candidate were rejected, the whole
// speculated list is scored in one launch, and the real enumerator is then replayed over the scores
// up to (and including) the first accepted candidate, where speculation restarts.  The replay makes
// exactly the calls the reference loop makes (next / feedback / observers, same order, same
// arguments), so enumerator quirks are inherited rather than re-implemented.  Batches = accepts + 1.
template <typename PE>
class CudaPoseEnumerationScanMatcher : public CudaScanMatcherBase {
public:
  CudaPoseEnumerationScanMatcher(std::shared_ptr<Context> ctx, std::shared_ptr<ScanProbabilityEstimator> spe,
                                 std::shared_ptr<ScanPointWeighting> spw, const PE &pe, int trig_mode = SLAMGPU_TRIG_DEVICE,
                                 std::size_t max_batch = 1 << 16)
    : CudaScanMatcherBase{ctx, spe, spw, trig_mode}, _pe{pe}, _max_batch{max_batch} {}

  void set_pose_enumerator(const PE &pe) { _pe = pe; }
  void reset_state() override { _pe.reset(); }

  double process_scan(const TransformedLaserScan &raw_scan, const RobotPose &init_pose, const GridMap &map,
                      RobotPoseDelta &pose_delta) override {
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_start(init_pose, raw_scan, map); });
    _poses_tested = 0;
    auto prep = prepare(raw_scan, init_pose, map);
    const LaserScan2D &scan = prep.scan;
    auto best_pose = init_pose;
    double best_pose_prob = score(prep, {best_pose})[0];
    consumed(0);
    do_for_each_observer([&](ObsPtr obs) {
      obs->on_scan_test(best_pose, scan, best_pose_prob);
      obs->on_pose_update(best_pose, scan, best_pose_prob);
    });
    _pe.reset();
    // A copy of an enumerator is only state-identical when it is taken right after reset(): the
    // Gaussian enumerator clones its random variables without their cached second variate
    // (random_utils.h:23-25).  So the run-ahead twin is always rebuilt from that clean snapshot by
    // replaying the calls the real enumerator has received since.
    const PE clean = _pe;
    struct Call { RobotPose best; bool accepted; };
    std::vector<Call> history;
    std::vector<RobotPose> batch;
    while (_pe.has_next()) {
      PE ahead = clean;
      for (const Call &c : history) { ahead.next(c.best); ahead.feedback(c.accepted); }
      // speculate: everything after the current state, assuming rejections only
      batch.clear();
      while (ahead.has_next() && batch.size() < _max_batch) {
        batch.push_back(ahead.next(best_pose));
        ahead.feedback(false);
      }
      auto probs = score(prep, batch);
      // replay on the real enumerator
      for (std::size_t k = 0; k < batch.size(); ++k) {
        const RobotPose arg = best_pose;
        auto sampled_pose = _pe.next(arg);
        if (sampled_pose.x != batch[k].x || sampled_pose.y != batch[k].y || sampled_pose.theta != batch[k].theta) {
          contract_violation("slamgpu: the pose enumerator is not reproducible from a copy taken after reset()");
        }
        double sampled_scan_prob = probs[k];
        do_for_each_observer([&](ObsPtr obs) { obs->on_scan_test(sampled_pose, scan, sampled_scan_prob); });
        auto pose_is_acceptable = best_pose_prob < sampled_scan_prob;
        _pe.feedback(pose_is_acceptable);
        history.push_back(Call{arg, pose_is_acceptable});
        consumed(k);
        if (!pose_is_acceptable) { continue; }
        best_pose_prob = sampled_scan_prob;
        best_pose = sampled_pose;
        do_for_each_observer([&](ObsPtr obs) { obs->on_pose_update(best_pose, scan, best_pose_prob); });
        break;  // the speculation past this point assumed a rejection
      }
    }
    pose_delta = best_pose - init_pose;
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_end(pose_delta, scan, best_pose_prob); });
    return best_pose_prob;
  }
protected:
  PE _pe;
private:
  std::size_t _max_batch;
};

using HillClimbingPoseEnumerator = FailedRoundsLimitedPoseEnumerator<Distorsion1DPoseEnumerator>;

// MonteCarloScanMatcher (monte_carlo_scan_matcher.h:84-100) on the device.  The Gaussian enumerator's pose shifts come
// from libstdc++'s normal_distribution and do not depend on the scores, only on how many were drawn and on the
// dispersion resets; so a copy of the enumerator samples the shifts ahead, slamgpu_match_mc runs the accept loop over
// them in one launch (re-basing at every accept), and only a dispersion reset -- an accept after more than a third of
// the failed-attempts budget -- brings the loop back to the host for the next list.  The real enumerator is advanced
// afterwards by replaying the recorded decisions (next / feedback / observers in the reference's order), so a match
// the device declines (host trig, border guard, ...) falls back to the speculative batches with nothing consumed.
class CudaMonteCarloScanMatcher : public CudaPoseEnumerationScanMatcher<GaussianPoseEnumerator> {
  using Base = CudaPoseEnumerationScanMatcher<GaussianPoseEnumerator>;
public:
  CudaMonteCarloScanMatcher(std::shared_ptr<Context> ctx, std::shared_ptr<ScanProbabilityEstimator> spe,
                            std::shared_ptr<ScanPointWeighting> spw, unsigned seed, double translation_dispersion,
                            double rotation_dispersion, unsigned failed_attempts_per_dispersion, unsigned total_attempts)
    : Base{ctx, spe, spw,
           GaussianPoseEnumerator{seed, translation_dispersion, rotation_dispersion, failed_attempts_per_dispersion, total_attempts}}
    , _max_failed{failed_attempts_per_dispersion}, _max_poses{total_attempts} {}

  double process_scan(const TransformedLaserScan &raw_scan, const RobotPose &init_pose, const GridMap &map,
                      RobotPoseDelta &pose_delta) override {
    if (_setup.oope == SLAMGPU_OOPE_GMAPPING && _setup.gm_cache == 2) { return Base::process_scan(raw_scan, init_pose, map, pose_delta); }
    auto prep = prepare(raw_scan, init_pose, map);
    const LaserScan2D &scan = prep.scan;
    GaussianPoseEnumerator clean = _pe;
    clean.reset();  // what process_scan's own reset() leaves: counters cleared, base dispersion, the engine where it is
    struct Step { RobotPose arg, pose; double prob; bool accepted; };
    std::vector<Step> steps;  // every candidate the reference loop would have tested, in order
    RobotPose best = init_pose;
    double best_prob = std::numeric_limits<double>::quiet_NaN(), init_prob = best_prob;
    bool have_best = false;
    unsigned failed = 0, poses = 0;
    std::vector<double> noise, log;
    for (;;) {
      if (have_best && !(failed < _max_failed && poses < _max_poses)) { break; }  // has_next()
      GaussianPoseEnumerator ahead = clean;  // a copy is state-identical only right after reset(): replay the history on it
      for (const Step &st : steps) { ahead.next(st.arg); ahead.feedback(st.accepted); }
      const int32_t K = (int32_t)(_max_poses - poses);
      noise.resize((std::size_t)3 * K);
      for (int32_t j = 0; j < K; ++j) {  // 0 + shift = the shift itself
        const RobotPose nz = ahead.next(RobotPose{0, 0, 0});
        noise[3 * j] = nz.x; noise[3 * j + 1] = nz.y; noise[3 * j + 2] = nz.theta;
      }
      log.resize((std::size_t)4 * (K + 1));
      const double b3[3] = {best.x, best.y, best.theta};
      double out[10];
      int32_t served = 0;
      _ctx->check(slamgpu_match_mc(_ctx->handle(), prep.map, prep.dscan, &prep.params, b3, best_prob, have_best ? 1 : 0, noise.data(), K,
                                   failed, poses, _max_failed, _max_poses, out, log.data(), K + 1, &served));
      if (!served) { return Base::process_scan(raw_scan, init_pose, map, pose_delta); }  // nothing was consumed from _pe
      const int32_t entries = (int32_t)out[9];
      int32_t e = 0;
      if (!have_best) { init_prob = best_prob = log[3]; have_best = true; e = 1; }
      for (; e < entries; ++e) {  // the same accept rule, to label the steps
        const double *l = log.data() + 4 * (std::size_t)e;
        const bool ok = best_prob < l[3];
        steps.push_back(Step{best, RobotPose{l[0], l[1], l[2]}, l[3], ok});
        if (ok) { best = RobotPose{l[0], l[1], l[2]}; best_prob = l[3]; }
      }
      if (best.x != out[0] || best.y != out[1] || best.theta != out[2] || !(best_prob == out[3] || (best_prob != best_prob && out[3] != out[3]))) {
        contract_violation("slamgpu: the Monte-Carlo log does not reproduce the device's result");
      }
      failed = (unsigned)out[5]; poses = (unsigned)out[6];
      if ((int32_t)out[4] == 0 && out[7] == 0) { break; }  // nothing consumed and no reset: the budget is spent
    }
    // ---- the reference's calls, in its order, on the real enumerator
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_start(init_pose, raw_scan, map); });
    _poses_tested = 1 + steps.size();
    do_for_each_observer([&](ObsPtr obs) {
      obs->on_scan_test(init_pose, scan, init_prob);
      obs->on_pose_update(init_pose, scan, init_prob);
    });
    _pe.reset();
    for (const Step &st : steps) {
      if (!_pe.has_next()) { contract_violation("slamgpu: the Monte-Carlo replay ran past the enumerator's budget"); }
      const RobotPose sampled = _pe.next(st.arg);
      if (sampled.x != st.pose.x || sampled.y != st.pose.y || sampled.theta != st.pose.theta) {
        contract_violation("slamgpu: the pose enumerator is not reproducible from a copy taken after reset()");
      }
      do_for_each_observer([&](ObsPtr obs) { obs->on_scan_test(sampled, scan, st.prob); });
      _pe.feedback(st.accepted);
      if (st.accepted) { do_for_each_observer([&](ObsPtr obs) { obs->on_pose_update(sampled, scan, st.prob); }); }
    }
    if (_pe.has_next()) { contract_violation("slamgpu: the Monte-Carlo replay stopped before the enumerator's budget was spent"); }
    pose_delta = best - init_pose;
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_end(pose_delta, scan, best_prob); });
    return best_prob;
  }
private:
  unsigned _max_failed, _max_poses;
};

// HillClimbingScanMatcher (hill_climbing_scan_matcher.h:128-170) on the device.  The enumerator is deterministic, so
// the whole match -- every round, the accept loop included -- runs in one call (slamgpu_match_hc: one launch for the
// obstacle / max / mean OOPEs); observers are replayed afterwards from the log of scored poses, in the reference's
// order.  The GMapping OOPE's carried cache rides along in the same launch.
class CudaHillClimbingScanMatcher : public CudaPoseEnumerationScanMatcher<HillClimbingPoseEnumerator> {
  using Base = CudaPoseEnumerationScanMatcher<HillClimbingPoseEnumerator>;
public:
  CudaHillClimbingScanMatcher(std::shared_ptr<Context> ctx, std::shared_ptr<ScanProbabilityEstimator> spe,
                              std::shared_ptr<ScanPointWeighting> spw, unsigned max_lookup_failed_attempts,
                              double translation_delta, double rotation_delta)
    : Base{ctx, spe, spw, HillClimbingPoseEnumerator{max_lookup_failed_attempts, translation_delta, rotation_delta}}
    , _max_failed{max_lookup_failed_attempts}, _tr{translation_delta}, _rot{rotation_delta} {}

  double process_scan(const TransformedLaserScan &raw_scan, const RobotPose &init_pose, const GridMap &map,
                      RobotPoseDelta &pose_delta) override {
    auto prep = prepare(raw_scan, init_pose, map);
    const bool observed = has_observers();
    const int32_t cap = observed ? 8192 : 0;
    _log.resize((std::size_t)4 * cap);
    const double init[3] = {init_pose.x, init_pose.y, init_pose.theta};
    double out[3], prob = 0;
    int64_t tested = 0;
    int32_t count = 0;
    const slamgpu_gm_cache gm_before = _gm_state;
    _ctx->check(slamgpu_match_hc(_ctx->handle(), prep.map, prep.dscan, &prep.params, init, _max_failed, _tr, _rot, out, &prob, &tested,
                                 cap ? _log.data() : nullptr, cap, &count, &_gm_state));  // (the GMapping OOPE's cache rides along)
    if (observed && (count < 0 || count > cap)) {
      // no (complete) log to replay: the speculative path calls the observers itself (from the cache state before the match)
      _gm_state = gm_before;
      return Base::process_scan(raw_scan, init_pose, map, pose_delta);
    }
    _poses_tested = (std::size_t)tested;
    const LaserScan2D &scan = prep.scan;
    if (observed) {
      do_for_each_observer([&](ObsPtr obs) { obs->on_matching_start(init_pose, raw_scan, map); });
      double best = 0;
      for (int32_t k = 0; k < count; ++k) {
        const double *l = _log.data() + 4 * (std::size_t)k;
        const RobotPose pose{l[0], l[1], l[2]};
        const double p = l[3];
        do_for_each_observer([&](ObsPtr obs) { obs->on_scan_test(pose, scan, p); });
        if (k > 0 && !(best < p)) { continue; }
        best = p;
        do_for_each_observer([&](ObsPtr obs) { obs->on_pose_update(pose, scan, p); });
      }
    }
    pose_delta = RobotPose{out[0], out[1], out[2]} - init_pose;
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_end(pose_delta, scan, prob); });
    return prob;
  }
private:
  unsigned _max_failed;
  double _tr, _rot;
  std::vector<double> _log;
};

//============================================================================//
// BruteForceScanMatcher (brute_force_scan_matcher.h:68-80) on the device.  The candidate set of
// BruteForcePoseEnumerator is the Cartesian product of three value lists, each built by FP
// accumulation with its own stop rule (:42-54); it ignores the feedback, so the whole set is scored by
// one launch of the grid kernel and the accept loop collapses to its arg-max (lowest index on ties).
// Observers, if any, still get on_scan_test for every candidate in order.
class CudaBruteForceScanMatcher : public CudaScanMatcherBase {
public:
  CudaBruteForceScanMatcher(std::shared_ptr<Context> ctx, std::shared_ptr<ScanProbabilityEstimator> spe,
                            std::shared_ptr<ScanPointWeighting> spw, double from_x, double to_x, double step_x, double from_y,
                            double to_y, double step_y, double from_t, double to_t, double step_t,
                            int trig_mode = SLAMGPU_TRIG_DEVICE)
    : CudaScanMatcherBase{ctx, spe, spw, trig_mode}
    , _dx{axis(from_x, to_x, step_x, false)}, _dy{axis(from_y, to_y, step_y, false)}, _dt{axis(from_t, to_t, step_t, true)} {}

  double process_scan(const TransformedLaserScan &raw_scan, const RobotPose &init_pose, const GridMap &map,
                      RobotPoseDelta &pose_delta) override {
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_start(init_pose, raw_scan, map); });
    _poses_tested = 0;
    auto prep = prepare(raw_scan, init_pose, map);
    const LaserScan2D &scan = prep.scan;
    auto best_pose = init_pose;
    double best_pose_prob = score(prep, {best_pose})[0];
    do_for_each_observer([&](ObsPtr obs) {
      obs->on_scan_test(best_pose, scan, best_pose_prob);
      obs->on_pose_update(best_pose, scan, best_pose_prob);
    });
    // the enumerator keeps the base pose of its first ever use (reset() does not clear it)
    if (!_base_pose_is_set) { _base_pose = init_pose; _base_pose_is_set = true; }
    std::vector<double> xs(_dx.size()), ys(_dy.size()), ts(_dt.size());
    for (std::size_t i = 0; i < xs.size(); ++i) xs[i] = _base_pose.x + _dx[i];
    for (std::size_t i = 0; i < ys.size(); ++i) ys[i] = _base_pose.y + _dy[i];
    for (std::size_t i = 0; i < ts.size(); ++i) ts[i] = _base_pose.theta + _dt[i];
    const std::size_t P = xs.size() * ys.size() * ts.size();
    const bool observed = has_observers();
    std::vector<double> probs(observed ? P : 0);
    int64_t idx = -1;
    double best = best_pose_prob;
    // every OOPE goes through the grid entry point: obstacle / max / mean are tabulated per cell (one gather per evaluation),
    // overlap and GMapping are expanded to a pose list on the device
    _ctx->check(slamgpu_score_grid(_ctx->handle(), prep.map, prep.dscan, &prep.params, xs.data(), (int32_t)xs.size(), ys.data(),
                                   (int32_t)ys.size(), ts.data(), (int32_t)ts.size(), best_pose_prob,
                                   observed ? probs.data() : nullptr, &idx, &best));
    _poses_tested += P;
    auto pose_at = [&](std::size_t i) {
      const std::size_t nx = xs.size(), ny = ys.size();
      return RobotPose{xs[i % nx], ys[(i / nx) % ny], ts[i / (nx * ny)]};
    };
    if (observed) {  // the reference's notifications, candidate by candidate
      for (std::size_t i = 0; i < P; ++i) {
        auto sampled_pose = pose_at(i);
        double sampled_scan_prob = probs[i];
        do_for_each_observer([&](ObsPtr obs) { obs->on_scan_test(sampled_pose, scan, sampled_scan_prob); });
        if (!(best_pose_prob < sampled_scan_prob)) { continue; }
        best_pose_prob = sampled_scan_prob;
        best_pose = sampled_pose;
        do_for_each_observer([&](ObsPtr obs) { obs->on_pose_update(best_pose, scan, best_pose_prob); });
      }
    } else if (idx >= 0) {
      best_pose = pose_at((std::size_t)idx);
      best_pose_prob = best;
    }
    pose_delta = best_pose - init_pose;
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_end(pose_delta, scan, best_pose_prob); });
    return best_pose_prob;
  }
private:
  // the values one axis of BruteForcePoseEnumerator takes: x / y are stepped while v < to (after the
  // value was used), theta is used while t <= to
  static std::vector<double> axis(double from, double to, double step, bool top_level) {
    std::vector<double> v;
    double cur = from;
    if (top_level) {
      while (cur <= to) { v.push_back(cur); cur += step; }
    } else {
      for (;;) {
        v.push_back(cur);
        if (!(cur < to)) { break; }
        cur += step;
      }
    }
    return v;
  }
  std::vector<double> _dx, _dy, _dt;
  bool _base_pose_is_set = false;
  RobotPose _base_pose;
};

//============================================================================//
// CudaPyramidGridMap: M3RSMRescalableGridMap<UnboundedPlainGridMap> on the device
// (src/core/scan_matchers/m3rsm_engine.h:17-131): the fine map plus its max-pyramid; scan insertion keeps
// every level up to date with the reference's incremental rule.
class CudaPyramidGridMap : public CudaGridMap {
public:
  CudaPyramidGridMap(std::shared_ptr<Context> ctx, std::shared_ptr<ObservationImpactEstimator> oie,
                     std::shared_ptr<GridCell> prototype, const GridMapParams &params = MapValues::gmp,
                     int grow = SLAMGPU_GROW_PLAIN)
    : CudaGridMap{ctx, prototype, params, grow} {
    int oie_id = SLAMGPU_OIE_DISCREPANCY;
    if (dynamic_cast<const OccupancyOIE *>(oie.get())) oie_id = SLAMGPU_OIE_OCCUPANCY;
    else if (!dynamic_cast<const DiscrepancyOIE *>(oie.get())) contract_violation("CudaPyramidGridMap: unknown OIE class");
    context()->check(slamgpu_pyramid_create(context()->handle(), raw_device_map(), oie_id, &_pyr));
  }
  ~CudaPyramidGridMap() override { slamgpu_pyramid_destroy(_pyr); }
  slamgpu_pyramid *pyramid() const { flush(); return _pyr; }
  int levels() const { flush(); return slamgpu_pyramid_levels(_pyr); }
  void update(const Coord &, const AreaOccupancyObservation &) override {
    contract_violation("CudaPyramidGridMap: cells are updated through a scan adder (whole scans)");
  }
  void reset(const Coord &, const GridCell &) override {
    contract_violation("CudaPyramidGridMap: cells are updated through a scan adder (whole scans)");
  }
protected:
  int64_t append_beams_on_device(int32_t n, const double *beams, const uint8_t *occ, const double *quality,
                                 const AdderParams &p) const override {
    int64_t cells = 0;
    context()->check(slamgpu_pyramid_append_beams(_pyr, n, beams, occ, quality, &p.est, p.blur, p.max_range, &cells));
    return cells;
  }
private:
  slamgpu_pyramid *_pyr = nullptr;
};

// BruteForceMultiResolutionScanMatcher (src/core/scan_matchers/bf_multi_res_scan_matcher.h:9-71) on the device:
// the branch-and-bound engine runs in the library, every batch of Match bounds is one K5 launch.
class CudaBfMultiResScanMatcher : public GridScanMatcher {
public:
  CudaBfMultiResScanMatcher(std::shared_ptr<Context> ctx, SPE est, std::shared_ptr<ScanPointWeighting> spw, double x_limit = 1,
                            double y_limit = 1, double rot_limit = deg2rad(5), double ang_step = deg2rad(0.1),
                            double transl_step = 0.05)
    : GridScanMatcher{est, x_limit, y_limit, rot_limit}, _ctx{ctx}, _spw{spw}, _ang_step{ang_step}, _transl_step{transl_step} {}

  void set_target_accuracy(double angle_step, double translation_step) { _ang_step = angle_step; _transl_step = translation_step; }

  double process_scan(const TransformedLaserScan &raw_scan, const RobotPose &pose, const GridMap &map,
                      RobotPoseDelta &result_pose_delta) override {
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_start(pose, raw_scan, map); });
    auto pm = dynamic_cast<const CudaPyramidGridMap *>(&map);
    if (!pm) { contract_violation("CudaBfMultiResScanMatcher needs a CudaPyramidGridMap"); }
    auto setup = detect_score_setup(*scan_probability_estimator());
    if (setup.oope != SLAMGPU_OOPE_MAX || setup.generic_oie) { contract_violation("CudaBfMultiResScanMatcher: max OOPE expected"); }
    auto fscan = filter_scan(raw_scan.scan, pose, map);
    const auto &pts = fscan.points();
    std::vector<double> r(pts.size()), a(pts.size()), w(pts.size());
    for (std::size_t i = 0; i < pts.size(); ++i) { r[i] = pts[i].range(); a[i] = pts[i].angle(); w[i] = _spw->weight(pts, i); }
    auto params = make_spe_params(setup, SLAMGPU_TRIG_HOST);
    params.prerotated = 1;
    const double p3[3] = {pose.x, pose.y, pose.theta};
    double delta[3], prob = 0;
    _ctx->check(slamgpu_match_m3rsm(pm->pyramid(), (int32_t)pts.size(), r.data(), a.data(), w.data(), p3, &params, max_x_error(),
                                    max_y_error(), max_th_error(), _ang_step, _transl_step, 0.0, delta, &prob, _stats));
    result_pose_delta = RobotPoseDelta{delta[0], delta[1], delta[2]};
    do_for_each_observer([&](ObsPtr obs) { obs->on_matching_end(result_pose_delta, fscan, prob); });
    return prob;
  }
  const int64_t *stats() const { return _stats; }  // matches scored, K5 calls, branches, rotations
private:
  std::shared_ptr<Context> _ctx;
  std::shared_ptr<ScanPointWeighting> _spw;
  double _ang_step, _transl_step;
  int64_t _stats[4] = {0, 0, 0, 0};
};

}  // namespace slamgpu
