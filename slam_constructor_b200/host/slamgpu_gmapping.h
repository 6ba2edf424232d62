// slamgpu_gmapping.h -- the GMapping particle filter with the per-particle loop on the device (K6).
//
// slamgpu::CudaGmappingParticleFilter : LaserScanGridWorld mirrors GmappingParticleFilter
// (src/slams/gmapping/gmapping_particle_filter.h:29-124) over GmappingWorld particles
// (src/slams/gmapping/gmapping_world.h:36-136), with every particle owning its OWN device map (the per-particle
// semantics GMapping is defined by; upstream's particles alias one map object, SURVEY quirk Q6 -- that aliased
// behaviour is what slamgpu::init_cuda_gmapping reproduces with the reference's own filter class).
//
// Per scan, instead of n x (hill climbing + scan insertion) run one after the other:
//   * the host keeps what upstream keeps per particle -- pose, raw odometry pose, weight, master flag, the mt19937
//     engine and the two pose-delta random variables (libstdc++ distributions: they cannot move to the device) --
//     and makes the same decisions in the same order (scan-matching gate, pose noise, weight product, N_eff test,
//     multinomial draw, master bookkeeping);
//   * all particles that scan-match this step hill-climb in lock step against their own maps
//     (slamgpu_particles_match_hc: one launch per hill-climbing round for all of them);
//   * all maps take the scan in one batched insertion (slamgpu_particles_append_scan);
//   * resampling moves / copies maps on the device (slamgpu_particles_resample).
// Nothing here scores a pose or updates a cell on the CPU.
#pragma once

#include <algorithm>
#include <chrono>
#include <functional>
#include <random>
#include <unordered_set>

#include "slamgpu_backend.h"
#include "src/core/states/laser_scan_grid_world.h"
#include "src/slams/gmapping/gmapping_world.h"

namespace slamgpu {

class CudaGmappingParticleFilter : public LaserScanGridWorld {
public:
  using RandomEngine = std::mt19937;
  struct Properties {
    GridMapParams map_params = MapValues::gmp;
    std::shared_ptr<ScanProbabilityEstimator> spe;   // WeightedMeanPointProbabilitySPE over GmappingOccupancyObservationPE
    std::shared_ptr<ScanPointWeighting> spw;          // the SPE's own weighting object
    // the GMapping OOPE with its cache carried from candidate to candidate as the estimator object does (the setup
    // detect_score_setup() gives that estimator); any other OOPE is refused by the constructor
    ScoreSetup setup = [] { ScoreSetup s; s.oope = SLAMGPU_OOPE_GMAPPING; s.gm_cache = 2; return s; }();
    CudaScanAdder::Properties adder;
    unsigned max_failed_rounds = 6;                   // HillClimbingScanMatcher(spe, 6, 0.1, 0.1), init_gmapping.h:58-60
    double translation_delta = 0.1, rotation_delta = 0.1;
    unsigned particles = 30;
    // A particle's first scan is matched against its still empty map, so its probability is 0 and upstream's
    // `weight = scan_prob * weight` (gmapping_world.h:98) would zero the particle for good.  Upstream never notices:
    // its particles alias ONE map (quirk Q6), which the master has already filled.  With a map per particle the
    // first scan therefore only registers the map and leaves the weight alone (as in the original GMapping);
    // false = upstream's literal arithmetic (weights degenerate to 0/0 once every particle has matched once).
    bool first_scan_keeps_weight = true;
    // upstream seeds every particle's engine and every resampling draw from std::random_device
    // (gmapping_world.h:48, particle_filter.h:49-50); tests pass a deterministic source instead
    std::function<std::uint32_t()> seed_source = [] { return (std::uint32_t)std::random_device{}(); };
  };

  CudaGmappingParticleFilter(std::shared_ptr<Context> ctx, const Properties &props, const GMappingParams &gparams)
    : _ctx{ctx}, _props{props}, _scan_binding{ctx}, _raw_binding{ctx} {
    if (!props.spe || !props.spw || props.particles == 0) { contract_violation("CudaGmappingParticleFilter: incomplete properties"); }
    if (props.setup.oope != SLAMGPU_OOPE_GMAPPING) { contract_violation("CudaGmappingParticleFilter: the GMapping OOPE is the only one this filter scores with"); }
    const auto &mp = props.map_params;
    _ctx->check(slamgpu_particles_create(_ctx->handle(), (int32_t)props.particles, mp.width_cells, mp.height_cells, mp.meters_per_cell,
                                         SLAMGPU_CELL_GMAPPING, SLAMGPU_GROW_TILED, nullptr, &_parts));
    _particles.reserve(props.particles);
    for (unsigned i = 0; i < props.particles; ++i) {
      _particles.push_back(std::make_shared<Particle>(gparams, _props.seed_source()));
      _particles.back()->weight = 1.0 / props.particles;  // ParticleFilter::ParticleFilter, particle_filter.h:80-84
    }
    // GmappingParticleFilter::GmappingParticleFilter: sample() every particle, the heaviest becomes the master
    for (auto &p : _particles) { p->is_master = false; }
    _particles[heaviest()]->mark_master();
    std::memset(&_est, 0, sizeof _est);
    _est.type = props.adder.estimator;
    _est.occ_p = props.adder.base_occupied.prob_occ; _est.occ_q = props.adder.base_occupied.estimation_quality;
    _est.empty_p = props.adder.base_empty.prob_occ; _est.empty_q = props.adder.base_empty.estimation_quality;
    _est.low_qual = props.adder.low_qual; _est.unknown_qual = props.adder.unknown_qual;
    _est.shift_amount = -1;
    _view = std::make_shared<CudaGridMap>(_ctx, std::make_shared<GmappingBaseCell>(), mp, SLAMGPU_GROW_TILED,
                                          slamgpu_particles_map(_parts, (int32_t)heaviest()));
    _view_of = heaviest();
  }
  ~CudaGmappingParticleFilter() override {
    _view.reset();
    slamgpu_particles_destroy(_parts);
  }
  CudaGmappingParticleFilter(const CudaGmappingParticleFilter &) = delete;
  CudaGmappingParticleFilter &operator=(const CudaGmappingParticleFilter &) = delete;

  // ---- GmappingParticleFilter's interface
  void handle_sensor_data(TransformedLaserScan &scan) override {
    update_robot_pose(scan.pose_delta);
    handle_observation(scan);
    notify_with_pose(pose());
    notify_with_map(map());
  }
  void update_robot_pose(const RobotPoseDelta &delta) override {
    for (auto &p : _particles) { p->update_robot_pose(delta); }
    _traversed_since_last_resample += delta.abs();
  }
  const LaserScanGridWorld &world() const override { return *this; }
  const RobotPose &pose() const override { return _particles[heaviest()]->pose; }
  const LaserScanGridWorld::MapType &map() const override {
    const std::size_t h = heaviest();
    if (h != _view_of) { _view->rebind(slamgpu_particles_map(_parts, (int32_t)h)); _view_of = h; }
    return *_view;
  }

  // ---- introspection (tests, publishers)
  std::size_t particles_nm() const { return _particles.size(); }
  const RobotPose &particle_pose(std::size_t i) const { return _particles[i]->pose; }
  double particle_weight(std::size_t i) const { return _particles[i]->weight; }
  bool particle_is_master(std::size_t i) const { return _particles[i]->is_master; }
  slamgpu_map *particle_map(std::size_t i) const { return slamgpu_particles_map(_parts, (int32_t)i); }
  std::size_t heaviest_particle() const { return heaviest(); }
  std::size_t resamplings() const { return _resamplings; }
  // device work of the last handle_observation: particles matched, hill-climbing candidates scored, cells updated
  struct StepStats {
    std::size_t matched = 0;
    int64_t poses_tested = 0, cells_updated = 0;
    double match_ms = 0, insert_ms = 0, resample_ms = 0;  // wall clock of the three device phases
  };
  const StepStats &last_step() const { return _last; }

protected:
  void handle_observation(TransformedLaserScan &obs) override {
    const std::size_t n = _particles.size();
    _last = StepStats{};
    // ---- GmappingWorld::handle_observation, first half, per particle (host): the gate and the pose noise
    std::vector<uint8_t> active(n, 0);
    for (std::size_t i = 0; i < n; ++i) {
      Particle &p = *_particles[i];
      if (p.delta_since_last_sm.sq_dist() < p.next_sm_delta.sq_dist() &&
          std::fabs(p.delta_since_last_sm.theta) < p.next_sm_delta.theta) { continue; }
      active[i] = 1;
      ++_last.matched;
      if (!p.scan_is_first) { p.pose += p.pose_guess_rv.sample(p.rnd_engine); }
    }
    if (_last.matched) {
      // ---- scan matching, all active particles in lock step.  filter_scan: particle maps are unbounded
      // (has_cell is always true), so one filtered scan serves every particle
      LaserScan2D filtered = _props.spe->filter_scan(obs.scan, _particles[first_active(active)]->pose, *_view);
      slamgpu_scan *dscan = _scan_binding.upload(filtered, *_props.spw);
      slamgpu_spe_params params = make_spe_params(_props.setup, SLAMGPU_TRIG_DEVICE);
      std::vector<double> init(3 * n), best(3 * n), probs(n);
      std::vector<int64_t> tested(n);
      for (std::size_t i = 0; i < n; ++i) { init[3 * i] = _particles[i]->pose.x; init[3 * i + 1] = _particles[i]->pose.y; init[3 * i + 2] = _particles[i]->pose.theta; }
      const auto t_match = std::chrono::steady_clock::now();
      _ctx->check(slamgpu_particles_match_hc(_parts, dscan, &params, init.data(), active.data(), _props.max_failed_rounds,
                                             _props.translation_delta, _props.rotation_delta, best.data(), probs.data(), tested.data()));
      _last.match_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_match).count();
      // ---- second half: pose correction, map update decision, weight
      std::vector<uint8_t> do_update(n, 0);
      std::vector<double> poses(3 * n);
      bool any_update = false;
      for (std::size_t i = 0; i < n; ++i) {
        Particle &p = *_particles[i];
        if (active[i]) {
          _last.poses_tested += tested[i];
          // pose_delta = best - init (pose_enumeration_scan_matcher.h:67-69), then pose += pose_delta
          p.pose += RobotPoseDelta{best[3 * i] - init[3 * i], best[3 * i + 1] - init[3 * i + 1], best[3 * i + 2] - init[3 * i + 2]};
          const bool was_first = p.scan_is_first;
          if (0.0 < probs[i] || p.scan_is_first) { do_update[i] = 1; any_update = true; p.scan_is_first = false; }
          if (!(was_first && _props.first_scan_keeps_weight)) { p.weight = probs[i] * p.weight; }
          p.reset_scan_matching_delta();
        }
        poses[3 * i] = p.pose.x; poses[3 * i + 1] = p.pose.y; poses[3 * i + 2] = p.pose.theta;
      }
      if (any_update) {
        // scan_adder()->append_scan(map(), pose(), scan.scan, scan.quality, 0): the RAW scan, per-point mapping quality
        // from the adder's OMQE (grid_map_scan_adders.h:54-75)
        const auto &pts = obs.scan.points();
        _even.reset(obs.scan);  // EvenSPW::_common_weight is set by reset() only (weighted_mean_point_probability_spe.h:21-32)
        slamgpu_scan *raw = _raw_binding.upload(obs.scan, _even);
        std::vector<double> pq(pts.size());
        _props.adder.observation_quality_estimator->reset(obs.scan);
        for (std::size_t k = 0; k < pts.size(); ++k) { pq[k] = _props.adder.observation_quality_estimator->quality(pts, k); }
        std::vector<int64_t> cells(n);
        const auto t_ins = std::chrono::steady_clock::now();
        _ctx->check(slamgpu_particles_append_scan(_parts, raw, poses.data(), do_update.data(), obs.quality, 0, &_est,
                                                  _props.adder.blur_distance, _props.adder.max_usable_range, pq.data(), cells.data()));
        _last.insert_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_ins).count();
        for (int64_t c : cells) { _last.cells_updated += c; }
        _view->touched();
      }
    }
    // NB (upstream): weights are updated during scan update
    normalize_weights();
    try_resample();
  }

private:
  // the host half of a GmappingWorld (gmapping_world.h:36-136)
  struct Particle {
    Particle(const GMappingParams &g, std::uint32_t seed)
      : rnd_engine(seed), pose_guess_rv{g.pose_guess_rv}, next_sm_delta_rv{g.next_sm_delta_rv} { reset_scan_matching_delta(); }
    void update_robot_pose(const RobotPoseDelta &delta) {  // :59-73
      auto d_th = (pose - raw_odom_pose).theta;
      auto s = std::sin(d_th), c = std::cos(d_th);
      auto corrected_delta = RobotPoseDelta{c * delta.x - s * delta.y, s * delta.x + c * delta.y, delta.theta};
      raw_odom_pose += delta;
      delta_since_last_sm += corrected_delta.abs();
      pose += corrected_delta;
    }
    void mark_master() {  // :103-110
      using GRV1D = GaussianRV1D<RandomEngine>;
      is_master = true;
      pose_guess_rv = RobotPoseDeltaRV<RandomEngine>{GRV1D{0, 0}, GRV1D{0, 0}, GRV1D{0, 0}};
      next_sm_delta_rv = RobotPoseDeltaRV<RandomEngine>{GRV1D{0, 0}, GRV1D{0, 0}, GRV1D{0, 0}};
    }
    void reset_scan_matching_delta() {  // :116-119
      delta_since_last_sm.reset();
      next_sm_delta = next_sm_delta_rv.sample(rnd_engine);
    }
    RobotPose pose{0, 0, 0}, raw_odom_pose{0, 0, 0};
    double weight = 1.0;
    bool is_master = false, scan_is_first = true;
    RandomEngine rnd_engine;
    RobotPoseDeltaRV<RandomEngine> pose_guess_rv, next_sm_delta_rv;
    RobotPoseDelta delta_since_last_sm, next_sm_delta;
  };

  static std::size_t first_active(const std::vector<uint8_t> &a) {
    return (std::size_t)(std::find(a.begin(), a.end(), 1) - a.begin());
  }
  std::size_t heaviest() const {  // ParticleFilter::heaviest_particle, particle_filter.h:114-121 (the last of equals)
    std::size_t best = 0;
    bool have = false;
    for (std::size_t i = 0; i < _particles.size(); ++i) {
      if (have && _particles[i]->weight < _particles[best]->weight) { continue; }
      best = i; have = true;
    }
    return best;
  }
  void normalize_weights() {  // particle_filter.h:108-112
    double total_weight = 0;
    for (auto &p : _particles) { total_weight += p->weight; }
    for (auto &p : _particles) { p->weight = p->weight / total_weight; }
  }
  bool try_resample() {  // gmapping_particle_filter.h:88-100 over particle_filter.h:86-106
    // (the second test is upstream's `std::fabs(theta <= 0.2)`: the comparison, not the angle, goes through fabs)
    if (_traversed_since_last_resample.sq_dist() <= 0.5 && std::fabs(_traversed_since_last_resample.theta <= 0.2)) { return false; }
    const std::size_t n = _particles.size();
    double sq_sum = 0;  // UniformResamling::resampling_is_required :34-43
    for (auto &p : _particles) { sq_sum += p->weight * p->weight; }
    double effective_particles_cnt = 1.0 / sq_sum;
    if (!(effective_particles_cnt * 2 < n)) { return false; }
    std::vector<unsigned> inds(n);  // UniformResamling::resample :45-66 (value-initialised: 0 when no prefix exceeds the sample)
    std::mt19937 engine(_props.seed_source());
    std::uniform_real_distribution<> uniform_distr(0, 1);
    for (std::size_t i = 0; i < n; i++) {
      double sample = uniform_distr(engine);
      double total_w = 0;
      for (std::size_t j = 0; j < n; j++) {
        total_w += _particles[j]->weight;
        if (sample < total_w) { inds[i] = (unsigned)j; break; }
      }
    }
    std::vector<std::shared_ptr<Particle>> next;
    std::unordered_set<unsigned> seen;
    std::vector<int32_t> src(n);
    for (std::size_t i = 0; i < n; ++i) {
      src[i] = (int32_t)inds[i];
      std::shared_ptr<Particle> sampled = _particles[inds[i]];  // the first draw of a particle keeps the object itself ...
      if (seen.count(inds[i])) {                                // ... later ones are copies, sample()d (:93-98)
        sampled = std::make_shared<Particle>(*sampled);
        sampled->is_master = false;
      } else {
        seen.insert(inds[i]);
      }
      next.push_back(sampled);
    }
    const auto t_res = std::chrono::steady_clock::now();
    _ctx->check(slamgpu_particles_resample(_parts, src.data()));
    _last.resample_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_res).count();
    _particles = std::move(next);
    normalize_weights();
    ++_resamplings;
    _traversed_since_last_resample.reset();
    // ensure_master_exists :103-115
    bool master_survived = std::any_of(_particles.begin(), _particles.end(), [](const std::shared_ptr<Particle> &p) { return p->is_master; });
    if (!master_survived) { _particles[heaviest()]->mark_master(); }
    _view_of = (std::size_t)-1;  // the maps moved: re-bind the published view on the next map()
    return true;
  }

  std::shared_ptr<Context> _ctx;
  Properties _props;
  slamgpu_particles *_parts = nullptr;
  std::vector<std::shared_ptr<Particle>> _particles;
  RobotPoseDelta _traversed_since_last_resample;
  slamgpu_estimator _est;
  ScanBinding _scan_binding, _raw_binding;
  EvenSPW _even;
  mutable std::shared_ptr<CudaGridMap> _view;  // the heaviest particle's map as a GridMap
  mutable std::size_t _view_of = 0;
  std::size_t _resamplings = 0;
  StepStats _last;
};

}  // namespace slamgpu
