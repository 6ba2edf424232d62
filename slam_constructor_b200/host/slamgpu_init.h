// slamgpu_init.h -- the reference's property-driven factories with the CUDA back end selected
// (SURVEY section 8, row f1).  Same keys and defaults as src/utils/init_scan_matching.h:193-218,
// src/utils/init_occupancy_mapping.h:82-146 and src/utils/init_slam.h:12-24, so the shipped presets
// (config/slams/tiny_slam_base.properties, viny_slam_base.properties, config/common/*.properties) select the GPU
// with no other change:  auto slam = slamgpu::init_cuda_1h_slam(props, ctx);
#pragma once

#include "slamgpu_backend.h"
#include "slamgpu_gmapping.h"
#include "slamgpu_factory_hooks.h"
#include "src/core/states/single_state_hypothesis_laser_scan_grid_world.h"
#include "src/utils/init_occupancy_mapping.h"
#include "src/utils/init_scan_matching.h"
#include "src/slams/gmapping/init_gmapping.h"
#include "src/utils/properties_providers.h"

namespace slamgpu {

struct CudaSpe {
  std::shared_ptr<ScanProbabilityEstimator> spe;
  std::shared_ptr<ScanPointWeighting> spw;  // the SPE's own weighting object (filter_scan resets it)
};

// init_spe (init_scan_matching.h:85-104), keeping hold of the weighting object
inline CudaSpe init_cuda_spe(const PropertiesProvider &props) {
  auto type = props.get_str(Slam_SM_NS + "spe/type", "<undefined>");
  if (type != "wmpp") {
    std::cerr << "Unknown Scan Probability Estimator type (" << Slam_SM_NS << "spe/type): " << type << std::endl;
    std::exit(-1);
  }
  const std::string WMPP_Prefix = Slam_SM_NS + "spe/wmpp";
  auto skip_rate = props.get_uint(WMPP_Prefix + "/sp_skip_rate", 0);
  auto max_range = props.get_dbl(WMPP_Prefix + "/sp_max_usable_range", -1);
  CudaSpe out;
  out.spw = init_swp(props);
  out.spe = std::make_shared<WeightedMeanPointProbabilitySPE>(init_oope(props), out.spw, skip_rate, max_range);
  return out;
}

// init_scan_matcher (init_scan_matching.h:193-218)
inline std::shared_ptr<GridScanMatcher> init_cuda_scan_matcher(const PropertiesProvider &props, std::shared_ptr<Context> ctx) {
  auto s = init_cuda_spe(props);
  auto sm_type = scan_matcher_type(props);
  std::cout << "Used Scan Matcher: " << sm_type << " (CUDA)" << std::endl;
  if (sm_type == "MC") {
    const std::string SM_NS = Slam_SM_NS + "MC/", DISP_NS = SM_NS + "dispersion/";
    auto seed = props.get_int(SM_NS + "seed", std::random_device{}());
    return std::make_shared<CudaMonteCarloScanMatcher>(ctx, s.spe, s.spw, seed, props.get_dbl(DISP_NS + "translation", 0.2),
                                                       props.get_dbl(DISP_NS + "rotation", 0.1),
                                                       props.get_uint(DISP_NS + "failed_attempts_limit", 20),
                                                       props.get_uint(SM_NS + "attempts_limit", 100));
  }
  if (sm_type == "HC") {
    // use_frame_alignement only rebuilds the enumerator upstream; the frame rotation itself is dropped there
    const std::string DIST_NS = Slam_SM_NS + "HC/distortion/";
    return std::make_shared<CudaHillClimbingScanMatcher>(ctx, s.spe, s.spw, props.get_uint(DIST_NS + "failed_attempts_limit", 6),
                                                         props.get_dbl(DIST_NS + "translation", 0.1),
                                                         props.get_dbl(DIST_NS + "rotation", 0.1));
  }
  if (sm_type == "BF") {
    const std::string SM_NS = Slam_SM_NS + "BF/";
    auto rng = [&](const char *dim, double limit, double step, double out[3]) {
      out[0] = props.get_dbl(SM_NS + dim + "/from", -limit);
      out[1] = props.get_dbl(SM_NS + dim + "/to", limit);
      out[2] = props.get_dbl(SM_NS + dim + "/step", step);
    };
    double x[3], y[3], t[3];
    rng("x", 0.5, 0.1, x); rng("y", 0.5, 0.1, y); rng("t", deg2rad(5), deg2rad(1), t);
    return std::make_shared<CudaBruteForceScanMatcher>(ctx, s.spe, s.spw, x[0], x[1], x[2], y[0], y[1], y[2], t[0], t[1], t[2]);
  }
  if (is_m3rsm(sm_type)) {
    const std::string SM_NS = Slam_SM_NS + "BF_M3RSM/";
    return std::make_shared<CudaBfMultiResScanMatcher>(ctx, s.spe, s.spw, props.get_dbl(SM_NS + "limits/x_translation", 1),
                                                       props.get_dbl(SM_NS + "limits/y_translation", 1),
                                                       props.get_dbl(SM_NS + "limits/rotation", deg2rad(5)),
                                                       props.get_dbl(SM_NS + "accuracy/rotation", deg2rad(0.1)),
                                                       props.get_dbl(SM_NS + "accuracy/translation", 0.05));
  }
  std::cerr << "Scan matcher type without a CUDA back end: " << sm_type << std::endl;
  std::exit(-1);
}

// init_grid_map (init_occupancy_mapping.h:126-146)
inline std::shared_ptr<GridMap> init_cuda_grid_map(const PropertiesProvider &props, std::shared_ptr<Context> ctx,
                                                   std::shared_ptr<GridCell> area_model = nullptr) {
  if (!area_model) area_model = init_occupied_area_model(props);
  auto map_params = init_grid_map_params(props);
  auto map_type = props.get_str("slam/mapping/grid/type", "<undefined>");
  int grow;
  if (map_type == "plain" || map_type == "lazy_tiled") grow = SLAMGPU_GROW_NONE;
  else if (map_type == "unbounded_plain") grow = SLAMGPU_GROW_PLAIN;
  else if (map_type == "unbounded_lazy_tiled") grow = SLAMGPU_GROW_TILED;
  else { std::cerr << "Unknown grid map type (slam/mapping/grid/type): " << map_type << std::endl; std::exit(-1); }
  if (is_m3rsm(scan_matcher_type(props))) {
    return std::make_shared<CudaPyramidGridMap>(ctx, init_oie(props, false), area_model, map_params, grow);
  }
  return std::make_shared<CudaGridMap>(ctx, area_model, map_params, grow);
}

// init_scan_adder (init_occupancy_mapping.h:82-93)
inline CudaScanAdder::Properties init_cuda_scan_adder_properties(const PropertiesProvider &props) {
  static const auto COE_NS = std::string{"slam/occupancy_estimator/"};
  CudaScanAdder::Properties p;
  p.base_occupied = Occupancy{props.get_dbl(COE_NS + "base_occupied/prob", 0.95), props.get_dbl(COE_NS + "base_occupied/qual", 1.0)};
  p.base_empty = Occupancy{props.get_dbl(COE_NS + "base_empty/prob", 0.01), props.get_dbl(COE_NS + "base_empty/qual", 1.0)};
  auto type = props.get_str(COE_NS + "type", "const");
  if (type == "const") p.estimator = SLAMGPU_EST_CONST;
  else if (type == "area") p.estimator = SLAMGPU_EST_AREA;
  else { std::cerr << "Unknown estimator type: " << type << std::endl; std::exit(-1); }
  p.observation_quality_estimator = init_omqe(props);
  p.blur_distance = props.get_dbl("slam/mapping/blur", 0.0);
  p.max_usable_range = props.get_dbl("slam/mapping/max_range", std::numeric_limits<double>::infinity());
  return p;
}
inline std::shared_ptr<GridMapScanAdder> init_cuda_scan_adder(const PropertiesProvider &props) {
  return std::make_shared<CudaScanAdder>(init_cuda_scan_adder_properties(props));
}

// `slam/backend=cuda` inside the reference's OWN factories: with integration/slam_backend_key.patch applied to
// src/utils/init_scan_matching.h / init_occupancy_mapping.h, init_scan_matcher / init_grid_map / init_scan_adder (and so
// init_1h_slam, init_gmapping, ...) hand the preset to these hooks when it carries that key -- no code change in the
// application beyond this one call.  (slamgpu_factory_hooks.h is what the patched headers include.)
inline void install_backend(std::shared_ptr<Context> ctx) {
  auto &h = slamgpu_hooks::hooks();
  h.scan_matcher = [ctx](const PropertiesProvider &props) { return init_cuda_scan_matcher(props, ctx); };
  h.grid_map = [ctx](const PropertiesProvider &props, std::shared_ptr<GridCell> area_model) { return init_cuda_grid_map(props, ctx, area_model); };
  h.scan_adder = [](const PropertiesProvider &props) { return init_cuda_scan_adder(props); };
}

// init_1h_slam (init_slam.h:12-24): the tinySLAM / vinySLAM world on the CUDA back end
inline std::shared_ptr<SingleStateHypothesisLaserScanGridWorld> init_cuda_1h_slam(const PropertiesProvider &props,
                                                                                   std::shared_ptr<Context> ctx) {
  auto slam_props = SingleStateHypothesisLSGWProperties{};
  double loc, raw;
  std::tie(loc, raw) = init_pose_quality_estimators(props);
  slam_props.localized_scan_quality = loc;
  slam_props.raw_scan_quality = raw;
  slam_props.grid_map = init_cuda_grid_map(props, ctx);
  slam_props.gsm = init_cuda_scan_matcher(props, ctx);
  slam_props.gmsa = init_cuda_scan_adder(props);
  return std::make_shared<SingleStateHypothesisLaserScanGridWorld>(slam_props);
}

// init_gmapping (src/slams/gmapping/init_gmapping.h:49-65): GmappingParticleFilter on the CUDA back end.  As upstream,
// every particle is built from the same properties (so they share one map object, quirk Q6); pose noise, weights and
// resampling stay in the reference's own particle filter.
inline std::shared_ptr<GmappingParticleFilter> init_cuda_gmapping(const PropertiesProvider &props, std::shared_ptr<Context> ctx) {
  const std::string OOPE_Pfx = "slam/scmtch/oope/";
  ScoreSetup setup;
  setup.oope = SLAMGPU_OOPE_GMAPPING; setup.gm_cache = 2;
  setup.gm_fullness_th = props.get_dbl(OOPE_Pfx + "custom/fullness_threshold", 0.1);
  setup.gm_window = (int)props.get_uint(OOPE_Pfx + "cutrom/window_size", 1);  // key spelled as upstream
  auto oope = std::make_shared<GmappingOccupancyObservationPE>(setup.gm_fullness_th, setup.gm_window);
  const std::string WMPP_Prefix = Slam_SM_NS + "spe/wmpp";
  auto spw = init_swp(props);
  auto spe = std::make_shared<WeightedMeanPointProbabilitySPE>(oope, spw, props.get_uint(WMPP_Prefix + "/sp_skip_rate", 0),
                                                                props.get_dbl(WMPP_Prefix + "/sp_max_usable_range", -1));
  auto matcher = std::make_shared<CudaHillClimbingScanMatcher>(ctx, spe, spw, 6, 0.1, 0.1);
  matcher->set_score_setup(setup);
  auto map = std::make_shared<CudaGridMap>(ctx, std::make_shared<GmappingBaseCell>(), init_grid_map_params(props), SLAMGPU_GROW_TILED);
  auto shw_params = SingleStateHypothesisLSGWProperties{1.0, 1.0, 0, map, matcher, init_cuda_scan_adder(props)};
  return std::make_shared<GmappingParticleFilter>(shw_params, init_gmapping_params(props), init_particles_nm(props));
}

// The same properties on the batched filter: every particle owns its own device map, all particles hill-climb in
// lock step and insert the scan in one batched pass (slamgpu_gmapping.h).
inline std::shared_ptr<CudaGmappingParticleFilter> init_cuda_gmapping_batched(const PropertiesProvider &props, std::shared_ptr<Context> ctx) {
  const std::string OOPE_Pfx = "slam/scmtch/oope/";
  CudaGmappingParticleFilter::Properties p;
  p.first_scan_keeps_weight = props.get_bool("slam/particles/first_scan_keeps_weight", true);  // key of this back end only
  p.setup.oope = SLAMGPU_OOPE_GMAPPING; p.setup.gm_cache = 2;
  p.setup.gm_fullness_th = props.get_dbl(OOPE_Pfx + "custom/fullness_threshold", 0.1);
  p.setup.gm_window = (int)props.get_uint(OOPE_Pfx + "cutrom/window_size", 1);
  auto oope = std::make_shared<GmappingOccupancyObservationPE>(p.setup.gm_fullness_th, p.setup.gm_window);
  const std::string WMPP_Prefix = Slam_SM_NS + "spe/wmpp";
  p.spw = init_swp(props);
  p.spe = std::make_shared<WeightedMeanPointProbabilitySPE>(oope, p.spw, props.get_uint(WMPP_Prefix + "/sp_skip_rate", 0),
                                                             props.get_dbl(WMPP_Prefix + "/sp_max_usable_range", -1));
  p.map_params = init_grid_map_params(props);
  p.adder = init_cuda_scan_adder_properties(props);
  p.particles = init_particles_nm(props);
  return std::make_shared<CudaGmappingParticleFilter>(ctx, p, init_gmapping_params(props));
}

}  // namespace slamgpu
