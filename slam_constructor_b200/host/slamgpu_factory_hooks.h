// slamgpu_factory_hooks.h -- what the reference's own factories call when a preset says `slam/backend=cuda`.
//
// integration/slam_backend_key.patch adds three early returns to src/utils/init_scan_matching.h:193-218 and
// src/utils/init_occupancy_mapping.h:82-93,126-146:
//     if (slamgpu_hooks::wants_cuda(props)) return slamgpu_hooks::scan_matcher(props);
// This header is all the patched files include: it knows the reference's interfaces only (no CUDA, no libslamgpu).  The
// application links the plug-ins by including slamgpu_init.h and calling slamgpu::install_backend(ctx) once; a preset
// that asks for the CUDA back end in a binary that never installed it stops with the factories' usual error + exit(-1).
#pragma once
#include <cstdlib>
#include <functional>
#include <iostream>
#include <memory>

class PropertiesProvider;
class GridScanMatcher;
class GridMap;
class GridCell;
class GridMapScanAdder;

namespace slamgpu_hooks {

struct Hooks {
  std::function<std::shared_ptr<GridScanMatcher>(const PropertiesProvider &)> scan_matcher;
  std::function<std::shared_ptr<GridMap>(const PropertiesProvider &, std::shared_ptr<GridCell>)> grid_map;
  std::function<std::shared_ptr<GridMapScanAdder>(const PropertiesProvider &)> scan_adder;
};
inline Hooks &hooks() { static Hooks h; return h; }

template <class Props>
inline bool wants_cuda(const Props &props) { return props.get_str("slam/backend", "cpu") == "cuda"; }

[[noreturn]] inline void not_installed(const char *what) {
  std::cerr << "slam/backend=cuda: no CUDA " << what << " factory installed (call slamgpu::install_backend(ctx))" << std::endl;
  std::exit(-1);
}
inline std::shared_ptr<GridScanMatcher> scan_matcher(const PropertiesProvider &props) {
  if (!hooks().scan_matcher) not_installed("scan matcher");
  return hooks().scan_matcher(props);
}
inline std::shared_ptr<GridMap> grid_map(const PropertiesProvider &props, std::shared_ptr<GridCell> area_model) {
  if (!hooks().grid_map) not_installed("grid map");
  return hooks().grid_map(props, area_model);
}
inline std::shared_ptr<GridMapScanAdder> scan_adder(const PropertiesProvider &props) {
  if (!hooks().scan_adder) not_installed("scan adder");
  return hooks().scan_adder(props);
}

}  // namespace slamgpu_hooks
