// mapping.h -- structures shared by mapping.cu (K2/K3) and pyramid.cu (K4/K5).
#pragma once
#include <vector>

#include "internal.h"

struct BeamRec {  // host-prepared, one per scan point
  double wx, wy;      // beam end point in the world (libm trig on the host)
  double quality;     // scan_quality * per-point mapping quality
  double hole_sq;     // blur radius in cells, squared (0 for free points)
  double obst_sq;     // squared cell distance robot -> obstacle cell
  int obx, oby;       // obstacle cell
  int is_occ, active; // active: inside the margin and the range gate
  int map_id, pad;    // which map of a batched insertion (0 for a single map)
};
struct BeamOut {  // device-produced, one per scan point
  int count;          // ray-cast cells of this beam
  int lx, ly;         // last (obstacle) cell
  int pad;
  double base_p, base_q;  // occupancy estimate of the obstacle cell
};
struct BeamPlan {
  std::vector<BeamRec> beams;
  std::vector<long long> offsets;  // N + 1 slot offsets (upper bounds)
  long long M = 0;
  double px = 0, py = 0;
  int rx = 0, ry = 0;  // robot cell
};
// one map of a (possibly batched) insertion as the kernels see it
struct MapSlot {
  double *cells;
  double px, py;       // robot position the beams of this map start at
  double shift;        // the area estimator's Shift_Amount for this map's beams
  int w, h, ox, oy;
  unsigned key_base;   // first sort key of this map (maps are laid end to end in key space)
  int tw;              // tiled map: tiles per row
  double *const *tiles;  // tiled map (copy-on-write particle maps): tile pointers [th][tw], cells == NULL
  int rx, ry;          // the robot's own cell: every beam updates it once; the window around it is updated apart (k_apply_ring)
  int beam_begin, beam_end;  // this map's beams in the batch
};
struct GrowState {
  int w, h, ox, oy, grow;
  bool ensure_inside(int x, int y);
};
struct AppendTrace {  // what a pyramid needs from a scan insertion (device pointers valid until the next call)
  int oie = 0;
  long long M = 0;
  int64_t applied = 0;
  const int2 *cells = nullptr;           // per slot: external cell
  const unsigned *keys_sorted = nullptr; // per sorted position
  const unsigned *vals_sorted = nullptr; // per sorted position: slot id
  const double *impact = nullptr;        // per slot: impact of the cell right after this update
  const double *rec = nullptr;           // per slot: the cell record right after this update
  int N = 0;                             // beams
  const int *slot_beam = nullptr;
  const BeamOut *d_bout = nullptr;
  const long long *d_offsets = nullptr;
};

int sg_map_regrow(slamgpu_map *m, const GrowState &g);
int sg_prepare_beams(const slamgpu_map *m, const slamgpu_scan *s, const double pose[3], double scan_quality, int scan_margin,
                     double blur, double max_range, const double *point_quality, bool gate, BeamPlan *plan);
int sg_plan_from_beams(slamgpu_ctx *ctx, const slamgpu_map *map, int32_t n, const double *beams, const uint8_t *is_occ,
                       const double *quality, double blur, double max_range, BeamPlan *out);
// deferred (optional, pinned, 2 values per map): the call returns with its work queued -- the cell counts land there when the
// stream gets to them, the caller synchronises -- so that the host work of the next batch overlaps this batch's kernels
int sg_append_plans(slamgpu_ctx *ctx, slamgpu_map *const *maps, const BeamPlan *plans, int n, const slamgpu_estimator *est,
                    int64_t *cells_updated, AppendTrace *trace, unsigned long long *deferred = nullptr);
int sg_append_plan(slamgpu_ctx *ctx, slamgpu_map *map, const BeamPlan &plan, const slamgpu_estimator *est,
                   int64_t *cells_updated, AppendTrace *trace);
int sg_append_scan_impl(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3], double scan_quality,
                        int32_t scan_margin, const slamgpu_estimator *est, double blur, double max_range,
                        const double *point_quality, int64_t *cells_updated, AppendTrace *trace);

// stable LSD radix sort of (key, value) pairs on the ctx stream; *_sorted point at whichever of the
// two buffer pairs holds the result
int sg_radix_sort(slamgpu_ctx *ctx, unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, long long n,
                  unsigned max_key, unsigned **keys_sorted, unsigned **vals_sorted);
