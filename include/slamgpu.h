/*
 * slamgpu.h -- C ABI of libslamgpu.so: the B200 (sm_100a) scan-scoring, grid-update
 * and max-pyramid engine for slam-constructor.
 *
 * This is the drop-in boundary.  Everything above it (the C++ plugin adapters in
 * slam_constructor_b200/host/, the ctypes binding in slam_constructor_b200/capi.py,
 * or a binding a maintainer adds to the reference) talks to the
 * GPU only through these entry points: plain pointers and sizes, caller-owned host
 * arrays, no C++ or torch types.  There is NO CPU fallback behind it: every compute
 * entry point runs hand-written CUDA kernels or fails with a negative status.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * upstream repository root, OSLL/slam-constructor).
 *
 * Conventions
 *   - return value: 0 = ok, negative = SLAMGPU_E_*; slamgpu_last_error(ctx) gives
 *     a message.  Nothing throws or aborts across the boundary (the reference uses
 *     assert / std::exit, src/utils/init_scan_matching.h:39-43).
 *   - one slamgpu_ctx per world (or per GPU rank), externally serialised: the
 *     reference core is single-threaded and not re-entrant.
 *   - calls are synchronous unless named *_launch (then slamgpu_*_fetch / slamgpu_sync).
 *   - a grid map is a dense row-major array cells[h][w][stride] of doubles plus
 *     (w, h, scale, origin): internal index = external cell + origin
 *     (src/core/maps/regular_squares_grid.h:120-143, plain_grid_map.h:40-44).
 *   - all floating point on the parity-critical path is IEEE double without FMA
 *     contraction, in the reference's operation order.
 */
#ifndef SLAMGPU_H
#define SLAMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAMGPU_ABI_VERSION 1

enum {
  SLAMGPU_OK = 0,
  SLAMGPU_E_INVALID = -1, /* bad argument */
  SLAMGPU_E_CUDA = -2,    /* CUDA runtime / driver error (message has the code) */
  SLAMGPU_E_NOMEM = -3,
  SLAMGPU_E_NCCL = -4,
  SLAMGPU_E_STATE = -5,   /* call order (e.g. fetch without launch) */
  SLAMGPU_E_NODEVICE = -6 /* no usable sm_100 device: there is no CPU fallback */
};

/* ---- cell models: record layouts (doubles), same as the reference's cell classes ----
 * LWW      {p, q, known}            GridCell base          src/core/maps/grid_cell.h:9-57
 * AFFINE   {p, known}               AffineQualityMergeCell src/core/maps/naive_grid_cells.h:6-21
 * MEAN     {p, n}                   MeanProbabilityCell    src/core/maps/naive_grid_cells.h:25-44
 * TBM_*    {p, q, u, e, o, known}   TbmBaseCell            src/core/maps/tbm_grid_cells.h:8-79
 * GMAPPING {p, ox, oy, hits, tries} GmappingBaseCell       src/slams/gmapping/gmapping_grid_cell.h:9-43
 */
enum {
  SLAMGPU_CELL_LWW = 0,
  SLAMGPU_CELL_AFFINE = 1,
  SLAMGPU_CELL_MEAN = 2,
  SLAMGPU_CELL_TBM_CONSISTENT = 3,
  SLAMGPU_CELL_TBM_UNKNOWN_EVEN = 4,
  SLAMGPU_CELL_GMAPPING = 5,
  SLAMGPU_CELL_CREDIBILIST = 6, /* src/slams/credibilist/grid_cell.h:10-68: TBM belief, disjunctive score */
  SLAMGPU_CELL_MODELS = 7
};
#define SLAMGPU_MAX_STRIDE 8

/* ObservationImpactEstimator: src/core/scan_matchers/observation_impact_estimators.h:14-28 */
enum { SLAMGPU_OIE_DISCREPANCY = 0, SLAMGPU_OIE_OCCUPANCY = 1 };
/* OccupancyObservationProbabilityEstimator: src/core/scan_matchers/occupancy_observation_probability.h:12-99,
 * src/slams/gmapping/gmapping_occupancy_observation_pe.h:11-45 */
enum {
  SLAMGPU_OOPE_OBSTACLE = 0,
  SLAMGPU_OOPE_MAX = 1,
  SLAMGPU_OOPE_MEAN = 2,
  SLAMGPU_OOPE_OVERLAP = 3,
  SLAMGPU_OOPE_GMAPPING = 4
};
/* map growth rule on update: PlainGridMap (bounded), UnboundedPlainGridMap (x1.2,
 * src/core/maps/plain_grid_map.h:133-173), UnboundedLazyTiledGridMap (128-cell tiles,
 * src/core/maps/lazy_tiled_grid_map.h:151-181) */
enum { SLAMGPU_GROW_NONE = 0, SLAMGPU_GROW_PLAIN = 1, SLAMGPU_GROW_TILED = 2 };
/* CellOccupancyEstimator: src/core/maps/const_occupancy_estimator.h:6-17, area_occupancy_estimator.h:27-240 */
enum { SLAMGPU_EST_CONST = 0, SLAMGPU_EST_AREA = 1 };
/* how beam trigonometry cos/sin(theta + angle) is produced for a candidate set:
 *   HOST   : libm on the host, once per distinct theta -- bit-identical to the reference
 *   DEVICE : CUDA sincos, once per distinct theta, with a boundary guard: any world
 *            point that lands within a guard band of a cell border is recomputed with
 *            HOST trig before the result is returned, so cell indices stay bit-exact */
enum { SLAMGPU_TRIG_DEVICE = 0, SLAMGPU_TRIG_HOST = 1 };
/* ScanPointWeighting: src/core/scan_matchers/weighted_mean_point_probability_spe.h:21-60 */
enum { SLAMGPU_SPW_EVEN = 0, SLAMGPU_SPW_VINY = 1, SLAMGPU_SPW_AHR = 2 };
/* ObservationMappingQualityEstimator: src/core/maps/grid_map_scan_adders.h:22-46 */
enum { SLAMGPU_OMQE_IDLE = 0, SLAMGPU_OMQE_AHR = 1 };

typedef struct slamgpu_ctx slamgpu_ctx;
typedef struct slamgpu_map slamgpu_map;
typedef struct slamgpu_scan slamgpu_scan;
typedef struct slamgpu_pyramid slamgpu_pyramid;
typedef struct slamgpu_particles slamgpu_particles;

/* SPEParams + the OOPE/OIE choice: src/core/scan_matchers/grid_scan_matcher.h:105-136 */
typedef struct slamgpu_spe_params {
  int32_t oope;          /* SLAMGPU_OOPE_* */
  int32_t oie;           /* SLAMGPU_OIE_* */
  double win_v, win_h;   /* sp_analysis_area side lengths (vside, hside), metres */
  int32_t prerotated;    /* SPEParams::scan_is_prerotated (scan must be Cartesian) */
  int32_t trig_mode;     /* SLAMGPU_TRIG_* */
  double gm_fullness_th; /* GMapping OOPE fullness threshold */
  int32_t gm_window;     /* GMapping OOPE half window, cells */
  int32_t gm_cache;      /* GMapping OOPE one-entry cache (gmapping_occupancy_observation_pe.h:21-24,37-38):
                            0 none (a pure function of the point), 1 restarted at every pose,
                            2 carried from pose to pose as upstream does (slamgpu_score_poses_chained, particles) */
} slamgpu_spe_params;

/* the state of that cache: the cell and the probability last computed; prob == -1: nothing cached yet */
typedef struct slamgpu_gm_cache {
  int32_t cx, cy;
  double prob;
} slamgpu_gm_cache;

/* CellOccupancyEstimator parameters (both kinds), src/utils/init_occupancy_mapping.h:42-80 */
typedef struct slamgpu_estimator {
  int32_t type; /* SLAMGPU_EST_* */
  int32_t reserved;
  double occ_p, occ_q;     /* base_occupied */
  double empty_p, empty_q; /* base_empty */
  double low_qual, unknown_qual;
  double shift_amount; /* area estimator's function-static Shift_Amount
                          (area_occupancy_estimator.h:71); <0: low_qual * cell side */
} slamgpu_estimator;

/* ------------------------------------------------------------------ context */
int slamgpu_abi_version(void);
/* number of visible CUDA devices with compute capability 10.x (0 on a CPU box) */
int slamgpu_device_count(void);
int slamgpu_ctx_create(int device, slamgpu_ctx **out);
/* one rank of an n-rank job (one process per GPU).  `nccl_id` is the 128-byte
 * ncclUniqueId made by slamgpu_nccl_unique_id on rank 0 and shipped to the others by
 * the launcher.  Candidate sets are sharded over ranks by contiguous index range and
 * merged through NVLink peer mailboxes inside the finalize kernel (one 32-byte-per-rank
 * NCCL all-gather where peer mapping is unavailable). */
int slamgpu_nccl_unique_id(void *id128);
int slamgpu_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id128, slamgpu_ctx **out);
void slamgpu_ctx_destroy(slamgpu_ctx *ctx);
const char *slamgpu_last_error(const slamgpu_ctx *ctx); /* ctx may be NULL: last create error */
int slamgpu_sync(slamgpu_ctx *ctx);
/* tuning knobs (defaults are right for production):
 *   "grid_variant"  highest brute-force grid kernel allowed: 0 = automatic (5 when the candidate set leaves the GPU
 *                   partly empty, 4 otherwise); 5 = variant 4's arithmetic behind a per-warp cp.async ring (LUT patches
 *                   and index records staged in shared memory many beams ahead; needs at most 4 cell rows under 8
 *                   consecutive y, else 4); 4 = every distinct map row gathered once per thread + indexed-branch
 *                   accumulate (needs ascending y, unit point factors and at most 8 cell rows under 8 consecutive y,
 *                   else 2); 2 = packed row words + L1-resident gathers; 3 = map patches staged in shared memory by TMA
 *                   bulk copies (kept for comparison: 2.3x slower than 2 at configs[2]); 1 = explicit row table.
 *                   Window OOPEs (max / mean) run on variant 1 over window tables, overlap / GMapping as a pose list.
 *   "warm_l2"       1 = stream the score LUT through L2 on a side stream before big grid launches (default 0)
 *   "p2p_timeout_ms" how long a rank waits for its peers' results in the fused exchange before the ctx falls back to
 *                   NCCL all-gathers (default 20000)
 *   "grid_rows"     rows per thread of variant 2: 0 = automatic, 2 / 4 / 8 */
int slamgpu_ctx_set_option(slamgpu_ctx *ctx, const char *name, int64_t value);
/* device-side stop watch on the ctx stream (CUDA events) for bench.py */
int slamgpu_timer_begin(slamgpu_ctx *ctx);
int slamgpu_timer_end(slamgpu_ctx *ctx, float *ms);
/* device time of the dominant kernel of the last compute call (the scoring kernel of K1, the
 * apply kernel of K3, ...), from CUDA events recorded around that one launch */
int slamgpu_last_kernel_ms(slamgpu_ctx *ctx, float *ms);
/* kernels launched by this ctx since creation (bench.py's gpu_launches) */
int64_t slamgpu_launch_count(const slamgpu_ctx *ctx);
/* write (and discard) a buffer larger than L2 on the ctx stream */
int slamgpu_flush_l2(slamgpu_ctx *ctx);

/* ------------------------------------------------------------------ grid map
 * replaces GridMap / PlainGridMap / UnboundedPlainGridMap / UnboundedLazyTiledGridMap storage
 * (src/core/maps/grid_map.h:22-74, plain_grid_map.h:13-190, lazy_tiled_grid_map.h:18-187) */
int slamgpu_model_stride(int model);
void slamgpu_default_unknown(int model, double *rec /* SLAMGPU_MAX_STRIDE */);
int slamgpu_map_create(slamgpu_ctx *ctx, int32_t w, int32_t h, double scale, int32_t model, int32_t grow,
                       const double *unknown_rec /* NULL: the model's default prototype */, slamgpu_map **out);
void slamgpu_map_destroy(slamgpu_map *m);
int slamgpu_map_info(const slamgpu_map *m, int32_t *w, int32_t *h, double *scale, int32_t *ox, int32_t *oy,
                     int32_t *stride);
/* replace the whole content (dims and origin may change); cells = h*w*stride doubles */
int slamgpu_map_upload(slamgpu_map *m, const double *cells, int32_t w, int32_t h, int32_t ox, int32_t oy);
int slamgpu_map_download(slamgpu_map *m, double *cells /* h*w*stride */);
/* GridMap::operator[] / update / reset for single cells (host mirror plumbing;
 * src/core/maps/grid_map.h:41-57).  (x, y) are external cell coordinates. */
int slamgpu_map_read_cell(slamgpu_map *m, int32_t x, int32_t y, double *rec);
int slamgpu_map_reset_cell(slamgpu_map *m, int32_t x, int32_t y, const double *rec);
int slamgpu_map_update_cell(slamgpu_map *m, int32_t x, int32_t y, int32_t aoo_is_occ, double aoo_p, double aoo_q,
                            double obst_x, double obst_y, double quality);
/* the per-cell observation impact ("score LUT") the scoring kernels gather:
 * OIE(cell, obstacle AOO {true,{1,1},.,1}), grid_scan_matcher.h:44-46 */
int slamgpu_map_lut_download(slamgpu_map *m, int32_t oie, double *lut /* h*w */, double *unknown_value);

/* score-only snapshot of a host-side map: `lut` holds ObservationImpactEstimator::estimate_obstacle_impact
 * of every cell (h*w, row major), evaluated by the caller with the reference's own classes, so any
 * GridCell / OIE type can be scored on the device; cell records are left unknown.  The next cell
 * update or upload invalidates it. */
int slamgpu_map_upload_lut(slamgpu_map *m, int32_t oie, const double *lut /* h*w */, double unknown_value, int32_t w,
                           int32_t h, int32_t ox, int32_t oy);

/* ------------------------------------------------------------------ laser scan
 * replaces LaserScan2D / ScanPoint2D (src/core/states/sensor_data.h:14-208) as the
 * kernels see it: filtered points with their pose-independent weights (the output of
 * WeightedMeanPointProbabilitySPE::filter_scan + ScanPointWeighting,
 * weighted_mean_point_probability_spe.h:21-95). */
int slamgpu_scan_create(slamgpu_ctx *ctx, slamgpu_scan **out);
void slamgpu_scan_destroy(slamgpu_scan *s);
int slamgpu_scan_upload(slamgpu_scan *s, int32_t n, int32_t cartesian, const double *a /* range | x */,
                        const double *b /* angle | y */, const uint8_t *occ /* NULL: all occupied */,
                        const double *factor /* NULL: 1.0 */, const double *weight /* NULL: 1/n */);

/* per-scan host preparation for callers that do not link the reference's C++ classes (O(n), libm, no GPU):
 * WeightedMeanPointProbabilitySPE::filter_scan + should_skip_point (weighted_mean_point_probability_spe.h:75-95,
 * 136-141) -> indices of the points kept, returns their count; the point weights of EvenSPW / VinySlamSPW /
 * AngleHistogramReciprocalSPW (:21-60) for the FILTERED points; the per-point mapping quality of IdleOMQE /
 * AngleHistogramResiprocalOMQE (grid_map_scan_adders.h:22-46) for the RAW points. */
int slamgpu_scan_filter(const slamgpu_map *m, int32_t n, const double *range, const double *angle, const uint8_t *occ,
                        const double pose[3], uint32_t skip_rate, double max_range, int32_t *keep_idx /* n */);
int slamgpu_point_weights(int32_t kind /* SLAMGPU_SPW_* */, int32_t n, const double *range, const double *angle, double *out_w);
int slamgpu_mapping_quality(int32_t kind /* SLAMGPU_OMQE_* */, int32_t n, const double *range, const double *angle,
                            double *out_q);

/* ------------------------------------------------------------------ K1: batched scan likelihood
 * replaces the candidate loop of PoseEnumerationScanMatcher::process_scan
 * (src/core/scan_matchers/pose_enumeration_scan_matcher.h:48-65) around
 * WeightedMeanPointProbabilitySPE::estimate_scan_probability
 * (weighted_mean_point_probability_spe.h:97-133): scores[k] = scan probability under
 * poses[k]; (best_idx, best_score) = the sequential strict-'<' accept loop started from
 * init_score, i.e. the lowest index among the maxima that beat init_score, or -1.
 * Under a dist ctx every rank passes the same full candidate set, scores its own
 * contiguous slice (out_scores is filled for that slice only) and gets the global best. */
int slamgpu_score_poses(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                        const double *poses /* 3*P: x, y, theta */, int64_t P, double init_score,
                        double *out_scores /* P or NULL */, int64_t *best_idx, double *best_score);
/* the same for the candidate set of BruteForcePoseEnumerator
 * (src/core/scan_matchers/brute_force_scan_matcher.h:10-64): the Cartesian product
 * thetas x ys x xs in the enumerator's order, index = (t*ny + y)*nx + x.  The axis
 * value lists are the enumerator's own accumulated values (base pose included). */
int slamgpu_score_grid(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                       const double *xs, int32_t nx, const double *ys, int32_t ny, const double *thetas, int32_t nt,
                       double init_score, double *out_scores /* nt*ny*nx or NULL */, int64_t *best_idx,
                       double *best_score);

/* GMapping OOPE with its cache carried along: the poses are scored as ONE sequence, in the order given, the way a
 * reference matcher evaluates its candidates one after the other -- pose k starts from the cache pose k-1 left,
 * pose 0 from *state_in ({0, 0, -1} for a new estimator object).  out_states[k] = the cache after pose k: pass the
 * state after the last candidate the matcher really consumed into the next call.  No arg-max: the accept rule of a
 * cached estimator is order-dependent and stays with the caller.  At most 8192 poses and 4 Mi pose-points a call. */
int slamgpu_score_poses_chained(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                                const double *poses, int64_t P, const slamgpu_gm_cache *state_in, double *out_scores /* P */,
                                slamgpu_gm_cache *out_states /* P or NULL */);
/* split form used by bench.py to time the device part with inputs resident in HBM:
 * stage a candidate set once, launch any number of times, fetch the last result */
int slamgpu_stage_poses(slamgpu_ctx *ctx, slamgpu_scan *scan, const slamgpu_spe_params *p, const double *poses,
                        int64_t P);
int slamgpu_stage_grid(slamgpu_ctx *ctx, slamgpu_scan *scan, const slamgpu_spe_params *p, const double *xs,
                       int32_t nx, const double *ys, int32_t ny, const double *thetas, int32_t nt);
int slamgpu_score_launch(slamgpu_ctx *ctx, slamgpu_map *map, double init_score);
int slamgpu_score_fetch(slamgpu_ctx *ctx, double *out_scores /* NULL ok */, int64_t *best_idx, double *best_score);
/* counters of the last scoring call: [0] guard hits (points re-done with host trig),
 * [1] kernel variant used (after slamgpu_stage_grid: 1..4 = grid kernel variant; after slamgpu_stage_poses: 0 list, 3 two-phase small batch, 4 fused one-launch small batch; 5 / 6 whole hill-climbing / Monte-Carlo match), [2] evaluations (poses*points) on this rank,
 * [3] first candidate index of this rank's slice, [4] slice length, [5] grid kernel rows per thread,
 * [6] 1 when the per-rank results were exchanged through peer memory (NVLink mailboxes), 0 for ncclAllGather / one rank */
int slamgpu_score_stats(const slamgpu_ctx *ctx, int64_t stats[8]);

/* ------------------------------------------------------------------ K2: ray casting
 * replaces RegularSquaresGrid::world_to_cells (src/core/maps/regular_squares_grid.h:56-101,
 * fail-over DiscreteSegment2D src/core/geometry_discrete_primitives.h:55-104) for every
 * beam of a scan seen from `pose`.  out_offsets[n+1] are prefix sums; out_cells holds
 * (x, y) pairs, external coordinates, in the reference's order.  Pass out_cells = NULL
 * (cap 0) to get only the offsets/total.  Bit-exact. */
int slamgpu_raycast(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3],
                    int64_t *out_offsets /* n+1 */, int32_t *out_cells /* 2*cap */, int64_t cap, int64_t *total);

/* the same for explicit world segments {bx, by, ex, ey} on a grid of cell size `scale`: the unit
 * interface world_to_cells(Segment2D) itself (the reference's golden vectors are stated on it) */
int slamgpu_raycast_segments(slamgpu_ctx *ctx, double scale, const double *segments /* 4*n */, int32_t n,
                             int64_t *out_offsets /* n+1 */, int32_t *out_cells /* 2*cap */, int64_t cap, int64_t *total);
/* batched CellOccupancyEstimator::estimate_occupancy(beam, cell_bounds, is_occupied) -> Occupancy
 * (src/core/maps/cell_occupancy_estimator.h:13-15): beams {bx, by, ex, ey}, cell_bounds
 * {bot, top, left, right}; out_pq {prob, quality}, (NaN, NaN) = Occupancy::invalid() */
int slamgpu_estimate_occupancy(slamgpu_ctx *ctx, const slamgpu_estimator *est, int32_t n, const double *beams /* 4*n */,
                               const double *cell_bounds /* 4*n */, const uint8_t *is_occ /* n */, double *out_pq /* 2*n */);

/* ------------------------------------------------------------------ K2+K3: scan insertion
 * replaces GridMapScanAdder::append_scan + WallDistanceBlurringScanAdder::handle_scan_point
 * (src/core/maps/grid_map_scan_adders.h:54-75, 138-189), the const/area occupancy estimators
 * and the cell models' operator+= with the reference's update order (beam index, then the
 * obstacle cell first).  `point_quality` (NULL: 1.0) is the per-point factor of the
 * ObservationMappingQualityEstimator (grid_map_scan_adders.h:22-46), a host precompute.
 * The scan here is the RAW scan (all points, occupied flags).  *cells_updated counts
 * ray-cast cells. */
int slamgpu_append_scan(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3],
                        double scan_quality, int32_t scan_margin, const slamgpu_estimator *est, double blur,
                        double max_range, const double *point_quality, int64_t *cells_updated);

/* the same for the beams of one scan given explicitly as world segments {bx, by, ex, ey} (all
 * starting at the robot position), with per-beam is_occupied and quality: exactly the arguments of
 * the pure virtual GridMapScanAdder::handle_scan_point(map, is_occ, scan_quality, beam)
 * (grid_map_scan_adders.h:85-86), which a drop-in adder queues per point and flushes in one call */
int slamgpu_append_beams(slamgpu_ctx *ctx, slamgpu_map *map, int32_t n, const double *beams /* 4*n */,
                         const uint8_t *is_occ /* n */, const double *quality /* n */, const slamgpu_estimator *est,
                         double blur, double max_range, int64_t *cells_updated);

/* ------------------------------------------------------------------ K4/K5: max-pyramid
 * replaces RescalableCachingGridMap + M3RSMRescalableGridMap
 * (src/core/maps/rescalable_caching_grid_map.h:16-200, src/core/scan_matchers/m3rsm_engine.h:17-131)
 * and the Match upper bound (m3rsm_engine.h:156-180). */
int slamgpu_pyramid_create(slamgpu_ctx *ctx, slamgpu_map *fine, int32_t oie, slamgpu_pyramid **out);
void slamgpu_pyramid_destroy(slamgpu_pyramid *p);
int slamgpu_pyramid_levels(slamgpu_pyramid *p);
int slamgpu_pyramid_level_info(slamgpu_pyramid *p, int32_t level, int32_t *w, int32_t *h, double *scale, int32_t *ox,
                               int32_t *oy);
/* rebuild every coarse level from the fine map (what the reference's incremental rule
 * yields for a map whose cells were each written once, or whose impacts only grew) */
int slamgpu_pyramid_build(slamgpu_pyramid *p);
int slamgpu_pyramid_level_download(slamgpu_pyramid *p, int32_t level, double *cells /* h*w*stride */,
                                   double *impact /* h*w or NULL */);
int slamgpu_pyramid_rescale(slamgpu_pyramid *p, double target_scale); /* level id, rescale :79-91 */
/* scan insertion through the pyramid (update :100-105 -> update_coarser_maps :101-126) */
int slamgpu_pyramid_append_scan(slamgpu_pyramid *p, slamgpu_scan *scan, const double pose[3], double scan_quality,
                                int32_t scan_margin, const slamgpu_estimator *est, double blur, double max_range,
                                const double *point_quality, int64_t *cells_updated);
/* the same for explicit beams (see slamgpu_append_beams): what a drop-in scan adder flushes */
int slamgpu_pyramid_append_beams(slamgpu_pyramid *p, int32_t n, const double *beams /* 4*n */, const uint8_t *is_occ,
                                 const double *quality, const slamgpu_estimator *est, double blur, double max_range,
                                 int64_t *cells_updated);
/* batched Match bounds: M windows, each over one of the pre-rotated scans.  windows =
 * {bot, top, left, right} per match (metres, relative to pose); bound[m] = scan probability
 * of scans[scan_id[m]] at the window centre on the level rescale(max side) with the `max`
 * OOPE over the window re-centred at each point. */
int slamgpu_score_windows(slamgpu_pyramid *p, slamgpu_scan *const *scans, int32_t n_scans, const int32_t *scan_id,
                          const double *windows /* 4*M */, int64_t M, const double pose[3],
                          const slamgpu_spe_params *spe, double *out_bounds /* M */);

/* BruteForceMultiResolutionScanMatcher::process_scan (src/core/scan_matchers/bf_multi_res_scan_matcher.h:24-66)
 * over M3RSMEngine (m3rsm_engine.h:252-365): best-first branch and bound over (rotation, translation
 * window).  The scan is the FILTERED polar scan (output of the SPE's filter_scan) with one weight per
 * point (NULL: 1/n); it is pre-rotated per rotation hypothesis on the host (libm) like upstream.
 * out_delta = {dx, dy, dtheta} of the first finest match popped, *out_prob its probability.
 * stats: [0] matches scored, [1] K5 calls, [2] branches, [3] rotation hypotheses. */
int slamgpu_match_m3rsm(slamgpu_pyramid *p, int32_t n, const double *range, const double *angle, const double *weight,
                        const double pose[3], const slamgpu_spe_params *spe, double x_limit, double y_limit, double rot_limit,
                        double ang_step, double transl_step, double max_finest_prob_diff, double out_delta[3],
                        double *out_prob, int64_t stats[4]);

/* ------------------------------------------------------------------ K6: GMapping particles
 * replaces the per-particle loop of GmappingParticleFilter::handle_observation
 * (src/slams/gmapping/gmapping_particle_filter.h:70-85): n particles, each with its OWN device map.
 * Pose noise, weights, N_eff and the resampling draw stay on the host (libstdc++ <random>,
 * src/slams/gmapping/gmapping_world.h:81-85, src/core/particle_filter.h:34-106).
 * On a distributed ctx (slamgpu_ctx_create_dist) the particles are sharded: rank r owns the contiguous range
 * [r*ceil(n/R), (r+1)*ceil(n/R)) with those maps on its GPU.  Every call below is collective: all ranks pass the
 * same arrays over all n particles and all receive the complete result (one NCCL all-gather per call); resample
 * ships maps whose source lives on another rank with ncclSend/ncclRecv.  slamgpu_particles_map returns NULL for a
 * particle of another rank. */
int slamgpu_particles_create(slamgpu_ctx *ctx, int32_t n, int32_t w, int32_t h, double scale, int32_t model, int32_t grow,
                             const double *unknown_rec, slamgpu_particles **out);
void slamgpu_particles_destroy(slamgpu_particles *p);
int slamgpu_particles_count(const slamgpu_particles *p);
/* Storage of the particle maps.  With SLAMGPU_GROW_TILED (UnboundedLazyTiledGridMap upstream) the maps share copy-on-write
 * 128 x 128 tiles out of one pool (LazyTiledGridMap, src/core/maps/lazy_tiled_grid_map.h:18-118: shared_ptr tiles cloned on
 * first write :57-71): a new map is a table of references to one all-unknown tile, scan insertion first makes the tiles
 * under the scan private, resampling copies tables.  stats: [0] 1 if tiled, [1] tiles in use, [2] tiles cloned so far,
 * [3] bytes per tile, [4] bytes of device memory held by the pool; of the last resampling: [5] bytes copied or received,
 * [6] tile references handed to copies instead, [7] device time in microseconds.  (SLAMGPU_DENSE_PARTICLES=1 in the
 * environment keeps round 1's dense per-particle arrays, for comparison.) */
int slamgpu_particles_tile_stats(const slamgpu_particles *p, int64_t stats[8]);
/* borrowed handle of particle i's map: valid for every slamgpu_map_* call until the next resample */
slamgpu_map *slamgpu_particles_map(slamgpu_particles *p, int32_t i);
/* scores[i*c + k] = scan probability of poses[i*c + k] on particle i's map; one launch for all particles */
int slamgpu_particles_score(slamgpu_particles *p, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                            const double *poses /* 3*n*c */, int32_t per_particle /* c */, double *out_scores /* n*c */);
/* HillClimbingScanMatcher::process_scan (src/core/scan_matchers/hill_climbing_scan_matcher.h:128-170) for
 * every active particle from its own initial pose against its own map, all particles in lock step
 * (one launch per hill-climbing round).  out_poses = best poses, out_probs = their probabilities,
 * out_tested = poses scored per particle (initial pose included). */
int slamgpu_particles_match_hc(slamgpu_particles *p, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                               const double *init_poses /* 3*n */, const uint8_t *active /* n or NULL */,
                               uint32_t max_failed_rounds, double translation_delta, double rotation_delta,
                               double *out_poses /* 3*n */, double *out_probs /* n */, int64_t *out_tested /* n or NULL */);
/* The same for one matcher and one map -- HillClimbingScanMatcher::process_scan as a single call (and, for the
 * obstacle / max / mean OOPEs with device trig, a single launch: the rounds run inside one thread block).  log, when
 * log_cap > 0, receives {x, y, theta, probability} of every pose scored, in evaluation order (initial pose first), so
 * that observers can be replayed; *log_count = entries produced (may exceed log_cap: the log was cut), or -1 when
 * the call took the round-by-round path, which keeps no log. */
int slamgpu_match_hc(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                     const double init_pose[3], uint32_t max_failed_rounds, double translation_delta, double rotation_delta,
                     double out_pose[3], double *out_prob, int64_t *out_tested, double *log /* 4*log_cap or NULL */,
                     int32_t log_cap, int32_t *log_count /* or NULL */,
                     slamgpu_gm_cache *gm_state /* in/out: the estimator's cache when spe->gm_cache == 2; NULL = a new estimator */);
/* A segment of MonteCarloScanMatcher::process_scan (GaussianPoseEnumerator, monte_carlo_scan_matcher.h:10-100) in one
 * launch.  The enumerator's noise is libstdc++'s and stays with the caller: noise[3k..3k+2] is the pose shift candidate k
 * would get (it does not depend on the scores), candidate k = best pose so far + noise[k].  The device runs the accept
 * loop with the enumerator's counters (failed attempts, poses tested) from the state given, until the budget is spent,
 * the K shifts are used up or an accept resets the dispersion (the noise that follows is then different: the caller
 * samples it and calls again).  have_best = 0: best_pose is the initial pose and is scored first.
 * out[10] = best x, y, theta, probability, shifts consumed, failed attempts, poses tested, 1 if the dispersion was
 * reset, guard hits, log entries; log = {x, y, theta, probability} of every pose scored, in order.
 * *served = 0: not covered by the device kernel (host trig, overlap OOPE, carried GMapping cache, very long scans) or a
 * device-trig border guard fired: nothing was consumed, use slamgpu_score_poses batches instead. */
int slamgpu_match_mc(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                     const double best_pose[3], double best_prob, int32_t have_best, const double *noise /* 3*K */, int32_t K,
                     uint32_t failed, uint32_t poses_nm, uint32_t max_failed, uint32_t max_poses, double out[10],
                     double *log /* 4*log_cap or NULL */, int32_t log_cap, int32_t *served);
/* GridMapScanAdder::append_scan into every particle's own map from its own pose (do_update NULL: all); the beams of
 * all particles go through one batched ray-cast, one sort keyed by (map, cell) and one ordered apply */
int slamgpu_particles_append_scan(slamgpu_particles *p, slamgpu_scan *scan, const double *poses /* 3*n */,
                                  const uint8_t *do_update /* n or NULL */, double scan_quality, int32_t scan_margin,
                                  const slamgpu_estimator *est, double blur, double max_range, const double *point_quality,
                                  int64_t *cells_updated /* n or NULL */);
/* the copy step of ParticleFilter::try_resample (src/core/particle_filter.h:92-98): particle i := particle src[i];
 * a source drawn once is moved, the others are copied device to device (or rank to rank) */
int slamgpu_particles_resample(slamgpu_particles *p, const int32_t *src /* n */);

/* ------------------------------------------------------------------ measurement aid
 * rate (loads/s) of uniformly random 8-byte loads over an L2-resident table of table_bytes: the gather roofline of
 * K1 when no two gathers share a sector (bench.py reports K1 against it next to the HBM figure) */
int slamgpu_probe_gather(slamgpu_ctx *ctx, int64_t table_bytes, int32_t loads_per_thread, double *gathers_per_s);

/* ------------------------------------------------------------------ test hook
 * out[i] = a[i] / b[i] through the division the TBM cell update uses (with its exact shortcut for subnormal
 * numerators): lets the parity tests hold it against the host's IEEE division. */
int slamgpu_debug_div(slamgpu_ctx *ctx, int32_t n, const double *a, const double *b, double *out);
/* the hill-climbing enumerator + accept loop (HillClimbingScanMatcher, hill_climbing_scan_matcher.h:10-170) that the
 * device kernel and the round-by-round path share, run on the host over a caller-supplied scoring function: no GPU
 * and no ctx needed, so CPU-only tests can hold it against the oracle's sequential matcher */
typedef double (*slamgpu_score_fn)(const double pose[3], void *user);
int slamgpu_debug_hill_climb(const double init_pose[3], uint32_t max_failed_rounds, double translation_delta,
                             double rotation_delta, slamgpu_score_fn score, void *user, double out_pose[3],
                             double *out_prob, int64_t *out_tested);

/* M3RSMEngine's best-first search (m3rsm_engine.h:252-365) as slamgpu_match_m3rsm runs it, over a caller-supplied
 * bound function: bounds[k] = Match bound of (rotations[k], windows[4k..4k+3] = bot, top, left, right) */
typedef void (*slamgpu_bounds_fn)(int32_t count, const double *rotations, const double *windows, double *out_bounds, void *user);
int slamgpu_debug_m3rsm(double x_limit, double y_limit, double rot_limit, double ang_step, double transl_step,
                        double max_finest_prob_diff, slamgpu_bounds_fn fn, void *user, double out_delta[3], double *out_prob,
                        int64_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif /* SLAMGPU_H */
