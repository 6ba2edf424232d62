/*
 * slam_oracle.h -- CPU restatement (plain C99) of the slam-constructor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker.  The product (libslamgpu.so) never
 * links, loads or calls this code and has no CPU fallback.
 *
 * Parity status: PINNED.  Every routine here is checked (tests/test_oracle_pin.py)
 *   (1) against the golden vectors the reference's own unit tests hold for this
 *       path (tests/golden/upstream_*.json, extracted from /root/reference/test),
 *   (2) against oracle/_ref/libslamref.so -- the unmodified reference headers
 *       compiled in place by oracle/Makefile -- on seeded random inputs, and
 *   (3) against committed fixtures generated from (2) (tests/golden/ref_*.npz).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference).  All arithmetic is IEEE double in the reference's operation
 * order; build with -O2 -ffp-contract=off (no FMA contraction), like the
 * reference's own x86-64 -O3 build.
 *
 * Data model shared with include/slamgpu.h (the product C ABI):
 *   a grid map is a dense row-major array cells[h][w][stride] of doubles, one
 *   record per cell, plus (w, h, scale, origin) where internal = external +
 *   origin (src/core/maps/regular_squares_grid.h:120-143).
 */
#ifndef SLAM_ORACLE_H
#define SLAM_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- cell models (record layouts, doubles) --------------------------------
 * LWW      {p, q, known}            GridCell base / test MockGridCell
 *                                   (src/core/maps/grid_cell.h:9-57)
 * AFFINE   {p, known}               AffineQualityMergeCell (naive_grid_cells.h:6-21)
 * MEAN     {p, n}                   MeanProbabilityCell (naive_grid_cells.h:25-44)
 * TBM_*    {p, q, u, e, o, known}   TbmBaseCell (tbm_grid_cells.h:8-79); the
 *                                   conflict mass is always 0 after an update
 * GMAPPING {p, ox, oy, hits, tries} GmappingBaseCell (slams/gmapping/gmapping_grid_cell.h:9-43)
 */
enum { ORC_CELL_LWW = 0, ORC_CELL_AFFINE = 1, ORC_CELL_MEAN = 2,
       ORC_CELL_TBM_CONSISTENT = 3, ORC_CELL_TBM_UNKNOWN_EVEN = 4,
       ORC_CELL_GMAPPING = 5, ORC_CELL_CREDIBILIST = 6 /* src/slams/credibilist/grid_cell.h */, ORC_CELL_MODELS = 7 };
#define ORC_MAX_STRIDE 8

enum { ORC_OIE_DISCREPANCY = 0, ORC_OIE_OCCUPANCY = 1 };
enum { ORC_OOPE_OBSTACLE = 0, ORC_OOPE_MAX = 1, ORC_OOPE_MEAN = 2,
       ORC_OOPE_OVERLAP = 3, ORC_OOPE_GMAPPING = 4 };
enum { ORC_GROW_NONE = 0,      /* PlainGridMap / LazyTiledGridMap (bounded)   */
       ORC_GROW_PLAIN = 1,     /* UnboundedPlainGridMap, x1.2 growth           */
       ORC_GROW_TILED = 2 };   /* UnboundedLazyTiledGridMap, 128-cell tiles    */
enum { ORC_EST_CONST = 0, ORC_EST_AREA = 1 };
enum { ORC_SPW_EVEN = 0, ORC_SPW_VINY = 1, ORC_SPW_AHR = 2 };
enum { ORC_OMQE_IDLE = 0, ORC_OMQE_AHR = 1 };

typedef struct orc_map {
  int32_t w, h;
  double scale;
  int32_t ox, oy;
  int32_t model, stride, grow;
  double unknown[ORC_MAX_STRIDE]; /* prototype ("unknown") record */
  double *cells;                  /* owned, h*w*stride */
  int64_t oob_updates;            /* updates dropped on a bounded map (reference asserts) */
} orc_map;

typedef struct orc_scan {
  int32_t n;
  int32_t cartesian;    /* 0: (range, angle)   1: (x, y)  -- sensor_data.h:14-136 */
  const double *a;      /* range | x */
  const double *b;      /* angle | y */
  const uint8_t *occ;   /* is_occupied */
  const double *factor; /* NULL -> 1.0 */
  const double *weight; /* ScanPointWeighting output, one per point */
} orc_scan;

typedef struct orc_spe_params {
  int32_t oope, oie;
  double win_v, win_h;  /* SPEParams::sp_analysis_area side lengths (vside, hside) */
  int32_t prerotated;   /* SPEParams::scan_is_prerotated */
  double gm_fullness_th;
  int32_t gm_window;
} orc_spe_params;

/* mutable 1-entry cache of GmappingOccupancyObservationPE (quirk Q7) */
typedef struct orc_gm_cache { int32_t cx, cy; double prob; } orc_gm_cache;

typedef struct orc_estimator {
  int32_t type;                 /* ORC_EST_* */
  double occ_p, occ_q;          /* base_occupied */
  double empty_p, empty_q;      /* base_empty */
  double low_qual, unknown_qual;
  double shift_amount;          /* function-static Shift_Amount (Q10); <0: low_qual*side of the first cell */
} orc_estimator;

/* ---- map life cycle ---- */
int  orc_model_stride(int model);
void orc_default_unknown(int model, double *rec);
orc_map *orc_map_create(int w, int h, double scale, int model, int grow, const double *unknown_rec);
void orc_map_destroy(orc_map *m);
orc_map *orc_map_clone(const orc_map *m);
void orc_map_info(const orc_map *m, int32_t *w, int32_t *h, double *scale, int32_t *ox, int32_t *oy, int32_t *stride);
double *orc_map_cells(orc_map *m);
/* record of external cell (x, y); the unknown record when outside */
const double *orc_map_at(const orc_map *m, int x, int y);
int  orc_map_has_cell(const orc_map *m, int x, int y);
int  orc_map_ensure_inside(orc_map *m, int x, int y);
void orc_map_reset_cell(orc_map *m, int x, int y, const double *rec);

/* ---- primitives ---- */
int    orc_are_equal(double a, double b);
int    orc_less(double a, double b);
int    orc_less_or_equal(double a, double b);
int    orc_world_to_cell(double v, double scale);
int    orc_raycast(double bx, double by, double ex, double ey, double scale, int32_t *xy, int cap);
int    orc_bresenham(int bx, int by, int ex, int ey, int32_t *xy, int cap);
int    orc_rasterize_rect(double scale, int w, int h, int ox, int oy, double bot, double top, double left, double right,
                          int include_border, int32_t *lbrt);
double orc_rect_overlap(double ab, double at, double al, double ar, double bb, double bt, double bl, double br);
void   orc_estimate_occupancy(const orc_estimator *e, double bx, double by, double ex, double ey,
                              double cbot, double ctop, double cleft, double cright, int is_occ, double *out_pq);

/* ---- cells ---- */
void   orc_cell_update(int model, double *rec, int aoo_is_occ, double aoo_p, double aoo_q,
                       double obst_x, double obst_y, double quality);
double orc_cell_discrepancy(int model, const double *rec, double aoo_p, double aoo_q, double obst_x, double obst_y, double quality);
double orc_cell_impact(int model, int oie, const double *rec, double obst_x, double obst_y);
void   orc_build_lut(const orc_map *m, int oie, double *lut /* h*w */, double *unknown_value);

/* ---- scan preparation (host side of the path) ---- */
int    orc_filter_scan(const orc_map *m, int n, const double *range, const double *angle, const uint8_t *occ,
                       double px, double py, double pth, unsigned skip_rate, double max_range, int32_t *keep_idx);
void   orc_point_weights(int spw, int n, const double *range, const double *angle, double *w);
void   orc_angle_histogram_values(int n, const double *range, const double *angle, uint32_t *values);

/* ---- scoring ---- */
double orc_point_probability(const orc_map *m, const orc_spe_params *p, double X, double Y, orc_gm_cache *cache);
double orc_scan_probability(const orc_map *m, const orc_scan *s, const orc_spe_params *p,
                            double px, double py, double pth, orc_gm_cache *cache);
void   orc_score_poses(const orc_map *m, const orc_scan *s, const orc_spe_params *p,
                       const double *poses /* 3*P */, int64_t P, double *scores, orc_gm_cache *cache);
/* sequential accept loop of PoseEnumerationScanMatcher over a fixed list */
int64_t orc_argbest(const double *scores, int64_t P, double init_score, double *best_score);

/* ---- map update ---- */
int64_t orc_append_scan(orc_map *m, const orc_scan *s, double px, double py, double pth, double scan_quality,
                        double scan_margin, const orc_estimator *est, double blur, double max_range, int omqe,
                        int32_t *log_xy, int64_t log_cap);

/* ---- pyramid (M3RSMRescalableGridMap) ---- */
typedef struct orc_pyramid orc_pyramid;
orc_pyramid *orc_pyramid_create(int w, int h, double scale, int model, int grow, const double *unknown_rec, int oie);
void  orc_pyramid_destroy(orc_pyramid *p);
int   orc_pyramid_levels(orc_pyramid *p);
orc_map *orc_pyramid_level(orc_pyramid *p, int level);
int   orc_pyramid_rescale(orc_pyramid *p, double target_scale);
void  orc_pyramid_update(orc_pyramid *p, int x, int y, int aoo_is_occ, double aoo_p, double aoo_q,
                         double obst_x, double obst_y, double quality);
int64_t orc_pyramid_append_scan(orc_pyramid *p, const orc_scan *s, double px, double py, double pth,
                                double scan_quality, double scan_margin, const orc_estimator *est, double blur,
                                double max_range, int omqe);
double orc_match_bound(orc_pyramid *p, const orc_scan *s, const orc_spe_params *spe, double px, double py, double pth,
                       double rotation, double wbot, double wtop, double wleft, double wright);

/* ---- pose enumerators / matchers (host logic, restated for end-to-end pins) ---- */
int64_t orc_bf_enumerate(double bx, double by, double bth, double fx, double tx, double sx, double fy, double ty,
                         double sy, double ft, double tt, double st, double *poses, int64_t cap,
                         double *xs, int32_t *nx, double *ys, int32_t *ny, double *ts, int32_t *nt);
typedef struct orc_match_result { double best_prob, dx, dy, dth; int64_t poses_tested; } orc_match_result;
void orc_match_list(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy, double ith,
                    const double *poses, int64_t P, orc_match_result *out);
void orc_match_hill_climbing(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy, double ith,
                             unsigned max_failed_rounds, double tr_delta, double rot_delta, orc_match_result *out,
                             orc_gm_cache *cache);
void orc_match_monte_carlo(const orc_map *m, const orc_scan *s, const orc_spe_params *p, double ix, double iy, double ith,
                           unsigned seed, double tr_disp, double rot_disp, unsigned fal, unsigned attempts,
                           orc_match_result *out);

/* std::mt19937 + libstdc++ std::normal_distribution<double>, restated (monte_carlo_scan_matcher.h:35-66) */
typedef struct orc_mt19937 { uint32_t mt[624]; int idx; } orc_mt19937;
void   orc_mt_seed(orc_mt19937 *g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937 *g);
typedef struct orc_normal { double mean, stddev, saved; int saved_available; } orc_normal;
double orc_normal_sample(orc_normal *d, orc_mt19937 *g);

#ifdef __cplusplus
}
#endif
#endif
