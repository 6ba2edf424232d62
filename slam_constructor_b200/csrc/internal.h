// internal.h -- host-side structures of libslamgpu.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/slamgpu.h"

// NVTX ranges around the K1..K6 entry points (header-only nvtx3: a no-op until a tool such as nsys / ncu --nvtx attaches)
#include <nvtx3/nvToolsExt.h>
struct SgNvtxRange {
  explicit SgNvtxRange(const char *name) { nvtxRangePushA(name); }
  ~SgNvtxRange() { nvtxRangePop(); }
};
#define SG_NVTX(name) SgNvtxRange sg_nvtx_range_(name)

#define SG_LUT_PAD 1  // ring of "unknown" cells around the padded score LUT
#define SG_LUT_SLACK 512  // doubles allocated past the LUT: a staged patch row may run past the last row
#define SG_LUT_SLACK_ROWS 7  // whole rows allocated past the LUT (k_score_grid4 loads rows it does not use)

struct DevBuf {  // grow-only device scratch buffer
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <class T> T *as() const { return (T *)p; }
};

struct NcclApi;  // nccl_dyn.cc

struct Candidates {  // a staged candidate set (device resident)
  int kind = -1;     // 0 list, 1 grid
  int64_t P = 0;     // global candidate count
  int64_t p0 = 0, p1 = 0;  // this rank's slice [p0, p1)
  slamgpu_spe_params spe{};
  slamgpu_scan *scan = nullptr;
  // list
  DevBuf poses;      // 3*Ploc doubles (slice only)
  DevBuf theta_id;   // int32 per local pose
  int32_t T = 0;     // distinct thetas
  std::vector<double> h_thetas;  // distinct theta values (host)
  DevBuf d_thetas;
  // grid
  int32_t nx = 0, ny = 0, nt = 0, nyp = 0;
  std::vector<double> h_xs, h_ys, h_ts;
  DevBuf d_xs, d_ys;
  DevBuf groups;     // int32 {t, k0, cnt, pad} per row group of this rank
  int32_t n_groups = 0, rows_per_group = 0;
  int32_t t_lo = 0, t_hi = 0;  // theta planes touched by this rank
  DevBuf cxp, cyp;   // int32 index tables
  DevBuf cyw;        // v2: packed row words, one per (theta, beam, y-group)
  DevBuf sm_cnt;     // v4: per-SM task counters
  DevBuf colrec, w2;  // v5: column records per (theta, beam, band); weight pairs
  std::vector<double> h_w2;
  int32_t nb5 = 0, bw5 = 0;  // v5: bands per y-group, columns per band
  DevBuf wtask;      // v4: warp table {first y-group, band | -groups}
  std::vector<int32_t> h_wtask;
  int32_t n_warps4 = 0;
  int32_t ngys = 0;  // row stride of cyw (v4 / v5: ngy rounded up to even)
  int32_t grid_R = 8, ngy = 0;  // rows per thread of the grid kernel, y-groups per (theta, beam)
  bool grid_v2 = true, force_v1 = false, force_list = false;
  // 5: v4's arithmetic behind a cp.async pipeline (default), 4: each distinct row gathered once per thread + indexed-branch accumulate, 3: TMA-staged patches (experimental,
  // slower: see DESIGN.md), 2: packed rows + L1 gathers (fallback of 4), 1: explicit row table
  int win_mode = 0;       // SLAMGPU_OOPE_MAX / _MEAN: the grid is scored out of the map's window LUTs (v1 kernel)
  int grid_variant = 5;   // variant of the staged set
  int user_variant = 0;   // requested through slamgpu_ctx_set_option / SLAMGPU_GRID_VARIANT (0: default = 5)
  int max_variant = 5;    // temporary cap while a launch falls back to a simpler variant
  DevBuf blocks, blk_rows, porg;          // v3: block table, per (theta, block) y range, patch origins
  int32_t nbt = 0, box_w = 0, box_h = 0, n_blocks3 = 0;
  double h_extent_x = 0, h_extent_y = 0;  // metres spanned by the x sweep / by the y rows of one block
  bool uniform_w = false;
  int64_t shape_key[8] = {0};  // shape of the last grid staging (see slamgpu_stage_grid)
  bool shape_valid = false;
  bool warm_l2 = false;   // stream the score LUT through L2 before a big grid launch (slamgpu_ctx_set_option "warm_l2"; off: measured
                          // 10 us slower per step than letting the scoring kernel pull the LUT in itself)
  // K6: every pose scored against its own particle's map
  bool multi = false;
  DevBuf views, view_id;
  // GMapping OOPE cache chain (spe.gm_cache == 2): predecessor of each pose, entry states, state after each pose
  std::vector<int32_t> h_groups;          // grid staging kept alive past the asynchronous uploads
  std::vector<double> h_axes;             // {xs | ys | thetas}
  const double *p_xs = nullptr, *p_ys = nullptr, *p_ts = nullptr;  // views into d_xs
  int user_rows = 0;  // slamgpu_ctx_set_option("grid_rows")
  bool gm_chain = false;
  DevBuf gm_pred, gm_in, gm_out;
  // trig tables (device), layout given by strides
  DevBuf trc, trs;
  bool trig_is_host = false;
  // results
  DevBuf scores;     // Ploc doubles
  DevBuf blk_best;   // per block {double score; int64 idx}
  DevBuf result;     // {double score; int64 idx; int64 guard}
  bool launched = false;
  slamgpu_map *last_map = nullptr;  // map of the last launch (a guard-triggered redo needs it)
  double init_score = 0;
  int64_t stats[8] = {0};
};

struct slamgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // peer-memory exchange of the per-rank results (score.cu: k_exchange_finalize): every rank owns a mailbox that its
  // peers write through NVLink (CUDA IPC mappings); NULL when unavailable -> the NCCL all-gather is used instead
  void *mailbox = nullptr;         // this rank's mailbox: 2 parities x nranks x 64 B
  void **d_peer_mailbox = nullptr; // device array [nranks]: every rank's mailbox as seen from this GPU
  void *peer_mapped[64] = {nullptr};
  unsigned long long p2p_seq = 0;
  int *d_p2p_status = nullptr;     // set by the kernel when a peer did not answer in time
  bool p2p_broken = false;         // after a timeout: results go through ncclAllGather
  void *h_slots = nullptr, *h_counters = nullptr;  // small pinned blocks of the scan insertion (map slots up, cell counts down)
  size_t h_slots_cap = 0, h_counters_cap = 0;
  double p2p_timeout_ms = 20000.0; // slamgpu_ctx_set_option("p2p_timeout_ms")
  double clock_hz = 1.965e9;       // SM clock (cudaDevAttrClockRate) for clock64 deadlines
  cudaStream_t side = nullptr;  // the robot cell's update chain runs here, next to the sort (mapping.cu)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_staged = nullptr;  // recorded behind the uploads of a deferred batched insertion (sg_append_plans): the next batch
  bool staged_pending = false;      // may refill the pinned staging block once it has fired, while the kernels still run
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evk0 = nullptr, evk1 = nullptr;  // around the dominant kernel of the last call
  bool evk_valid = false;
  int rank = 0, nranks = 1;
  void *comm = nullptr;  // ncclComm_t
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;
  Candidates cand;
  DevBuf flush;      // L2 flush target
  DevBuf gather;     // all-gather staging: nranks * 16 B
  void *h_pinned = nullptr;  // small pinned staging (results)
  size_t h_pinned_cap = 0;
  DevBuf scratch[8];  // generic scratch (append_scan etc.)
};

// Copy-on-write tiles for the per-particle maps of K6 (LazyTiledGridMap, src/core/maps/lazy_tiled_grid_map.h:18-118: 128 x 128
// tiles behind shared_ptr, cloned on first write :57-71).  A pool hands out tiles of 128*128*stride doubles from device
// chunks; maps of the pool hold a table of tile ids (host) and of tile pointers (device).  Tile 0 is the shared "all
// unknown" tile every new map starts with; a tile referenced more than once (or tile 0) is cloned before a map writes to it,
// so resampling a particle is a table copy.
#define SG_TILE_BITS 7
#define SG_TILE (1 << SG_TILE_BITS)
struct SgTilePool {
  slamgpu_ctx *ctx = nullptr;
  int stride = 0;
  size_t tile_doubles = 0;
  int tiles_per_chunk = 64;
  std::vector<double *> chunks;
  std::vector<int32_t> refcnt;    // per tile id
  std::vector<int32_t> free_ids;
  int64_t tiles_cloned = 0, tiles_live = 0;  // statistics (slamgpu_particles_tile_stats)
  DevBuf jobs;                    // device copy of a batch of clone jobs
  double *ptr(int32_t id) const { return chunks[(size_t)id / tiles_per_chunk] + (size_t)(id % tiles_per_chunk) * tile_doubles; }
};

struct slamgpu_map {
  slamgpu_ctx *ctx = nullptr;
  // tiled storage (pool != NULL): d_cells stays NULL, the cells live in pool tiles
  SgTilePool *pool = nullptr;
  int32_t tw = 0, th = 0;               // tiles per row / column = ceil(w / 128), ceil(h / 128)
  std::vector<int32_t> tile_ids;        // [th][tw]
  std::vector<double *> h_tile_ptrs;    // the same as device pointers (kept for the uploads)
  double **d_tile_ptrs = nullptr;       // device copy of h_tile_ptrs
  size_t d_tile_cap = 0;
  bool tiles_dirty = false;             // h_tile_ptrs changed since the last upload
  int32_t w = 0, h = 0, ox = 0, oy = 0;
  double scale = 1;
  int32_t model = 0, stride = 0, grow = 0;
  double unknown[SLAMGPU_MAX_STRIDE] = {0};
  double *d_cells = nullptr;  // h*w*stride
  size_t cells_cap = 0;       // doubles
  // padded score LUT per OIE: (h + 2*PAD) rows of `pitch` doubles
  double *d_lut[2] = {nullptr, nullptr};
  size_t lut_cap[2] = {0, 0};
  bool lut_valid[2] = {false, false};
  int32_t pitch = 0;
  double unknown_lut[2] = {0, 0};
  // window LUTs of the brute-force grid path (score.cu: sg_map_ensure_wlut): for the max / mean OOPEs the probability of a
  // window depends only on its first cell and on how many cells it spans per axis (two possibilities each) -> 4 tables
  double *d_wlut = nullptr;
  size_t wlut_cap = 0;
  bool wlut_valid = false;
  int wl_oie = 0, wl_mode = 0, wl_nxmin = 0, wl_nymin = 0, wl_pitch = 0, wl_rows = 0;
  double wl_v = 0, wl_h = 0;
  size_t wl_T = 0;  // doubles per table
  struct slamgpu_pyramid *pyr = nullptr;  // owning pyramid if this is its level 0
};

struct slamgpu_scan {
  slamgpu_ctx *ctx = nullptr;
  int32_t n = 0, cartesian = 0;
  bool has_factor = false;
  // host copies (libm-derived fields are computed on the host so they are bit-identical
  // to the reference's: range/angle of Cartesian points, x/y of polar points)
  std::vector<double> range, angle, x, y, weight, factor;
  std::vector<uint8_t> occ;
  double wsum = 0;  // sequential sum of weights (pose independent)
  bool xy_valid = true;  // polar uploads derive x, y (libm, host) only when a pre-rotated / window path asks: sg_scan_ensure_xy
  // device: one block of 6*n doubles + n bytes
  DevBuf d;
  double *d_range = nullptr, *d_angle = nullptr, *d_x = nullptr, *d_y = nullptr, *d_w = nullptr, *d_f = nullptr;
  uint8_t *d_occ = nullptr;
};

// ---- error plumbing ----
int sg_fail(slamgpu_ctx *ctx, int code, const char *fmt, ...);
void sg_set_global_error(const char *msg);
#define SG_CUDA(ctx, call)                                                                             \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return sg_fail(ctx, SLAMGPU_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)
#define SG_TRY(call)             \
  do {                           \
    int r__ = (call);            \
    if (r__ != SLAMGPU_OK) return r__; \
  } while (0)
#define SG_LAUNCHED(ctx) (++(ctx)->launches)

int sg_pinned(slamgpu_ctx *ctx, size_t bytes, void **out);
int sg_map_ensure_lut(slamgpu_map *m, int oie);
void sg_map_invalidate_lut(slamgpu_map *m);
int sg_map_realloc(slamgpu_map *m, int32_t w, int32_t h);
// tiled maps (core.cu)
int sg_pool_create(slamgpu_ctx *ctx, int model, const double *unknown_rec, SgTilePool **out);
void sg_pool_destroy(SgTilePool *pool);
int sg_map_create_tiled(slamgpu_ctx *ctx, SgTilePool *pool, int32_t w, int32_t h, double scale, int32_t model, int32_t grow,
                        const double *unknown_rec, slamgpu_map **out);
int sg_map_sync_tiles(slamgpu_map *m);  // upload the tile pointer table if it changed
// make the tiles covering internal cells [x0, x1] x [y0, y1] private to this map (clones are queued in `copies` as
// {src id, dst id}; run them with sg_pool_run_copies before anything writes)
int sg_map_make_writable(slamgpu_map *m, int x0, int y0, int x1, int y1, std::vector<int32_t> *copies);
int sg_pool_run_copies(SgTilePool *pool, const std::vector<int32_t> &copies);
int sg_map_share_tiles(slamgpu_map *to, const slamgpu_map *from);  // `to` becomes a copy-on-write copy of `from`
int sg_map_gather_dense(slamgpu_map *m, double *d_dst);            // tiles -> dense [h][w][stride] on the device
int sg_map_scatter_dense(slamgpu_map *m, const double *d_src);     // dense -> private tiles
// K6 (score.cu): poses[k] is scored against maps[view_id[k]]; out_scores may be NULL
int sg_score_poses_multi(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, slamgpu_scan *scan,
                         const slamgpu_spe_params *p, const double *poses, int64_t P, double *out_scores);
int sg_allgather16(slamgpu_ctx *ctx, const void *d_send16, void *d_recv);

// NCCL through dlopen (nccl_dyn.cc)
int sg_nccl_unique_id(void *id128, std::string *err);
int sg_nccl_init(int nranks, int rank, const void *id128, void **comm, std::string *err);
int sg_nccl_allgather(void *comm, const void *send, void *recv, size_t bytes, cudaStream_t s, std::string *err);
// rank-local work on a distributed ctx (per-particle scoring: the particles, not the candidates, are sharded)
struct SgLocalScope {
  slamgpu_ctx *c; int rank, nranks;
  explicit SgLocalScope(slamgpu_ctx *ctx) : c(ctx), rank(ctx->rank), nranks(ctx->nranks) { c->rank = 0; c->nranks = 1; }
  ~SgLocalScope() { c->rank = rank; c->nranks = nranks; }
};
struct SgXfer { int peer; void *ptr; size_t bytes; };
int sg_nccl_exchange(void *comm, const SgXfer *sends, int n_sends, const SgXfer *recvs, int n_recvs, cudaStream_t s, std::string *err);
void sg_nccl_destroy(void *comm);
// all-gather of host data: `host` holds nranks chunks of chunk_bytes, this rank's chunk filled in; on return all are
// poses scored as the sequences the reference would evaluate them in, carrying GmappingOccupancyObservationPE's
// one-entry cache from pose to pose: pred[k] >= 0 is the pose evaluated just before pose k, pred[k] < 0 means pose k
// starts from states_in[-1 - pred[k]]; pred == NULL: one sequence, pose k after pose k-1, pose 0 from states_in[0]
int sg_score_chained(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, slamgpu_scan *scan,
                     const slamgpu_spe_params *p, const double *poses, int64_t P, const int32_t *pred,
                     const slamgpu_gm_cache *states_in, int n_states, double *out_scores, slamgpu_gm_cache *out_states);
// a whole hill-climbing match per instance in one launch; out8 = 8 doubles per instance {x, y, theta, prob, tested, ...}
int sg_hill_climb_device(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n, slamgpu_scan *scan, const slamgpu_spe_params *p,
                         const double *init, const uint8_t *active, uint32_t max_failed_rounds, double tr, double rot,
                         double *out8, double *log, int log_cap, slamgpu_gm_cache *gm_states /* n, in/out: gm_cache == 2 */,
                         int *served);
int sg_allgather_host(slamgpu_ctx *ctx, void *host, size_t chunk_bytes);
int sg_scans_upload_xy(slamgpu_ctx *ctx, slamgpu_scan *const *scans, int count, int32_t n, const double *xs, const double *ys,
                       const double *weight);
int sg_scan_ensure_xy(slamgpu_scan *s);
void sg_preload_score();
void sg_preload_mapping();
void sg_preload_pyramid();
void sg_p2p_setup(slamgpu_ctx *ctx);
void sg_p2p_teardown(slamgpu_ctx *ctx);
