// pyramid.cu -- K4 (multi-resolution max-pyramid) and K5 (batched Match upper bounds).
//
// Replaces RescalableCachingGridMap (src/core/maps/rescalable_caching_grid_map.h:16-200),
// M3RSMRescalableGridMap (src/core/scan_matchers/m3rsm_engine.h:17-131) and the scoring done by
// Match::Match (m3rsm_engine.h:156-180).
//
// Level 0 is the caller's fine map; level k has cells of 2^k times the fine size and the last
// level is a single cell of infinite size.  Every level is its own map with its own origin; all
// levels share world alignment through floor(x / scale_k).
//
// The reference maintains coarse cells incrementally (update_coarser_maps :101-126): after each
// fine cell update, walking up the levels, a coarse cell is replaced by a clone of the fine cell
// iff it is still unknown or the fine cell's impact is not less_or_equal (eps compare) to the
// coarse one, and the walk stops at the first level that did not change.  That is a sequential,
// order dependent fold, reproduced here exactly for a whole scan insertion:
//   per level:  k_level_coords  coarse cell of every still-alive update, in the reference's
//                               update order (beam order, obstacle cell first)
//               radix sort      stable, by coarse cell
//               k_level_fold    one thread per coarse cell walks its run in order with the
//                               reference's accept rule; updates that did not change the level
//                               die, the others go on to the next level
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "dev_window.cuh"
#include "mapping.h"

struct slamgpu_pyramid {
  slamgpu_ctx *ctx = nullptr;
  int oie = 0;
  std::vector<slamgpu_map *> lv;  // lv[0] is the caller's fine map (not owned)
  DevBuf ent;                     // per-update working arrays of the incremental fold
  DevBuf views, matches, scans, terms, bounds;  // K5 staging
  std::vector<slamgpu_scan *> rot_scans;        // pre-rotated copies of the scan being matched (slamgpu_match_m3rsm)
  // one-GPU fast path of slamgpu_match_m3rsm: the rotated point sets, level views and scan views of one match in ONE
  // device block; match records and bounds travel through a pinned host block the kernel reads and writes directly
  DevBuf m3_dev;
  void *m3_io = nullptr;
  size_t m3_io_cap = 0;
};

namespace {

// record copy with compile-time indices (a loop bounded by the runtime stride would push the record into local memory)
#ifndef SG_COPY_REC
#define SG_COPY_REC(dst, src, stride)                                   \
  do {                                                                  \
    _Pragma("unroll") for (int k_ = 0; k_ < SLAMGPU_MAX_STRIDE; ++k_)   \
      if (k_ < (stride)) (dst)[k_] = (src)[k_];                         \
  } while (0)
#endif

#define SG_INVALID_KEY 0xFFFFFFFFu

int ge_pow2(int i) {
  int p = 1;
  while (p < i) p *= 2;
  return p;
}

// RescalableCachingGridMap::ensure_map_cache_is_continuous :171-194
int ensure_continuous(slamgpu_pyramid *p) {
  if (p->lv.size() < 2) return SLAMGPU_OK;
  const slamgpu_map *pc = p->lv[p->lv.size() - 2];
  int pc_w = pc->w, pc_h = pc->h;
  double pc_scale = pc->scale;
  if (pc_w <= 2 && pc_h <= 2) return SLAMGPU_OK;
  pc_w = ge_pow2(pc_w); pc_h = ge_pow2(pc_h);
  const slamgpu_map *fine = p->lv[0];
  while (2 < pc_w || 2 < pc_h) {
    // std::ceil(w / factor) with an unsigned factor upstream: integer division
    pc_w = std::max(2, (int)std::ceil((double)((unsigned)pc_w / 2u)));
    pc_h = std::max(2, (int)std::ceil((double)((unsigned)pc_h / 2u)));
    pc_scale *= 2;
    slamgpu_map *m = nullptr;
    SG_TRY(slamgpu_map_create(p->ctx, pc_w, pc_h, pc_scale, fine->model, fine->grow, fine->unknown, &m));
    p->lv.insert(p->lv.end() - 1, m);
  }
  return SLAMGPU_OK;
}

// ---------------------------------------------------------------- incremental update (exact)
struct OrderArgs {
  const long long *offsets;  // N + 1 slot offsets
  const BeamOut *bout;
  const int2 *cells;         // per slot
  int N;
  long long M;
  int w, h, ox, oy;          // fine map bounds (updates outside a bounded map were dropped)
  int *ent_slot;             // per position in update order: slot, or -1
  unsigned char *alive;      // per position
};

// position pos of the reference's update sequence -> slot: within a beam the obstacle (last) cell
// comes first, then cells 0 .. n-2 (grid_map_scan_adders.h:151-171)
__global__ void k_order_entries(OrderArgs a) {
  long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= a.M) return;
  int lo = 0, hi = a.N;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (a.offsets[mid] <= pos) lo = mid; else hi = mid;
  }
  const int j = (int)(pos - a.offsets[lo]);
  const int count = a.bout[lo].count;
  int slot = -1;
  unsigned char alive = 0;
  if (j < count) {
    const int k = j == 0 ? count - 1 : j - 1;
    slot = (int)(a.offsets[lo] + k);
    const int2 c = a.cells[slot];
    const int ix = c.x + a.ox, iy = c.y + a.oy;
    alive = (ix >= 0 && ix < a.w && iy >= 0 && iy < a.h) ? 1 : 0;
  }
  a.ent_slot[pos] = slot;
  a.alive[pos] = alive;
}

struct CoordArgs {
  const int *ent_slot;
  const unsigned char *alive;
  const int2 *cells;
  long long M;
  double fine_scale, scale;  // fine cell size, this level's cell size
  int w, h, ox, oy;          // this level's bounds
  int2 *coords;              // per position: this level's (external) cell
  unsigned long long *counters;  // [0] alive, [1] outside the level, [2] fine cell spans several coarse cells
};

// cells of this level covered by the fine cell: GridRasterizedRectangle(level, fine bounds, border excluded)
// (grid_rasterization.h:26-47 with the 1e-9 inset)
__global__ void k_level_coords(CoordArgs a) {
  long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= a.M) return;
  if (!a.alive[pos]) return;
  const int2 c = a.cells[a.ent_slot[pos]];
  const double s = a.fine_scale;
  const double bot = sg::mul(s, (double)c.y), top = sg::mul(s, (double)(c.y + 1));
  const double left = sg::mul(s, (double)c.x), right = sg::mul(s, (double)(c.x + 1));
  const double off = 1e-9;
  const int lx = sg::world_to_cell(sg::add(left, off), a.scale), ly = sg::world_to_cell(sg::add(bot, off), a.scale);
  const int rx = sg::world_to_cell(sg::sub(right, off), a.scale), ry = sg::world_to_cell(sg::sub(top, off), a.scale);
  a.coords[pos] = make_int2(lx, ly);
  atomicAdd(a.counters, 1ull);
  if (lx != rx || ly != ry) atomicAdd(a.counters + 2, 1ull);
  const int ix = lx + a.ox, iy = ly + a.oy;
  if (ix < 0 || ix >= a.w || iy < 0 || iy >= a.h) atomicAdd(a.counters + 1, 1ull);
}

struct KeyArgs {
  const unsigned char *alive;
  const int2 *coords;
  long long M;
  int w, h, ox, oy;
  unsigned *keys, *vals;
  unsigned char *alive_next;
};

__global__ void k_level_keys(KeyArgs a) {
  long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= a.M) return;
  unsigned key = SG_INVALID_KEY;
  unsigned char next = 0;
  if (a.alive[pos]) {
    const int2 c = a.coords[pos];
    const int ix = c.x + a.ox, iy = c.y + a.oy;
    if (ix >= 0 && ix < a.w && iy >= 0 && iy < a.h) key = (unsigned)iy * (unsigned)a.w + (unsigned)ix;
    else next = 1;  // outside a bounded level: reads as unknown, nothing is stored, the walk goes on
  }
  a.keys[pos] = key;
  a.vals[pos] = (unsigned)pos;
  a.alive_next[pos] = next;
}

struct FoldGatherArgs {
  const unsigned *keys, *vals;  // sorted by level cell; vals = positions in update order
  long long M;
  const int *ent_slot;
  const double *impact;  // per slot: impact of the fine cell right after its update
  const double *rec;     // per slot: fine cell record right after its update
  int stride;
  double *g_impact;      // per sorted position
  double *g_rec;         // per sorted position, stride doubles
};

// operands of the fold in sorted order, so the sequential walk streams through contiguous memory
__global__ void k_level_gather(FoldGatherArgs a) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= a.M) return;
  if (a.keys[t] == SG_INVALID_KEY) return;
  const int slot = a.ent_slot[a.vals[t]];
  a.g_impact[t] = a.impact[slot];
  const double *r = a.rec + (size_t)slot * a.stride;
  double *o = a.g_rec + (size_t)t * a.stride;
  SG_COPY_REC(o, r, a.stride);
}

struct FoldArgs {
  const unsigned *keys, *vals;  // sorted by level cell; vals = positions in update order
  long long M;
  const double *g_impact, *g_rec;  // sorted
  double *cells;         // this level
  int stride, model, oie;
  unsigned char *alive_next;
  unsigned *n_long;      // runs of SG_FOLD_LONG updates or more are queued for k_level_fold_long (a warp per run)
  unsigned *long_runs;   // their first sorted positions
};

#define SG_FOLD_LONG 64

// M3RSMRescalableGridMap::update_coarser_maps :101-126 for one level, one thread per level cell
__global__ void __launch_bounds__(128) k_level_fold(FoldArgs a) {
  long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= a.M) return;
  const unsigned key = a.keys[j];
  if (key == SG_INVALID_KEY) return;
  if (j > 0 && a.keys[j - 1] == key) return;
  if (j + SG_FOLD_LONG - 1 < a.M && a.keys[j + SG_FOLD_LONG - 1] == key) {  // (sorted: equal ends = one run)
    a.long_runs[atomicAdd(a.n_long, 1u)] = (unsigned)j;
    return;
  }
  double cur[SLAMGPU_MAX_STRIDE];
  double *cell = a.cells + (size_t)key * a.stride;
  SG_COPY_REC(cur, cell, a.stride);
  bool cur_unknown = sg::rec_is_unknown(a.model, cur);
  double cur_impact = sg::cell_impact(a.model, a.oie, cur, 0.0, 0.0);
  bool changed = false;
  for (long long t = j; t < a.M && a.keys[t] == key; ++t) {
    const double x = a.g_impact[t];
    if (!cur_unknown && sg::less_or_equal(x, cur_impact)) { a.alive_next[a.vals[t]] = 0; continue; }
    const double *r = a.g_rec + (size_t)t * a.stride;
    SG_COPY_REC(cur, r, a.stride);
    cur_unknown = sg::rec_is_unknown(a.model, cur);
    cur_impact = sg::cell_impact(a.model, a.oie, cur, 0.0, 0.0);
    changed = true;
    a.alive_next[a.vals[t]] = 1;
  }
  if (changed)
    SG_COPY_REC(cell, cur, a.stride);
}

// The same fold for a long run (the coarse cell under the robot collects an update from every beam), a warp per run.  The
// fold is a running maximum: an update survives only if it beats the value the cell holds when its turn comes, so in 32
// consecutive updates few survive.  The lanes test their updates against the current value at once; the first that beats
// it is the next survivor (everything before it dies, exactly as in the sequential walk), the value moves on to it, and
// the lanes behind it test again: a handful of ballots per 32 updates instead of 32 dependent steps.
__global__ void __launch_bounds__(128) k_level_fold_long(FoldArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned n_runs = *a.n_long;
  for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_runs; w += (gridDim.x * blockDim.x) >> 5) {
    const long long j = a.long_runs[w];
    const unsigned key = a.keys[j];
    double cur[SLAMGPU_MAX_STRIDE];
    double *cell = a.cells + (size_t)key * a.stride;
    SG_COPY_REC(cur, cell, a.stride);
    bool cur_unknown = sg::rec_is_unknown(a.model, cur);
    double cur_impact = sg::cell_impact(a.model, a.oie, cur, 0.0, 0.0);
    bool changed = false;
    for (long long t0 = j;; t0 += 32) {
      const long long t = t0 + lane;
      const bool in = t < a.M && a.keys[t] == key;
      const double x = in ? a.g_impact[t] : 0.0;
      const unsigned pending = __ballot_sync(0xffffffffu, in);
      if (!pending) break;
      unsigned survivors = 0, rem = pending;
      while (rem) {
        const bool beats_cur = ((rem >> lane) & 1u) && (cur_unknown || !sg::less_or_equal(x, cur_impact));
        const unsigned b = __ballot_sync(0xffffffffu, beats_cur);
        if (!b) break;
        const int l = __ffs(b) - 1;
        survivors |= 1u << l;
        const double *r = a.g_rec + (size_t)(t0 + l) * a.stride;
        SG_COPY_REC(cur, r, a.stride);
        cur_unknown = sg::rec_is_unknown(a.model, cur);
        cur_impact = sg::cell_impact(a.model, a.oie, cur, 0.0, 0.0);
        changed = true;
        rem &= l == 31 ? 0u : ~((2u << l) - 1u);
      }
      if (in) a.alive_next[a.vals[t]] = (survivors >> lane) & 1u;
      if (pending != 0xffffffffu) break;  // the run ended inside this chunk
    }
    if (changed && lane == 0)
      SG_COPY_REC(cell, cur, a.stride);
  }
}

// ---------------------------------------------------------------- from-scratch build
struct BuildArgs {
  const double *src;  // finer level
  int sw, sh, sox, soy;
  double *dst;        // this level
  int dw, dh, dox, doy;
  int stride, model, oie;
  int last;           // this is the single infinite cell: reduce the whole finer level
};

SG_DEV void build_take(const BuildArgs &a, const double *r, double *cur, bool *cur_unknown, double *cur_impact) {
  double rec[SLAMGPU_MAX_STRIDE];
  SG_COPY_REC(rec, r, a.stride);
  if (sg::rec_is_unknown(a.model, rec)) return;  // a cell nobody wrote never propagated
  double x = sg::cell_impact(a.model, a.oie, rec, 0.0, 0.0);
  if (!*cur_unknown && sg::less_or_equal(x, *cur_impact)) return;
  SG_COPY_REC(cur, rec, a.stride);
  *cur_unknown = false;
  *cur_impact = x;
}

__global__ void k_build_level(BuildArgs a) {
  int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ix >= a.dw || iy >= a.dh) return;
  double *cell = a.dst + ((size_t)iy * a.dw + ix) * a.stride;
  double cur[SLAMGPU_MAX_STRIDE];
  SG_COPY_REC(cur, cell, a.stride);
  bool cur_unknown = true;
  double cur_impact = 0;
  if (a.last) {
    for (int sx = 0; sx < a.sw; ++sx)
      for (int sy = 0; sy < a.sh; ++sy) build_take(a, a.src + ((size_t)sy * a.sw + sx) * a.stride, cur, &cur_unknown, &cur_impact);
  } else {
    const int X = ix - a.dox, Y = iy - a.doy;  // external cell of this level
    for (int dx = 0; dx < 2; ++dx)
      for (int dy = 0; dy < 2; ++dy) {
        const int sx = 2 * X + dx + a.sox, sy = 2 * Y + dy + a.soy;
        if (sx < 0 || sx >= a.sw || sy < 0 || sy >= a.sh) continue;
        build_take(a, a.src + ((size_t)sy * a.sw + sx) * a.stride, cur, &cur_unknown, &cur_impact);
      }
  }
  if (!cur_unknown)
    SG_COPY_REC(cell, cur, a.stride);
}

// From-scratch build, five levels per pass.  k_build_level re-reads every level it has just written (and k_fill_level writes
// each level once more before it): 4096^2 mean cells moved 625 MB for 357 MB of compulsory traffic, in 24 launches.  Here a
// block owns one cell of level L0+5 = a 32 x 32 tile of the source level L0 (in external coordinates, so the 2 x 2 groups are
// the reference's at every level whatever the origins).  Each of its 256 threads reads one 2 x 2 group of the source and
// folds it in registers (level L0+1); the 16 x 16 results {record, impact, known} go to shared memory and are folded
// 16x16 -> 8x8 -> 4x4 -> 2x2 -> 1, every level's cells written as they appear (the unknown prototype where no known child
// exists).  The fold of one cell is build_take's: children in x-outer / y-inner order, a child replaces the current pick only
// if its impact is greater by more than the reference's eps (rescalable_caching_grid_map.h:171-194, m3rsm_engine.h:101-126).
// A cell that lies outside its level's array does not exist for the next level, exactly as when that level is read back
// from memory.
#define SG_FUSED_LEVELS 5
struct FusedLevel { double *cells; int w, h, ox, oy; double unknown[SLAMGPU_MAX_STRIDE]; };
struct FusedArgs {
  const double *src;
  int sw, sh, sox, soy;
  FusedLevel lv[SG_FUSED_LEVELS];
  int nlv, model, oie, tx0, ty0;
};

// one fold step of build_take on {have, impact}: true if the candidate replaces the current pick
SG_DEV bool fused_takes(bool have, double best_imp, bool known, double xi) {
  return known && !(have && sg::less_or_equal(xi, best_imp));
}

template <int STRIDE>
SG_DEV void fused_store(const FusedLevel &L, int ix, int iy, bool have, double (&rec)[STRIDE], bool *inside) {
  *inside = ix >= 0 && ix < L.w && iy >= 0 && iy < L.h;
#pragma unroll
  for (int k = 0; k < STRIDE; ++k) rec[k] = have ? rec[k] : L.unknown[k];
  if (*inside) {
    double *d = L.cells + ((size_t)iy * L.w + ix) * STRIDE;
    if (STRIDE == 2) *reinterpret_cast<double2 *>(d) = make_double2(rec[0], rec[1]);
    else {
#pragma unroll
      for (int k = 0; k < STRIDE; ++k) d[k] = rec[k];
    }
  }
}

template <int MODEL>
__global__ void __launch_bounds__(256, 6) k_build_fused(FusedArgs a) {
  constexpr int STRIDE = MODEL == SLAMGPU_CELL_LWW ? 3 : (MODEL == SLAMGPU_CELL_AFFINE || MODEL == SLAMGPU_CELL_MEAN) ? 2 :
                         MODEL == SLAMGPU_CELL_GMAPPING ? 5 : 6;
  // levels L0+3 .. L0+5 are folded by warp 0 alone out of this buffer (8 x 8 cells of level L0+2, then 4 x 4, 2 x 2)
  __shared__ double s_rec[STRIDE][64 + 16 + 4];
  __shared__ double s_imp[64 + 16 + 4];
  __shared__ unsigned char s_known[64 + 16 + 4];
  const int t = threadIdx.x, lane = t & 31;
  const int TX = a.tx0 + (int)blockIdx.x, TY = a.ty0 + (int)blockIdx.y;
  // a warp holds 16 x 2 cells of level L0+1: lane = (qy & 1) * 16 + qx
  const int qx = t & 15, qy = t >> 4;
  // ---- level L0+1 from this thread's 2 x 2 group of the source, in registers
  bool have = false;
  double imp = 0.0, pick[STRIDE];
#pragma unroll
  for (int k = 0; k < STRIDE; ++k) pick[k] = 0.0;
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int sx = TX * 32 + 2 * qx + dx + a.sox, sy = TY * 32 + 2 * qy + dy + a.soy;
      if (sx < 0 || sx >= a.sw || sy < 0 || sy >= a.sh) continue;
      double rec[SLAMGPU_MAX_STRIDE];
      const double *r = a.src + ((size_t)sy * a.sw + sx) * STRIDE;
      if (STRIDE == 2) {  // one 16-byte load per cell
        const double2 v = __ldg(reinterpret_cast<const double2 *>(r));
        rec[0] = v.x; rec[1] = v.y;
      } else {
#pragma unroll
        for (int k = 0; k < STRIDE; ++k) rec[k] = __ldg(r + k);
      }
      const bool known = !sg::rec_is_unknown(MODEL, rec);
      const double xi = sg::cell_impact(MODEL, a.oie, rec, 0.0, 0.0);
      if (!fused_takes(have, imp, known, xi)) continue;
      have = true; imp = xi;
#pragma unroll
      for (int k = 0; k < STRIDE; ++k) pick[k] = rec[k];
    }
  bool inside;
  fused_store<STRIDE>(a.lv[0], TX * 16 + qx + a.lv[0].ox, TY * 16 + qy + a.lv[0].oy, have, pick, &inside);
  have = have && inside;
  if (a.nlv < 2) return;
  // ---- level L0+2 by shuffles: the 2 x 2 group of a level-(L0+2) cell sits in one warp; its leader is the lane with even qx, qy
  {
    const int base = lane & 14;  // lane of child (dx, dy) = base + dx + 16 * dy
    bool h2 = false;
    double i2 = 0.0;
    int src = base;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int l = base + dx + 16 * dy;
        const bool known = __shfl_sync(0xffffffffu, (int)have, l) != 0;
        const double xi = __shfl_sync(0xffffffffu, imp, l);
        if (!fused_takes(h2, i2, known, xi)) continue;
        h2 = true; i2 = xi; src = l;
      }
    double rec2[STRIDE];
#pragma unroll
    for (int k = 0; k < STRIDE; ++k) rec2[k] = __shfl_sync(0xffffffffu, pick[k], src);
    if ((qx & 1) == 0 && (qy & 1) == 0) {
      const int x = qx >> 1, y = qy >> 1;
      bool in2;
      fused_store<STRIDE>(a.lv[1], TX * 8 + x + a.lv[1].ox, TY * 8 + y + a.lv[1].oy, h2, rec2, &in2);
      const int c = y * 8 + x;
#pragma unroll
      for (int k = 0; k < STRIDE; ++k) s_rec[k][c] = rec2[k];
      s_imp[c] = i2; s_known[c] = (h2 && in2) ? 1 : 0;
    }
  }
  if (a.nlv < 3) return;
  __syncthreads();
  if (t >= 32) return;
  // ---- levels L0+3 .. L0+5: warp 0, out of shared memory
  int side = 8, off = 0;
#pragma unroll 1
  for (int j = 2; j < a.nlv; ++j) {
    const int n = side >> 1, noff = off + side * side;
    if (t < n * n) {
      const int x = t % n, y = t / n;
      bool h = false;
      double bi = 0.0;
      int best = off;
      for (int dx = 0; dx < 2; ++dx)
        for (int dy = 0; dy < 2; ++dy) {
          const int c = off + (2 * y + dy) * side + 2 * x + dx;
          if (!fused_takes(h, bi, s_known[c] != 0, s_imp[c])) continue;
          h = true; bi = s_imp[c]; best = c;
        }
      double rec[STRIDE];
#pragma unroll
      for (int k = 0; k < STRIDE; ++k) rec[k] = s_rec[k][best];
      bool in;
      fused_store<STRIDE>(a.lv[j], TX * n + x + a.lv[j].ox, TY * n + y + a.lv[j].oy, h, rec, &in);
#pragma unroll
      for (int k = 0; k < STRIDE; ++k) s_rec[k][noff + t] = rec[k];
      s_imp[noff + t] = bi; s_known[noff + t] = (h && in) ? 1 : 0;
    }
    __syncwarp();
    side = n; off = noff;
  }
}

struct RecParam { double v[SLAMGPU_MAX_STRIDE]; };
__global__ void k_fill_level(double *cells, size_t n_cells, int stride, RecParam rec) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  size_t n = n_cells * stride;
  for (; i < n; i += st) cells[i] = rec.v[i % stride];
}

// ---------------------------------------------------------------- K5
struct ScanView { const double *sx, *sy, *w, *f; int n, has_factor; double wsum; };
struct MatchRec { double dx, dy, vside, hside; int scan_id, level; };

// phase 1: one thread per (match, point): the `max` OOPE over the window re-centred at the point
// (occupancy_observation_probability.h:35-48) on the level Match::Match picked (:176)
__global__ void k_window_terms(const MatchRec *__restrict__ matches, const ScanView *__restrict__ scans,
                               const MapView *__restrict__ views, int n_max, double *__restrict__ terms) {
  const int m = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const MatchRec mr = matches[m];
  const ScanView sv = scans[mr.scan_id];
  if (i >= sv.n) return;
  const MapView &mv = views[mr.level];
  // pre-rotated scan: point + pose offset (sensor_data.h:103-105)
  const double X = sg::add(sv.sx[i], mr.dx), Y = sg::add(sv.sy[i], mr.dy);
  const double prob = window_probability<SLAMGPU_OOPE_MAX>(mv, X, Y, mr.vside, mr.hside);
  double term = sg::mul(prob, sv.w[i]);
  if (sv.has_factor) term = sg::mul(term, sv.f[i]);
  terms[(size_t)m * n_max + i] = term;
}

// phase 2: the reference's sequential FP64 sum in point order (weighted_mean_point_probability_spe.h:107-132)
__global__ void k_ordered_sums(const MatchRec *__restrict__ matches, const ScanView *__restrict__ scans, int n_max,
                               const double *__restrict__ terms, long long M, double *__restrict__ out) {
  long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const ScanView sv = scans[matches[m].scan_id];
  double total = 0;
  const double *t = terms + (size_t)m * n_max;
  for (int i = 0; i < sv.n; ++i) total = sg::add(total, t[i]);
  out[m] = sv.wsum == 0 ? NAN : sg::div(total, sv.wsum);
}

// the same with the rows staged in shared memory: G matches per block, coalesced loads, then one lane per match (each
// in its own warp) runs the chain of adds from shared memory with the loads issued eight ahead
__global__ void __launch_bounds__(256) k_ordered_sums_staged(const MatchRec *__restrict__ matches, const ScanView *__restrict__ scans,
                                                             int n_max, const double *__restrict__ terms, long long M, int G,
                                                             double *__restrict__ out) {
  extern __shared__ double sh_rows[];  // [G][n_max]
  const long long m0 = (long long)blockIdx.x * G;
  const int nm = (int)min((long long)G, M - m0);
  const double *src = terms + (size_t)m0 * n_max;
  for (int e = threadIdx.x; e < nm * n_max; e += blockDim.x) sh_rows[e] = src[e];
  __syncthreads();
  const int g = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0 || g >= nm) return;
  const ScanView sv = scans[matches[m0 + g].scan_id];
  const double *t = sh_rows + (size_t)g * n_max;
  double total = 0;
  int i = 0;
  for (; i + 8 <= sv.n; i += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = t[i + u];
#pragma unroll
    for (int u = 0; u < 8; ++u) total = sg::add(total, v[u]);
  }
  for (; i < sv.n; ++i) total = sg::add(total, t[i]);
  out[m0 + g] = sv.wsum == 0 ? NAN : sg::div(total, sv.wsum);
}

// Match::Match in one launch: one block per match; the threads take the window maxima of the points into shared memory,
// then one lane adds them in point order (the loads eight ahead of the chain of adds).  `matches` and `out` may live in
// pinned host memory: a record is read once per block, a bound written once.
__global__ void __launch_bounds__(512) k_match_bounds(const MatchRec *__restrict__ matches, const ScanView *__restrict__ scans,
                                                      const MapView *__restrict__ views, double *__restrict__ out,
                                                      unsigned *done_count, volatile unsigned *host_flag, unsigned seq) {
  extern __shared__ double sh_terms[];
  __shared__ MatchRec s_mr;
  if (threadIdx.x == 0) s_mr = matches[blockIdx.x];
  __syncthreads();
  const MatchRec mr = s_mr;
  const ScanView sv = scans[mr.scan_id];
  const MapView &mv = views[mr.level];
  for (int i = threadIdx.x; i < sv.n; i += blockDim.x) {
    const double X = sg::add(sv.sx[i], mr.dx), Y = sg::add(sv.sy[i], mr.dy);
    const double prob = window_probability<SLAMGPU_OOPE_MAX>(mv, X, Y, mr.vside, mr.hside);
    double term = sg::mul(prob, sv.w[i]);
    if (sv.has_factor) term = sg::mul(term, sv.f[i]);
    sh_terms[i] = term;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double total = 0;
  int i = 0;
  for (; i + 8 <= sv.n; i += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = sh_terms[i + u];
#pragma unroll
    for (int u = 0; u < 8; ++u) total = sg::add(total, v[u]);
  }
  for (; i < sv.n; ++i) total = sg::add(total, sh_terms[i]);
  out[blockIdx.x] = sv.wsum == 0 ? NAN : sg::div(total, sv.wsum);
  // the last block to finish tells the host, which polls the flag instead of paying for a stream synchronisation
  __threadfence_system();
  if (atomicAdd(done_count, 1u) == gridDim.x - 1) {
    *done_count = 0;
    __threadfence_system();
    *host_flag = seq;
  }
}

MapView level_view(const slamgpu_map *m, int oie) {
  MapView v;
  v.lut = m->d_lut[oie]; v.cells = m->d_cells; v.tiles = nullptr; v.tw = 0;
  v.w = m->w; v.h = m->h; v.ox = m->ox; v.oy = m->oy; v.pitch = m->pitch; v.stride = m->stride; v.model = m->model;
  v.scale = m->scale; v.unknown_lut = m->unknown_lut[oie];
  memcpy(v.unknown_rec, m->unknown, sizeof v.unknown_rec);
  return v;
}

int rescale_id(slamgpu_pyramid *p, double target) {  // rescale :79-91
  int id = 0;
  while (id + 1 < (int)p->lv.size() && !(target <= p->lv[id]->scale)) ++id;
  return id;
}

}  // namespace

// ---------------------------------------------------------------- ABI
extern "C" int slamgpu_pyramid_create(slamgpu_ctx *ctx, slamgpu_map *fine, int32_t oie, slamgpu_pyramid **out) {
  if (!ctx || !fine || !out) return sg_fail(ctx, SLAMGPU_E_INVALID, "pyramid_create: NULL argument");
  if (fine->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map belongs to another ctx");
  if (oie < 0 || oie > 1) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad oie %d", oie);
  if (fine->pyr) return sg_fail(ctx, SLAMGPU_E_STATE, "this map already is level 0 of a pyramid");
  *out = nullptr;
  slamgpu_pyramid *p = new slamgpu_pyramid();
  p->ctx = ctx; p->oie = oie;
  p->lv.push_back(fine);
  slamgpu_map *top = nullptr;
  int r = slamgpu_map_create(ctx, 1, 1, INFINITY, fine->model, fine->grow, fine->unknown, &top);
  if (r != SLAMGPU_OK) { delete p; return r; }
  p->lv.push_back(top);
  r = ensure_continuous(p);
  if (r != SLAMGPU_OK) { slamgpu_pyramid_destroy(p); return r; }
  fine->pyr = p;
  *out = p;
  return SLAMGPU_OK;
}

extern "C" void slamgpu_pyramid_destroy(slamgpu_pyramid *p) {
  if (p && p->m3_io) { cudaFreeHost(p->m3_io); p->m3_io = nullptr; }
  if (p) p->m3_dev.release();
  if (!p) return;
  if (!p->lv.empty() && p->lv[0]) p->lv[0]->pyr = nullptr;
  for (size_t i = 1; i < p->lv.size(); ++i) slamgpu_map_destroy(p->lv[i]);
  for (slamgpu_scan *s : p->rot_scans) slamgpu_scan_destroy(s);
  DevBuf *bufs[] = {&p->ent, &p->views, &p->matches, &p->scans, &p->terms, &p->bounds};
  for (DevBuf *b : bufs) b->release();
  delete p;
}

extern "C" int slamgpu_pyramid_levels(slamgpu_pyramid *p) {
  if (!p) return SLAMGPU_E_INVALID;
  if (ensure_continuous(p) != SLAMGPU_OK) return SLAMGPU_E_NOMEM;
  return (int)p->lv.size();
}

extern "C" int slamgpu_pyramid_level_info(slamgpu_pyramid *p, int32_t level, int32_t *w, int32_t *h, double *scale,
                                          int32_t *ox, int32_t *oy) {
  if (!p || level < 0 || level >= (int)p->lv.size()) return SLAMGPU_E_INVALID;
  return slamgpu_map_info(p->lv[level], w, h, scale, ox, oy, nullptr);
}

extern "C" int slamgpu_pyramid_rescale(slamgpu_pyramid *p, double target_scale) {
  if (!p) return SLAMGPU_E_INVALID;
  if (ensure_continuous(p) != SLAMGPU_OK) return SLAMGPU_E_NOMEM;
  return rescale_id(p, target_scale);
}

extern "C" int slamgpu_pyramid_level_download(slamgpu_pyramid *p, int32_t level, double *cells, double *impact) {
  if (!p || level < 0 || level >= (int)p->lv.size()) return SLAMGPU_E_INVALID;
  if (cells) SG_TRY(slamgpu_map_download(p->lv[level], cells));
  if (impact) SG_TRY(slamgpu_map_lut_download(p->lv[level], p->oie, impact, nullptr));
  return SLAMGPU_OK;
}

static int floor_div_i(int v, int d) { return v >= 0 ? v / d : -((-v + d - 1) / d); }

extern "C" int slamgpu_pyramid_build(slamgpu_pyramid *p) {
  SG_NVTX("K4 pyramid_build");
  if (!p) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  SG_TRY(ensure_continuous(p));
  // geometry first (it depends on the finer level's extent only): every level covers the finer one, as the reference's would
  for (size_t id = 1; id + 1 < p->lv.size(); ++id) {
    slamgpu_map *src = p->lv[id - 1], *dst = p->lv[id];
    if (dst->grow != SLAMGPU_GROW_NONE && src->w > 0 && src->h > 0) {
      GrowState g{dst->w, dst->h, dst->ox, dst->oy, dst->grow};
      auto half = [](int v) { return (int)std::floor(v / 2.0); };
      bool grew = g.ensure_inside(half(-src->ox), half(-src->oy));
      grew |= g.ensure_inside(half(src->w - 1 - src->ox), half(src->h - 1 - src->oy));
      if (grew) SG_TRY(sg_map_regrow(dst, g));
    }
    SG_TRY(ensure_continuous(p));
  }
  const size_t nlev = p->lv.size();
  // levels 1 .. nlev-2, five per pass (k_build_fused); the last level is the single infinite cell
  bool first = true;
  for (size_t l0 = 0; l0 + 2 < nlev; l0 += SG_FUSED_LEVELS) {
    slamgpu_map *src = p->lv[l0];
    FusedArgs a;
    a.src = src->d_cells; a.sw = src->w; a.sh = src->h; a.sox = src->ox; a.soy = src->oy;
    a.model = src->model; a.oie = p->oie;
    a.nlv = (int)std::min<size_t>(SG_FUSED_LEVELS, nlev - 2 - l0);
    int tx0 = floor_div_i(-src->ox, 32), tx1 = floor_div_i(src->w - 1 - src->ox, 32);
    int ty0 = floor_div_i(-src->oy, 32), ty1 = floor_div_i(src->h - 1 - src->oy, 32);
    for (int j = 0; j < a.nlv; ++j) {
      slamgpu_map *m = p->lv[l0 + 1 + j];
      FusedLevel &L = a.lv[j];
      L.cells = m->d_cells; L.w = m->w; L.h = m->h; L.ox = m->ox; L.oy = m->oy;
      memcpy(L.unknown, m->unknown, sizeof L.unknown);
      const int per = 32 >> (j + 1);  // cells of this level per tile side
      if (m->w > 0 && m->h > 0) {
        tx0 = std::min(tx0, floor_div_i(-m->ox, per)); tx1 = std::max(tx1, floor_div_i(m->w - 1 - m->ox, per));
        ty0 = std::min(ty0, floor_div_i(-m->oy, per)); ty1 = std::max(ty1, floor_div_i(m->h - 1 - m->oy, per));
      }
      sg_map_invalidate_lut(m);
    }
    a.tx0 = tx0; a.ty0 = ty0;
    dim3 grd((unsigned)(tx1 - tx0 + 1), (unsigned)(ty1 - ty0 + 1));
    if (first) cudaEventRecord(ctx->evk0, ctx->stream);
    switch (src->model) {
      case SLAMGPU_CELL_LWW: k_build_fused<SLAMGPU_CELL_LWW><<<grd, 256, 0, ctx->stream>>>(a); break;
      case SLAMGPU_CELL_AFFINE: k_build_fused<SLAMGPU_CELL_AFFINE><<<grd, 256, 0, ctx->stream>>>(a); break;
      case SLAMGPU_CELL_MEAN: k_build_fused<SLAMGPU_CELL_MEAN><<<grd, 256, 0, ctx->stream>>>(a); break;
      case SLAMGPU_CELL_TBM_CONSISTENT: k_build_fused<SLAMGPU_CELL_TBM_CONSISTENT><<<grd, 256, 0, ctx->stream>>>(a); break;
      case SLAMGPU_CELL_TBM_UNKNOWN_EVEN: k_build_fused<SLAMGPU_CELL_TBM_UNKNOWN_EVEN><<<grd, 256, 0, ctx->stream>>>(a); break;
      case SLAMGPU_CELL_CREDIBILIST: k_build_fused<SLAMGPU_CELL_CREDIBILIST><<<grd, 256, 0, ctx->stream>>>(a); break;
      default: k_build_fused<SLAMGPU_CELL_GMAPPING><<<grd, 256, 0, ctx->stream>>>(a); break;
    }
    if (first) { cudaEventRecord(ctx->evk1, ctx->stream); ctx->evk_valid = true; first = false; }
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
  }
  if (nlev >= 2) {
    slamgpu_map *src = p->lv[nlev - 2], *dst = p->lv[nlev - 1];
    RecParam rp;
    memcpy(rp.v, dst->unknown, sizeof rp.v);
    k_fill_level<<<1, 32, 0, ctx->stream>>>(dst->d_cells, (size_t)dst->w * dst->h, dst->stride, rp);
    BuildArgs a;
    a.src = src->d_cells; a.sw = src->w; a.sh = src->h; a.sox = src->ox; a.soy = src->oy;
    a.dst = dst->d_cells; a.dw = dst->w; a.dh = dst->h; a.dox = dst->ox; a.doy = dst->oy;
    a.stride = dst->stride; a.model = dst->model; a.oie = p->oie; a.last = 1;
    dim3 blk(32, 8), grd((dst->w + 31) / 32, (dst->h + 7) / 8);
    k_build_level<<<grd, blk, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    SG_CUDA(ctx, cudaGetLastError());
    sg_map_invalidate_lut(dst);
  }
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

// walk the coarser levels for one scan insertion into the fine map (trace = what K3 recorded)
static int pyramid_propagate(slamgpu_pyramid *p, const AppendTrace &tr) {
  slamgpu_ctx *ctx = p->ctx;
  slamgpu_map *fine = p->lv[0];
  const long long M = tr.M;
  if (M == 0) return SLAMGPU_OK;
  // per-position arrays: ent_slot (i32) coords (int2) keys vals keys_tmp vals_tmp (u32) alive alive_next (u8) counters
  const size_t bytes = (size_t)M * (4 + 8 + 16 + 2) + 256 + (size_t)M * sizeof(double) * (1 + fine->stride) + 64 +
                       ((size_t)(M / SG_FOLD_LONG) + 2) * sizeof(unsigned) + 64;
  if (p->ent.reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "pyramid update buffers (%lld updates)", M);
  int2 *coords = p->ent.as<int2>();
  int *ent_slot = (int *)(coords + M);
  unsigned *keys = (unsigned *)(ent_slot + M), *vals = keys + M, *keys_tmp = vals + M, *vals_tmp = keys_tmp + M;
  unsigned char *alive = (unsigned char *)(vals_tmp + M), *alive_next = alive + M;
  unsigned long long *counters = (unsigned long long *)(((uintptr_t)(alive_next + M) + 63) & ~(uintptr_t)63);
  double *g_impact = (double *)(counters + 8), *g_rec = g_impact + M;
  unsigned *n_long = (unsigned *)(g_rec + (size_t)M * fine->stride), *long_runs = n_long + 16;  // counters[4] is cleared with the level's counters
  const unsigned nblk = (unsigned)((M + 127) / 128);
  OrderArgs oa;
  oa.offsets = tr.d_offsets; oa.bout = tr.d_bout; oa.cells = tr.cells; oa.N = tr.N; oa.M = M;
  oa.w = fine->w; oa.h = fine->h; oa.ox = fine->ox; oa.oy = fine->oy; oa.ent_slot = ent_slot; oa.alive = alive;
  k_order_entries<<<nblk, 128, 0, ctx->stream>>>(oa);
  SG_LAUNCHED(ctx);
  for (size_t id = 1; id < p->lv.size(); ++id) {
    slamgpu_map *lvl = p->lv[id];
    SG_CUDA(ctx, cudaMemsetAsync(counters, 0, 32, ctx->stream));
    SG_CUDA(ctx, cudaMemsetAsync(n_long, 0, 64, ctx->stream));
    CoordArgs ca;
    ca.ent_slot = ent_slot; ca.alive = alive; ca.cells = tr.cells; ca.M = M; ca.fine_scale = fine->scale; ca.scale = lvl->scale;
    ca.w = lvl->w; ca.h = lvl->h; ca.ox = lvl->ox; ca.oy = lvl->oy; ca.coords = coords; ca.counters = counters;
    k_level_coords<<<nblk, 128, 0, ctx->stream>>>(ca);
    SG_LAUNCHED(ctx);
    unsigned long long hc[4];
    SG_CUDA(ctx, cudaMemcpyAsync(hc, counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hc[0] == 0) break;  // every update stopped below this level
    if (hc[2] != 0) return sg_fail(ctx, SLAMGPU_E_STATE, "a fine cell spans several cells of level %zu (unsupported grid scale)", id);
    if (hc[1] != 0 && lvl->grow != SLAMGPU_GROW_NONE) {
      // the level grows exactly as the reference's would: replay ensure_inside over the updates in order
      std::vector<int2> hcoords((size_t)M);
      std::vector<unsigned char> halive((size_t)M);
      SG_CUDA(ctx, cudaMemcpyAsync(hcoords.data(), coords, sizeof(int2) * M, cudaMemcpyDeviceToHost, ctx->stream));
      SG_CUDA(ctx, cudaMemcpyAsync(halive.data(), alive, M, cudaMemcpyDeviceToHost, ctx->stream));
      SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      GrowState g{lvl->w, lvl->h, lvl->ox, lvl->oy, lvl->grow};
      for (long long pos = 0; pos < M; ++pos)
        if (halive[pos]) g.ensure_inside(hcoords[pos].x, hcoords[pos].y);
      SG_TRY(sg_map_regrow(lvl, g));
    }
    if ((long long)lvl->w * lvl->h >= 0xFFFFFFFFll) return sg_fail(ctx, SLAMGPU_E_NOMEM, "level too large for 32-bit cell keys");
    KeyArgs ka;
    ka.alive = alive; ka.coords = coords; ka.M = M; ka.w = lvl->w; ka.h = lvl->h; ka.ox = lvl->ox; ka.oy = lvl->oy;
    ka.keys = keys; ka.vals = vals; ka.alive_next = alive_next;
    k_level_keys<<<nblk, 128, 0, ctx->stream>>>(ka);
    SG_LAUNCHED(ctx);
    unsigned *ks, *vs;
    SG_TRY(sg_radix_sort(ctx, keys, vals, keys_tmp, vals_tmp, M, (unsigned)((long long)lvl->w * lvl->h), &ks, &vs));
    FoldGatherArgs fg;
    fg.keys = ks; fg.vals = vs; fg.M = M; fg.ent_slot = ent_slot; fg.impact = tr.impact; fg.rec = tr.rec; fg.stride = lvl->stride;
    fg.g_impact = g_impact; fg.g_rec = g_rec;
    k_level_gather<<<nblk, 128, 0, ctx->stream>>>(fg);
    SG_LAUNCHED(ctx);
    FoldArgs fa;
    fa.keys = ks; fa.vals = vs; fa.M = M; fa.g_impact = g_impact; fa.g_rec = g_rec;
    fa.cells = lvl->d_cells; fa.stride = lvl->stride; fa.model = lvl->model; fa.oie = p->oie; fa.alive_next = alive_next;
    fa.n_long = n_long; fa.long_runs = long_runs;
    k_level_fold<<<nblk, 128, 0, ctx->stream>>>(fa);
    k_level_fold_long<<<ctx->sm_count, 128, 0, ctx->stream>>>(fa);
    SG_LAUNCHED(ctx);
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
    sg_map_invalidate_lut(lvl);
    std::swap(alive, alive_next);
    SG_TRY(ensure_continuous(p));  // a grown 2x2 level gets coarser levels above it
  }
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

extern "C" int slamgpu_pyramid_append_scan(slamgpu_pyramid *p, slamgpu_scan *scan, const double pose[3], double scan_quality,
                                           int32_t scan_margin, const slamgpu_estimator *est, double blur, double max_range,
                                           const double *point_quality, int64_t *cells_updated) {
  SG_NVTX("K4 pyramid_append_scan");
  if (!p) return SLAMGPU_E_INVALID;
  SG_TRY(ensure_continuous(p));
  AppendTrace tr;
  tr.oie = p->oie;
  SG_TRY(sg_append_scan_impl(p->ctx, p->lv[0], scan, pose, scan_quality, scan_margin, est, blur, max_range, point_quality,
                             cells_updated, &tr));
  return pyramid_propagate(p, tr);
}

extern "C" int slamgpu_pyramid_append_beams(slamgpu_pyramid *p, int32_t n, const double *beams, const uint8_t *is_occ,
                                            const double *quality, const slamgpu_estimator *est, double blur, double max_range,
                                            int64_t *cells_updated) {
  if (!p) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  if (n < 0 || (n > 0 && (!beams || !is_occ || !quality)) || !est) return sg_fail(ctx, SLAMGPU_E_INVALID, "pyramid_append_beams: bad argument");
  if (cells_updated) *cells_updated = 0;
  if (n == 0) return SLAMGPU_OK;
  SG_TRY(ensure_continuous(p));
  BeamPlan plan;
  SG_TRY(sg_plan_from_beams(ctx, p->lv[0], n, beams, is_occ, quality, blur, max_range, &plan));
  AppendTrace tr;
  tr.oie = p->oie;
  SG_TRY(sg_append_plan(ctx, p->lv[0], plan, est, cells_updated, &tr));
  return pyramid_propagate(p, tr);
}

extern "C" int slamgpu_score_windows(slamgpu_pyramid *p, slamgpu_scan *const *scans, int32_t n_scans, const int32_t *scan_id,
                                     const double *windows, int64_t M, const double pose[3], const slamgpu_spe_params *spe,
                                     double *out_bounds) {
  if (!p) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  if (!scans || n_scans <= 0 || M < 0 || (M > 0 && (!scan_id || !windows || !out_bounds)) || !pose || !spe)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "score_windows: bad argument");
  if (spe->oope != SLAMGPU_OOPE_MAX || !spe->prerotated)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "score_windows: the Match bound is defined for the max OOPE on pre-rotated scans");
  if (M == 0) return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  SG_TRY(ensure_continuous(p));
  const int L = (int)p->lv.size();
  std::vector<ScanView> sv(n_scans);
  int n_max = 1;
  for (int k = 0; k < n_scans; ++k) {
    const slamgpu_scan *s = scans[k];
    if (!s || s->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "scan %d is NULL or belongs to another ctx", k);
    SG_TRY(sg_scan_ensure_xy(scans[k]));
    sv[k] = ScanView{s->d_x, s->d_y, s->d_w, s->d_f, s->n, s->has_factor ? 1 : 0, s->wsum};
    n_max = std::max(n_max, s->n);
  }
  std::vector<MatchRec> mr((size_t)M);
  std::vector<char> used(L, 0);
  for (int64_t m = 0; m < M; ++m) {
    const double bot = windows[4 * m], top = windows[4 * m + 1], left = windows[4 * m + 2], right = windows[4 * m + 3];
    const double vside = top - bot, hside = right - left;  // LightWeightRectangle::vside/hside
    const double cx = left + hside / 2, cy = bot + vside / 2;  // ::center, geometry_primitives.h:191-193
    if (scan_id[m] < 0 || scan_id[m] >= n_scans) return sg_fail(ctx, SLAMGPU_E_INVALID, "match %lld: bad scan id", (long long)m);
    MatchRec &r = mr[m];
    r.dx = pose[0] + cx; r.dy = pose[1] + cy; r.vside = vside; r.hside = hside; r.scan_id = scan_id[m];
    r.level = rescale_id(p, std::max(vside, hside));  // Match::Match :176
    used[r.level] = 1;
  }
  std::vector<MapView> views(L);
  for (int l = 0; l < L; ++l) {
    if (used[l]) SG_TRY(sg_map_ensure_lut(p->lv[l], spe->oie));
    views[l] = level_view(p->lv[l], spe->oie);
  }
  // multi-GPU: a large batch (the root matches of all rotations) is split into contiguous chunks, one per rank, on the
  // replicated pyramid; one all-gather of the bounds (8 B x M) puts the whole batch on every rank, so the host engine
  // advances identically everywhere.  Small batches (children of a branch) are scored redundantly: no collective.
  const bool shard = ctx->nranks > 1 && M >= 4 * (int64_t)ctx->nranks;
  const int64_t chunk = shard ? (M + ctx->nranks - 1) / ctx->nranks : M;
  const int64_t m0 = shard ? std::min<int64_t>(M, chunk * ctx->rank) : 0;
  const int64_t m1 = shard ? std::min<int64_t>(M, m0 + chunk) : M;
  const int64_t Ml = m1 - m0;
  const int64_t Mpad = shard ? chunk * ctx->nranks : M;
  // one staging block {views | scans | matches} through pinned memory: one copy per call
  auto up8 = [](size_t b) { return (b + 63) & ~(size_t)63; };
  const size_t off_s = up8(sizeof(MapView) * L), off_m = off_s + up8(sizeof(ScanView) * n_scans);
  const size_t stage_bytes = off_m + sizeof(MatchRec) * M;
  if (p->views.reserve(stage_bytes) != SLAMGPU_OK || p->terms.reserve(sizeof(double) * std::max<int64_t>(Ml, 1) * n_max) != SLAMGPU_OK ||
      p->bounds.reserve(sizeof(double) * Mpad) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "score_windows buffers");
  void *hp;
  SG_TRY(sg_pinned(ctx, stage_bytes + sizeof(double) * M, &hp));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging may still be in flight
  memcpy(hp, views.data(), sizeof(MapView) * L);
  memcpy((char *)hp + off_s, sv.data(), sizeof(ScanView) * n_scans);
  memcpy((char *)hp + off_m, mr.data(), sizeof(MatchRec) * M);
  SG_CUDA(ctx, cudaMemcpyAsync(p->views.p, hp, stage_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const MapView *d_views = p->views.as<MapView>();
  const ScanView *d_scans = (const ScanView *)((const char *)p->views.p + off_s);
  const MatchRec *d_matches = (const MatchRec *)((const char *)p->views.p + off_m);
  double *h_bounds = (double *)((char *)hp + stage_bytes);
  if (Ml > 0) {
    dim3 grd((n_max + 127) / 128, (unsigned)Ml);
    cudaEventRecord(ctx->evk0, ctx->stream);
    k_window_terms<<<grd, 128, 0, ctx->stream>>>(d_matches + m0, d_scans, d_views, n_max, p->terms.as<double>());
    cudaEventRecord(ctx->evk1, ctx->stream);
    ctx->evk_valid = true;
    // rows staged in shared memory when up to 8 of them fit (they do below ~5600 points); else straight from global
    const int G = (int)std::min<size_t>(8, (size_t)(200 * 1024) / ((size_t)n_max * sizeof(double)));
    if (G >= 1) {
      const size_t smem = (size_t)G * n_max * sizeof(double);
      static bool attr_set = false;
      if (!attr_set) { cudaFuncSetAttribute(k_ordered_sums_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
      k_ordered_sums_staged<<<(unsigned)((Ml + G - 1) / G), 256, smem, ctx->stream>>>(d_matches + m0, d_scans, n_max, p->terms.as<double>(), Ml, G,
                                                                                   p->bounds.as<double>() + m0);
    } else {
      k_ordered_sums<<<(unsigned)((Ml + 63) / 64), 64, 0, ctx->stream>>>(d_matches + m0, d_scans, n_max, p->terms.as<double>(), Ml,
                                                                        p->bounds.as<double>() + m0);
    }
    ctx->launches += 2;
    SG_CUDA(ctx, cudaGetLastError());
  }
  if (shard) {
    std::string err;
    double *b = p->bounds.as<double>();
    int r = sg_nccl_allgather(ctx->comm, b + chunk * ctx->rank, b, sizeof(double) * chunk, ctx->stream, &err);  // in place
    if (r != SLAMGPU_OK) return sg_fail(ctx, r, "%s", err.c_str());
  }
  SG_CUDA(ctx, cudaMemcpyAsync(h_bounds, p->bounds.p, sizeof(double) * M, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out_bounds, h_bounds, sizeof(double) * M);
  return SLAMGPU_OK;
}

// ---------------------------------------------------------------- BF-M3RSM matcher (host engine over K5)
// BruteForceMultiResolutionScanMatcher::process_scan (src/core/scan_matchers/bf_multi_res_scan_matcher.h:24-66)
// over M3RSMEngine (m3rsm_engine.h:252-365): a best-first branch and bound over (rotation, translation
// window) matches.  The engine logic (priority queue with the reference's eps comparator, pruning,
// branching order) runs on the host exactly as upstream; every batch of new matches -- all roots at
// once, then the 2/4 children of a branch or the 5 point hypotheses of a leaf -- is scored by one K5 call.
namespace {

bool h_are_equal(double a, double b) {
  double sc = std::max(1.0, std::max(std::fabs(a), std::fabs(b)));
  return std::fabs(a - b) <= 1e-7 * sc;
}
bool h_less(double a, double b) { return a < b + 2.220446049250313e-16; }
bool h_less_or_equal(double a, double b) { return h_are_equal(a, b) || h_less(a, b); }

struct HMatch {
  double bound, rotation, bot, top, left, right;
  int scan_id;
  double abs_rotation, drift_amount;
  double hside() const { return right - left; }
  double vside() const { return top - bot; }
  bool is_finest() const { return drift_amount <= 0; }
  // Match::operator< (m3rsm_engine.h:192-203): true if *this is less preferable
  bool operator<(const HMatch &that) const {
    if (!h_are_equal(bound, that.bound)) return h_less(bound, that.bound);
    if (!h_are_equal(drift_amount, that.drift_amount)) return drift_amount > that.drift_amount;
    return abs_rotation > that.abs_rotation;
  }
};

HMatch make_match(double rot, double bot, double top, double left, double right, int scan_id) {
  HMatch m;
  m.bound = 0; m.rotation = rot; m.bot = bot; m.top = top; m.left = left; m.right = right; m.scan_id = scan_id;
  m.abs_rotation = std::fabs(rot);
  m.drift_amount = (right - left) + (top - bot);  // hside_len() + vside_len()
  return m;
}

}  // namespace

#include <atomic>
#include <chrono>
#include <functional>
#include <queue>
#include <set>
#include <unordered_map>

namespace {
struct M3Rot { double rot; int scan_id; };
// scores M matches: scan (= rotation) id and window {bot, top, left, right} per match -> bounds
using M3RawScore = std::function<int(const int32_t *scan_id, const double *windows, int64_t M, double *bounds)>;

// M3RSMEngine's best-first search over a scoring function (the K5 kernels, or a host callback in the CPU-only test hook)
int m3rsm_search(const std::vector<M3Rot> &rots, double x_limit, double y_limit, double transl_step, double max_finest_prob_diff,
                 const M3RawScore &raw_score, double out_delta[3], double *out_prob, int64_t stats[4]) {
  int64_t n_scored = 0, n_calls = 0, n_branches = 0;
  std::vector<int32_t> sid;
  std::vector<double> win, bounds;
  // what the engine does with a match it pops: split a window that is still wider than the target accuracy
  // (branch :337-357), or test the corners and the centre of a leaf as point windows; nothing for a point
  auto expansions = [&](const HMatch &m, std::vector<HMatch> &out) {
    const bool horz = h_less(transl_step, m.hside()), vert = h_less(transl_step, m.vside());
    const double cx = m.left + m.hside() / 2, cy = m.bot + m.vside() / 2;  // center()
    auto child = [&](double b, double t, double l, double r) { out.push_back(make_match(m.rotation, b, t, l, r, m.scan_id)); };
    if (horz && vert) {  // split4_evenly: left-bot, left-top, right-bot, right-top
      child(m.bot, cy, m.left, cx); child(cy, m.top, m.left, cx);
      child(m.bot, cy, cx, m.right); child(cy, m.top, cx, m.right);
    } else if (horz) {   // split_horz
      child(m.bot, m.top, m.left, cx); child(m.bot, m.top, cx, m.right);
    } else if (vert) {   // split_vert
      child(m.bot, cy, m.left, m.right); child(cy, m.top, m.left, m.right);
    } else if (!m.is_finest()) {  // exact translation hypotheses: the four corners and the centre
      const double px[5] = {m.left, m.left, m.right, m.right, cx};
      const double py[5] = {m.bot, m.top, m.bot, m.top, cy};
      for (int k = 0; k < 5; ++k) child(py[k], py[k], px[k], px[k]);
    }
  };
  // A bound is a pure function of (rotation, window), so it may be computed before the engine asks for it.  Every K5
  // call therefore also scores what the engine is likely to ask next -- the expansions of the matches being added and
  // of the best matches already queued -- and keeps the bounds; the engine itself (queue order, pruning, branching)
  // is untouched and asks for exactly the matches upstream scores, most of which are then already known.
  struct Key {
    int scan_id; double b, t, l, r;
    bool operator==(const Key &o) const { return scan_id == o.scan_id && b == o.b && t == o.t && l == o.l && r == o.r; }
  };
  struct KeyHash {
    size_t operator()(const Key &k) const {
      uint64_t h = (uint64_t)k.scan_id * 0x9E3779B97F4A7C15ull;
      for (double d : {k.b, k.t, k.l, k.r}) { uint64_t u; memcpy(&u, &d, 8); h = (h ^ u) * 0x100000001B3ull; h ^= h >> 29; }
      return (size_t)h;
    }
  };
  std::unordered_map<Key, double, KeyHash> known;
  known.reserve(8192);  // a match scores ~1000 windows: no rehash on the way
  auto key_of = [](const HMatch &m) { return Key{m.scan_id, m.bot, m.top, m.left, m.right}; };
  std::vector<HMatch> heap;  // std::priority_queue's own algorithm (push_heap / pop_heap), kept open for peeking
  std::vector<HMatch> ask, spec;
  heap.reserve(2048); ask.reserve(1024); spec.reserve(4096);
  // (a call costs ~25 us however many matches it carries up to a few thousand: 768 / 96 halves the calls of 192 / 12)
  const size_t kSpeculate = 768, kPeek = 96;
  auto score = [&](std::vector<HMatch> &ms) -> int {
    if (ms.empty()) return (int)SLAMGPU_OK;
    ask.clear();
    auto want = [&](const HMatch &m) {
      const Key k = key_of(m);
      if (known.count(k)) return;
      known.emplace(k, NAN);
      ask.push_back(m);
    };
    for (const HMatch &m : ms) want(m);
    if (!ask.empty()) {  // a call is being made anyway: fill it up with likely next requests
      spec.clear();
      for (const HMatch &m : ms) expansions(m, spec);
      // the first entries of the heap array are its top levels: not exactly the kPeek best matches, but close enough for
      // a guess, and free (no copy of the queue, no pops)
      for (size_t k = 0; k < kPeek && k < heap.size(); ++k) expansions(heap[k], spec);
      for (const HMatch &m : spec) {
        if (ask.size() >= ms.size() + kSpeculate) break;
        want(m);
      }
      sid.clear(); win.clear();
      for (const HMatch &m : ask) {
        sid.push_back(m.scan_id);
        win.push_back(m.bot); win.push_back(m.top); win.push_back(m.left); win.push_back(m.right);
      }
      bounds.resize(ask.size());
      { int rc_ = raw_score(sid.data(), win.data(), (int64_t)ask.size(), bounds.data()); if (rc_ != SLAMGPU_OK) return rc_; }
      for (size_t k = 0; k < ask.size(); ++k) known[key_of(ask[k])] = bounds[k];
      n_scored += (int64_t)ask.size(); ++n_calls;
    }
    for (HMatch &m : ms) m.bound = known[key_of(m)];
    return (int)SLAMGPU_OK;
  };
  // ---- engine state (M3RSMEngine :252-289)
  double best_finest = 0.0;
  auto add_match = [&](const HMatch &m) {
    if (m.bound < best_finest) return;
    if (m.is_finest()) best_finest = std::max(best_finest, m.bound - max_finest_prob_diff);
    heap.push_back(m);
    std::push_heap(heap.begin(), heap.end());
  };
  std::vector<HMatch> batch;
  for (const M3Rot &r : rots) {
    batch.push_back(make_match(r.rot, 0, 0, 0, 0, r.scan_id));
    batch.push_back(make_match(r.rot, -y_limit, y_limit, -x_limit, x_limit, r.scan_id));
  }
  { int rc_ = score(batch); if (rc_ != SLAMGPU_OK) return rc_; }
  for (const HMatch &m : batch) add_match(m);
  // ---- best-first search (bf_multi_res_scan_matcher.h:44-65, next_best_match :318-333, branch :337-357)
  for (;;) {
    HMatch best;
    bool found = false;
    while (!heap.empty()) {
      std::pop_heap(heap.begin(), heap.end());
      best = heap.back();
      heap.pop_back();
      const bool horz = h_less(transl_step, best.hside()), vert = h_less(transl_step, best.vside());
      if (!horz && !vert) { found = true; break; }
      ++n_branches;
      batch.clear();
      expansions(best, batch);
      { int rc_ = score(batch); if (rc_ != SLAMGPU_OK) return rc_; }
      for (const HMatch &m : batch) add_match(m);
    }
    if (!found) return SLAMGPU_E_STATE;  // the match queue ran empty
    if (best.is_finest()) {
      out_delta[0] = best.left + best.hside() / 2;
      out_delta[1] = best.bot + best.vside() / 2;
      out_delta[2] = best.rotation;
      *out_prob = best.bound;
      break;
    }
    batch.clear();
    expansions(best, batch);
    { int rc_ = score(batch); if (rc_ != SLAMGPU_OK) return rc_; }
    for (const HMatch &m : batch) add_match(m);
  }
  if (stats) { stats[0] = n_scored; stats[1] = n_calls; stats[2] = n_branches; stats[3] = (int64_t)rots.size(); }
  return SLAMGPU_OK;
}
}  // namespace

extern "C" int slamgpu_match_m3rsm(slamgpu_pyramid *p, int32_t n, const double *range, const double *angle,
                                   const double *weight, const double pose[3], const slamgpu_spe_params *spe, double x_limit,
                                   double y_limit, double rot_limit, double ang_step, double transl_step,
                                   double max_finest_prob_diff, double out_delta[3], double *out_prob, int64_t stats[4]) {
  SG_NVTX("K5 match_m3rsm");
  if (!p) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  if (n < 0 || (n > 0 && (!range || !angle)) || !pose || !spe || !out_delta || !out_prob || !(ang_step > 0))
    return sg_fail(ctx, SLAMGPU_E_INVALID, "match_m3rsm: bad argument");
  slamgpu_spe_params sp = *spe;
  sp.oope = SLAMGPU_OOPE_MAX; sp.prerotated = 1;
  // ---- rotations in the engine's order (add_scan_matching_request :291-316)
  std::vector<M3Rot> rots;
  const double sector = 2 * rot_limit;
  for (double rd = 0; h_less_or_equal(2 * rd, sector); rd += ang_step)
    for (double rot : std::set<double>{rd, -rd}) rots.push_back(M3Rot{rot, (int)rots.size()});
  const bool timing = getenv("SLAMGPU_M3_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_a = now();
  double t_b = t_a, t_c = t_a, t_raw = 0;
  int rc;
  const size_t R = rots.size();
  static const bool env_fast = [] { const char *e = getenv("SLAMGPU_M3_FAST"); return !e || atoi(e) != 0; }();
  if (ctx->nranks == 1 && n > 0 && n <= 6000 && env_fast) {
    // ---- one GPU: everything the match needs goes up in ONE copy -- {level views | scan views | x | y | weights} -- and
    // every scoring call is one launch of k_match_bounds reading its records from, and writing its bounds to, pinned
    // host memory (no copies, no staging kernels: ~20 us per call instead of ~50)
    SG_CUDA(ctx, cudaSetDevice(ctx->device));
    SG_TRY(ensure_continuous(p));
    const int L = (int)p->lv.size();
    auto up64 = [](size_t b) { return (b + 63) & ~(size_t)63; };
    const size_t off_sv = up64(sizeof(MapView) * L), off_x = off_sv + up64(sizeof(ScanView) * R);
    const size_t off_y = off_x + sizeof(double) * n * R, off_w = off_y + sizeof(double) * n * R;
    const size_t dev_bytes = off_w + sizeof(double) * n;
    const size_t max_batch = 4096;
    const size_t io_bytes = dev_bytes + max_batch * (sizeof(MatchRec) + sizeof(double)) + 64;
    if (p->m3_io_cap < io_bytes) {
      if (p->m3_io) cudaFreeHost(p->m3_io);
      p->m3_io = nullptr; p->m3_io_cap = 0;
      SG_CUDA(ctx, cudaHostAlloc(&p->m3_io, io_bytes, cudaHostAllocMapped));
      p->m3_io_cap = io_bytes;
    }
    if (p->m3_dev.reserve(dev_bytes + 64) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "match_m3rsm buffers");
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the io block may still be read by an earlier call
    char *h = (char *)p->m3_io, *d = p->m3_dev.as<char>();
    MapView *hv = (MapView *)h;
    for (int l = 0; l < L; ++l) {
      SG_TRY(sg_map_ensure_lut(p->lv[l], sp.oie));
      hv[l] = level_view(p->lv[l], sp.oie);
    }
    // pre-rotated Cartesian copies of the scan: LaserScan2D::to_cartesian(rot + pose.theta) (src/core/states/sensor_data.h:
    // 156-167, RawTrigonometryProvider: cos(base + angle)); libm on the host, as the reference computes them
    double *hx = (double *)(h + off_x), *hy = (double *)(h + off_y), *hw = (double *)(h + off_w);
    for (size_t k = 0; k < R; ++k) {
      const double base = rots[k].rot + pose[2];
      double *x = hx + k * (size_t)n, *y = hy + k * (size_t)n;
      for (int i = 0; i < n; ++i) {
        x[i] = 0 + range[i] * std::cos(base + angle[i]);
        y[i] = 0 + range[i] * std::sin(base + angle[i]);
      }
    }
    double ws = 0;
    for (int i = 0; i < n; ++i) { hw[i] = weight ? weight[i] : 1.0 / n; ws += hw[i]; }  // weighted_mean_point_probability_spe.h:125
    ScanView *hs = (ScanView *)(h + off_sv);
    for (size_t k = 0; k < R; ++k)
      hs[k] = ScanView{(const double *)(d + off_x) + k * (size_t)n, (const double *)(d + off_y) + k * (size_t)n, (const double *)(d + off_w),
                       nullptr, n, 0, ws};
    t_b = now();
    SG_CUDA(ctx, cudaMemcpyAsync(d, h, dev_bytes, cudaMemcpyHostToDevice, ctx->stream));
    unsigned *d_done = (unsigned *)(d + dev_bytes);
    SG_CUDA(ctx, cudaMemsetAsync(d_done, 0, 64, ctx->stream));
    t_c = now();
    MatchRec *h_mr = (MatchRec *)(h + dev_bytes);
    double *h_bounds = (double *)(h + dev_bytes + max_batch * sizeof(MatchRec));
    volatile unsigned *h_flag = (volatile unsigned *)(h + dev_bytes + max_batch * (sizeof(MatchRec) + sizeof(double)));
    *h_flag = 0;
    unsigned seq = 0;
    const size_t smem = sizeof(double) * (size_t)n;
    M3RawScore raw = [&](const int32_t *sid, const double *win, int64_t M, double *bounds) -> int {
      const double t0 = timing ? now() : 0;
      for (int64_t m0 = 0; m0 < M; m0 += (int64_t)max_batch) {
        const int64_t mb = std::min<int64_t>((int64_t)max_batch, M - m0);
        for (int64_t m = 0; m < mb; ++m) {
          const double *wn = win + 4 * (m0 + m);
          const double vside = wn[1] - wn[0], hside = wn[3] - wn[2];         // LightWeightRectangle::vside/hside
          const double cx = wn[2] + hside / 2, cy = wn[0] + vside / 2;       // ::center, geometry_primitives.h:191-193
          MatchRec &r = h_mr[m];
          r.dx = pose[0] + cx; r.dy = pose[1] + cy; r.vside = vside; r.hside = hside; r.scan_id = sid[m0 + m];
          r.level = rescale_id(p, std::max(vside, hside));  // Match::Match :176
        }
        ++seq;
        k_match_bounds<<<(unsigned)mb, 512, smem, ctx->stream>>>(h_mr, (const ScanView *)(d + off_sv), (const MapView *)d, h_bounds, d_done,
                                                               h_flag, seq);
        ctx->launches += 1;
        SG_CUDA(ctx, cudaGetLastError());
        // poll the completion flag (a few microseconds cheaper than a stream synchronisation); a failed launch or a
        // device fault never sets it, so look at the stream now and then
        int idle_seen = 0;
        for (unsigned spins = 0; *h_flag != seq; ++spins) {
          if ((spins & 0xFFFFu) == 0xFFFFu) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) return sg_fail(ctx, SLAMGPU_E_CUDA, "match_m3rsm: %s", cudaGetErrorString(q));
            // the stream is idle: the flag must be visible by the next look or two
            if (++idle_seen > 2 && *h_flag != seq) return sg_fail(ctx, SLAMGPU_E_CUDA, "match_m3rsm: the scoring kernel finished without reporting");
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        memcpy(bounds + m0, h_bounds, sizeof(double) * mb);
      }
      if (timing) t_raw += now() - t0;
      return SLAMGPU_OK;
    };
    rc = m3rsm_search(rots, x_limit, y_limit, transl_step, max_finest_prob_diff, raw, out_delta, out_prob, stats);
  } else {
    // ---- sharded over ranks (or a scan too long for one block's shared memory): device scan objects + slamgpu_score_windows
    std::vector<slamgpu_scan *> &pool = p->rot_scans;
    while (pool.size() < R) {
      slamgpu_scan *s = nullptr;
      SG_TRY(slamgpu_scan_create(ctx, &s));
      pool.push_back(s);
    }
    std::vector<double> xs(std::max<size_t>((size_t)n * R, 1)), ys(xs.size());
    for (size_t k = 0; k < R; ++k) {
      const double base = rots[k].rot + pose[2];
      double *x = xs.data() + k * (size_t)n, *y = ys.data() + k * (size_t)n;
      for (int i = 0; i < n; ++i) {
        x[i] = 0 + range[i] * std::cos(base + angle[i]);
        y[i] = 0 + range[i] * std::sin(base + angle[i]);
      }
    }
    t_b = now();
    SG_TRY(sg_scans_upload_xy(ctx, pool.data(), (int)R, n, xs.data(), ys.data(), weight));
    t_c = now();
    M3RawScore raw = [&](const int32_t *sid, const double *win, int64_t M, double *bounds) -> int {
      const double t0 = timing ? now() : 0;
      int r = slamgpu_score_windows(p, pool.data(), (int32_t)R, sid, win, M, pose, &sp, bounds);
      if (timing) t_raw += now() - t0;
      return r;
    };
    rc = m3rsm_search(rots, x_limit, y_limit, transl_step, max_finest_prob_diff, raw, out_delta, out_prob, stats);
  }
  if (timing)
    fprintf(stderr, "[m3rsm] trig + views %.0f us, upload %.0f us, search %.0f us of which K5 calls %.0f us (%lld calls, %lld matches)\n",
            t_b - t_a, t_c - t_b, now() - t_c, t_raw, stats ? (long long)stats[1] : -1ll, stats ? (long long)stats[0] : -1ll);
  if (rc == SLAMGPU_E_STATE) return sg_fail(ctx, SLAMGPU_E_STATE, "match_m3rsm: the match queue ran empty");
  return rc;
}

// Test hook (no GPU, no ctx): the same search over a caller-supplied bound function, so that CPU-only tests can hold
// the engine -- queue order, pruning, branching, the speculative filling of every scoring call -- against the
// reference's BruteForceMultiResolutionScanMatcher.
extern "C" int slamgpu_debug_m3rsm(double x_limit, double y_limit, double rot_limit, double ang_step, double transl_step,
                                   double max_finest_prob_diff, slamgpu_bounds_fn fn, void *user, double out_delta[3],
                                   double *out_prob, int64_t stats[4]) {
  if (!fn || !out_delta || !out_prob || !(ang_step > 0)) return SLAMGPU_E_INVALID;
  std::vector<M3Rot> rots;
  const double sector = 2 * rot_limit;
  for (double rd = 0; h_less_or_equal(2 * rd, sector); rd += ang_step)
    for (double rot : std::set<double>{rd, -rd}) rots.push_back(M3Rot{rot, (int)rots.size()});
  std::vector<double> rv;
  M3RawScore raw = [&](const int32_t *sid, const double *win, int64_t M, double *bounds) -> int {
    rv.resize((size_t)M);
    for (int64_t k = 0; k < M; ++k) rv[k] = rots[sid[k]].rot;
    fn((int32_t)M, rv.data(), win, bounds, user);
    return SLAMGPU_OK;
  };
  return m3rsm_search(rots, x_limit, y_limit, transl_step, max_finest_prob_diff, raw, out_delta, out_prob, stats);
}

#define SG_TOUCH(k) do { cudaFuncAttributes fa_; (void)cudaFuncGetAttributes(&fa_, k); } while (0)
void sg_preload_pyramid() {  // see sg_preload_score
  SG_TOUCH(k_order_entries); SG_TOUCH(k_level_coords); SG_TOUCH(k_level_keys); SG_TOUCH(k_level_gather); SG_TOUCH(k_level_fold); SG_TOUCH(k_level_fold_long);
  SG_TOUCH(k_build_level); SG_TOUCH(k_build_fused<SLAMGPU_CELL_LWW>); SG_TOUCH(k_build_fused<SLAMGPU_CELL_MEAN>); SG_TOUCH(k_build_fused<SLAMGPU_CELL_TBM_CONSISTENT>); SG_TOUCH(k_build_fused<SLAMGPU_CELL_GMAPPING>); SG_TOUCH(k_fill_level); SG_TOUCH(k_window_terms); SG_TOUCH(k_ordered_sums); SG_TOUCH(k_ordered_sums_staged);
  (void)cudaGetLastError();
}
