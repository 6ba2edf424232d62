// mapping.cu -- K2 (ray casting) and K3 (occupancy estimation + ordered cell update).
//
// Replaces GridMapScanAdder::append_scan + WallDistanceBlurringScanAdder::handle_scan_point
// (src/core/maps/grid_map_scan_adders.h:54-75, 138-189), RegularSquaresGrid::world_to_cells
// (src/core/maps/regular_squares_grid.h:56-101), the const / area occupancy estimators and the
// cell models' operator+=.
//
// Pipeline of one scan insertion (all on the ctx stream, one host sync at the end):
//   host   per-beam record: end point (libm trig, bit-identical to the reference), range gate,
//          blur radius, slot offsets from the |dx|+|dy|+1 upper bound of each beam's cell count
//   K2     k_raycast       one thread per beam walks its cells in the reference's order into the
//                          beam's slot range and estimates the obstacle (last) cell
//   K3a    k_estimate      one thread per (beam, cell) slot: occupancy estimate + wall blur -> AOO,
//                          sort key = internal cell index
//   sort   k_radix_*       stable LSD radix sort of (cell, slot): slots are generated in beam order,
//                          so every cell's run comes out in the reference's update order
//   K3b    k_apply         one thread per run walks it sequentially through the cell model's +=
// The reference updates cells beam by beam (obstacle cell first, then along the beam); a cell
// appears at most once per beam, so per-cell order == beam order, which the stable sort keeps.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "dev_geometry.cuh"
#include "mapping.h"

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

// ---------------------------------------------------------------- K2
struct RaycastArgs {
  const BeamRec *beams;
  const long long *offsets;  // slot offset of each beam (exclusive prefix of the upper bounds)
  int N;
  const MapSlot *maps;
  double scale;
  slamgpu_estimator est;
  int2 *cells;      // per slot
  BeamOut *out;     // per beam
  unsigned long long *counters;  // per map: [2*id] cells visited (one add per beam)
};

__global__ void k_raycast(RaycastArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const BeamRec b = a.beams[i];
  BeamOut o;
  o.count = 0; o.base_p = 0; o.base_q = 0; o.lx = o.ly = 0;
  if (b.active) {
    int2 *dst = a.cells + a.offsets[i];
    int lx = 0, ly = 0;
    const double px = a.maps[b.map_id].px, py = a.maps[b.map_id].py;
    o.count = sg::raycast(px, py, b.wx, b.wy, a.scale, [&](int k, int x, int y) {
      dst[k] = make_int2(x, y);
      lx = x; ly = y;
    });
    o.lx = lx; o.ly = ly;
    atomicAdd(a.counters + 2 * b.map_id, (unsigned long long)o.count);
    // the obstacle cell is estimated (and, in the reference, updated) first: grid_map_scan_adders.h:151-154
    const double s = a.scale;
    sg::estimate_occupancy(a.est, a.maps[b.map_id].shift, px, py, b.wx, b.wy, sg::mul(s, (double)ly), sg::mul(s, (double)(ly + 1)),
                           sg::mul(s, (double)lx), sg::mul(s, (double)(lx + 1)), b.is_occ != 0, &o.base_p, &o.base_q);
  }
  a.out[i] = o;
}

// ---------------------------------------------------------------- K3a
struct EstimateArgs {
  const BeamRec *beams;
  const BeamOut *bout;
  const long long *offsets;  // N + 1
  int N;
  long long M;  // total slots
  const MapSlot *maps;
  double scale;
  slamgpu_estimator est;
  const int2 *cells;
  double *aoo_p, *aoo_q;  // per slot
  unsigned *keys;         // per slot: internal cell index or SG_INVALID_KEY
  unsigned *vals;         // per slot: slot id
  int *slot_beam;         // per slot
  unsigned long long *counters;  // per map: [2*id+1] updates dropped outside a bounded map ([2*id]: see k_raycast)
  int ring;               // >= 0: cells within this many cells of the robot are taken out of the sort (k_apply_ring)
  int head;               // k_estimate skips the first `head` cells of every ray (k_estimate_head has done them)
};

// AREA = false: the const estimator only (a select); the kernel then does not carry the area estimator's registers
// everything k_estimate does for slot s = cell k of beam i
template <bool AREA>
SG_DEV void estimate_slot(const EstimateArgs &a, long long s, int i, int k) {
  const BeamOut bo = a.bout[i];
  a.vals[s] = (unsigned)s;
  a.slot_beam[s] = i;
  if (k >= bo.count) { a.keys[s] = SG_INVALID_KEY; return; }
  const BeamRec b = a.beams[i];
  const MapSlot ms = a.maps[b.map_id];
  const int2 c = a.cells[s];
  double p, q;
  if (k == bo.count - 1) {
    p = bo.base_p; q = bo.base_q;
  } else {
    if (AREA) {
      const double sc = a.scale;
      sg::estimate_occupancy(a.est, ms.shift, ms.px, ms.py, b.wx, b.wy, sg::mul(sc, (double)c.y), sg::mul(sc, (double)(c.y + 1)),
                             sg::mul(sc, (double)c.x), sg::mul(sc, (double)(c.x + 1)), false, &p, &q);
    } else {
      p = a.est.empty_p; q = a.est.empty_q;  // ConstOccupancyEstimator: a free cell of the ray
    }
    // wall blur ("hole"), grid_map_scan_adders.h:160-169; distances in cells, squared (exact in double)
    double ddx = (double)(c.x - b.obx), ddy = (double)(c.y - b.oby);
    double d_sq = sg::add(sg::mul(ddx, ddx), sg::mul(ddy, ddy));
    if (d_sq < b.hole_sq && b.hole_sq < b.obst_sq) {
      double prob_scale = sg::sub(1.0, sg::div(d_sq, b.hole_sq));
      p = sg::mul(bo.base_p, prob_scale);
    }
  }
  a.aoo_p[s] = p; a.aoo_q[s] = q;
  int ix = c.x + ms.ox, iy = c.y + ms.oy;
  if (ix < 0 || ix >= ms.w || iy < 0 || iy >= ms.h) {
    a.keys[s] = SG_INVALID_KEY;
    atomicAdd(a.counters + 2 * b.map_id + 1, 1ull);
  } else {
    if (a.ring >= 0 && abs(c.x - ms.rx) <= a.ring && abs(c.y - ms.ry) <= a.ring) {
      a.keys[s] = SG_INVALID_KEY;  // applied by k_apply_ring, in beam order
    } else {
      a.keys[s] = ms.key_base + (unsigned)iy * (unsigned)ms.w + (unsigned)ix;
    }
  }
}

template <bool AREA>
__global__ void __launch_bounds__(128) k_estimate(EstimateArgs a) {
  const long long s0 = blockIdx.x * (long long)blockDim.x;
  long long s = s0 + threadIdx.x;
  // beam of a slot: last i with offsets[i] <= s.  Two threads bracket the block's slots with a full binary search,
  // the others search inside that bracket (a block rarely spans more than a couple of beams)
  __shared__ int s_range[2];
  if (threadIdx.x < 2) {
    long long t = threadIdx.x == 0 ? s0 : min(s0 + (long long)blockDim.x - 1, a.M - 1);
    int lo = 0, hi = a.N;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (a.offsets[mid] <= t) lo = mid; else hi = mid;
    }
    s_range[threadIdx.x] = lo;
  }
  __syncthreads();
  if (s >= a.M) return;
  int lo = s_range[0], hi = s_range[1] + 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (a.offsets[mid] <= s) lo = mid; else hi = mid;
  }
  const int i = lo;
  const int k = (int)(s - a.offsets[i]);
  if (k < a.head) return;  // done by k_estimate_head, ahead of the ring kernel
  estimate_slot<AREA>(a, s, i, k);
}

// The first `head` cells of every ray, one thread each: all the ring kernel needs (a ray can be inside the window around the
// robot only during its first 2W+1 cells), so that it can start beside the estimate of the other ~95 % of the slots
template <bool AREA>
__global__ void __launch_bounds__(128) k_estimate_head(EstimateArgs a) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int i = (int)(t / a.head), k = (int)(t - (long long)i * a.head);
  if (i >= a.N) return;
  const long long s = a.offsets[i] + k;
  if (s >= a.offsets[i + 1]) return;
  estimate_slot<AREA>(a, s, i, k);
}

// ---------------------------------------------------------------- stable LSD radix sort (8-bit digits)
#define SG_SORT_THREADS 256
#define SG_SORT_ITEMS 8
#define SG_SORT_TILE (SG_SORT_THREADS * SG_SORT_ITEMS)

__global__ void __launch_bounds__(SG_SORT_THREADS) k_radix_hist(const unsigned *__restrict__ keys, long long n, int shift,
                                                                unsigned *__restrict__ ghist, int nb) {
  __shared__ unsigned hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * SG_SORT_TILE;
  for (int it = 0; it < SG_SORT_ITEMS; ++it) {
    long long idx = base + it * SG_SORT_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  ghist[(size_t)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(1024) k_radix_scan(unsigned *data, int n) {  // exclusive scan, one block
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    unsigned v = i < n ? data[i] : 0, x = v;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      unsigned ws = warp_sums[lane], sx = ws;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, sx, off);
        if (lane >= off) sx += y;
      }
      warp_sums[lane] = sx - ws;  // exclusive
    }
    __syncthreads();
    unsigned excl = x - v + warp_sums[wid] + carry;
    if (i < n) data[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

// large histograms: per-chunk exclusive scan (4096 counters a block) + chunk totals, one-block scan of the
// totals (k_radix_scan), then the chunk offsets are added back
#define SG_SCAN_CHUNK 4096
__global__ void __launch_bounds__(1024) k_scan_chunks(unsigned *data, int n, unsigned *totals) {
  __shared__ unsigned warp_sums[32];
  const int base = blockIdx.x * SG_SCAN_CHUNK + threadIdx.x * 4;
  unsigned v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = base + k < n ? data[base + k] : 0u;
  const unsigned mine = v[0] + v[1] + v[2] + v[3];
  unsigned x = mine;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    unsigned ws = warp_sums[lane], sx = ws;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, sx, off);
      if (lane >= off) sx += y;
    }
    warp_sums[lane] = sx - ws;
    if (lane == 31) totals[blockIdx.x] = sx;
  }
  __syncthreads();
  unsigned run = x - mine + warp_sums[wid];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}

__global__ void __launch_bounds__(1024) k_scan_add(unsigned *data, int n, const unsigned *__restrict__ totals) {
  const unsigned add = totals[blockIdx.x];
  const int base = blockIdx.x * SG_SCAN_CHUNK + threadIdx.x * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n) data[base + k] += add;
}

__global__ void __launch_bounds__(SG_SORT_THREADS) k_radix_scatter(const unsigned *__restrict__ keys_in,
                                                                   const unsigned *__restrict__ vals_in,
                                                                   unsigned *__restrict__ keys_out,
                                                                   unsigned *__restrict__ vals_out, long long n, int shift,
                                                                   const unsigned *__restrict__ ghist, int nb) {
  constexpr int NW = SG_SORT_THREADS / 32;
  __shared__ unsigned wcount[NW][256];
  __shared__ unsigned gbase[256];     // global position of this tile's first element of each digit
  __shared__ unsigned dstart[257];    // position of each digit's run inside the tile, once sorted
  __shared__ unsigned skey[SG_SORT_TILE], sval[SG_SORT_TILE];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < NW * 256; k += SG_SORT_THREADS) (&wcount[0][0])[k] = 0;
  gbase[threadIdx.x] = ghist[(size_t)threadIdx.x * nb + blockIdx.x];
  __syncthreads();
  // each warp owns a contiguous 256-element span of the tile and walks it in order, 32 at a time,
  // so ranks within a digit follow the input order (stability)
  const long long tbase = (long long)blockIdx.x * SG_SORT_TILE;
  const long long wbase = tbase + (long long)w * (32 * SG_SORT_ITEMS);
  unsigned key[SG_SORT_ITEMS], val[SG_SORT_ITEMS], rank[SG_SORT_ITEMS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < SG_SORT_ITEMS; ++c) {
    long long idx = wbase + c * 32 + lane;
    bool valid = idx < n;
    key[c] = valid ? keys_in[idx] : 0u;
    val[c] = valid ? vals_in[idx] : 0u;
    unsigned digit = valid ? ((key[c] >> shift) & 255u) : 256u;
    unsigned peers = __match_any_sync(0xffffffffu, digit);
    unsigned r = __popc(peers & lt);
    unsigned pre = valid ? wcount[w][digit] : 0u;
    __syncwarp();
    if (valid && r == 0) wcount[w][digit] = pre + __popc(peers);
    __syncwarp();
    rank[c] = pre + r;
  }
  __syncthreads();
  {  // exclusive prefix over the warps of this block, per digit; its total is the digit's count in the tile
    unsigned run = 0;
    const int d = threadIdx.x;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      unsigned t = wcount[ww][d];
      wcount[ww][d] = run;
      run += t;
    }
    dstart[d + 1] = run;
  }
  if (threadIdx.x == 0) dstart[0] = 0;
  __syncthreads();
  if (threadIdx.x < 32) {  // inclusive scan of the 256 digit counts by one warp, 8 per lane
    unsigned v[8], sum = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { v[u] = dstart[1 + lane * 8 + u]; sum += v[u]; }
    unsigned x = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    unsigned run = x - sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) { run += v[u]; dstart[1 + lane * 8 + u] = run; }
  }
  __syncthreads();
  // the tile, sorted by this digit, in shared memory ...
#pragma unroll
  for (int c = 0; c < SG_SORT_ITEMS; ++c) {
    long long idx = wbase + c * 32 + lane;
    if (idx < n) {
      unsigned digit = (key[c] >> shift) & 255u;
      unsigned pos = dstart[digit] + wcount[w][digit] + rank[c];
      skey[pos] = key[c];
      sval[pos] = val[c];
    }
  }
  __syncthreads();
  // ... then out: neighbouring threads write neighbouring elements of a digit's run (whole sectors, not 4-byte scatters)
  const int cnt = (int)min((long long)SG_SORT_TILE, n - tbase);
  for (int e = threadIdx.x; e < cnt; e += SG_SORT_THREADS) {
    const unsigned k = skey[e];
    const unsigned digit = (k >> shift) & 255u;
    const unsigned dst = gbase[digit] + ((unsigned)e - dstart[digit]);
    keys_out[dst] = k;
    vals_out[dst] = sval[e];
  }
}

// ---------------------------------------------------------------- K3b
// per sorted position: everything one cell update needs, gathered so that the (sequential) apply walk
// streams through contiguous memory instead of chasing slot -> beam -> AOO pointers
struct SortedAoo {
  double p, q, quality, wx, wy;
  unsigned slot;
  int map_id;
};

struct GatherArgs {
  const unsigned *keys, *vals;  // sorted
  long long M;
  const double *aoo_p, *aoo_q;
  const int *slot_beam;
  const BeamRec *beams;
  SortedAoo *out;
};

__global__ void k_gather_sorted(GatherArgs a) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= a.M) return;
  if (a.keys[t] == SG_INVALID_KEY) return;
  const unsigned s = a.vals[t];
  const BeamRec &b = a.beams[a.slot_beam[s]];
  SortedAoo o;
  o.p = a.aoo_p[s]; o.q = a.aoo_q[s]; o.quality = b.quality; o.wx = b.wx; o.wy = b.wy; o.slot = s; o.map_id = b.map_id;
  a.out[t] = o;
}

struct ApplyArgs {
  const unsigned *keys;  // sorted
  const SortedAoo *aoo;  // sorted
  long long M;
  const MapSlot *maps;
  int stride, model;
  // optional trace for the pyramid: post-update impact of every applied slot (indexed by slot)
  double *trace_impact;
  double *trace_rec;  // per slot: the cell record right after this update (stride doubles)
  int trace_oie;
  struct LongRun *long_runs;  // runs handed to k_apply_long
  unsigned *n_long;
};

// One thread per cell run (the updates of one cell, in beam order).  Short runs are walked here; a run of
// SG_LONG_RUN updates or more -- the robot's own cell takes one update from EVERY beam, the cells around it dozens --
// is handed to k_apply_long, where a whole warp streams its operands.
#define SG_LONG_RUN 24
struct LongRun { long long head; int len; int pad; };

// the record of internal cell (ix, iy) of one map of the batch: dense array or copy-on-write tile
SG_DEV double *slot_cell_xy(const MapSlot &ms, int ix, int iy, int stride) {
  if (ms.tiles) {
    double *t = ms.tiles[(iy >> SG_TILE_BITS) * ms.tw + (ix >> SG_TILE_BITS)];
    return t + ((size_t)(iy & (SG_TILE - 1)) * SG_TILE + (ix & (SG_TILE - 1))) * stride;
  }
  return ms.cells + ((size_t)iy * ms.w + ix) * stride;
}
SG_DEV double *slot_cell_key(const MapSlot &ms, unsigned key, int stride) {  // key = key_base + iy * w + ix
  const unsigned local = key - ms.key_base;
  if (ms.tiles) {
    const int iy = (int)(local / (unsigned)ms.w), ix = (int)(local - (unsigned)iy * (unsigned)ms.w);
    return slot_cell_xy(ms, ix, iy, stride);
  }
  return ms.cells + (size_t)local * stride;
}

__global__ void __launch_bounds__(128) k_apply(ApplyArgs a) {
  long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= a.M) return;
  const unsigned key = a.keys[j];
  if (key == SG_INVALID_KEY) return;
  if (j > 0 && a.keys[j - 1] == key) return;  // not the head of this cell's run
  if (j + SG_LONG_RUN - 1 < a.M && a.keys[j + SG_LONG_RUN - 1] == key) {
    // long run: find its end (the keys are sorted) and queue it
    long long lo = j + SG_LONG_RUN - 1, hi = a.M;  // keys[lo] == key, keys[hi] > key (or hi == M)
    while (hi - lo > 1) {
      long long mid = (lo + hi) >> 1;
      if (a.keys[mid] == key) lo = mid; else hi = mid;
    }
    unsigned slot = atomicAdd(a.n_long, 1u);
    a.long_runs[slot] = LongRun{j, (int)(hi - j), 0};
    return;
  }
  double r[SLAMGPU_MAX_STRIDE];
  SortedAoo cur = a.aoo[j];
  const MapSlot &ms = a.maps[cur.map_id];
  double *cell = slot_cell_key(ms, key, a.stride);
  SG_COPY_REC(r, cell, a.stride);
  for (long long t = j;;) {
    const bool more = t + 1 < a.M && a.keys[t + 1] == key;
    SortedAoo nxt = cur;
    if (more) nxt = a.aoo[t + 1];  // next update's operands are in flight while this one is computed
    sg::cell_update(a.model, r, cur.p, cur.q, cur.wx, cur.wy, cur.quality);
    if (a.trace_impact) {
      a.trace_impact[cur.slot] = sg::cell_impact(a.model, a.trace_oie, r, 0.0, 0.0);
      double *tr = a.trace_rec + (size_t)cur.slot * a.stride;
      SG_COPY_REC(tr, r, a.stride);
    }
    if (!more) break;
    cur = nxt;
    ++t;
  }
  SG_COPY_REC(cell, r, a.stride);
}

// A chain of updates often repeats one observation (const estimator: every beam leaves the same "empty" estimate in
// the cells near the robot).  Once such an update maps a TBM record onto itself -- the belief has saturated, typically
// with the occupied and unknown masses stuck at the smallest subnormals -- every further update with the same operands
// is the identity, exactly: it is skipped after a three-double compare instead of being recomputed (~220 dependent
// instructions in that regime).  TBM cells only: the other models count their updates or use the obstacle position.
struct FixedPoint {
  double p, q, quality;
  bool have;
};
// The TBM update (tbm_grid_cells.h:12-19 over transferable_belief_model.h:63-143) for a WARP that carries one cell:
// every lane holds the belief and does the cheap part (16 products, the ordered sums), but the divisions -- four
// masses by their total in the conjunction, three by the non-conflict weight in the normalisation, each a ~20
// instruction dependent sequence behind a slow-path branch -- are spread over lanes 0..3 and handed back by shuffles:
// one division sequence where the scalar code runs four, then three.  Same operations, same operands, same order of
// every sum: bit-identical results.  The published fields (probability, quality) are functions of the masses and are
// derived when somebody needs them (tbm_publish), not once per update.
SG_DEV void tbm_update_warp(double *r, double p, double q, double quality, int lane) {
  using namespace sg;
  // aoo2tbm (the caller has already skipped invalid occupancies)
  const double est = mul(q, quality);
  const double mo = mul(p, est), me = mul(sub(1.0, p), est);
  const double mu = sub(sub(1.0, mo), me);
  const double bu = r[2], be = r[3], bo = r[4];
  const bool careful = is_tiny(bu) || is_tiny(be) || is_tiny(bo);
  // conjunction: t[i|j] += l[i]*r[j] over i, j in {u, e, o, c} in (i, j) order, with c = 0 on both sides.  The products
  // with a zero conflict mass are +0 (the masses are finite and non-negative) and adding +0 changes nothing, except
  // that it turns a -0 sum into +0 -- which a sum of non-negative products never is.  So they are left out.
  const double tu = add(0.0, mul(bu, mu));
  const double te = add(add(add(0.0, mul(bu, me)), mul(be, mu)), mul(be, me));
  const double to = add(add(add(0.0, mul(bu, mo)), mul(bo, mu)), mul(bo, mo));
  const double tc = add(add(0.0, mul(be, mo)), mul(bo, me));
  const double tot = add(add(add(tu, te), to), tc);
  const int k = lane & 3;
  double num = k == 0 ? tu : (k == 1 ? te : (k == 2 ? to : tc));
  // (x / 1 = x exactly: the masses of both sides sum to one up to rounding, so the total is exactly 1.0 for many updates,
  // and the test is uniform over the warp)
  double qv = num;
  if (tot != 1.0) qv = careful ? tbm_div<true>(num, tot, near_one(tot)) : __ddiv_rn(num, tot);
  double ou = __shfl_sync(0xffffffffu, qv, 0), oe = __shfl_sync(0xffffffffu, qv, 1), oo = __shfl_sync(0xffffffffu, qv, 2);
  if (tot == 0.0) { ou = 1.0; oe = 0.0; oo = 0.0; }
  const double w = add(add(ou, oe), oo);  // normalize_conflict
  num = k == 0 ? ou : (k == 1 ? oe : oo);
  qv = num;
  if (w != 1.0) qv = careful ? tbm_div<true>(num, w, near_one(w)) : __ddiv_rn(num, w);
  double nu = __shfl_sync(0xffffffffu, qv, 0), ne = __shfl_sync(0xffffffffu, qv, 1), no = __shfl_sync(0xffffffffu, qv, 2);
  if (w == 0.0) { nu = 1.0; ne = 0.0; no = 0.0; }
  r[2] = nu; r[3] = ne; r[4] = no;
}
// probability / quality / "known" flag of a TBM record from its masses, as the scalar update leaves them
SG_DEV void tbm_publish(int model, double *r) {
  using namespace sg;
  if (model == SLAMGPU_CELL_TBM_CONSISTENT) {
    const double qual = add(r[4], r[3]);
    r[0] = (is_tiny(r[4]) || r[4] == 0.0) ? tbm_div<true>(r[4], qual, near_one(qual)) : div(r[4], qual);
    r[1] = qual;
  } else {
    r[0] = add(r[4], mul(0.5, r[2])); r[1] = 1.0;
  }
  r[5] = 1;
}

// One TBM update of a warp-carried chain.  *dirty: the record's published fields are behind its masses.
SG_DEV void chain_update_tbm(double *r, double p, double q, double quality, FixedPoint &fx, int lane, bool *dirty) {
  const bool repeated = __double_as_longlong(p) == __double_as_longlong(fx.p) && __double_as_longlong(q) == __double_as_longlong(fx.q) &&
                        __double_as_longlong(quality) == __double_as_longlong(fx.quality);
  if (repeated && fx.have) return;  // same observation on a saturated belief: the identity
  // the belief masses (record fields 2..4) are the whole state of a TBM cell: the other fields are functions of them
  const long long u0 = __double_as_longlong(r[2]), e0 = __double_as_longlong(r[3]), o0 = __double_as_longlong(r[4]);
  if (!(isnan(p) || isnan(q))) {  // an invalid occupancy is skipped (uniform over the warp)
    tbm_update_warp(r, p, q, quality, lane);
    *dirty = true;
  }
  fx.have = repeated && u0 == __double_as_longlong(r[2]) && e0 == __double_as_longlong(r[3]) && o0 == __double_as_longlong(r[4]);
  fx.p = p; fx.q = q; fx.quality = quality;
}

// A warp per long run: the lanes fetch 32 consecutive operand records at once (coalesced, the next 32 already in
// flight), then the updates are applied in order with the operands handed round by shuffles.  The chain of dependent
// cell updates stays sequential (it is the reference's arithmetic), but it no longer waits on a memory round trip per
// update.  Every lane carries the record, so no lane idles on a broadcast; lane l writes the trace of update l.
template <bool TBM, bool TRACE>
__global__ void __launch_bounds__(128) k_apply_long(ApplyArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned n = *a.n_long;
  for (unsigned e = warp; e < n; e += nwarps) {
    const LongRun run = a.long_runs[e];
    const unsigned key = a.keys[run.head];
    const SortedAoo *src = a.aoo + run.head;
    SortedAoo mine = src[min(lane, run.len - 1)];
    const MapSlot &ms = a.maps[mine.map_id];
    double *cell = slot_cell_key(ms, key, a.stride);
    double r[SLAMGPU_MAX_STRIDE];
    SG_COPY_REC(r, cell, a.stride);
    FixedPoint fx{NAN, NAN, NAN, false};
    bool dirty = false;
    for (int base = 0; base < run.len; base += 32) {
      const int cnt = min(32, run.len - base);
      SortedAoo nxt = mine;
      if (base + 32 < run.len) nxt = src[min(base + 32 + lane, run.len - 1)];
      for (int l = 0; l < cnt; ++l) {
        const double p = __shfl_sync(0xffffffffu, mine.p, l), q = __shfl_sync(0xffffffffu, mine.q, l);
        const double quality = __shfl_sync(0xffffffffu, mine.quality, l);
        if (TBM) {  // (the obstacle position is not part of a TBM update)
          chain_update_tbm(r, p, q, quality, fx, lane, &dirty);
        } else {
          const double wx = __shfl_sync(0xffffffffu, mine.wx, l), wy = __shfl_sync(0xffffffffu, mine.wy, l);
          sg::cell_update(a.model, r, p, q, wx, wy, quality);
        }
        if (TRACE && dirty) { tbm_publish(a.model, r); dirty = false; }  // a pyramid reads every intermediate record
        if (TRACE && lane == l) {
          a.trace_impact[mine.slot] = sg::cell_impact(a.model, a.trace_oie, r, 0.0, 0.0);
          double *tr = a.trace_rec + (size_t)mine.slot * a.stride;
          SG_COPY_REC(tr, r, a.stride);
        }
      }
      mine = nxt;
    }
    if (dirty) tbm_publish(a.model, r);
    if (lane == 0)
      SG_COPY_REC(cell, r, a.stride);
  }
}

// The robot's own cell is the first cell of EVERY ray, so its run is as long as the scan and its chain of dependent
// updates is the critical path of an insertion (1081 TBM updates ~ 250 us).  It does not need the sort: its updates
// are one per beam, in beam order.  One warp per map applies them straight from the per-slot estimates, on a side
// stream, while the main stream sorts and applies every other cell.
struct RobotArgs {
  const MapSlot *maps;
  int n_maps;
  const BeamRec *beams;
  const BeamOut *bout;         // per beam: cells of its ray
  const long long *offsets;    // per beam: first slot
  const int2 *cells;           // per slot: external cell
  const double *aoo_p, *aoo_q; // per slot: the occupancy estimate
  int stride, model;
  int ring;                    // half size W of the window of cells around the robot handled here (0: the robot cell only)
  double *trace_impact, *trace_rec;  // pyramid: per slot, impact and record right after the update (NULL: none)
  int trace_oie;
};

// One warp per cell of the (2W+1)^2 window around the robot of each map.  A ray is monotone in x and y, so it can only
// be inside the window during its first 2W+1 cells, and it visits a cell at most once: the updates of a window cell
// are one per beam at most, in beam order -- no sort needed.  The lanes test 32 beams at a time for "does this ray
// pass through my cell", then the hits are applied in beam order (the chain of dependent updates), W = 0 being the
// robot's own cell, which every ray starts in.
// MODEL: the cell model as a compile-time constant (the update's switch and the operands it does not read drop out of the
// chain) or -1 for the run-time a.model
template <bool TBM, int MODEL>
__global__ void __launch_bounds__(128) k_apply_ring(RobotArgs a) {
  const int model = MODEL >= 0 ? MODEL : a.model;
  constexpr bool WANT_XY = MODEL < 0 || MODEL == SLAMGPU_CELL_GMAPPING;  // only the GMapping cell reads the obstacle point
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int side = 2 * a.ring + 1, per_map = side * side;
  const int m = wid / per_map;
  if (m >= a.n_maps) return;
  const int t = wid - m * per_map;
  const MapSlot ms = a.maps[m];
  const int tx = ms.rx + (t % side) - a.ring, ty = ms.ry + (t / side) - a.ring;  // the cell of this warp
  const int ix = tx + ms.ox, iy = ty + ms.oy;
  if (ix < 0 || ix >= ms.w || iy < 0 || iy >= ms.h) return;
  double *cell = slot_cell_xy(ms, ix, iy, a.stride);
  double r[SLAMGPU_MAX_STRIDE];
  SG_COPY_REC(r, cell, a.stride);
  bool any = false, dirty = false;
  FixedPoint fx{NAN, NAN, NAN, false};
  for (int base = ms.beam_begin; base < ms.beam_end; base += 32) {
    const int i = base + lane;
    long long rs = -1;
    double p = 0, q = 0, quality = 0, wx = 0, wy = 0;
    if (i < ms.beam_end) {
      const int cnt = min(a.bout[i].count, side);  // cells of this ray that can lie in the window
      const long long off = a.offsets[i];
      if (tx == ms.rx && ty == ms.ry) {  // the robot's own cell (the longest chain): every ray starts there
        if (cnt > 0) { const int2 c = a.cells[off]; if (c.x == tx && c.y == ty) rs = off; }
      } else {
        // a ray visits a cell at most once: all candidate slots are loaded at once (independent loads), no early exit
#pragma unroll
        for (int k = 0; k < 13; ++k) {
          if (k < cnt) {
            const int2 c = a.cells[off + k];
            if (c.x == tx && c.y == ty) rs = off + k;
          }
        }
      }
      if (rs >= 0) {
        const BeamRec &b = a.beams[i];
        p = a.aoo_p[rs]; q = a.aoo_q[rs]; quality = b.quality; wx = b.wx; wy = b.wy;
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, rs >= 0);
    any |= todo != 0;
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      if (TBM)
        chain_update_tbm(r, __shfl_sync(0xffffffffu, p, l), __shfl_sync(0xffffffffu, q, l), __shfl_sync(0xffffffffu, quality, l), fx, lane, &dirty);
      else
        sg::cell_update(model, r, __shfl_sync(0xffffffffu, p, l), __shfl_sync(0xffffffffu, q, l), WANT_XY ? __shfl_sync(0xffffffffu, wx, l) : 0.0,
                        WANT_XY ? __shfl_sync(0xffffffffu, wy, l) : 0.0, __shfl_sync(0xffffffffu, quality, l));
      if (!TBM && a.trace_impact) {  // what the pyramid folds upwards: the cell right after this update
        const long long slot = __shfl_sync(0xffffffffu, rs, l);
        if (lane == 0) {
          a.trace_impact[slot] = sg::cell_impact(model, a.trace_oie, r, 0.0, 0.0);
          double *tr = a.trace_rec + (size_t)slot * a.stride;
          SG_COPY_REC(tr, r, a.stride);
        }
      }
    }
  }
  if (dirty) tbm_publish(model, r);
  if (any && lane == 0)
    SG_COPY_REC(cell, r, a.stride);
}

// ---------------------------------------------------------------- map growth (device side)
__global__ void k_copy_block(const double *__restrict__ src, int sw, int sh, double *__restrict__ dst, int dw, int offx,
                             int offy, int stride) {
  size_t n = (size_t)sw * sh * stride;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    size_t cell = i / stride;
    int k = (int)(i - cell * stride);
    int y = (int)(cell / sw), x = (int)(cell - (size_t)y * sw);
    dst[(((size_t)(y + offy)) * dw + (x + offx)) * stride + k] = src[i];
  }
}
struct RecParam { double v[SLAMGPU_MAX_STRIDE]; };
__global__ void k_fill(double *cells, size_t n_cells, int stride, RecParam rec) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  size_t n = n_cells * stride;
  for (; i < n; i += st) cells[i] = rec.v[i % stride];
}

__global__ void k_reset_cell(double *cell, int stride, RecParam rec) {
  for (int k = 0; k < stride; ++k) cell[k] = rec.v[k];
}
__global__ void k_update_cell(double *cell, int stride, int model, double p, double q, double obx, double oby, double quality) {
  double r[SLAMGPU_MAX_STRIDE];
  for (int k = 0; k < stride; ++k) r[k] = cell[k];
  sg::cell_update(model, r, p, q, obx, oby, quality);
  for (int k = 0; k < stride; ++k) cell[k] = r[k];
}

int host_world_to_cell(double v, double scale) { return (int)std::floor(v / scale); }

}  // namespace

// ---------------------------------------------------------------- growth rules (host logic)
// UnboundedPlainGridMap::ensure_inside (src/core/maps/plain_grid_map.h:133-173, with its
// unsigned/double conversions) and UnboundedLazyTiledGridMap::ensure_inside
// (src/core/maps/lazy_tiled_grid_map.h:151-181)
bool GrowState::ensure_inside(int x, int y) {
  int cx = x + ox, cy = y + oy;
  if (0 <= cx && cx < w && 0 <= cy && cy < h) return false;
  if (grow == SLAMGPU_GROW_PLAIN) {
    unsigned uw = w, uh = h, prep_x = 0, app_x = 0, prep_y = 0, app_y = 0;
    if (cx < 0) prep_x = 0 - cx; else if ((int)uw <= cx) app_x = cx - uw + 1;
    if (cy < 0) prep_y = 0 - cy; else if ((int)uh <= cy) app_y = cy - uh + 1;
    unsigned new_w = prep_x + uw + app_x, new_h = prep_y + uh + app_y;
    const double rate = 1.2;
    if (uw < new_w && new_w < rate * uw) {
      double sc = prep_x / (new_w - uw);  // unsigned division, as upstream
      prep_x += (rate * uw - new_w) * sc;
      new_w = rate * uw;
    }
    if (uh < new_h && new_h < rate * uh) {
      double sc = prep_y / (new_h - uh);
      prep_y += (rate * uh - new_h) * sc;
      new_h = rate * uh;
    }
    w = (int)new_w; h = (int)new_h; ox += (int)prep_x; oy += (int)prep_y;
    return true;
  }
  if (grow == SLAMGPU_GROW_TILED) {
    const unsigned bits = 7, tile = 1u << bits;
    unsigned tx = (w + tile - 1) / tile, ty = (h + tile - 1) / tile;
    unsigned prep_x = 0, app_x = 0, prep_y = 0, app_y = 0;
    if (cx < 0) prep_x = 1 + ((0 - cx) >> bits); else if (w <= cx) app_x = 1 + ((cx - w) >> bits);
    if (cy < 0) prep_y = 1 + ((0 - cy) >> bits); else if (h <= cy) app_y = 1 + ((cy - h) >> bits);
    unsigned ntx = prep_x + tx + app_x, nty = prep_y + ty + app_y;
    w = (int)(ntx * tile); h = (int)(nty * tile); ox += (int)(prep_x * tile); oy += (int)(prep_y * tile);
    return true;
  }
  return false;
}

// move the device map into a (larger) array laid out for `g`
int sg_map_regrow(slamgpu_map *m, const GrowState &g) {
  slamgpu_ctx *ctx = m->ctx;
  if (g.w == m->w && g.h == m->h && g.ox == m->ox && g.oy == m->oy) return SLAMGPU_OK;
  const int offx = g.ox - m->ox, offy = g.oy - m->oy;
  if (offx < 0 || offy < 0 || offx + m->w > g.w || offy + m->h > g.h) return sg_fail(ctx, SLAMGPU_E_STATE, "map growth would shrink the map");
  if (m->pool) {
    // UnboundedLazyTiledGridMap grows by whole tiles (lazy_tiled_grid_map.h:151-181): the tile table is re-laid, no cell moves
    if ((offx | offy) & (SG_TILE - 1)) return sg_fail(ctx, SLAMGPU_E_STATE, "a tiled map grows by whole tiles");
    const int ntw = (g.w + SG_TILE - 1) >> SG_TILE_BITS, nth = (g.h + SG_TILE - 1) >> SG_TILE_BITS;
    std::vector<int32_t> ids((size_t)ntw * nth, 0);
    std::vector<double *> ptrs((size_t)ntw * nth, m->pool->ptr(0));
    const int tx0 = offx >> SG_TILE_BITS, ty0 = offy >> SG_TILE_BITS;
    for (int ty = 0; ty < m->th; ++ty)
      for (int tx = 0; tx < m->tw; ++tx) {
        ids[(size_t)(ty + ty0) * ntw + tx + tx0] = m->tile_ids[(size_t)ty * m->tw + tx];
        ptrs[(size_t)(ty + ty0) * ntw + tx + tx0] = m->h_tile_ptrs[(size_t)ty * m->tw + tx];
      }
    m->tile_ids.swap(ids); m->h_tile_ptrs.swap(ptrs); m->tw = ntw; m->th = nth; m->tiles_dirty = true;
    m->w = g.w; m->h = g.h; m->ox = g.ox; m->oy = g.oy;
    m->pitch = (m->w + 2 * SG_LUT_PAD + 1) & ~1;
    sg_map_invalidate_lut(m);
    return SLAMGPU_OK;
  }
  size_t need = (size_t)g.w * g.h * m->stride;
  double *nc = nullptr;
  SG_CUDA(ctx, cudaMalloc(&nc, std::max<size_t>(need, 1) * sizeof(double)));
  RecParam rp;
  memcpy(rp.v, m->unknown, sizeof rp.v);
  k_fill<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(nc, (size_t)g.w * g.h, m->stride, rp);
  SG_LAUNCHED(ctx);
  if ((size_t)m->w * m->h > 0) {
    k_copy_block<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(m->d_cells, m->w, m->h, nc, g.w, offx, offy, m->stride);
    SG_LAUNCHED(ctx);
  }
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (m->d_cells) cudaFree(m->d_cells);
  m->d_cells = nc; m->cells_cap = need;
  m->w = g.w; m->h = g.h; m->ox = g.ox; m->oy = g.oy;
  m->pitch = (m->w + 2 * SG_LUT_PAD + 1) & ~1;
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}

// device address of one internal cell for the single-cell entry points; a tiled map first makes the cell's tile private
static int writable_cell_ptr(slamgpu_map *m, int ix, int iy, double **out) {
  if (!m->pool) { *out = m->d_cells + ((size_t)iy * m->w + ix) * m->stride; return SLAMGPU_OK; }
  std::vector<int32_t> copies;
  SG_TRY(sg_map_make_writable(m, ix, iy, ix, iy, &copies));
  SG_TRY(sg_pool_run_copies(m->pool, copies));
  *out = m->pool->ptr(m->tile_ids[(size_t)(iy >> SG_TILE_BITS) * m->tw + (ix >> SG_TILE_BITS)]) +
         ((size_t)(iy & (SG_TILE - 1)) * SG_TILE + (ix & (SG_TILE - 1))) * m->stride;
  return SLAMGPU_OK;
}

// ---------------------------------------------------------------- single-cell plumbing
extern "C" int slamgpu_map_reset_cell(slamgpu_map *m, int32_t x, int32_t y, const double *rec) {
  if (!m || !rec) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  GrowState g{m->w, m->h, m->ox, m->oy, m->grow};
  if (g.ensure_inside(x, y)) SG_TRY(sg_map_regrow(m, g));
  int ix = x + m->ox, iy = y + m->oy;
  if (ix < 0 || ix >= m->w || iy < 0 || iy >= m->h) return sg_fail(ctx, SLAMGPU_E_INVALID, "cell (%d, %d) is outside the bounded map", x, y);
  RecParam rp;
  memset(rp.v, 0, sizeof rp.v);
  memcpy(rp.v, rec, sizeof(double) * m->stride);
  double *cell = nullptr;
  SG_TRY(writable_cell_ptr(m, ix, iy, &cell));
  k_reset_cell<<<1, 1, 0, ctx->stream>>>(cell, m->stride, rp);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_update_cell(slamgpu_map *m, int32_t x, int32_t y, int32_t aoo_is_occ, double aoo_p, double aoo_q,
                                       double obst_x, double obst_y, double quality) {
  (void)aoo_is_occ;
  if (!m) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  GrowState g{m->w, m->h, m->ox, m->oy, m->grow};
  if (g.ensure_inside(x, y)) SG_TRY(sg_map_regrow(m, g));
  int ix = x + m->ox, iy = y + m->oy;
  if (ix < 0 || ix >= m->w || iy < 0 || iy >= m->h) return sg_fail(ctx, SLAMGPU_E_INVALID, "cell (%d, %d) is outside the bounded map", x, y);
  double *cell = nullptr;
  SG_TRY(writable_cell_ptr(m, ix, iy, &cell));
  k_update_cell<<<1, 1, 0, ctx->stream>>>(cell, m->stride, m->model, aoo_p, aoo_q,
                                         obst_x, obst_y, quality);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}

// ---------------------------------------------------------------- beam preparation (host, libm)
// GridMapScanAdder::append_scan :54-75 + the per-beam part of handle_scan_point :138-150, 176-189
int sg_prepare_beams(const slamgpu_map *m, const slamgpu_scan *s, const double pose[3], double scan_quality, int scan_margin,
                     double blur, double max_range, const double *point_quality, bool gate, BeamPlan *plan) {
  const int N = s->n;
  const double px = pose[0], py = pose[1], pth = pose[2], scale = m->scale;
  plan->beams.assign(N, BeamRec{});
  plan->offsets.assign(N + 1, 0);
  plan->px = px; plan->py = py;
  plan->rx = host_world_to_cell(px, scale); plan->ry = host_world_to_cell(py, scale);
  const double max_range_sq = std::pow(max_range, 2);
  long long total = 0;
  const long long first = scan_margin, last = (long long)N - scan_margin - 1;  // grid_map_scan_adders.h:65-66
  for (int i = 0; i < N; ++i) {
    BeamRec &b = plan->beams[i];
    plan->offsets[i] = total;
    bool in_margin = i >= first && i <= last;
    if (gate && !in_margin) continue;
    const double r = s->range[i], a = s->angle[i];
    b.wx = px + r * std::cos(pth + a);
    b.wy = py + r * std::sin(pth + a);
    b.quality = scan_quality * (point_quality ? point_quality[i] : 1.0);
    b.is_occ = s->occ[i] ? 1 : 0;
    const double len_sq = std::pow(b.wx - px, 2) + std::pow(b.wy - py, 2);
    if (gate && max_range_sq < len_sq) continue;
    if (!std::isfinite(b.wx) || !std::isfinite(b.wy)) continue;
    b.obx = host_world_to_cell(b.wx, scale); b.oby = host_world_to_cell(b.wy, scale);
    b.obst_sq = std::pow(plan->rx - b.obx, 2) + std::pow(plan->ry - b.oby, 2);
    double blur_dist = 0;
    if (b.is_occ) {
      blur_dist = blur / scale;
      if (blur_dist < 0) blur_dist *= -len_sq;
    }
    b.hole_sq = std::pow(blur_dist, 2);
    b.active = 1;
    // |dx| + |dy| + 1 cells at most, plus the one extra cell the walk emits before it detects an
    // overshoot and fails over to Bresenham (regular_squares_grid.h:84-86)
    long long ub = std::llabs((long long)b.obx - plan->rx) + std::llabs((long long)b.oby - plan->ry) + 2;
    if (ub > (1ll << 26)) return SLAMGPU_E_INVALID;
    total += ub;
  }
  plan->offsets[N] = total;
  plan->M = total;
  return SLAMGPU_OK;
}

namespace {

int upload_async(slamgpu_ctx *ctx, DevBuf &dst, const void *src, size_t bytes) {
  if (dst.reserve(std::max<size_t>(bytes, 16)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "device buffer of %zu bytes", bytes);
  if (bytes) SG_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SLAMGPU_OK;
}

size_t map_counters_offset(size_t n) { return (n * sizeof(MapSlot) + 63) & ~(size_t)63; }

double est_shift(const slamgpu_estimator &e, double scale, const BeamPlan &plan) {
  // Shift_Amount is a function-static of the reference, fixed by the first cell it ever estimates
  // (quirk Q10): the obstacle cell of the first inserted beam
  if (e.shift_amount >= 0) return e.shift_amount;
  int ly = plan.ry;
  for (const BeamRec &b : plan.beams)
    if (b.active) { ly = b.oby; break; }
  return e.low_qual * (scale * (ly + 1) - scale * ly);
}

// K2 for prepared beams (one or several maps): fills scratch[0] (beams), [1] (offsets), [2] (cells), [3] (BeamOut),
// scratch[7] (MapSlot array, px/py valid; dims are refreshed after the growth check)
int run_raycast_multi(slamgpu_ctx *ctx, double scale, const BeamRec *beams, int N, const long long *offsets,
                      long long M, const std::vector<MapSlot> &slots, const slamgpu_estimator &est, bool beams_pinned = false) {
  // scratch[7]: MapSlot[n] | counters[2n] (zeroed by the same copy)
  const size_t coff = map_counters_offset(slots.size());
  const size_t slot_bytes = coff + 16 * slots.size();
  const size_t bb = (sizeof(BeamRec) * (size_t)N + 63) & ~(size_t)63, ob = (sizeof(long long) * ((size_t)N + 1) + 63) & ~(size_t)63;
  if (ctx->scratch[0].reserve(std::max<size_t>(sizeof(BeamRec) * N, 16)) != SLAMGPU_OK ||
      ctx->scratch[1].reserve(sizeof(long long) * ((size_t)N + 1)) != SLAMGPU_OK || ctx->scratch[7].reserve(slot_bytes) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "ray-cast inputs");
  // everything goes up from pinned memory (a copy from pageable memory is staged by the driver, ~8 us apiece): the slot
  // block always, beams and offsets unless the caller already assembled them in the ctx's pinned block
  void *hp;
  if (beams_pinned) {
    if (ctx->h_slots_cap < slot_bytes) {
      if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
      ctx->h_slots = nullptr; ctx->h_slots_cap = 0;
      SG_CUDA(ctx, cudaMallocHost(&ctx->h_slots, slot_bytes * 2));
      ctx->h_slots_cap = slot_bytes * 2;
    }
    if (ctx->staged_pending) { SG_CUDA(ctx, cudaEventSynchronize(ctx->ev_staged)); ctx->staged_pending = false; }  // (the slot block of the batch before)
    hp = ctx->h_slots;
  } else {
    SG_TRY(sg_pinned(ctx, bb + ob + slot_bytes, &hp));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging may still be in flight
    if (N) memcpy(hp, beams, sizeof(BeamRec) * N);
    memcpy((char *)hp + bb, offsets, sizeof(long long) * ((size_t)N + 1));
    beams = (const BeamRec *)hp; offsets = (const long long *)((char *)hp + bb);
    hp = (char *)hp + bb + ob;
  }
  memset(hp, 0, slot_bytes);
  memcpy(hp, slots.data(), sizeof(MapSlot) * slots.size());
  if (N) SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0].p, beams, sizeof(BeamRec) * N, cudaMemcpyHostToDevice, ctx->stream));
  SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[1].p, offsets, sizeof(long long) * ((size_t)N + 1), cudaMemcpyHostToDevice, ctx->stream));
  SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[7].p, hp, slot_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (beams_pinned) {
    if (!ctx->ev_staged) SG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_staged, cudaEventDisableTiming));
    SG_CUDA(ctx, cudaEventRecord(ctx->ev_staged, ctx->stream));
    ctx->staged_pending = true;
  }
  if (ctx->scratch[2].reserve(std::max<size_t>(M, 1) * sizeof(int2)) != SLAMGPU_OK ||
      ctx->scratch[3].reserve(std::max<size_t>(N, 1) * sizeof(BeamOut)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "ray-cast buffers (%lld slots)", M);
  if (N == 0) return SLAMGPU_OK;
  RaycastArgs a;
  a.beams = ctx->scratch[0].as<BeamRec>(); a.offsets = ctx->scratch[1].as<long long>(); a.N = N;
  a.maps = ctx->scratch[7].as<MapSlot>(); a.scale = scale; a.est = est;
  a.cells = ctx->scratch[2].as<int2>(); a.out = ctx->scratch[3].as<BeamOut>();
  a.counters = (unsigned long long *)((char *)ctx->scratch[7].p + coff);
  cudaEventRecord(ctx->evk0, ctx->stream);
  k_raycast<<<(N + 31) / 32, 32, 0, ctx->stream>>>(a);
  cudaEventRecord(ctx->evk1, ctx->stream);
  ctx->evk_valid = true;
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  return SLAMGPU_OK;
}

MapSlot slot_of(const slamgpu_map *m, double px, double py, double shift, unsigned key_base) {
  MapSlot s;
  s.cells = m->d_cells; s.px = px; s.py = py; s.shift = shift; s.w = m->w; s.h = m->h; s.ox = m->ox; s.oy = m->oy; s.key_base = key_base;
  s.tiles = m->pool ? m->d_tile_ptrs : nullptr; s.tw = m->tw;
  s.rx = host_world_to_cell(px, m->scale); s.ry = host_world_to_cell(py, m->scale);
  s.beam_begin = s.beam_end = 0;
  return s;
}

int run_raycast(slamgpu_ctx *ctx, const slamgpu_map *m, const BeamPlan &plan, const slamgpu_estimator &est) {
  std::vector<MapSlot> slots{slot_of(m, plan.px, plan.py, est_shift(est, m->scale, plan), 0)};
  return run_raycast_multi(ctx, m->scale, plan.beams.data(), (int)plan.beams.size(), plan.offsets.data(), plan.M, slots, est);
}

}  // namespace

int sg_radix_sort(slamgpu_ctx *ctx, unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, long long n,
               unsigned max_key, unsigned **keys_sorted, unsigned **vals_sorted) {
  int bits = 1;
  while (bits < 32 && (max_key >> bits) != 0) ++bits;
  const int passes = (bits + 7) / 8;
  const int nb = (int)((n + SG_SORT_TILE - 1) / SG_SORT_TILE);
  const int nh = 256 * std::max(nb, 1);
  const int nchunks = (nh + SG_SCAN_CHUNK - 1) / SG_SCAN_CHUNK;
  if (ctx->scratch[6].reserve(((size_t)nh + nchunks) * sizeof(unsigned)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "sort histogram");
  unsigned *ghist = ctx->scratch[6].as<unsigned>();
  unsigned *totals = ghist + nh;
  unsigned *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
  for (int p = 0; p < passes && nb > 0; ++p) {
    k_radix_hist<<<nb, SG_SORT_THREADS, 0, ctx->stream>>>(ki, n, 8 * p, ghist, nb);
    if (nchunks <= 4) {
      k_radix_scan<<<1, 1024, 0, ctx->stream>>>(ghist, 256 * nb);
    } else {
      k_scan_chunks<<<nchunks, 1024, 0, ctx->stream>>>(ghist, 256 * nb, totals);
      k_radix_scan<<<1, 1024, 0, ctx->stream>>>(totals, nchunks);
      k_scan_add<<<nchunks, 1024, 0, ctx->stream>>>(ghist, 256 * nb, totals);
      ctx->launches += 2;
    }
    k_radix_scatter<<<nb, SG_SORT_THREADS, 0, ctx->stream>>>(ki, vi, ko, vo, n, 8 * p, ghist, nb);
    ctx->launches += 3;
    std::swap(ki, ko); std::swap(vi, vo);
  }
  SG_CUDA(ctx, cudaGetLastError());
  *keys_sorted = ki; *vals_sorted = vi;
  return SLAMGPU_OK;
}


// the whole insertion; `trace` (optional) receives what the pyramid needs
int sg_append_scan_impl(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3], double scan_quality,
                        int32_t scan_margin, const slamgpu_estimator *est, double blur, double max_range,
                        const double *point_quality, int64_t *cells_updated, AppendTrace *trace) {
  if (!ctx || !map || !scan || !pose || !est) return sg_fail(ctx, SLAMGPU_E_INVALID, "append_scan: NULL argument");
  if (map->ctx != ctx || scan->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map/scan belongs to another ctx");
  if (scan_margin < 0) return sg_fail(ctx, SLAMGPU_E_INVALID, "negative scan margin");
  if (cells_updated) *cells_updated = 0;
  if (trace) { trace->M = 0; trace->applied = 0; }
  if (scan->n == 0) return SLAMGPU_OK;  // grid_map_scan_adders.h:59
  BeamPlan plan;
  if (sg_prepare_beams(map, scan, pose, scan_quality, scan_margin, blur, max_range, point_quality, true, &plan) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "a beam spans more than 2^26 cells");
  return sg_append_plan(ctx, map, plan, est, cells_updated, trace);
}

// K2 + K3 for prepared beams (all sharing the begin point plan.px, plan.py)
int sg_append_plan(slamgpu_ctx *ctx, slamgpu_map *map, const BeamPlan &plan, const slamgpu_estimator *est,
                   int64_t *cells_updated, AppendTrace *trace) {
  return sg_append_plans(ctx, &map, &plan, 1, est, cells_updated, trace);
}

// K2 + K3 for n maps at once (GMapping particles: every particle inserts the scan into its own map from its own
// pose): one ray-cast launch, one estimate launch, ONE sort over (map, cell) keys and one apply launch for all of
// them.  Maps must share cell model, cell size and ctx.  `trace` is only available for n == 1.
int sg_append_plans(slamgpu_ctx *ctx, slamgpu_map *const *maps, const BeamPlan *plans, int n, const slamgpu_estimator *est,
                    int64_t *cells_updated, AppendTrace *trace, unsigned long long *deferred) {
  if (!ctx || !maps || !plans || n <= 0 || !est) return sg_fail(ctx, SLAMGPU_E_INVALID, "append: NULL argument");
  if (est->type != SLAMGPU_EST_CONST && est->type != SLAMGPU_EST_AREA) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad estimator type");
  if (trace && n != 1) return sg_fail(ctx, SLAMGPU_E_INVALID, "append: a trace needs a single map");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (cells_updated) for (int k = 0; k < n; ++k) cells_updated[k] = 0;
  if (trace) { trace->M = 0; trace->applied = 0; }
  for (int k = 0; k < n; ++k) {
    if (!maps[k] || maps[k]->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map %d is NULL or belongs to another ctx", k);
    if (maps[k]->model != maps[0]->model || maps[k]->scale != maps[0]->scale) return sg_fail(ctx, SLAMGPU_E_INVALID, "batched maps must share cell model and scale");
  }
  // ---- concatenate the plans
  std::vector<long long> beam0(n + 1, 0), slot0(n + 1, 0);
  for (int k = 0; k < n; ++k) { beam0[k + 1] = beam0[k] + (long long)plans[k].beams.size(); slot0[k + 1] = slot0[k] + plans[k].M; }
  const long long M = slot0[n];
  const int N = (int)beam0[n];
  if (M == 0) return SLAMGPU_OK;
  if (M >= (1ll << 31) || beam0[n] >= (1ll << 31)) return sg_fail(ctx, SLAMGPU_E_NOMEM, "scan insertion needs %lld cell slots", M);
  const BeamRec *beams_ptr = plans[0].beams.data();
  const long long *offs_ptr = plans[0].offsets.data();
  if (n > 1) {
    // concatenated straight into pinned memory: 72 B per beam (13 MB for 256 particles x 720 beams) go up at DMA speed
    const size_t bb = (sizeof(BeamRec) * (size_t)N + 63) & ~(size_t)63;
    void *hp;
    SG_TRY(sg_pinned(ctx, bb + sizeof(long long) * ((size_t)N + 1), &hp));
    // the pinned staging block may still be in flight: behind a deferred batch only its uploads have to be over (their
    // event), not its kernels
    if (ctx->staged_pending) { SG_CUDA(ctx, cudaEventSynchronize(ctx->ev_staged)); ctx->staged_pending = false; }
    else SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BeamRec *all_beams = (BeamRec *)hp;
    long long *all_offs = (long long *)((char *)hp + bb);
    for (int k = 0; k < n; ++k) {
      BeamRec *dst = all_beams + beam0[k];
      long long *od = all_offs + beam0[k];
      const size_t nk = plans[k].beams.size();
      if (nk) memcpy(dst, plans[k].beams.data(), sizeof(BeamRec) * nk);
      for (size_t i = 0; i < nk; ++i) { dst[i].map_id = k; od[i] = slot0[k] + plans[k].offsets[i]; }
    }
    all_offs[N] = M;
    beams_ptr = all_beams; offs_ptr = all_offs;
  }
  std::vector<MapSlot> slots(n);
  // Shift_Amount per map, exactly as one call per map would fix it; key bases and beam ranges as far as they are known
  // (growth and copy-on-write below may still change a map: then the slots go up a second time)
  {
    unsigned long long kb = 0;
    for (int k = 0; k < n; ++k) {
      slots[k] = slot_of(maps[k], plans[k].px, plans[k].py, est_shift(*est, maps[0]->scale, plans[k]), (unsigned)std::min<unsigned long long>(kb, 0xFFFFFFFFull));
      slots[k].beam_begin = (int)beam0[k]; slots[k].beam_end = (int)beam0[k + 1];
      kb += (unsigned long long)maps[k]->w * maps[k]->h;
    }
  }
  const std::vector<MapSlot> slots_sent = slots;
  SG_TRY(run_raycast_multi(ctx, maps[0]->scale, beams_ptr, N, offs_ptr, M, slots, *est, n > 1));

  // ---- map growth (Q9): only when some beam leaves a map's current bounds; replays the reference's
  // ensure_inside sequence over that map's cells in update order
  for (int k = 0; k < n; ++k) {
    slamgpu_map *map = maps[k];
    const BeamPlan &plan = plans[k];
    if (map->grow == SLAMGPU_GROW_NONE || plan.M == 0) continue;
    auto outside = [&](int x, int y) { int ix = x + map->ox, iy = y + map->oy; return ix < 0 || ix >= map->w || iy < 0 || iy >= map->h; };
    bool need = false;
    const int Nk = (int)plan.beams.size();
    for (int i = 0; i < Nk && !need; ++i)
      if (plan.beams[i].active && (outside(plan.rx, plan.ry) || outside(plan.beams[i].obx, plan.beams[i].oby))) need = true;
    if (!need) continue;
    std::vector<int2> cells((size_t)plan.M);
    std::vector<BeamOut> bout(Nk);
    SG_CUDA(ctx, cudaMemcpyAsync(cells.data(), ctx->scratch[2].as<int2>() + slot0[k], sizeof(int2) * plan.M, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(bout.data(), ctx->scratch[3].as<BeamOut>() + beam0[k], sizeof(BeamOut) * Nk, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GrowState g{map->w, map->h, map->ox, map->oy, map->grow};
    for (int i = 0; i < Nk; ++i) {
      if (!plan.beams[i].active) continue;
      const int2 *c = cells.data() + plan.offsets[i];
      const int cnt = bout[i].count;
      g.ensure_inside(c[cnt - 1].x, c[cnt - 1].y);  // the obstacle cell is updated first
      for (int q = 0; q < cnt - 1; ++q) g.ensure_inside(c[q].x, c[q].y);
    }
    SG_TRY(sg_map_regrow(map, g));
  }
  // ---- copy-on-write maps: the tiles under the bounding box of each map's beams (robot cell .. end cells, which holds every
  // ray-cast cell and the ring around the robot) become private to the map before anything is written
  {
    std::vector<int32_t> copies;
    SgTilePool *pool = nullptr;
    for (int k = 0; k < n; ++k) {
      slamgpu_map *map = maps[k];
      if (!map->pool || plans[k].M == 0) continue;
      pool = map->pool;
      const BeamPlan &plan = plans[k];
      int x0 = plan.rx - 7, x1 = plan.rx + 7, y0 = plan.ry - 7, y1 = plan.ry + 7;  // k_apply_ring's window
      bool any = false;
      for (const BeamRec &b : plan.beams) {
        if (!b.active) continue;
        any = true;
        x0 = std::min(x0, b.obx - 1); x1 = std::max(x1, b.obx + 1); y0 = std::min(y0, b.oby - 1); y1 = std::max(y1, b.oby + 1);
      }
      if (any) SG_TRY(sg_map_make_writable(map, x0 + map->ox, y0 + map->oy, x1 + map->ox, y1 + map->oy, &copies));
    }
    if (pool) SG_TRY(sg_pool_run_copies(pool, copies));
    for (int k = 0; k < n; ++k) SG_TRY(sg_map_sync_tiles(maps[k]));
  }
  // ---- key space: the maps end to end
  unsigned long long key_total = 0;
  for (int k = 0; k < n; ++k) {
    slots[k] = slot_of(maps[k], plans[k].px, plans[k].py, slots[k].shift, (unsigned)key_total);
    slots[k].beam_begin = (int)beam0[k]; slots[k].beam_end = (int)beam0[k + 1];
    key_total += (unsigned long long)maps[k]->w * maps[k]->h;
  }
  if (key_total >= 0xFFFFFFFFull) return sg_fail(ctx, SLAMGPU_E_NOMEM, "maps too large for 32-bit cell keys (%llu cells): insert in smaller batches", key_total);
  if (memcmp(slots.data(), slots_sent.data(), sizeof(MapSlot) * n) != 0) {  // a map grew or cloned tiles since the first upload
    if (ctx->h_slots_cap < sizeof(MapSlot) * n) {
      if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
      ctx->h_slots = nullptr; ctx->h_slots_cap = 0;
      SG_CUDA(ctx, cudaMallocHost(&ctx->h_slots, sizeof(MapSlot) * n * 2));
      ctx->h_slots_cap = sizeof(MapSlot) * n * 2;
    } else {
      SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the first upload may still read the block
    }
    memcpy(ctx->h_slots, slots.data(), sizeof(MapSlot) * n);
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[7].p, ctx->h_slots, sizeof(MapSlot) * n, cudaMemcpyHostToDevice, ctx->stream));
  }

  // ---- K3a: per-slot AOO + sort keys
  DevBuf &slotbuf = ctx->scratch[4];
  // layout: aoo_p[M] aoo_q[M] (double) | keys[M] vals[M] keys_tmp[M] vals_tmp[M] (u32) | slot_beam[M] (i32) | sorted aoo
  const size_t cbytes = 0;
  size_t bytes = (size_t)M * (2 * sizeof(double) + 5 * sizeof(unsigned)) + 64 + (size_t)M * sizeof(SortedAoo) + 64;
  if (slotbuf.reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "slot buffers (%lld slots)", M);
  double *aoo_p = slotbuf.as<double>(), *aoo_q = aoo_p + M;
  unsigned *keys = (unsigned *)(aoo_q + M), *vals = keys + M, *keys_tmp = vals + M, *vals_tmp = keys_tmp + M;
  int *slot_beam = (int *)(vals_tmp + M);
  size_t coff = ((size_t)M * (2 * sizeof(double) + 5 * sizeof(unsigned)) + 15) & ~(size_t)15;
  unsigned long long *counters = (unsigned long long *)((char *)ctx->scratch[7].p + map_counters_offset(n));  // zeroed + cells counted by K2
  SortedAoo *sorted_aoo = (SortedAoo *)((char *)slotbuf.p + coff + cbytes);
  EstimateArgs ea;
  ea.beams = ctx->scratch[0].as<BeamRec>(); ea.bout = ctx->scratch[3].as<BeamOut>(); ea.offsets = ctx->scratch[1].as<long long>();
  ea.N = N; ea.M = M; ea.maps = ctx->scratch[7].as<MapSlot>(); ea.scale = maps[0]->scale; ea.est = *est;
  ea.cells = ctx->scratch[2].as<int2>();
  ea.aoo_p = aoo_p; ea.aoo_q = aoo_q; ea.keys = keys; ea.vals = vals; ea.slot_beam = slot_beam; ea.counters = counters;
  // the cells around the robot leave the sort and run on the side stream (with a pyramid's per-slot trace only for non-TBM cells):
  // the robot's own cell always (every ray starts in it: the longest chain of an insertion); for a few maps also the
  // ring of cells around it, whose runs of dozens to hundreds of updates would otherwise follow the sort
  const bool tbm_model = maps[0]->model == SLAMGPU_CELL_TBM_CONSISTENT || maps[0]->model == SLAMGPU_CELL_TBM_UNKNOWN_EVEN ||
                         maps[0]->model == SLAMGPU_CELL_CREDIBILIST;
  const bool robot_split = trace == nullptr || !tbm_model;  // (TBM chains publish their fields once per run: no per-update trace there)
  static const int ring_env = getenv("SLAMGPU_RING") ? atoi(getenv("SLAMGPU_RING")) : -1;  // experiment: 0..6
  const int ring = !robot_split ? -1 : (ring_env >= 0 && ring_env <= 6 ? ring_env : (n <= 8 ? 6 : 0));  // (k_apply_ring unrolls 2 * 6 + 1 = 13 candidate slots per ray)
  ea.ring = ring;
  // the ring kernel reads only the first 2 * ring + 1 cells of every ray: those are estimated first (one small launch), the
  // ring forks off, and the other slots are estimated beside it
  ea.head = robot_split ? 2 * ring + 1 : 0;
  if (ea.head > 0) {
    const unsigned hblk = (unsigned)(((long long)N * ea.head + 127) / 128);
    if (est->type == SLAMGPU_EST_AREA) k_estimate_head<true><<<hblk, 128, 0, ctx->stream>>>(ea);
    else k_estimate_head<false><<<hblk, 128, 0, ctx->stream>>>(ea);
    SG_LAUNCHED(ctx);
  }
  if (trace && ctx->scratch[5].reserve((size_t)M * sizeof(double) * (1 + maps[0]->stride)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "trace buffer");
  if (robot_split) {
    RobotArgs ra;
    ra.trace_impact = trace ? ctx->scratch[5].as<double>() : nullptr; ra.trace_rec = trace ? ra.trace_impact + M : nullptr;
    ra.trace_oie = trace ? trace->oie : 0;
    ra.maps = ctx->scratch[7].as<MapSlot>(); ra.n_maps = n; ra.beams = ctx->scratch[0].as<BeamRec>();
    ra.bout = ctx->scratch[3].as<BeamOut>(); ra.offsets = ctx->scratch[1].as<long long>(); ra.cells = ctx->scratch[2].as<int2>();
    ra.aoo_p = aoo_p; ra.aoo_q = aoo_q; ra.stride = maps[0]->stride; ra.model = maps[0]->model; ra.ring = ring;
    SG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    SG_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    const bool tbm_cells = ra.model == SLAMGPU_CELL_TBM_CONSISTENT || ra.model == SLAMGPU_CELL_TBM_UNKNOWN_EVEN || ra.model == SLAMGPU_CELL_CREDIBILIST;
    const long long warps = (long long)n * (2 * ring + 1) * (2 * ring + 1);
    const unsigned rblk = (unsigned)((warps * 32 + 127) / 128);
    if (tbm_cells) k_apply_ring<true, -1><<<rblk, 128, 0, ctx->side>>>(ra);
    else switch (ra.model) {
      case SLAMGPU_CELL_MEAN: k_apply_ring<false, SLAMGPU_CELL_MEAN><<<rblk, 128, 0, ctx->side>>>(ra); break;
      case SLAMGPU_CELL_AFFINE: k_apply_ring<false, SLAMGPU_CELL_AFFINE><<<rblk, 128, 0, ctx->side>>>(ra); break;
      case SLAMGPU_CELL_GMAPPING: k_apply_ring<false, SLAMGPU_CELL_GMAPPING><<<rblk, 128, 0, ctx->side>>>(ra); break;
      default: k_apply_ring<false, -1><<<rblk, 128, 0, ctx->side>>>(ra); break;
    }
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
  }

  if (est->type == SLAMGPU_EST_AREA) k_estimate<true><<<(unsigned)((M + 127) / 128), 128, 0, ctx->stream>>>(ea);
  else k_estimate<false><<<(unsigned)((M + 127) / 128), 128, 0, ctx->stream>>>(ea);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  // ---- sort by (map, cell) (stable), then apply each cell's run in order
  unsigned *ks, *vs;
  SG_TRY(sg_radix_sort(ctx, keys, vals, keys_tmp, vals_tmp, M, (unsigned)key_total, &ks, &vs));
  GatherArgs ga;
  ga.keys = ks; ga.vals = vs; ga.M = M; ga.aoo_p = aoo_p; ga.aoo_q = aoo_q; ga.slot_beam = slot_beam;
  ga.beams = ctx->scratch[0].as<BeamRec>(); ga.out = sorted_aoo;
  k_gather_sorted<<<(unsigned)((M + 127) / 128), 128, 0, ctx->stream>>>(ga);
  SG_LAUNCHED(ctx);
  ApplyArgs aa;
  aa.keys = ks; aa.aoo = sorted_aoo; aa.M = M; aa.maps = ctx->scratch[7].as<MapSlot>(); aa.stride = maps[0]->stride; aa.model = maps[0]->model;
  aa.trace_impact = nullptr; aa.trace_rec = nullptr; aa.trace_oie = 0;
  if (trace) { aa.trace_impact = ctx->scratch[5].as<double>(); aa.trace_rec = aa.trace_impact + M; aa.trace_oie = trace->oie; }  // (reserved above)
  // long-run queue: at most M / SG_LONG_RUN entries, the counter in front
  const size_t lr_bytes = 64 + ((size_t)(M / SG_LONG_RUN) + 1) * sizeof(LongRun);
  if (ctx->scratch[6].reserve(lr_bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "long-run queue");
  aa.n_long = ctx->scratch[6].as<unsigned>();
  aa.long_runs = (LongRun *)((char *)ctx->scratch[6].p + 64);
  SG_CUDA(ctx, cudaMemsetAsync(aa.n_long, 0, 64, ctx->stream));
  cudaEventRecord(ctx->evk0, ctx->stream);
  k_apply<<<(unsigned)((M + 127) / 128), 128, 0, ctx->stream>>>(aa);
  {
    const bool tbm_cells = aa.model == SLAMGPU_CELL_TBM_CONSISTENT || aa.model == SLAMGPU_CELL_TBM_UNKNOWN_EVEN || aa.model == SLAMGPU_CELL_CREDIBILIST;
    const int blocks = ctx->sm_count * 2;
    if (tbm_cells && trace) k_apply_long<true, true><<<blocks, 128, 0, ctx->stream>>>(aa);
    else if (tbm_cells) k_apply_long<true, false><<<blocks, 128, 0, ctx->stream>>>(aa);
    else if (trace) k_apply_long<false, true><<<blocks, 128, 0, ctx->stream>>>(aa);
    else k_apply_long<false, false><<<blocks, 128, 0, ctx->stream>>>(aa);
  }
  SG_LAUNCHED(ctx);
  cudaEventRecord(ctx->evk1, ctx->stream);
  ctx->evk_valid = true;
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  if (robot_split) SG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  for (int k = 0; k < n; ++k)
    if (plans[k].M > 0) sg_map_invalidate_lut(maps[k]);
  if (deferred) {
    SG_CUDA(ctx, cudaMemcpyAsync(deferred, counters, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    return SLAMGPU_OK;
  }
  ctx->staged_pending = false;  // (the synchronisation below covers the uploads too)
  if (ctx->h_counters_cap < (size_t)n * 16) {
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    ctx->h_counters = nullptr; ctx->h_counters_cap = 0;
    SG_CUDA(ctx, cudaMallocHost(&ctx->h_counters, (size_t)n * 32));
    ctx->h_counters_cap = (size_t)n * 32;
  }
  unsigned long long *h_counters = (unsigned long long *)ctx->h_counters;
  SG_CUDA(ctx, cudaMemcpyAsync(h_counters, counters, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (cells_updated) for (int k = 0; k < n; ++k) cells_updated[k] = (int64_t)h_counters[2 * k];
  if (trace) {
    trace->M = M; trace->applied = (int64_t)h_counters[0];
    trace->cells = ctx->scratch[2].as<int2>(); trace->keys_sorted = ks; trace->vals_sorted = vs;
    trace->impact = aa.trace_impact; trace->rec = aa.trace_rec; trace->slot_beam = slot_beam; trace->N = N;
    trace->d_bout = ctx->scratch[3].as<BeamOut>(); trace->d_offsets = ctx->scratch[1].as<long long>();
  }
  return SLAMGPU_OK;
}

extern "C" int slamgpu_append_scan(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3],
                                   double scan_quality, int32_t scan_margin, const slamgpu_estimator *est, double blur,
                                   double max_range, const double *point_quality, int64_t *cells_updated) {
  SG_NVTX("K2/K3 append_scan");
  if (!ctx) return SLAMGPU_E_INVALID;
  if (map && map->pyr) return sg_fail(ctx, SLAMGPU_E_STATE, "this map is level 0 of a pyramid: use slamgpu_pyramid_append_scan");
  return sg_append_scan_impl(ctx, map, scan, pose, scan_quality, scan_margin, est, blur, max_range, point_quality,
                             cells_updated, nullptr);
}

// beams given explicitly (the arguments of GridMapScanAdder::handle_scan_point) -> plan
int sg_plan_from_beams(slamgpu_ctx *ctx, const slamgpu_map *map, int32_t n, const double *beams, const uint8_t *is_occ,
                       const double *quality, double blur, double max_range, BeamPlan *out) {
  BeamPlan &plan = *out;
  const double scale = map->scale;
  const double px = beams[0], py = beams[1];
  plan.px = px; plan.py = py;
  plan.rx = host_world_to_cell(px, scale); plan.ry = host_world_to_cell(py, scale);
  plan.beams.assign(n, BeamRec{});
  plan.offsets.assign(n + 1, 0);
  const double max_range_sq = std::pow(max_range, 2);
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    const double *s = beams + 4 * i;
    if (s[0] != px || s[1] != py) return sg_fail(ctx, SLAMGPU_E_INVALID, "append_beams: beam %d starts elsewhere (one call = one scan)", i);
    BeamRec &b = plan.beams[i];
    plan.offsets[i] = total;
    b.wx = s[2]; b.wy = s[3]; b.quality = quality[i]; b.is_occ = is_occ[i] ? 1 : 0;
    const double len_sq = std::pow(b.wx - px, 2) + std::pow(b.wy - py, 2);  // Segment2D::length_sq
    if (max_range_sq < len_sq) continue;
    if (!std::isfinite(b.wx) || !std::isfinite(b.wy)) continue;
    b.obx = host_world_to_cell(b.wx, scale); b.oby = host_world_to_cell(b.wy, scale);
    b.obst_sq = std::pow(plan.rx - b.obx, 2) + std::pow(plan.ry - b.oby, 2);
    double blur_dist = 0;
    if (b.is_occ) {
      blur_dist = blur / scale;
      if (blur_dist < 0) blur_dist *= -len_sq;
    }
    b.hole_sq = std::pow(blur_dist, 2);
    b.active = 1;
    long long ub = std::llabs((long long)b.obx - plan.rx) + std::llabs((long long)b.oby - plan.ry) + 2;
    if (ub > (1ll << 26)) return sg_fail(ctx, SLAMGPU_E_INVALID, "beam %d spans more than 2^26 cells", i);
    total += ub;
  }
  plan.offsets[n] = total;
  plan.M = total;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_append_beams(slamgpu_ctx *ctx, slamgpu_map *map, int32_t n, const double *beams, const uint8_t *is_occ,
                                    const double *quality, const slamgpu_estimator *est, double blur, double max_range,
                                    int64_t *cells_updated) {
  SG_NVTX("K2/K3 append_beams");
  if (!ctx || !map || n < 0 || (n > 0 && (!beams || !is_occ || !quality)) || !est) return sg_fail(ctx, SLAMGPU_E_INVALID, "append_beams: bad argument");
  if (map->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map belongs to another ctx");
  if (map->pyr) return sg_fail(ctx, SLAMGPU_E_STATE, "this map is level 0 of a pyramid: use slamgpu_pyramid_append_beams");
  if (cells_updated) *cells_updated = 0;
  if (n == 0) return SLAMGPU_OK;
  BeamPlan plan;
  SG_TRY(sg_plan_from_beams(ctx, map, n, beams, is_occ, quality, blur, max_range, &plan));
  return sg_append_plan(ctx, map, plan, est, cells_updated, nullptr);
}

extern "C" int slamgpu_raycast(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const double pose[3],
                               int64_t *out_offsets, int32_t *out_cells, int64_t cap, int64_t *total) {
  SG_NVTX("K2 raycast");
  if (!ctx || !map || !scan || !pose || !out_offsets) return sg_fail(ctx, SLAMGPU_E_INVALID, "raycast: NULL argument");
  if (map->ctx != ctx || scan->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map/scan belongs to another ctx");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  const int N = scan->n;
  BeamPlan plan;
  if (sg_prepare_beams(map, scan, pose, 1.0, 0, 0.0, INFINITY, nullptr, false, &plan) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "a beam spans more than 2^26 cells");
  slamgpu_estimator est;
  memset(&est, 0, sizeof est);
  est.type = SLAMGPU_EST_CONST;
  SG_TRY(run_raycast(ctx, map, plan, est));
  std::vector<BeamOut> bout(std::max(N, 1));
  if (N > 0) SG_CUDA(ctx, cudaMemcpyAsync(bout.data(), ctx->scratch[3].p, sizeof(BeamOut) * N, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t acc = 0;
  for (int i = 0; i < N; ++i) { out_offsets[i] = acc; acc += bout[i].count; }
  out_offsets[N] = acc;
  if (total) *total = acc;
  if (out_cells && cap > 0) {
    std::vector<int2> cells((size_t)std::max<long long>(plan.M, 1));
    if (plan.M > 0) SG_CUDA(ctx, cudaMemcpy(cells.data(), ctx->scratch[2].p, sizeof(int2) * plan.M, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < bout[i].count; ++k) {
        int64_t o = out_offsets[i] + k;
        if (o >= cap) break;
        out_cells[2 * o] = cells[plan.offsets[i] + k].x;
        out_cells[2 * o + 1] = cells[plan.offsets[i] + k].y;
      }
  }
  return SLAMGPU_OK;
}

// ---------------------------------------------------------------- unit-level entry points
// (the reference's own unit interfaces, used by the golden-vector tests and the C++ adapters)
namespace {
__global__ void k_raycast_segments(const double4 *__restrict__ segs, const long long *__restrict__ offsets, int n, double scale,
                                   int2 *cells, int *counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 s = segs[i];
  int2 *dst = cells + offsets[i];
  counts[i] = sg::raycast(s.x, s.y, s.z, s.w, scale, [&](int k, int x, int y) { dst[k] = make_int2(x, y); });
}
__global__ void k_estimate_batch(slamgpu_estimator est, double shift, const double4 *__restrict__ beams,
                                 const double4 *__restrict__ bounds, const unsigned char *__restrict__ is_occ, int n,
                                 double2 *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 b = beams[i], c = bounds[i];
  double p, q;
  sg::estimate_occupancy(est, shift, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, is_occ[i] != 0, &p, &q);
  out[i] = make_double2(p, q);
}
}  // namespace

extern "C" int slamgpu_raycast_segments(slamgpu_ctx *ctx, double scale, const double *segments, int32_t n,
                                        int64_t *out_offsets, int32_t *out_cells, int64_t cap, int64_t *total) {
  if (!ctx || n < 0 || (n > 0 && !segments) || !out_offsets || !(scale > 0)) return sg_fail(ctx, SLAMGPU_E_INVALID, "raycast_segments: bad argument");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<long long> offs(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    const double *s = segments + 4 * i;
    for (int k = 0; k < 4; ++k)
      if (!std::isfinite(s[k])) return sg_fail(ctx, SLAMGPU_E_INVALID, "segment %d is not finite", i);
    long long ub = std::llabs((long long)host_world_to_cell(s[2], scale) - host_world_to_cell(s[0], scale)) +
                   std::llabs((long long)host_world_to_cell(s[3], scale) - host_world_to_cell(s[1], scale)) + 2;
    if (ub > (1ll << 26)) return sg_fail(ctx, SLAMGPU_E_INVALID, "segment %d spans more than 2^26 cells", i);
    offs[i + 1] = offs[i] + ub;
  }
  const long long M = offs[n];
  SG_TRY(upload_async(ctx, ctx->scratch[0], segments, sizeof(double) * 4 * n));
  SG_TRY(upload_async(ctx, ctx->scratch[1], offs.data(), sizeof(long long) * (n + 1)));
  if (ctx->scratch[2].reserve(std::max<size_t>(M, 1) * sizeof(int2)) != SLAMGPU_OK ||
      ctx->scratch[3].reserve(std::max<size_t>(n, 1) * sizeof(int)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "ray-cast buffers");
  std::vector<int> counts(std::max(n, 1), 0);
  std::vector<int2> cells((size_t)std::max<long long>(M, 1));
  if (n > 0) {
    k_raycast_segments<<<(n + 31) / 32, 32, 0, ctx->stream>>>(ctx->scratch[0].as<double4>(), ctx->scratch[1].as<long long>(), n,
                                                              scale, ctx->scratch[2].as<int2>(), ctx->scratch[3].as<int>());
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
    SG_CUDA(ctx, cudaMemcpyAsync(counts.data(), ctx->scratch[3].p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(cells.data(), ctx->scratch[2].p, sizeof(int2) * M, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t acc = 0;
  for (int i = 0; i < n; ++i) { out_offsets[i] = acc; acc += counts[i]; }
  out_offsets[n] = acc;
  if (total) *total = acc;
  if (out_cells)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < counts[i]; ++k) {
        int64_t o = out_offsets[i] + k;
        if (o >= cap) break;
        out_cells[2 * o] = cells[offs[i] + k].x;
        out_cells[2 * o + 1] = cells[offs[i] + k].y;
      }
  return SLAMGPU_OK;
}

extern "C" int slamgpu_estimate_occupancy(slamgpu_ctx *ctx, const slamgpu_estimator *est, int32_t n, const double *beams,
                                          const double *cell_bounds, const uint8_t *is_occ, double *out_pq) {
  if (!ctx || !est || n < 0 || (n > 0 && (!beams || !cell_bounds || !is_occ || !out_pq))) return sg_fail(ctx, SLAMGPU_E_INVALID, "estimate_occupancy: bad argument");
  if (n == 0) return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  SG_TRY(upload_async(ctx, ctx->scratch[0], beams, sizeof(double) * 4 * n));
  SG_TRY(upload_async(ctx, ctx->scratch[1], cell_bounds, sizeof(double) * 4 * n));
  SG_TRY(upload_async(ctx, ctx->scratch[2], is_occ, n));
  if (ctx->scratch[3].reserve(sizeof(double2) * n) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "estimate buffers");
  // Shift_Amount: explicit, or low_qual * side of the first cell (quirk Q10)
  double shift = est->shift_amount >= 0 ? est->shift_amount : est->low_qual * (cell_bounds[1] - cell_bounds[0]);
  k_estimate_batch<<<(n + 63) / 64, 64, 0, ctx->stream>>>(*est, shift, ctx->scratch[0].as<double4>(), ctx->scratch[1].as<double4>(),
                                                         ctx->scratch[2].as<unsigned char>(), n, ctx->scratch[3].as<double2>());
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaMemcpyAsync(out_pq, ctx->scratch[3].p, sizeof(double2) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

#define SG_TOUCH(k) do { cudaFuncAttributes fa_; (void)cudaFuncGetAttributes(&fa_, k); } while (0)
void sg_preload_mapping() {  // see sg_preload_score
  SG_TOUCH(k_raycast); SG_TOUCH(k_estimate<true>); SG_TOUCH(k_estimate<false>); SG_TOUCH(k_estimate_head<true>); SG_TOUCH(k_estimate_head<false>); SG_TOUCH(k_radix_hist); SG_TOUCH(k_radix_scan);
  SG_TOUCH(k_scan_chunks); SG_TOUCH(k_scan_add); SG_TOUCH(k_radix_scatter); SG_TOUCH(k_gather_sorted); SG_TOUCH(k_apply);
  SG_TOUCH((k_apply_long<true, false>)); SG_TOUCH((k_apply_long<false, false>)); SG_TOUCH((k_apply_long<false, true>));
  SG_TOUCH((k_apply_ring<true, -1>)); SG_TOUCH((k_apply_ring<false, -1>)); SG_TOUCH((k_apply_ring<false, SLAMGPU_CELL_MEAN>));
  SG_TOUCH((k_apply_ring<false, SLAMGPU_CELL_AFFINE>)); SG_TOUCH((k_apply_ring<false, SLAMGPU_CELL_GMAPPING>)); SG_TOUCH(k_copy_block); SG_TOUCH(k_fill);
  (void)cudaGetLastError();
}
