// score.cu -- K1: batched likelihood of one laser scan under many candidate poses.
//
// Replaces the candidate loop of PoseEnumerationScanMatcher::process_scan
// (src/core/scan_matchers/pose_enumeration_scan_matcher.h:48-65) around
// WeightedMeanPointProbabilitySPE::estimate_scan_probability
// (src/core/scan_matchers/weighted_mean_point_probability_spe.h:97-133).
//
// Two kernels:
//   k_score_list  any pose list: one thread per pose walks the scan points in index order
//                 (the reference's FP64 summation order), beam trig comes from a per-theta
//                 table built once per candidate set;
//   k_score_grid  the Cartesian candidate set of BruteForcePoseEnumerator
//                 (brute_force_scan_matcher.h:10-64): cell indices are separable in
//                 (theta, beam, x) and (theta, beam, y), so they are precomputed into two
//                 small integer tables and the inner loop is gather + multiply + add.
// Both end in a warp-shuffle arg-max whose tie-break is the lowest candidate index, which
// is what the reference's sequential strict-'<' accept loop selects.
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>

#include "dev_window.cuh"
#include "hill_climb.h"
#include "internal.h"

namespace {

struct Best {
  double score;
  long long idx;
};
struct Result {
  double score;
  long long idx;
  long long guard;
  long long pad;
};

// a beats b in the reference's accept loop: strictly larger, or equal with a lower index
SG_DEV bool beats(double s, long long i, double bs, long long bi) { return (s > bs) || (s == bs && i < bi); }

SG_DEV void warp_argmax(double &s, long long &i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double os = __shfl_down_sync(0xffffffffu, s, off);
    long long oi = __shfl_down_sync(0xffffffffu, i, off);
    if (beats(os, oi, s, i)) { s = os; i = oi; }
  }
}

// block arg-max -> *dst (written by thread 0); every thread of the block must call
SG_DEV void block_argmax(double s, long long i, Best *dst) {
  __shared__ double sh_s[32];
  __shared__ long long sh_i[32];
  warp_argmax(s, i);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sh_s[wid] = s; sh_i[wid] = i; }
  __syncthreads();
  if (wid == 0) {
    s = lane < nw ? sh_s[lane] : -INFINITY;
    i = lane < nw ? sh_i[lane] : LLONG_MAX;
    warp_argmax(s, i);
    if (lane == 0) { dst->score = s; dst->idx = i; }
  }
}

__global__ void k_reduce_blocks(const Best *__restrict__ blk, int n, Result *out) {
  double s = -INFINITY;
  long long i = LLONG_MAX;
  for (int k = threadIdx.x; k < n; k += blockDim.x)  // ascending k per thread keeps the lowest index on ties
    if (beats(blk[k].score, blk[k].idx, s, i)) { s = blk[k].score; i = blk[k].idx; }
  block_argmax(s, i, reinterpret_cast<Best *>(out));  // Result starts with {score, idx}
}

// merge the per-rank results (after the all-gather) and apply the accept rule against
// init_score: the winner must be strictly better than the initial pose's score
__global__ void k_finalize(const Result *__restrict__ per_rank, int nranks, double init_score, Result *out) {
  double s = -INFINITY;
  long long i = LLONG_MAX, guard = 0, enc = 0;
  for (int r = 0; r < nranks; ++r) {
    guard += per_rank[r].guard;
    enc += per_rank[r].pad;
    if (beats(per_rank[r].score, per_rank[r].idx, s, i)) { s = per_rank[r].score; i = per_rank[r].idx; }
  }
  if (!(init_score < s) || i == LLONG_MAX) { s = init_score; i = -1; }
  out->score = s; out->idx = i; out->guard = guard; out->pad = enc;
}

// one rank: k_reduce_blocks and k_finalize in one launch (thread 0 holds the reduced result and applies the accept rule)
__global__ void __launch_bounds__(1024) k_reduce_finalize(const Best *__restrict__ blk, int n, Result *res, double init_score, Result *out) {
  double s = -INFINITY;
  long long i = LLONG_MAX;
  for (int k = threadIdx.x; k < n; k += blockDim.x)
    if (beats(blk[k].score, blk[k].idx, s, i)) { s = blk[k].score; i = blk[k].idx; }
  block_argmax(s, i, reinterpret_cast<Best *>(res));
  if (threadIdx.x != 0) return;
  s = res->score; i = res->idx;  // written by this thread
  if (!(init_score < s) || i == LLONG_MAX) { s = init_score; i = -1; }
  out->score = s; out->idx = i; out->guard = res->guard; out->pad = res->pad;
}

// The same merge fused with the exchange: instead of ncclAllGather + k_finalize, lane r of one warp stores this
// rank's result straight into rank r's mailbox (a peer mapping over NVLink; slot = this rank, double-buffered by the
// parity of the call number), publishes it with a system-scope fence + sequence number, then waits for rank r's result
// to land in this rank's own mailbox.  Lane 0 merges in rank order exactly as k_finalize does.  A peer cannot run two
// calls ahead (its next exchange needs this rank's next message), so two buffers suffice.
struct Mail { Result res; unsigned long long seq; unsigned long long pad[3]; };  // 64 B
// The block results are reduced here as well (blk != NULL: what k_reduce_blocks does, by the same 1024 threads), so a
// multi-rank step has one launch less.  A peer that has not delivered within `deadline` clock cycles (ctx option
// "p2p_timeout_ms", default 20 s: rank skew from module loads, map regrowth or a stalled host thread is not an error) makes the
// fetch fail and the ctx fall back to ncclAllGather for the following calls, so a late message can never be taken for a new one.
__global__ void __launch_bounds__(1024) k_exchange_finalize(const Best *__restrict__ blk, int n_blk, Result *mine, void *const *peer_mailbox, int rank,
                                                            int nranks, unsigned long long seq, long long deadline, double init_score, Result *out,
                                                            int *status) {
  const int r = threadIdx.x;
  const int parity = (int)(seq & 1ull);
  __shared__ Result got[64];
  __shared__ int timed_out;
  if (r == 0) timed_out = 0;
  if (blk) {
    double s = -INFINITY;
    long long i = LLONG_MAX;
    for (int k = r; k < n_blk; k += blockDim.x)
      if (beats(blk[k].score, blk[k].idx, s, i)) { s = blk[k].score; i = blk[k].idx; }
    block_argmax(s, i, reinterpret_cast<Best *>(mine));  // {score, idx} of this rank; guard / pad were accumulated by the kernels
  }
  __syncthreads();
  if (r < nranks) {
    volatile Mail *dst = reinterpret_cast<volatile Mail *>(peer_mailbox[r]) + (size_t)parity * nranks + rank;
    volatile Result *vm = mine;
    dst->res.score = vm->score; dst->res.idx = vm->idx; dst->res.guard = vm->guard; dst->res.pad = vm->pad;
    __threadfence_system();
    dst->seq = seq;
    volatile Mail *src = reinterpret_cast<volatile Mail *>(peer_mailbox[rank]) + (size_t)parity * nranks + r;
    const long long t0 = clock64();
    bool ok = true;
    while (src->seq != seq) {
      if (clock64() - t0 > deadline) { ok = false; break; }
    }
    __threadfence_system();
    if (!ok) { atomicExch(status, 1); timed_out = 1; }
    got[r].score = src->res.score; got[r].idx = src->res.idx; got[r].guard = src->res.guard; got[r].pad = src->res.pad;
  }
  __syncthreads();
  if (r == 0) {
    double s = -INFINITY;
    long long i = LLONG_MAX, guard = 0, enc = 0;
    for (int k = 0; k < nranks; ++k) {
      guard += got[k].guard;
      enc += got[k].pad;
      if (beats(got[k].score, got[k].idx, s, i)) { s = got[k].score; i = got[k].idx; }
    }
    if (!(init_score < s) || i == LLONG_MAX) { s = init_score; i = -1; }
    if (timed_out) { i = LLONG_MIN; guard = 0; enc = 0; }  // reported by the fetch
    out->score = s; out->idx = i; out->guard = guard; out->pad = enc;
  }
}

// Streams a table through L2 (ld.global.cg: L2 only).  A scoring kernel that follows an L2-cold LUT (the map was just
// rebuilt, or another working set went through the cache) would take its first touches from DRAM inside the latency-bound
// beam loop; reading the 32 MB score LUT once costs ~5 us of HBM time and runs on the side stream beside the index kernels.
__global__ void k_warm_l2(const double2 *__restrict__ p, size_t n16, double *sink) {
  double acc = 0.0;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += 4 * st) {
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u * st < n16) ? __ldcg(p + i + u * st) : make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y;
  }
  if (acc == -1.2345678e-300) *sink = acc;  // never true for LUT values (they are in [0, 1]); keeps the loads alive
}

// ------------------------------------------------------------------ trig table
// ScanPoint2D::move_origin + RawTrigonometryProvider: src/core/states/sensor_data.h:83-87,
// src/core/trigonometry_utils.h:21-27 -- r*cos(theta + a), r*sin(theta + a)
__global__ void k_trig_table(const double *__restrict__ thetas, int T, const double *__restrict__ range,
                             const double *__restrict__ angle, int N, long long st_t, long long st_i,
                             double *__restrict__ trc, double *__restrict__ trs) {
  long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (id >= (long long)T * N) return;
  int t = (int)(id / N), i = (int)(id - (long long)t * N);
  double s, c;
  sincos(sg::add(thetas[t], angle[i]), &s, &c);
  long long o = t * st_t + i * st_i;
  trc[o] = sg::mul(range[i], c);
  trs[o] = sg::mul(range[i], s);
}

// slack of a world coordinate built from device trig (see slamgpu.h, SLAMGPU_TRIG_DEVICE)
SG_DEV double trig_slack(double rc, double X) { return 4e-15 * fabs(rc) + 1e-15 * fabs(X); }

// GmappingOccupancyObservationPE::probability, src/slams/gmapping/gmapping_occupancy_observation_pe.h:17-38
SG_DEV double gmapping_probability(const MapView &m, int cx, int cy, double X, double Y, double th, int win) {
  double best = 0;
  for (int dx = -win; dx <= win; ++dx)
    for (int dy = -win; dy <= win; ++dy) {
      int ix = cx + dx + m.ox, iy = cy + dy + m.oy;
      double r[SLAMGPU_MAX_STRIDE];
      if (ix < 0 || ix >= m.w || iy < 0 || iy >= m.h) {
#pragma unroll
        for (int k = 0; k < SLAMGPU_MAX_STRIDE; ++k) r[k] = m.unknown_rec[k];
      } else {
        const double *src = view_cell(m, ix, iy);
        for (int k = 0; k < m.stride; ++k) r[k] = __ldg(src + k);
      }
      if (r[0] < th) continue;
      double v = sg::sub(1.0, sg::cell_discrepancy_obstacle(m.model, r, X, Y));
      best = sg::maxd(best, v);
    }
  return best;
}

// exact cell of a world coordinate (+ the device-trig border guard; the tolerance itself need not be exact)
SG_DEV int grid_cell(double v, double rc, double scale, double inv_scale, int guard, bool *unsafe) {
  {  // the common case: the quotient is nowhere near a cell border (see grid_axis_cell for the bound)
    const double qa = v * inv_scale, fa = floor(qa), da = qa - fa;
    if (da > 1e-6 && da < 1.0 - 1e-6 && fabs(qa) < 1e9) return cell_of(fa);
  }
  const double f = sg::floor_div(v, scale, inv_scale);
  if (guard) {
    const double q = v * inv_scale;
    const double tol = trig_slack(rc, v) * inv_scale + 8.0 * 2.220446049250313e-16 * fabs(q);
    if ((q - f) <= tol || ((f + 1.0) - q) <= tol) *unsafe = true;
  }
  return cell_of(f);
}

struct ListArgs {
  MapView map;
  const MapView *views;  // K6: per-particle maps (NULL: every pose against `map`)
  const int *view_id;    // K6: per local pose
  const double *poses;      // 3*Ploc
  const int *theta_id;      // Ploc (table column), unused when prerotated
  const double *trc, *trs;  // [i*T + tid]
  int T;
  const double *sx, *sy, *w, *f;
  int N;
  long long Ploc, p0;  // local count, global index of local pose 0
  double wsum, win_v, win_h, gm_th;
  int gm_win, gm_cache;
  double *scores;  // Ploc
  Best *blk;
  Result *result;  // guard counter lives here
};

template <int MODE, bool PREROT, bool GUARD, bool FACTOR>
__global__ void __launch_bounds__(128) k_score_list(ListArgs a) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if (p < a.Ploc) {
    const double px = a.poses[3 * p], py = a.poses[3 * p + 1];
    const int tid = PREROT ? 0 : a.theta_id[p];
    const MapView &mv = a.views ? a.views[a.view_id[p]] : a.map;
    const double s = mv.scale, inv_s = 1.0 / mv.scale;
    double total_probability = 0;
    bool unsafe_any = false;
    int cache_x = 0, cache_y = 0;
    double cache_p = -1;
    for (int i = 0; i < a.N; ++i) {
      double X, Y, rc = 0, rs = 0;
      if (PREROT) {  // sensor_data.h:103-105 -- point + pose offset
        X = sg::add(__ldg(a.sx + i), px); Y = sg::add(__ldg(a.sy + i), py);
      } else {
        rc = __ldg(a.trc + (size_t)i * a.T + tid); rs = __ldg(a.trs + (size_t)i * a.T + tid);
        X = sg::add(px, rc); Y = sg::add(py, rs);
      }
      double prob;
      if (MODE == SLAMGPU_OOPE_OBSTACLE || MODE == SLAMGPU_OOPE_GMAPPING) {
        int cx, cy;
        cx = grid_cell(X, rc, s, inv_s, GUARD ? 1 : 0, &unsafe_any);
        cy = grid_cell(Y, rs, s, inv_s, GUARD ? 1 : 0, &unsafe_any);
        if (MODE == SLAMGPU_OOPE_OBSTACLE) {
          prob = lut_at(mv, cx, cy);
        } else {
          if (a.gm_cache && cx == cache_x && cy == cache_y && cache_p != -1) {
            prob = cache_p;
          } else {
            prob = gmapping_probability(mv, cx, cy, X, Y, a.gm_th, a.gm_win);
            cache_x = cx; cache_y = cy; cache_p = prob;
          }
        }
      } else {
        if (GUARD) {  // window corners depend on X, Y too: guard the four borders
          double hv = sg::div(a.win_v, 2.0), hh = sg::div(a.win_h, 2.0);
          bool u;
          sg::world_to_cell_guard(sg::sub(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
          sg::world_to_cell_guard(sg::add(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
          sg::world_to_cell_guard(sg::sub(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
          sg::world_to_cell_guard(sg::add(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        }
        prob = window_probability<MODE>(mv, X, Y, a.win_v, a.win_h);
      }
      double term = sg::mul(prob, __ldg(a.w + i));
      if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
      total_probability = sg::add(total_probability, term);
    }
    double score = a.wsum == 0 ? NAN : sg::div(total_probability, a.wsum);
    a.scores[p] = score;
    if (GUARD && unsafe_any) atomicAdd((unsigned long long *)&a.result->guard, 1ull);
    if (score == score) { best_s = score; best_i = a.p0 + p; }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

// ---- small candidate sets (Monte-Carlo batches, hill-climbing rounds, particle rounds): latency, not
// throughput, is what matters, and one thread per pose would walk its ~1000 dependent gathers alone.
// Phase 1 spreads the (pose, point) pairs over the whole device -- trig computed in place, no table;
// phase 2 adds each pose's terms in point order (the reference's FP64 summation order) and feeds the
// same arg-max.
struct PointArgs {
  ListArgs l;
  const double *thetas;   // distinct thetas (indexed by theta_id)
  const double *range, *angle;
  int trig_is_table;      // 1: l.trc/l.trs hold a host (libm) table, 0: sincos here
  double *terms;          // [Ploc][N]
  int2 *cellids;          // [Ploc][N] (GMapping cache emulation only)
};

template <int MODE, bool PREROT, bool GUARD, bool FACTOR>
__global__ void __launch_bounds__(128) k_point_terms(PointArgs pa) {
  const ListArgs &a = pa.l;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const long long p = blockIdx.y;
  if (i >= a.N) return;
  const double px = a.poses[3 * p], py = a.poses[3 * p + 1];
  const MapView &mv = a.views ? a.views[a.view_id[p]] : a.map;
  const double s = mv.scale, inv_s = 1.0 / mv.scale;
  double X, Y, rc = 0, rs = 0;
  bool unsafe_any = false;
  if (PREROT) {
    X = sg::add(__ldg(a.sx + i), px); Y = sg::add(__ldg(a.sy + i), py);
  } else {
    const int tid = a.theta_id[p];
    if (pa.trig_is_table) {
      rc = __ldg(a.trc + (size_t)i * a.T + tid); rs = __ldg(a.trs + (size_t)i * a.T + tid);
    } else {
      double sn, cs;
      sincos(sg::add(pa.thetas[tid], pa.angle[i]), &sn, &cs);
      rc = sg::mul(pa.range[i], cs); rs = sg::mul(pa.range[i], sn);
    }
    X = sg::add(px, rc); Y = sg::add(py, rs);
  }
  double prob;
  if (MODE == SLAMGPU_OOPE_OBSTACLE || MODE == SLAMGPU_OOPE_GMAPPING) {
    const int cx = grid_cell(X, rc, s, inv_s, GUARD ? 1 : 0, &unsafe_any);
    const int cy = grid_cell(Y, rs, s, inv_s, GUARD ? 1 : 0, &unsafe_any);
    if (MODE == SLAMGPU_OOPE_OBSTACLE) {
      prob = lut_at(mv, cx, cy);
    } else {
      prob = gmapping_probability(mv, cx, cy, X, Y, a.gm_th, a.gm_win);
      if (a.gm_cache) pa.cellids[(size_t)p * a.N + i] = make_int2(cx, cy);
    }
  } else {
    if (GUARD) {
      double hv = sg::div(a.win_v, 2.0), hh = sg::div(a.win_h, 2.0);
      bool u;
      sg::world_to_cell_guard(sg::sub(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
      sg::world_to_cell_guard(sg::add(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
      sg::world_to_cell_guard(sg::sub(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
      sg::world_to_cell_guard(sg::add(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
    }
    prob = window_probability<MODE>(mv, X, Y, a.win_v, a.win_h);
  }
  // GMapping cache emulation needs the raw probability in phase 2 (the cached value replaces it)
  double term = prob;
  if (!(MODE == SLAMGPU_OOPE_GMAPPING && a.gm_cache)) {
    term = sg::mul(prob, __ldg(a.w + i));
    if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
  }
  pa.terms[(size_t)p * a.N + i] = term;
  if (GUARD && unsafe_any) atomicAdd((unsigned long long *)&a.result->guard, 1ull);
}

template <bool GMCACHE, bool FACTOR>
__global__ void __launch_bounds__(128) k_pose_sums(PointArgs pa) {
  const ListArgs &a = pa.l;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if (p < a.Ploc) {
    const double *t = pa.terms + (size_t)p * a.N;
    double total = 0;
    if (GMCACHE) {  // GmappingOccupancyObservationPE's 1-entry cell cache, restarted per pose (quirk Q7)
      const int2 *c = pa.cellids + (size_t)p * a.N;
      int cache_x = 0, cache_y = 0;
      double cache_p = -1;
      for (int i = 0; i < a.N; ++i) {
        double prob = t[i];
        if (c[i].x == cache_x && c[i].y == cache_y && cache_p != -1) prob = cache_p;
        else { cache_x = c[i].x; cache_y = c[i].y; cache_p = prob; }
        double term = sg::mul(prob, __ldg(a.w + i));
        if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
        total = sg::add(total, term);
      }
    } else {
      for (int i = 0; i < a.N; ++i) total = sg::add(total, t[i]);
    }
    double score = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
    a.scores[p] = score;
    if (score == score) { best_s = score; best_i = a.p0 + p; }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

// ---- a whole hill-climbing match in ONE launch: one block per matcher instance (a world's matcher, or one GMapping
// particle against its own map).  The block scores the candidates of a round (at most 6 poses x N points, terms in
// shared memory, each pose's terms added in point order by one thread), thread 0 replays the reference's accept loop
// and enumerator (hill_climb.h) and emits the next round, until the failed-rounds budget is spent.  No host round
// trip per round: the ~15 rounds of a match cost ~15 block-level iterations instead of ~15 launch + copy cycles.
struct HcArgs {
  MapView map;
  const MapView *views;  // per instance (NULL: every instance against `map`)
  int n_inst;
  const double *init;           // 3 per instance
  const unsigned char *active;  // per instance or NULL
  unsigned max_failed;
  double tr, rot;
  const double *range, *angle, *w, *f;
  int N;
  double wsum, win_v, win_h, gm_th;
  int gm_win;
  double *out;   // SG_HC_OUT per instance: best x, y, theta, probability, poses tested, guard hits, log entries, 0,
                 // then the GMapping OOPE cache after the match: cell x, cell y, probability
  double *log;   // per instance: log_cap x {x, y, theta, score}, in evaluation order
  int log_cap;
  const slamgpu_gm_cache *gm_in;  // per instance, or NULL: the GMapping OOPE's cache is carried through the match
};
#define SG_HC_OUT 12

template <int MODE, bool FACTOR>
__global__ void __launch_bounds__(1024) k_hill_climb(HcArgs a) {
  extern __shared__ double sh_terms[];  // [6][N] terms (raw probabilities when the cache is carried), then [6][N] cell ids
  __shared__ __align__(8) unsigned char hc_raw[sizeof(HillClimb)];  // (a __shared__ object cannot have initialisers)
  __shared__ int s_cx, s_cy, s_ecx[6], s_ecy[6];   // carried cache: at the round's entry, and after each of its poses
  __shared__ double s_cp, s_ecp[6];
  const bool chain = MODE == SLAMGPU_OOPE_GMAPPING && a.gm_in != nullptr;
  int2 *sh_cell = reinterpret_cast<int2 *>(sh_terms + (size_t)6 * a.N);
  HillClimb &hc = *reinterpret_cast<HillClimb *>(hc_raw);
  __shared__ double cand[6][3], scores[6];
  __shared__ int s_k, s_logged;
  __shared__ unsigned int s_guard;
  const int inst = blockIdx.x, tid = threadIdx.x, N = a.N;
  double *out = a.out + SG_HC_OUT * (size_t)inst;
  const double *init = a.init + 3 * (size_t)inst;
  if (a.active && !a.active[inst]) {
    if (tid == 0) {
      out[0] = init[0]; out[1] = init[1]; out[2] = init[2]; out[3] = NAN; out[4] = 0; out[5] = 0; out[6] = 0; out[7] = 0;
      if (chain) { out[8] = a.gm_in[inst].cx; out[9] = a.gm_in[inst].cy; out[10] = a.gm_in[inst].prob; }
    }
    return;
  }
  const MapView &mv = a.views ? a.views[inst] : a.map;
  const double s = mv.scale, inv_s = 1.0 / mv.scale;
  if (tid == 0) {
    cand[0][0] = init[0]; cand[0][1] = init[1]; cand[0][2] = init[2];
    s_k = 1; s_guard = 0; s_logged = 0;
    if (chain) { s_cx = a.gm_in[inst].cx; s_cy = a.gm_in[inst].cy; s_cp = a.gm_in[inst].prob; }
  }
  __syncthreads();
  bool first = true;
  for (;;) {
    const int k = s_k;
    bool unsafe_any = false;
    for (int e = tid; e < k * N; e += blockDim.x) {
      const int j = e / N, i = e - j * N;
      const double px = cand[j][0], py = cand[j][1];
      double sn, cs;
      sincos(sg::add(cand[j][2], __ldg(a.angle + i)), &sn, &cs);
      const double r = __ldg(a.range + i);
      const double rc = sg::mul(r, cs), rs = sg::mul(r, sn);
      const double X = sg::add(px, rc), Y = sg::add(py, rs);
      double prob;
      if (MODE == SLAMGPU_OOPE_OBSTACLE || MODE == SLAMGPU_OOPE_GMAPPING) {
        const int cx = grid_cell(X, rc, s, inv_s, 1, &unsafe_any);
        const int cy = grid_cell(Y, rs, s, inv_s, 1, &unsafe_any);
        prob = MODE == SLAMGPU_OOPE_OBSTACLE ? lut_at(mv, cx, cy) : gmapping_probability(mv, cx, cy, X, Y, a.gm_th, a.gm_win);
        if (chain) sh_cell[e] = make_int2(cx, cy);
      } else {
        double hv = sg::div(a.win_v, 2.0), hh = sg::div(a.win_h, 2.0);
        bool u;
        sg::world_to_cell_guard(sg::sub(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::sub(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        prob = window_probability<MODE>(mv, X, Y, a.win_v, a.win_h);
      }
      double term = prob;  // the carried cache replaces raw probabilities: weights are applied while summing
      if (!chain) {
        term = sg::mul(prob, __ldg(a.w + i));
        if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
      }
      sh_terms[e] = term;
    }
    if (unsafe_any) atomicAdd(&s_guard, 1u);
    __syncthreads();
    if (chain && (tid & 31) == 0 && (tid >> 5) < k) {
      // GmappingOccupancyObservationPE's cache through the round (see k_pose_sums_chained): pose j enters with the cache
      // its predecessor left -- found by walking the predecessors' cell ids backwards, the round's entry state at the end
      const int j = tid >> 5;
      int cache_x = s_cx, cache_y = s_cy;
      double cache_p = s_cp;
      int depth = 0;
      for (int q = j - 1; q >= 0; --q) {
        const int2 *c = sh_cell + (size_t)q * N;
        int jj = N - 1;
        while (jj > 0 && c[jj - 1].x == c[jj].x && c[jj - 1].y == c[jj].y) --jj;
        if (jj > 0) { cache_x = c[jj].x; cache_y = c[jj].y; cache_p = sh_terms[(size_t)q * N + jj]; break; }
        ++depth;
      }
      for (int d = depth; d > 0; --d) {  // poses lying in one cell between that state and pose j, oldest first
        const int r = j - d;
        const int2 c0 = sh_cell[(size_t)r * N];
        if (!(c0.x == cache_x && c0.y == cache_y && cache_p != -1)) { cache_x = c0.x; cache_y = c0.y; cache_p = sh_terms[(size_t)r * N]; }
      }
      const double *t = sh_terms + (size_t)j * N;
      const int2 *c = sh_cell + (size_t)j * N;
      double total = 0;
      for (int i = 0; i < N; ++i) {
        double prob = t[i];
        if (c[i].x == cache_x && c[i].y == cache_y && cache_p != -1) prob = cache_p;
        else { cache_x = c[i].x; cache_y = c[i].y; cache_p = prob; }
        double term = sg::mul(prob, __ldg(a.w + i));
        if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
        total = sg::add(total, term);
      }
      scores[j] = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
      s_ecx[j] = cache_x; s_ecy[j] = cache_y; s_ecp[j] = cache_p;
    } else if ((tid & 31) == 0 && (tid >> 5) < k) {  // one warp per pose: the sums run side by side
      const int j = tid >> 5;
      const double *t = sh_terms + (size_t)j * N;
      double total = 0;
      int i = 0;
      for (; i + 8 <= N; i += 8) {  // loads first, then the chain of adds: the chain never waits on shared memory
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = t[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) total = sg::add(total, v[u]);
      }
      for (; i < N; ++i) total = sg::add(total, t[i]);
      scores[j] = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
    }
    __syncthreads();
    if (tid == 0) {
      if (chain) { s_cx = s_ecx[k - 1]; s_cy = s_ecy[k - 1]; s_cp = s_ecp[k - 1]; }  // every pose of a round is evaluated
      for (int j = 0; j < k; ++j, ++s_logged)
        if (a.log && s_logged < a.log_cap) {
          double *l = a.log + ((size_t)inst * a.log_cap + s_logged) * 4;
          l[0] = cand[j][0]; l[1] = cand[j][1]; l[2] = cand[j][2]; l[3] = scores[j];
        }
      if (first) {  // probability of the initial pose (pose_enumeration_scan_matcher.h:40), then reset()
        hc = HillClimb();
        hc.bx = init[0]; hc.by = init[1]; hc.bt = init[2];
        hc.tr = a.tr; hc.rot = a.rot;
        hc.best = scores[0];
        hc.done = !(0 < a.max_failed);
      } else {
        hc.apply(a.max_failed, cand, scores, k);
      }
      int nk = 0;
      if (!hc.done) {
        nk = hc.next_round(a.max_failed, cand);
        if (nk == 0) hc.done = true;
      }
      s_k = nk;
    }
    first = false;
    __syncthreads();
    if (s_k == 0) break;
  }
  if (tid == 0) {
    out[0] = hc.bx; out[1] = hc.by; out[2] = hc.bt; out[3] = hc.best; out[4] = (double)hc.tested;
    out[5] = (double)s_guard; out[6] = (double)s_logged; out[7] = 0;
    if (chain) { out[8] = s_cx; out[9] = s_cy; out[10] = s_cp; }
  }
}

// ---- a Monte-Carlo match (GaussianPoseEnumerator, monte_carlo_scan_matcher.h:10-82, under the accept loop of
// PoseEnumerationScanMatcher) in one launch.  The enumerator's noise comes from libstdc++'s normal_distribution and
// cannot be produced here, but it does not depend on the scores: candidate k is (best pose so far) + noise[k], and the
// noise sequence only changes when an accept shrinks the dispersion (reset_shift).  The host passes the noise list,
// the block speculates a batch of candidates from the current best pose, scores them (terms in shared memory, one
// warp per pose adds them in point order), thread 0 walks the batch in order with the enumerator's counters, re-bases
// at the first accept, and stops at a dispersion reset (the host then supplies the next noise list) or when the
// enumerator's budget is spent.
struct McArgs {
  MapView map;
  const double *noise;  // 3 per candidate: the pose shifts the enumerator would sample, in order
  int K;                // candidates available
  double bx, by, bt, best;  // best pose so far and its probability (NaN with have_best = 0: score it first)
  int have_best;
  unsigned failed, poses_nm, max_failed, max_poses;  // the enumerator's counters and limits
  int batch;            // candidates per speculation round (shared memory holds batch x N terms)
  const double *range, *angle, *w, *f;
  int N;
  double wsum, win_v, win_h, gm_th;
  int gm_win;
  double *out;   // bx, by, bt, best, consumed, failed, poses_nm, reset (1: dispersion was reset), guard hits, log entries
  double *log;   // {x, y, theta, probability} of every pose scored and consumed, in order
  int log_cap;
};
#define SG_MC_OUT 10

template <int MODE, bool FACTOR>
__global__ void __launch_bounds__(1024) k_monte_carlo(McArgs a) {
  extern __shared__ double sh_terms[];  // [batch][N]
  __shared__ double cand[32][3], scores[32];
  __shared__ double s_bx, s_by, s_bt, s_best;
  __shared__ int s_b, s_k, s_logged, s_stop, s_reset;
  __shared__ unsigned s_failed, s_poses, s_guard;
  const int tid = threadIdx.x, N = a.N;
  const MapView &mv = a.map;
  const double s = mv.scale, inv_s = 1.0 / mv.scale;
  if (tid == 0) {
    s_bx = a.bx; s_by = a.by; s_bt = a.bt; s_best = a.best;
    s_failed = a.failed; s_poses = a.poses_nm; s_guard = 0; s_logged = 0; s_k = 0; s_stop = 0; s_reset = 0;
    if (!a.have_best) { cand[0][0] = a.bx; cand[0][1] = a.by; cand[0][2] = a.bt; s_b = 1; }
  }
  __syncthreads();
  bool init_round = !a.have_best;
  for (;;) {
    if (tid == 0 && !init_round) {
      // has_next(): failed < max_failed && poses_nm < max_poses; the batch = what would be tested if everything failed
      int b = 0;
      if (s_failed < a.max_failed && s_poses < a.max_poses && s_k < a.K) {
        b = min(min(a.batch, a.K - s_k), (int)min(a.max_failed - s_failed, a.max_poses - s_poses));
        for (int j = 0; j < b; ++j) {  // RobotPose + RobotPoseDelta: component-wise sums
          const double *nz = a.noise + 3 * (size_t)(s_k + j);
          cand[j][0] = sg::add(s_bx, nz[0]); cand[j][1] = sg::add(s_by, nz[1]); cand[j][2] = sg::add(s_bt, nz[2]);
        }
      }
      s_b = b;
    }
    __syncthreads();
    const int b = s_b;
    if (b == 0) break;
    bool unsafe_any = false;
    for (int e = tid; e < b * N; e += blockDim.x) {
      const int j = e / N, i = e - j * N;
      const double px = cand[j][0], py = cand[j][1];
      double sn, cs;
      sincos(sg::add(cand[j][2], __ldg(a.angle + i)), &sn, &cs);
      const double r = __ldg(a.range + i);
      const double rc = sg::mul(r, cs), rs = sg::mul(r, sn);
      const double X = sg::add(px, rc), Y = sg::add(py, rs);
      double prob;
      if (MODE == SLAMGPU_OOPE_OBSTACLE || MODE == SLAMGPU_OOPE_GMAPPING) {
        const int cx = grid_cell(X, rc, s, inv_s, 1, &unsafe_any);
        const int cy = grid_cell(Y, rs, s, inv_s, 1, &unsafe_any);
        prob = MODE == SLAMGPU_OOPE_OBSTACLE ? lut_at(mv, cx, cy) : gmapping_probability(mv, cx, cy, X, Y, a.gm_th, a.gm_win);
      } else {
        double hv = sg::div(a.win_v, 2.0), hh = sg::div(a.win_h, 2.0);
        bool u;
        sg::world_to_cell_guard(sg::sub(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::sub(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        prob = window_probability<MODE>(mv, X, Y, a.win_v, a.win_h);
      }
      double term = sg::mul(prob, __ldg(a.w + i));
      if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
      sh_terms[e] = term;
    }
    if (unsafe_any) atomicAdd(&s_guard, 1u);
    __syncthreads();
    if ((tid & 31) == 0 && (tid >> 5) < b) {  // one warp per pose
      const int j = tid >> 5;
      const double *t = sh_terms + (size_t)j * N;
      double total = 0;
      int i = 0;
      for (; i + 8 <= N; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = t[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) total = sg::add(total, v[u]);
      }
      for (; i < N; ++i) total = sg::add(total, t[i]);
      scores[j] = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
    }
    __syncthreads();
    if (tid == 0) {
      auto log_it = [&](int j) {
        if (a.log && s_logged < a.log_cap) {
          double *l = a.log + (size_t)s_logged * 4;
          l[0] = cand[j][0]; l[1] = cand[j][1]; l[2] = cand[j][2]; l[3] = scores[j];
        }
        ++s_logged;
      };
      if (init_round) {
        s_best = scores[0];
        log_it(0);
      } else {
        int used = b;
        for (int j = 0; j < b; ++j) {
          log_it(j);
          const bool ok = s_best < scores[j];
          ++s_poses;  // feedback(): ++_poses_nm
          if (!ok) { ++s_failed; continue; }
          s_best = scores[j]; s_bx = cand[j][0]; s_by = cand[j][1]; s_bt = cand[j][2];
          used = j + 1;  // the candidates after an accept were speculated from the old best pose
          if (!(s_failed <= a.max_failed / 3)) { s_failed = 0; s_reset = 1; s_stop = 1; }  // reset_shift(0.5): counter cleared, new dispersion, new noise
          break;
        }
        s_k += used;
      }
    }
    init_round = false;
    __syncthreads();
    if (s_stop) break;
  }
  if (tid == 0) {
    double *o = a.out;
    o[0] = s_bx; o[1] = s_by; o[2] = s_bt; o[3] = s_best; o[4] = (double)s_k; o[5] = (double)s_failed; o[6] = (double)s_poses;
    o[7] = (double)s_reset; o[8] = (double)s_guard; o[9] = (double)s_logged;
  }
}

// ---- GmappingOccupancyObservationPE's cache carried from pose to pose (gm_cache == 2).  The cache holds the cell
// of the point evaluated last and the probability computed at the last MISS, so a point reuses the value computed
// at the start of the run of consecutive same-cell points it belongs to -- a run that may begin in the pose
// evaluated before (or, when a whole pose lies in one cell, further back).  Every pose finds its entry state by
// walking its predecessors' cell ids backwards; no pose waits for another.
struct ChainArgs {
  const int *pred;                   // per pose: previous pose of its sequence, or -1 - (entry state index)
  const slamgpu_gm_cache *states_in;
  slamgpu_gm_cache *states_out;      // per pose: the cache after it
};

template <bool FACTOR>
__global__ void __launch_bounds__(128) k_pose_sums_chained(PointArgs pa, ChainArgs ch) {
  const ListArgs &a = pa.l;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if (p < a.Ploc) {
    const int N = a.N;
    // ---- entry state: the nearest predecessor whose final run starts inside it fixes the cache; the poses walked
    // over on the way (each a single run from its first to its last point) are then re-applied forwards
    int cache_x = 0, cache_y = 0;
    double cache_p = -1;
    int depth = 0;
    long long q = ch.pred[p];
    bool found = false;
    while (q >= 0) {
      if (N > 0) {
        const int2 *c = pa.cellids + (size_t)q * N;
        int j = N - 1;
        while (j > 0 && c[j - 1].x == c[j].x && c[j - 1].y == c[j].y) --j;
        if (j > 0) { cache_x = c[j].x; cache_y = c[j].y; cache_p = pa.terms[(size_t)q * N + j]; found = true; break; }
        ++depth;
      }
      q = ch.pred[q];
    }
    if (!found && q < 0) {
      const slamgpu_gm_cache st = ch.states_in[-1 - q];
      cache_x = st.cx; cache_y = st.cy; cache_p = st.prob;
    }
    for (int d = depth; d > 0 && N > 0; --d) {  // single-run poses between the state found and pose p, oldest first
      long long r = ch.pred[p];
      for (int k = 1; k < d; ++k) r = ch.pred[r];
      const int2 c0 = pa.cellids[(size_t)r * N];
      if (!(c0.x == cache_x && c0.y == cache_y && cache_p != -1)) { cache_x = c0.x; cache_y = c0.y; cache_p = pa.terms[(size_t)r * N]; }
    }
    // ---- this pose, in point order
    const double *t = pa.terms + (size_t)p * N;
    const int2 *c = pa.cellids + (size_t)p * N;
    double total = 0;
    for (int i = 0; i < N; ++i) {
      double prob = t[i];
      if (c[i].x == cache_x && c[i].y == cache_y && cache_p != -1) prob = cache_p;
      else { cache_x = c[i].x; cache_y = c[i].y; cache_p = prob; }
      double term = sg::mul(prob, __ldg(a.w + i));
      if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
      total = sg::add(total, term);
    }
    double score = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
    a.scores[p] = score;
    slamgpu_gm_cache out;
    out.cx = cache_x; out.cy = cache_y; out.prob = cache_p;
    ch.states_out[p] = out;
    if (score == score) { best_s = score; best_i = a.p0 + p; }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

// ---- the one-shot form of the small-batch path: two launches between ONE H2D and ONE D2H.  Inputs travel in
// one staging buffer {header | poses | views | view ids}, outputs come back as {header | scores}.
struct SmallHdr {
  Result best;             // out: final (score, index, guard)
  unsigned int ticket;     // in: 0
  unsigned int pad0;
  unsigned long long guard;  // in: 0
  double pad1[2];
};

struct SmallArgs {
  MapView map;
  const MapView *views;  // or NULL
  const int *view_id;
  const double *poses;   // 3*P
  const double *range, *angle, *sx, *sy, *w, *f;
  int N, P;
  double wsum, win_v, win_h, gm_th, init_score;
  int gm_win;
  SmallHdr *hdr;
  double *scores;        // P, directly after the header in the output buffer
  double *terms;         // P*N
};

template <int MODE, bool PREROT, bool FACTOR>
__global__ void __launch_bounds__(128) k_small_fused(SmallArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (i < a.N) {
    const double px = a.poses[3 * p], py = a.poses[3 * p + 1];
    const MapView &mv = a.views ? a.views[a.view_id[p]] : a.map;
    const double s = mv.scale, inv_s = 1.0 / mv.scale;
    double X, Y, rc = 0, rs = 0;
    bool unsafe_any = false;
    if (PREROT) {
      X = sg::add(__ldg(a.sx + i), px); Y = sg::add(__ldg(a.sy + i), py);
    } else {
      double sn, cs;
      sincos(sg::add(a.poses[3 * p + 2], __ldg(a.angle + i)), &sn, &cs);
      const double r = __ldg(a.range + i);
      rc = sg::mul(r, cs); rs = sg::mul(r, sn);
      X = sg::add(px, rc); Y = sg::add(py, rs);
    }
    double prob;
    if (MODE == SLAMGPU_OOPE_OBSTACLE || MODE == SLAMGPU_OOPE_GMAPPING) {
      const int cx = grid_cell(X, rc, s, inv_s, PREROT ? 0 : 1, &unsafe_any);
      const int cy = grid_cell(Y, rs, s, inv_s, PREROT ? 0 : 1, &unsafe_any);
      prob = MODE == SLAMGPU_OOPE_OBSTACLE ? lut_at(mv, cx, cy) : gmapping_probability(mv, cx, cy, X, Y, a.gm_th, a.gm_win);
    } else {
      if (!PREROT) {
        double hv = sg::div(a.win_v, 2.0), hh = sg::div(a.win_h, 2.0);
        bool u;
        sg::world_to_cell_guard(sg::sub(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(X, hh), s, trig_slack(rc, X), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::sub(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
        sg::world_to_cell_guard(sg::add(Y, hv), s, trig_slack(rs, Y), &u); unsafe_any |= u;
      }
      prob = window_probability<MODE>(mv, X, Y, a.win_v, a.win_h);
    }
    double term = sg::mul(prob, __ldg(a.w + i));
    if (FACTOR) term = sg::mul(term, __ldg(a.f + i));
    a.terms[(size_t)p * a.N + i] = term;
    if (unsafe_any) atomicAdd(&a.hdr->guard, 1ull);
  }
}

// phase 2 of the one-shot path: each block stages the terms of its G poses in shared memory (coalesced), one
// thread per pose adds them in point order, block arg-max; the last block to finish (ticket) merges the block
// results and applies the accept rule
__global__ void __launch_bounds__(256) k_small_final(SmallArgs a, int G, Best *blk) {
  extern __shared__ double sh_terms[];  // [G][N]
  const int q0 = blockIdx.x * G;
  const int nq = min(G, a.P - q0);
  const double *src = a.terms + (size_t)q0 * a.N;
  for (int e = threadIdx.x; e < nq * a.N; e += blockDim.x) sh_terms[e] = __ldcg(src + e);
  __syncthreads();
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nq) {
    // one pose per WARP (lane 0): neighbouring lanes walking rows N doubles apart would collide on the shared-memory
    // banks (8-way for N = 360); loads are issued eight ahead of the chain of adds
    const int j = threadIdx.x >> 5;
    const double *t = sh_terms + (size_t)j * a.N;
    double total = 0;
    int k = 0;
    for (; k + 8 <= a.N; k += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = t[k + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) total = sg::add(total, v[u]);
    }
    for (; k < a.N; ++k) total = sg::add(total, t[k]);
    const double score = a.wsum == 0 ? NAN : sg::div(total, a.wsum);
    const int q = q0 + j;
    a.scores[q] = score;
    if (score == score) { best_s = score; best_i = q; }
  }
  block_argmax(best_s, best_i, blk + blockIdx.x);
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&a.hdr->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  best_s = -INFINITY; best_i = LLONG_MAX;
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
    const double bs = __ldcg(&blk[k].score);
    const long long bi = __ldcg(&blk[k].idx);
    if (beats(bs, bi, best_s, best_i)) { best_s = bs; best_i = bi; }
  }
  __shared__ Best s_best;
  __syncthreads();
  block_argmax(best_s, best_i, &s_best);
  __syncthreads();
  if (threadIdx.x == 0) {
    double sc = s_best.score;
    long long id = s_best.idx;
    if (!(a.init_score < sc) || id == LLONG_MAX) { sc = a.init_score; id = -1; }
    a.hdr->best.score = sc; a.hdr->best.idx = id;
    a.hdr->best.guard = (long long)a.hdr->guard;
    a.hdr->best.pad = 0;
  }
}

template <int MODE>
void launch_small_m(slamgpu_ctx *ctx, const SmallArgs &a, dim3 grd, bool prerot, bool factor) {
  if (prerot) {
    if (factor) k_small_fused<MODE, true, true><<<grd, 128, 0, ctx->stream>>>(a);
    else k_small_fused<MODE, true, false><<<grd, 128, 0, ctx->stream>>>(a);
  } else {
    if (factor) k_small_fused<MODE, false, true><<<grd, 128, 0, ctx->stream>>>(a);
    else k_small_fused<MODE, false, false><<<grd, 128, 0, ctx->stream>>>(a);
  }
}

// ------------------------------------------------------------------ grid (brute force) path
struct GridIdxArgs {
  const double *trc, *trs;  // [t*N + i]
  const double *xs, *ys;
  int nx, ny, nyp, N, t_lo, nt_loc;
  const double *thetas, *range, *angle;  // device trig: r*cos(theta + a), r*sin(theta + a) computed in the kernel (NULL: trc / trs hold a host table)
  int w, h, ox, oy, pitch;
  double scale;
  int guard;
  int *cxp;  // [(tl*N + i)*nx + j]   padded LUT column
  int *cyp;  // [(tl*N + i)*nyp + k]  padded LUT row * pitch            (v1 kernel)
  unsigned long long *cyw;  // [(tl*N + i)*ngy + gy] packed rows of one y-group (v2 kernel)
  int v2, R, ngy;
  int ngys;      // v4 / v5: row stride of cyw (ngy rounded up to even: 16-byte copies of a pair of row words stay aligned)
  int v4, dmax;  // v4 row word: low 32 bits = element offset of the first y's padded LUT row, bits 32..38 = "new row" mask
  // window OOPEs (max / mean) out of the map's window LUTs: the x table holds (first column of the window) + vx * T, the y
  // table (first row) * pitch + vy * 2T, vx / vy = cells spanned beyond the minimum; see sg_map_ensure_wlut
  int win, wl_nxmin, wl_nymin, wl_pitch;
  double win_hh, win_hv;   // half window sides
  long long wl_T;
  int v5, nb, bw;   // v5: column records per (theta, beam, band of bw columns) instead of the cxp table
  uint4 *colrec;    // [(tl*N + i)*nb + band][2]: {first column (even), 0, 0, 0} {32 nibbles: column of lane - first column}
  // v3 (TMA-staged patches): per (theta, beam, block-of-theta) patch origin {first padded column, first padded row}
  int v3, nbt, box_w, box_h;
  const int2 *blk_rows;  // [nbt] y range {k_lo, k_hi} (inclusive) each block of a theta touches
  int2 *porg;            // [(tl*N + i)*nbt + b]
  Result *result;
};

// v2/v3 row word: low 32 bits = padded LUT row of the group's first y, then one signed nibble per
// further y = (cell row of y_m) - (cell row of y_{m-1}); 0 means "same cell row: reuse the gathered value"
SG_DEV int nib_sext(unsigned long long w, int m) {  // m in 1..7
  int v = (int)((w >> (32 + 4 * (m - 1))) & 15ull);
  return (v ^ 8) - 8;
}

// one thread per (theta, beam, axis value): the exact world_to_cell of the reference,
// regular_squares_grid.h:40-46, applied to x_j + r*cos(theta+a) and y_k + r*sin(theta+a)
#define SG_IDX_PAIRS 8     // beams per block of the index kernel
#define SG_IDX_THREADS 128

// cell of one axis value for the index tables: floor(RN(v / scale)) like the reference, computed as
// floor(v * (1/scale)) and verified against both cell borders with exact FMA residuals; only a value
// within `margin` of a border (a few ulps, or the device-trig slack when the guard is on) takes the
// real division, and with the guard on such a value also flags the call for a libm-trig redo.
SG_DEV int grid_axis_cell(double v, double rc, double scale, double inv_scale, int guard, bool *unsafe) {
  const double qa = v * inv_scale;
  int c = __double2int_rd(qa);  // saturating
  const double f = (double)c;
  // qa is within |qa| * 2.3e-16 of the exact quotient (two roundings), i.e. within 2.3e-7 cells below 1e9 cells; the
  // device-trig slack is ~1e-13 cells.  A fractional part farther than 1e-6 from both ends therefore fixes the cell
  // for the reference's floor(v / scale) and for the guard at once: no residuals needed (the common case by far)
  const double da = qa - f;
  if (da > 1e-6 && da < 1.0 - 1e-6 && fabs(qa) < 1e9) return c;
  const double lo = fma(-f, scale, v);        // v - f*scale
  const double hi = fma(f + 1.0, scale, -v);  // (f+1)*scale - v
  double margin = fabs(v) * 8.9e-16 + 1e-300;
  if (guard) margin += trig_slack(rc, v);
  if (!(lo >= margin && hi >= margin)) {
    c = cell_of(floor(sg::div(v, scale)));
    if (guard) {
      const double q = v * inv_scale, fl = floor(q);
      const double tol = trig_slack(rc, v) * inv_scale + 8.0 * 2.220446049250313e-16 * fabs(q);
      if ((q - fl) <= tol || ((fl + 1.0) - q) <= tol) *unsafe = true;
    }
  }
  return c;
}

// MODE: bit 0 window tables, 1 packed row words (v2+), 2 TMA block patches (v3), 3 row de-duplication (v4 / v5), 4 column
// records (v5) -- compile-time so that the per-item loops carry no mode tests
template <int MODE>
__global__ void __launch_bounds__(SG_IDX_THREADS) k_grid_indices(GridIdxArgs a) {
  constexpr bool M_WIN = (MODE & 1) != 0, M_V2 = (MODE & 2) != 0, M_V3 = (MODE & 4) != 0, M_V4 = (MODE & 8) != 0, M_V5 = (MODE & 16) != 0;
  // grid = (ceil(N / SG_IDX_PAIRS), nt_loc): each block handles SG_IDX_PAIRS beams of one theta; a thread
  // owns one axis value (an x or a y) and walks the beams, so the axis value is loaded once
  extern __shared__ int sh_rows[];  // [SG_IDX_PAIRS][ny] padded LUT rows (v2 packing)
  __shared__ double sh_rc[SG_IDX_PAIRS], sh_rs[SG_IDX_PAIRS];
  __shared__ int sh_xmin[SG_IDX_PAIRS], sh_xmax[SG_IDX_PAIRS];
  if (threadIdx.x < SG_IDX_PAIRS) { sh_xmin[threadIdx.x] = INT_MAX; sh_xmax[threadIdx.x] = INT_MIN; }
  const int tl = blockIdx.y;
  const int i0 = blockIdx.x * SG_IDX_PAIRS;
  const int np = min(SG_IDX_PAIRS, a.N - i0);
  const double inv_scale = 1.0 / a.scale;
  if (threadIdx.x < np) {
    if (a.thetas) {  // device trig: what k_trig_table computes, here (one launch and one table round trip less)
      double sn, cs;
      sincos(sg::add(a.thetas[a.t_lo + tl], a.angle[i0 + threadIdx.x]), &sn, &cs);
      sh_rc[threadIdx.x] = sg::mul(a.range[i0 + threadIdx.x], cs);
      sh_rs[threadIdx.x] = sg::mul(a.range[i0 + threadIdx.x], sn);
    } else {         // libm table from the host
      const long long src = (long long)(a.t_lo + tl) * a.N + i0 + threadIdx.x;
      sh_rc[threadIdx.x] = a.trc[src];
      sh_rs[threadIdx.x] = a.trs[src];
    }
  }
  __syncthreads();
  bool unsafe_any = false;
  const long long ti0 = (long long)tl * a.N + i0;
  const int ny_items = M_V2 ? a.ny : a.nyp;
  for (int c = threadIdx.x; c < a.nx + ny_items; c += blockDim.x) {
    if (c < a.nx) {
      const double x = a.xs[c];
      int *dst = a.cxp + ti0 * a.nx + c;
      for (int pr = 0; pr < np; ++pr) {
        const double rc = sh_rc[pr];
        if (M_WIN) {
          // window_probability (dev_window.cuh): lx = floor((X - hh) / s), rx = floor((X + hh) / s), x outer loop
          const double X = sg::add(x, rc);
          const int lx = grid_axis_cell(sg::sub(X, a.win_hh), rc, a.scale, inv_scale, a.guard, &unsafe_any);
          const int rx = grid_axis_cell(sg::add(X, a.win_hh), rc, a.scale, inv_scale, a.guard, &unsafe_any);
          int vx = rx - lx + 1 - a.wl_nxmin;
          if (vx < 0 || vx > 1) { atomicAdd((unsigned long long *)&a.result->pad, 1ull); vx = 0; }
          const int K = a.wl_nxmin + 1;
          dst[(size_t)pr * a.nx] = clampi(lx + a.ox, -K, a.w) + K + vx * (int)a.wl_T;
          continue;
        }
        const int cx = grid_axis_cell(sg::add(x, rc), rc, a.scale, inv_scale, a.guard, &unsafe_any);
        const int col = clampi(cx + a.ox, -1, a.w) + SG_LUT_PAD;
        if (M_V5) sh_rows[SG_IDX_PAIRS * a.ny + pr * a.nx + c] = col;
        else dst[(size_t)pr * a.nx] = col;
        if (M_V3) { atomicMin(&sh_xmin[pr], col); atomicMax(&sh_xmax[pr], col); }
      }
    } else {
      const int k = c - a.nx;
      const double y = k < a.ny ? a.ys[k] : 0.0;
      for (int pr = 0; pr < np; ++pr) {
        int prow = 0;  // v1 padding rows point at the ring row 0 (always valid memory)
        if (M_WIN) {
          int enc = 0;
          if (k < a.ny) {
            const double rs = sh_rs[pr];
            const double Y = sg::add(y, rs);
            const int ly = grid_axis_cell(sg::sub(Y, a.win_hv), rs, a.scale, inv_scale, a.guard, &unsafe_any);
            const int ry = grid_axis_cell(sg::add(Y, a.win_hv), rs, a.scale, inv_scale, a.guard, &unsafe_any);
            int vy = ry - ly + 1 - a.wl_nymin;
            if (vy < 0 || vy > 1) { atomicAdd((unsigned long long *)&a.result->pad, 1ull); vy = 0; }
            const int K = a.wl_nymin + 1;
            enc = (clampi(ly + a.oy, -K, a.h) + K) * a.wl_pitch + vy * 2 * (int)a.wl_T;
          }
          a.cyp[(ti0 + pr) * a.nyp + k] = enc;
          continue;
        }
        if (k < a.ny) {
          const double rs = sh_rs[pr];
          const int cy = grid_axis_cell(sg::add(y, rs), rs, a.scale, inv_scale, a.guard, &unsafe_any);
          prow = clampi(cy + a.oy, -1, a.h) + SG_LUT_PAD;
        }
        if (M_V2) sh_rows[pr * a.ny + k] = prow;
        else a.cyp[(ti0 + pr) * a.nyp + k] = prow * a.pitch;
      }
    }
  }
  if (M_V2) {
    __syncthreads();
    for (int e = threadIdx.x; e < np * a.ngy; e += blockDim.x) {
      const int pr = e / a.ngy, gy = e - pr * a.ngy;
      const int *rows = sh_rows + pr * a.ny;
      const int k0 = gy * a.R;
      if (M_V4) {
        // k_score_grid4: rows of the group must be base, base + 1, ... (steps of 0 or 1, fewer than dmax distinct ones)
        unsigned mask = 0;
        int prev4 = rows[k0], distinct = 1;
        bool ok = true;
        for (int m = 1; m < 8; ++m) {
          const int k = k0 + m;
          const int prow = k < a.ny ? rows[k] : prev4;
          const int d = prow - prev4;
          if (d == 1) { mask |= 1u << (m - 1); ++distinct; }
          else if (d != 0) ok = false;
          prev4 = prow;
        }
        if (!ok || distinct > a.dmax) atomicAdd((unsigned long long *)&a.result->pad, 1ull);
        a.cyw[(ti0 + pr) * a.ngys + gy] = ((unsigned long long)mask << 32) | (unsigned)(rows[k0] * a.pitch);
        continue;
      }
      unsigned long long word = (unsigned)rows[k0];
      int prev = rows[k0];
      for (int m = 1; m < a.R; ++m) {
        const int k = k0 + m;
        const int prow = k < a.ny ? rows[k] : prev;  // rows past the end repeat the last one (delta 0)
        int d = prow - prev;
        if (d < -8 || d > 7) { atomicAdd((unsigned long long *)&a.result->pad, 1ull); d = 0; }
        word |= (unsigned long long)(d & 15) << (32 + 4 * (m - 1));
        prev = prow;
      }
      a.cyw[(ti0 + pr) * a.ngy + gy] = word;
    }
    if (M_V5) {
      // k_score_grid5: one record per band of <= 30 consecutive x; a band's 14-column window must hold all its columns
      const int *cols_all = sh_rows + SG_IDX_PAIRS * a.ny;
      for (int e = threadIdx.x; e < np * a.nb; e += blockDim.x) {
        const int pr = e / a.nb, b = e - pr * a.nb;
        const int *cols = cols_all + pr * a.nx;
        const int j0 = b * a.bw, j1 = min(j0 + a.bw, a.nx);
        const int col0 = cols[j0] & ~1;
        unsigned nib[4] = {0u, 0u, 0u, 0u};
        bool ok = true;
        for (int jj = j0; jj < j1; ++jj) {
          const int d = cols[jj] - col0;
          if (d < 0 || d > 13) { ok = false; continue; }
          const int l = jj - j0;
          nib[l >> 3] |= (unsigned)d << ((l & 7) * 4);
        }
        if (!ok) atomicAdd((unsigned long long *)&a.result->pad, 1ull);
        uint4 *dst = a.colrec + ((ti0 + pr) * a.nb + b) * 2;
        dst[0] = make_uint4((unsigned)col0, 0u, 0u, 0u);
        dst[1] = make_uint4(nib[0], nib[1], nib[2], nib[3]);
      }
    }
    if (M_V3) {
      for (int e = threadIdx.x; e < np * a.nbt; e += blockDim.x) {
        const int pr = e / a.nbt, b = e - pr * a.nbt;
        const int *rows = sh_rows + pr * a.ny;
        const int2 kr = a.blk_rows[(size_t)tl * a.nbt + b];
        int ymin = INT_MAX, ymax = INT_MIN;
        for (int k = kr.x; k <= kr.y && k < a.ny; ++k) { ymin = min(ymin, rows[k]); ymax = max(ymax, rows[k]); }
        // the box must hold every cell this block can touch, else the call is redone with the v2 kernel
        const int x0 = sh_xmin[pr] & ~1;  // 16-byte aligned source rows for the bulk copies
        if (ymax - ymin + 1 > a.box_h || sh_xmax[pr] - x0 + 1 > a.box_w) atomicAdd((unsigned long long *)&a.result->pad, 1ull);
        a.porg[(ti0 + pr) * a.nbt + b] = make_int2(x0, ymin);
      }
    }
  }
  if (unsafe_any) atomicAdd((unsigned long long *)&a.result->guard, 1ull);
}

struct GridArgs {
  const double *lut;
  const int *cxp, *cyp;
  const int4 *groups;  // {t, k0, m_lo, m_hi}: rows k0+m for m in [m_lo, m_hi) belong to this rank
  int n_groups, nx, ny, nyp, N, t_lo;
  const double *w, *f;
  double wsum;
  long long p0;
  double *scores;  // local slice
  Best *blk;
};

#define SG_GRID_R 8

template <bool FACTOR>
__global__ void __launch_bounds__(128) k_score_grid(GridArgs a) {
  long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int g = (int)(q0 / a.nx);
  int j = (int)(q0 - (long long)g * a.nx);
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if (g < a.n_groups) {
    const int4 grp = __ldg(a.groups + g);
    const int tl = grp.x - a.t_lo;
    const int *pcx = a.cxp + (size_t)tl * a.N * a.nx + j;
    const int4 *pcy = reinterpret_cast<const int4 *>(a.cyp + (size_t)tl * a.N * a.nyp + grp.y);
    const int cy_step = a.nyp / 4;
    double acc[SG_GRID_R];
#pragma unroll
    for (int m = 0; m < SG_GRID_R; ++m) acc[m] = 0.0;
#pragma unroll 2
    for (int i = 0; i < a.N; ++i) {
      const int cx = __ldg(pcx);
      const int4 r0 = __ldg(pcy), r1 = __ldg(pcy + 1);
      const double wi = __ldg(a.w + i);
      const double fi = FACTOR ? __ldg(a.f + i) : 1.0;
      const int row[SG_GRID_R] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      double v[SG_GRID_R];
#pragma unroll
      for (int m = 0; m < SG_GRID_R; ++m) v[m] = __ldg(a.lut + (row[m] + cx));
#pragma unroll
      for (int m = 0; m < SG_GRID_R; ++m) {
        double term = sg::mul(v[m], wi);
        if (FACTOR) term = sg::mul(term, fi);
        acc[m] = sg::add(acc[m], term);
      }
      pcx += a.nx;
      pcy += cy_step;
    }
#pragma unroll
    for (int m = 0; m < SG_GRID_R; ++m) {
      if (m < grp.z || m >= grp.w) continue;
      double score = a.wsum == 0 ? NAN : sg::div(acc[m], a.wsum);
      long long idx = ((long long)grp.x * a.ny + (grp.y + m)) * a.nx + j;
      a.scores[idx - a.p0] = score;
      if (score == score && beats(score, idx, best_s, best_i)) { best_s = score; best_i = idx; }
    }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

// v2 of the grid kernel.  Same arithmetic and summation order; what changes is the data movement:
//  * the R consecutive y of a thread usually fall into ~R*step/scale+1 distinct cell rows; the packed row
//    word says (warp-uniformly) where the row changes, so a repeated row re-uses the value already in a
//    register instead of gathering (and multiplying) again;
//  * the row table shrinks to one 64-bit uniform load per beam (was two LDG.128);
//  * the per-beam index loads run PF beams ahead of the gathers that depend on them;
//  * equal point weights (EvenSPW) come from a kernel parameter instead of a load.
struct GridArgs2 {
  const double *lut;
  const int *cxp;
  const unsigned long long *cyw;
  const int4 *groups;  // {t, k0, m_lo, m_hi}
  int n_groups, nx, ny, ngy, N, t_lo, pitch;
  const double *w, *f;
  double w0, wsum;
  long long p0;
  double *scores;
  Best *blk;
};

template <int R, bool FACTOR, bool UNIW>
__global__ void __launch_bounds__(128) k_score_grid2(GridArgs2 a) {
  long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int g = (int)(q0 / a.nx);
  int j = (int)(q0 - (long long)g * a.nx);
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  if (g < a.n_groups) {
    const int4 grp = __ldg(a.groups + g);
    const int tl = grp.x - a.t_lo;
    const int *pcx = a.cxp + (size_t)tl * a.N * a.nx + j;
    const unsigned long long *pcw = a.cyw + (size_t)tl * a.N * a.ngy + grp.y / R;
    double acc[R];
#pragma unroll
    for (int m = 0; m < R; ++m) acc[m] = 0.0;
    // software pipeline: indices run two beams ahead, gathers one beam ahead of the adds
    int cx1 = 0, cx2 = 0;
    unsigned long long cw1 = 0, cw2 = 0;
    double v[R];
    {
      int cx0 = a.N > 0 ? __ldg(pcx) : 0;
      unsigned long long cw0 = a.N > 0 ? __ldg(pcw) : 0ull;
      if (a.N > 1) { cx1 = __ldg(pcx + a.nx); cw1 = __ldg(pcw + a.ngy); }
      if (a.N > 2) { cx2 = __ldg(pcx + 2 * (size_t)a.nx); cw2 = __ldg(pcw + 2 * (size_t)a.ngy); }
      int off = (int)(unsigned)(cw0 & 0xffffffffull) * a.pitch + cx0;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m > 0) off += nib_sext(cw0, m) * a.pitch;
        v[m] = a.N > 0 ? __ldg(a.lut + off) : 0.0;
      }
    }
    for (int i = 0; i < a.N; ++i) {
      // gathers of beam i+1 (indices already in registers)
      double vn[R];
      {
        int off = (int)(unsigned)(cw1 & 0xffffffffull) * a.pitch + cx1;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          if (m > 0) off += nib_sext(cw1, m) * a.pitch;
          vn[m] = (i + 1 < a.N) ? __ldg(a.lut + off) : 0.0;
        }
      }
      cx1 = cx2; cw1 = cw2;
      if (i + 3 < a.N) { cx2 = __ldg(pcx + (size_t)(i + 3) * a.nx); cw2 = __ldg(pcw + (size_t)(i + 3) * a.ngy); }
      const double wi = UNIW ? a.w0 : __ldg(a.w + i);
      const double fi = FACTOR ? __ldg(a.f + i) : 1.0;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        double term = sg::mul(v[m], wi);
        if (FACTOR) term = sg::mul(term, fi);
        acc[m] = sg::add(acc[m], term);
        v[m] = vn[m];
      }
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < grp.z || m >= grp.w) continue;
      double score = a.wsum == 0 ? NAN : sg::div(acc[m], a.wsum);
      long long idx = ((long long)grp.x * a.ny + (grp.y + m)) * a.nx + j;
      a.scores[idx - a.p0] = score;
      if (score == score && beats(score, idx, best_s, best_i)) { best_s = score; best_i = idx; }
    }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}


// v4 of the grid kernel: each distinct map row is gathered once per thread.
// The 8 consecutive y of a thread (0.02 m apart on 0.05 m cells in configs[2]) fall into 3-4 consecutive cell rows, so
// v2 gathers, multiplies and address-computes the same LUT entry two or three times.  Here the thread loads the DMAX
// consecutive rows base..base+DMAX-1 of its column once (one IMAD.WIDE + LDG each, row base pointers in uniform
// registers), multiplies each by the point weight once, and a 7-bit "new row" mask from the index kernel says which
// product every accumulator takes.  The mask is the same for a whole warp (lanes = 32 consecutive x of one y-group), so
// it selects one of 128 straight-line bodies of 8 DADDs through an indexed branch (brx.idx; grid4_dispatch.inc,
// generated by tools/gen_grid4_dispatch.py -- NVVM lowers a C++ switch to a 12-instruction compare tree instead).
// The adds are the same FP64 adds in the same (beam) order as in every other variant: scores are bit-identical.
// Columns are split into bands of 32 so that full bands give warp-uniform masks; the nx % 32 leftover columns are
// packed several y-groups to a warp (the branch diverges there, which is correct, only slower).
struct GridArgs4 {
  const double *lut;
  const unsigned *cxp;
  const uint2 *cyw;
  const int4 *groups;  // {t, k0, m_lo, m_hi}
  int n_groups, nx, ny, ngy, ngys, N, t_lo, pitch;
  int nb_full, wr;         // full 32-column bands per y-group, width of the leftover band
  const int2 *wtask;       // warp table (see grid4_task)
  int n_warps;
  const double *w;
  double w0, wsum;
  long long p0;
  double *scores;
  Best *blk;
  int *sm_cnt;  // per-chunk task counters, zeroed before the launch (NULL: task = blockIdx.x)
  int n_chunks, chunk;  // chunks (= SMs), tasks per chunk
  unsigned zero;        // 0 (a run-time value: see the beam loop)
};

SG_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
SG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
SG_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }

template <class T>
SG_DEV T *opaque_ptr(T *p, unsigned zero) {
  unsigned long long v = (unsigned long long)p;
  asm volatile("{ .reg .u64 z; cvt.u64.u32 z, %1; add.u64 %0, %0, z; }" : "+l"(v) : "r"(zero));
  return (T *)v;
}

template <int DMAX>
SG_DEV void add_pattern(unsigned mask, double (&acc)[8], const double (&t)[DMAX]) {
  // ranks past DMAX-1 never occur (k_grid_indices counts such groups and the call falls back to v2): any register will do
#define SG_G4T(d) t[(d) < DMAX ? (d) : DMAX - 1]
  asm(
#include "grid4_dispatch.inc"
      : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4]), "+d"(acc[5]), "+d"(acc[6]), "+d"(acc[7])
      : "d"(SG_G4T(0)), "d"(SG_G4T(1)), "d"(SG_G4T(2)), "d"(SG_G4T(3)), "d"(SG_G4T(4)), "d"(SG_G4T(5)), "d"(SG_G4T(6)),
        "d"(SG_G4T(7)), "r"(mask));
#undef SG_G4T
}

// thread -> (y-group, column) through the warp table: warp w of the launch (4 per task) takes wtask[w] = {first y-group, band}
// -- a full band of 32 columns of one y-group -- or {first y-group, -count}: the nx % 32 leftover columns of `count`
// consecutive y-groups packed into one warp.  false: no work, (0, 0) returned: a thread without work walks group 0 / column 0
// and drops the result, so that the beam loop sits in uniform control flow.
SG_DEV bool grid4_task(const GridArgs4 &a, int task, int &g, int &j) {
  const int w = task * 4 + (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  bool valid = false;
  g = 0; j = 0;
  if (task >= 0 && w < a.n_warps) {
    const int2 e = __ldg(a.wtask + w);
    if (e.y >= 0) {
      g = e.x; j = e.y * 32 + lane; valid = true;
    } else if (lane < -e.y * a.wr) {
      const int k = lane / a.wr;
      g = e.x + k; j = a.nb_full * 32 + (lane - k * a.wr); valid = true;
    }
  }
  return valid;
}

template <int DMAX, bool UNIW>
__global__ void __launch_bounds__(128, DMAX <= 4 ? 9 : (DMAX <= 6 ? 5 : 4)) k_score_grid4(GridArgs4 a) {
  // configs[2]'s 1030 blocks (7 per SM) must run as ONE wave -- 56 registers leave room for 9 --: every task (a warp walking all the
  // beams) takes the same time, so a second, nearly empty wave would double the kernel time
  // Which 128 candidates-columns this block takes.  Blocks are handed out per SM: the work is cut into one contiguous chunk
  // per SM (neighbouring y-groups of the same theta read the same LUT lines for a given beam), so that the warps resident on
  // an SM share their gathers in L1.  With blocks dealt round-robin (task = blockIdx.x) every SM holds ~7 different thetas,
  // every LUT line of a beam's patch is requested by most SMs at about the same time (all warps walk the beams in step) and
  // the L2 slices serve those bursts one sector at a time: ncu showed ~1000 cycles from gather to first use.  A block that
  // finds its SM's chunk empty takes a leftover task of another chunk (every task is taken exactly once).
  __shared__ int s_task;
  if (threadIdx.x == 0) {
    int task = -1;
    if (a.sm_cnt) {
      unsigned smid;
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      const int nsm = a.n_chunks, per = a.chunk;
      int s0 = (int)(smid % (unsigned)nsm);
      for (int k = 0; k < nsm && task < 0; ++k) {
        const int s2 = s0 + k < nsm ? s0 + k : s0 + k - nsm;
        if (*(volatile int *)(a.sm_cnt + s2) >= per) continue;
        const int c2 = atomicAdd(a.sm_cnt + s2, 1);
        if (c2 < per && s2 * per + c2 < (int)gridDim.x) task = s2 * per + c2;
      }
    } else {
      task = blockIdx.x;
    }
    s_task = task;
  }
  __syncthreads();
  int g, j;
  bool valid = grid4_task(a, *(volatile int *)&s_task, g, j);
  // element offsets into the index tables (the host checks that they fit 32 bits): one register each instead of a pointer pair
  unsigned ox, ow;
  {
    const int4 grp = __ldg(a.groups + g);
    const int tl = grp.x - a.t_lo;
    ox = (unsigned)tl * (unsigned)a.N * (unsigned)a.nx + (unsigned)j;
    ow = (unsigned)tl * (unsigned)a.N * (unsigned)a.ngys + (unsigned)(grp.y >> 3);
  }
  // Loop invariants are made opaque to ptxas with a real instruction (an add of a run-time zero): otherwise it re-derives
  // them from the constant bank every beam (LDC + IMAD.WIDE), and those LDCs occupy the scoreboards the loads need.
  const unsigned zero = a.zero;  // 0, but not to the compiler
  const double *rowp[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) rowp[d] = opaque_ptr(a.lut + (size_t)d * (size_t)a.pitch, zero);
  const unsigned *const cxp = opaque_ptr(a.cxp, zero);
  const uint2 *const cyw = opaque_ptr(a.cyw, zero);
  const unsigned sx = a.nx, sw = a.ngys;
  const int N = a.N;
  const double w0 = a.w0;
  double acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0;
  // ptxas puts EVERY global load of the loop on one scoreboard (SB5 in all our kernels), so the first use of any loaded
  // register waits for all loads in flight: the loads cannot run more than one beam ahead of their use whatever the source
  // says.  An iteration therefore does FIRST the products of its beam (the one place where the warp waits), THEN issues the
  // gathers of the next beam -- their address is made to depend on a product so that the scheduler cannot hoist them above
  // the wait --, reloads the index registers in place for the beam after that, and only then runs the adds, which overlap
  // the loads' flight.  One value set and one index set are enough for that order (a set is dead once its products / its
  // address are taken), which keeps the kernel at 56 registers (two sets: 72) and the loop short.
  // (Measured alternatives: gathers issued before the wait 0.67 ms; index pairs through a cp.async ring 0.91 ms -- LDGSTS
  // costs ~1 cycle per lane here; two beams per wait 0.59 ms; this order with two register sets, 72 registers 0.517 ms; one set 0.491 ms.)
  double v[DMAX];
  unsigned mcur;
  {
    const unsigned cx0 = __ldcg(cxp + ox);
    const uint2 cw0 = __ldcg(cyw + ow);
    mcur = cw0.y;
    const unsigned b = cw0.x + cx0;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) v[d] = __ldg(rowp[d] + b);
  }
  // the index tables are streamed (L2 only: they would only push LUT lines out of L1)
  unsigned cx = __ldcg(cxp + (ox + sx));
  uint2 cw = __ldcg(cyw + (ow + sw));
  ox += 2 * sx; ow += 2 * sw;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const double wi = UNIW ? w0 : __ldg(a.w + i);
    double t[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; ++d) t[d] = sg::mul(v[d], wi);
    const unsigned b = cw.x + cx + ((unsigned)__double2hiint(t[0]) & zero);
    const unsigned mnext = cw.y;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) v[d] = __ldg(rowp[d] + b);
    cx = __ldcg(cxp + ox); cw = __ldcg(cyw + ow);
    ox += sx; ow += sw;
    add_pattern<DMAX>(mcur, acc, t);
    mcur = mnext;
  }
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  // (group, column) are derived again from the task number rather than kept in registers across the beam loop
  valid = grid4_task(a, *(volatile int *)&s_task, g, j);
  if (valid) {
    const int4 grp = __ldg(a.groups + g);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < grp.z || m >= grp.w) continue;
      double score = a.wsum == 0 ? NAN : sg::div(acc[m], a.wsum);
      long long idx = ((long long)grp.x * a.ny + (grp.y + m)) * a.nx + j;
      a.scores[idx - a.p0] = score;
      if (score == score && beats(score, idx, best_s, best_i)) { best_s = score; best_i = idx; }
    }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

template <int DMAX>
void launch_grid4(slamgpu_ctx *ctx, const GridArgs4 &a, int nblk, bool uniw) {
  if (uniw) k_score_grid4<DMAX, true><<<nblk, 128, 0, ctx->stream>>>(a);
  else k_score_grid4<DMAX, false><<<nblk, 128, 0, ctx->stream>>>(a);
}


// v5 of the grid kernel: v4's arithmetic (each distinct map row once, indexed-branch accumulate) behind an asynchronous
// copy pipeline.  ptxas puts EVERY global load of a loop on one scoreboard (SB5 in all our kernels), so the first use of any
// loaded register waits for all loads in flight: register prefetching cannot run more than one beam ahead, and a warp pays a
// full L2 round trip per beam (ncu, lone warp: ~700 cycles per beam).  cp.async (LDGSTS) completes per commit group
// instead, so here each warp issues ONE 16-byte-per-lane copy per beam, SG_G5_P beams ahead of its use:
//   lanes 0 .. 7*DMAX-1   the LUT patch of beam i+P: DMAX consecutive rows x 14 columns (7 aligned pairs) holding every
//                         cell the warp's <= 30 consecutive x can touch in its 8 y;
//   lanes 28, 29          the column record of beam i+P+Q {first column, 32 nibbles};
//   lane 30               the row word pair of beam i+P+Q {row offset, new-row mask};
//   lane 31               the point weight pair of beam i+P (uneven weights only);
// into a ring of SG_G5_STAGES 512-byte stages private to the warp (no block barrier anywhere).  Every value the loop
// consumes comes from shared memory (LDS, ~30 cycles); cp.async.wait_group leaves the youngest copies in flight.
struct GridArgs5 {
  const double *lut;
  const uint4 *colrec;
  const uint2 *cyw;
  const double2 *w2;   // {w[i], w[i+1]} per beam (uneven weights)
  const int4 *groups;  // {t, k0, m_lo, m_hi}
  int n_groups, nx, ny, ngy, ngys, N, t_lo, pitch, nb, bw;
  double w0, wsum;
  long long p0;
  double *scores;
  Best *blk;
  unsigned zero;
};

SG_DEV void cp_async16(unsigned dst, const void *src, bool on) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 16; }" ::"r"(dst), "l"(src), "r"((unsigned)on) : "memory");
}
SG_DEV unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
SG_DEV uint2 lds64(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
SG_DEV double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

#define SG_G5_STAGE 512u  // bytes per stage: 32 lanes x 16
#define SG_G5_ROW 112u    // bytes per patch row in a stage: 7 pairs
#define SG_G5_STAGES 16   // ring depth (power of two, > SG_G5_P + SG_G5_Q)
#define SG_G5_P 8         // the patch (and weight) of a beam is copied this many beams ahead of its use
#define SG_G5_Q 7         // ... and its index records this many beams ahead of the patch copy that needs them

template <int DMAX, bool UNIW>
__global__ void __launch_bounds__(64, 13) k_score_grid5(GridArgs5 a) {
  __shared__ __align__(16) unsigned char ring[2][SG_G5_STAGES * SG_G5_STAGE];
  const int lane = (int)(threadIdx.x & 31), wib = (int)(threadIdx.x >> 5);
  int w = (int)blockIdx.x * 2 + wib;
  const bool wvalid = w < a.n_groups * a.nb;
  if (!wvalid) w = 0;  // walks task 0 and drops the result: the beam loop stays in uniform control flow
  const int g = w / a.nb, band = w - g * a.nb;
  int tl, gy;
  {
    const int4 grp = __ldg(a.groups + g);
    tl = grp.x - a.t_lo; gy = grp.y >> 3;
  }
  // what this lane copies every beam: src = base + 8 * u, u = (row offset + first column) for the patch lanes, a running
  // table offset for the others
  const bool is_patch = lane < 7 * DMAX;
  const bool is_idx = lane >= 28 && lane <= 30, is_w = lane == 31 && !UNIW;
  const char *base;
  unsigned uidx = 0, inc = 0;
  {
    const int r = lane / 7, cp = lane - r * 7;
    if (is_patch) base = reinterpret_cast<const char *>(a.lut + (size_t)r * (size_t)a.pitch + 2 * cp);
    else if (lane == 28 || lane == 29) {
      base = reinterpret_cast<const char *>(a.colrec) + (lane - 28) * 16;
      uidx = ((unsigned)tl * (unsigned)a.N * (unsigned)a.nb + (unsigned)band) * 4u; inc = (unsigned)a.nb * 4u;
    } else if (lane == 30) {
      base = reinterpret_cast<const char *>(a.cyw);
      uidx = (unsigned)tl * (unsigned)a.N * (unsigned)a.ngys + (unsigned)(gy & ~1); inc = (unsigned)a.ngys;
    } else if (is_w) {
      base = reinterpret_cast<const char *>(a.w2); inc = 2u;
    } else base = reinterpret_cast<const char *>(a.lut);  // idle lane: copies the first LUT pair
  }
  base = opaque_ptr(base, a.zero);
  const unsigned ring_w = smem_u32(&ring[wib][0]);
  const unsigned my_chunk = ring_w + (unsigned)lane * 16u;
  const unsigned nib_addr = ring_w + 29u * 16u + (unsigned)(lane >> 3) * 4u, nib_shift = (unsigned)(lane & 7) * 4u;
  const unsigned rw_addr = ring_w + 30u * 16u + (unsigned)(gy & 1) * 8u;
  const unsigned ring_mask = SG_G5_STAGES * SG_G5_STAGE - 1u;
  const int N = a.N;
  const double w0 = a.w0;
  double acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0;

  // The copy of iteration m goes to slot m % STAGES and carries {patch(m+P), weight(m+P), index records of beam m+P+Q}; negative
  // m are the prologue.  Iteration i reads: the records of beam i+P (slot (i-Q) % STAGES) to address the patch it requests, the
  // patch and weight of beam i (slot (i-P) % STAGES), and the column nibble / new-row mask of beam i (slot (i-P-Q) % STAGES, still
  // alive because the ring is deeper than P+Q).
  // ---- prologue 1: the records of beams 0 .. P+Q-1 (iterations -P-Q .. -1), nothing else
#pragma unroll
  for (int m = -(SG_G5_P + SG_G5_Q); m < 0; ++m) {
    cp_async16(my_chunk + ((unsigned)(m & (SG_G5_STAGES - 1))) * SG_G5_STAGE, base + (size_t)uidx * 8u, is_idx);
    if (is_idx) uidx += inc;
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncwarp();
  // ---- prologue 2: patches and weights of beams 0 .. P-1 (iterations -P .. -1) from the records just fetched
#pragma unroll
  for (int m = -SG_G5_P; m < 0; ++m) {
    const unsigned rs = ((unsigned)((m - SG_G5_Q) & (SG_G5_STAGES - 1))) * SG_G5_STAGE;  // records of beam m+P
    const unsigned col0 = lds32(ring_w + 28u * 16u + rs);
    const uint2 rw = lds64(rw_addr + rs);
    const unsigned u = is_patch ? rw.x + col0 : uidx;
    cp_async16(my_chunk + ((unsigned)(m & (SG_G5_STAGES - 1))) * SG_G5_STAGE, base + (size_t)u * 8u, is_patch || is_w);
    if (is_w) uidx += inc;
    cp_async_commit();
  }

  // ---- beam loop, software-pipelined through registers: iteration i adds beam i out of values it loaded from the ring
  // during iteration i-1 and meanwhile loads beam i+1's (the LDS -> FP64 -> indexed-branch chain of a lone warp was ~300
  // cycles per beam; with the loads a beam ahead only the arithmetic is left on it)
  cp_async_wait<SG_G5_P - 1>();  // the patch of beam 0 has landed
  __syncwarp();
  constexpr unsigned ST = SG_G5_STAGE;
  unsigned sw_ = 0;  // stage offset of slot i % STAGES
  // slots relative to sw_: patch / weight of beam i+k at sw_ + (k-P)*ST, records of beam i+k at sw_ + (k-P-Q)*ST
  // two register sets, A and B, alternate (the loop is unrolled by two so that nothing has to be moved between them): the
  // even beams are added out of A while B is loaded for the next beam, the odd ones the other way round
  unsigned msk_a = lds32(rw_addr + 4u + ((sw_ - (SG_G5_P + SG_G5_Q) * ST) & ring_mask)), msk_b = 0;
  unsigned nib_b = lds32(nib_addr + ((sw_ + ST - (SG_G5_P + SG_G5_Q) * ST) & ring_mask)), nib_a = 0;
  double t_a[DMAX], t_b[DMAX], wi_a = w0, wi_b = w0;
  {
    const unsigned nibw0 = lds32(nib_addr + ((sw_ - (SG_G5_P + SG_G5_Q) * ST) & ring_mask));
    const unsigned s_val = (sw_ - SG_G5_P * ST) & ring_mask;
    const unsigned va = ring_w + s_val + ((nibw0 >> nib_shift) & 15u) * 8u;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) t_a[d] = lds_f64(va + d * SG_G5_ROW);
    if (!UNIW) wi_a = lds_f64(ring_w + 31u * 16u + s_val);
  }
#pragma unroll
  for (int d = 0; d < DMAX; ++d) t_b[d] = 0.0;
  // one beam: TC / WC / MC hold beam i; TN / WN / MN receive beam i+1 (its column nibbles are in NC); NN receives the
  // nibbles of beam i+2
#define SG_G5_BEAM(TC, WC, MC, TN, WN, MN, NC, NN, WAIT)                                                                  \
  {                                                                                                                      \
    if (WAIT) { /* one wait serves both halves: all copies but the last Q-2 have landed, i.e. iterations <= i+1-Q */      \
      cp_async_wait<SG_G5_Q - 2>();                                                                                      \
      __syncwarp();                                                                                                      \
    }                                                                                                                    \
    const unsigned s_nxt = (sw_ + ST - SG_G5_P * ST) & ring_mask;                                                         \
    const unsigned va = ring_w + s_nxt + ((NC >> nib_shift) & 15u) * 8u;                                                  \
    _Pragma("unroll") for (int d = 0; d < DMAX; ++d) TN[d] = lds_f64(va + d * SG_G5_ROW);                                 \
    if (!UNIW) WN = lds_f64(ring_w + 31u * 16u + s_nxt);                                                                  \
    NN = lds32(nib_addr + ((sw_ + 2u * ST - (SG_G5_P + SG_G5_Q) * ST) & ring_mask));                                      \
    MN = lds32(rw_addr + 4u + ((sw_ + ST - (SG_G5_P + SG_G5_Q) * ST) & ring_mask));                                       \
    const unsigned s_adr = (sw_ - SG_G5_Q * ST) & ring_mask;                                                              \
    const unsigned col0 = lds32(ring_w + 28u * 16u + s_adr);                                                              \
    const uint2 rw = lds64(rw_addr + s_adr);                                                                              \
    const unsigned u = is_patch ? rw.x + col0 : uidx;                                                                     \
    uidx += inc;                                                                                                          \
    cp_async16(my_chunk + sw_, base + (size_t)u * 8u, true); /* patch + weight of beam i+P, records of beam i+P+Q */       \
    cp_async_commit();                                                                                                    \
    double t[DMAX];                                                                                                       \
    _Pragma("unroll") for (int d = 0; d < DMAX; ++d) t[d] = sg::mul(TC[d], WC);                                           \
    add_pattern<DMAX>(MC, acc, t);                                                                                        \
    sw_ = (sw_ + ST) & ring_mask;                                                                                         \
  }
#pragma unroll 1
  for (int i = 0; i < N; i += 2) {
    SG_G5_BEAM(t_a, wi_a, msk_a, t_b, wi_b, msk_b, nib_b, nib_a, true)
    if (i + 1 < N) SG_G5_BEAM(t_b, wi_b, msk_b, t_a, wi_a, msk_a, nib_a, nib_b, false)
  }
#undef SG_G5_BEAM
  cp_async_wait<0>();

  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  const int j = band * a.bw + lane;
  if (wvalid && lane < a.bw && j < a.nx) {
    const int4 grp = __ldg(a.groups + g);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < grp.z || m >= grp.w) continue;
      double score = a.wsum == 0 ? NAN : sg::div(acc[m], a.wsum);
      long long idx = ((long long)grp.x * a.ny + (grp.y + m)) * a.nx + j;
      a.scores[idx - a.p0] = score;
      if (score == score && beats(score, idx, best_s, best_i)) { best_s = score; best_i = idx; }
    }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

template <int DMAX>
void launch_grid5(slamgpu_ctx *ctx, const GridArgs5 &a, int nblk, bool uniw) {
  if (uniw) k_score_grid5<DMAX, true><<<nblk, 64, 0, ctx->stream>>>(a);
  else k_score_grid5<DMAX, false><<<nblk, 64, 0, ctx->stream>>>(a);
}

// ---------------------------------------------------------------- v3: map patches staged in shared memory by TMA
// For one beam, every candidate of a block reads cells from one small patch of the LUT (the block's y rows
// x the whole x sweep: about 10 x 44 cells).  Warp 0 asks the TMA unit for that patch S beams ahead, one
// bulk copy per patch row (cp.async.bulk global -> shared, SASS UBLKCP); completion is counted on an
// mbarrier per stage; the gathers then hit shared memory (fixed ~30-cycle latency) instead of chasing L1
// misses into L2.  (The 2-D tensor form, cp.async.bulk.tensor / UTMALDG, faults with "illegal
// instruction" on this pool's driver even in a minimal probe -- tools/scratch/tma_probe.cu -- so rows
// are copied individually; patch columns start at an even column to keep the 16-byte source alignment.)
SG_DEV void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
SG_DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SG_DEV void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok = 0;
  const unsigned addr = smem_u32(bar);
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  }
}
// one patch row: TMA bulk copy global -> shared, completion counted on the stage's mbarrier
SG_DEV void bulk_load_row(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct GridArgs3 {
  const double *lut;
  int pitch, lut_rows;
  const int *cxp;
  const unsigned long long *cyw;
  const int2 *porg;
  const int4 *groups;  // {t, k0, m_lo, m_hi}
  const int4 *blocks;  // per block {t, first group, end group, block index within its theta}
  int nx, ny, ngy, N, t_lo, nbt, box_w, box_h;
  const double *w, *f;
  double w0, wsum;
  long long p0;
  double *scores;
  Best *blk;
};

#define SG_TMA_STAGES 4

// warp 0 fills one stage: lane 0 arms the barrier with the byte count, lanes copy one row each
SG_DEV void fill_stage(const GridArgs3 &a, unsigned char *dst, unsigned long long *bar, int2 org, unsigned box_bytes) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) mbar_expect_tx(bar, box_bytes);
  __syncwarp();
  const unsigned row_bytes = (unsigned)a.box_w * 8u;
  for (int r = lane; r < a.box_h; r += 32) {
    const int row = min(org.y + r, a.lut_rows - 1);  // rows past the padded LUT are never indexed: any valid row will do
    bulk_load_row(dst + (size_t)r * row_bytes, a.lut + (size_t)row * a.pitch + org.x, row_bytes, bar);
  }
}

template <int R, bool FACTOR, bool UNIW>
__global__ void __launch_bounds__(128) k_score_grid3(GridArgs3 a) {
  extern __shared__ __align__(128) unsigned char sm_patches[];  // SG_TMA_STAGES boxes, 128-byte aligned each
  __shared__ __align__(8) unsigned long long full_bar[SG_TMA_STAGES];
  const int4 blkrec = __ldg(a.blocks + blockIdx.x);
  const int tl = blkrec.x - a.t_lo;
  const int q = blkrec.w * 128 + threadIdx.x;
  const int gq = q / a.nx;
  const int g = blkrec.y + gq;
  const int j = q - gq * a.nx;
  const bool active = g < blkrec.z;
  const unsigned box_bytes = (unsigned)(a.box_w * a.box_h * 8);
  const unsigned stage_bytes = (box_bytes + 127u) & ~127u;
  const int2 *porg = a.porg + (size_t)tl * a.N * a.nbt + blkrec.w;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SG_TMA_STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int s = 0; s < SG_TMA_STAGES && s < a.N; ++s)
      fill_stage(a, sm_patches + s * stage_bytes, &full_bar[s], __ldg(porg + (size_t)s * a.nbt), box_bytes);
  }
  double best_s = -INFINITY;
  long long best_i = LLONG_MAX;
  int4 grp = make_int4(0, 0, 0, 0);
  const int *pcx = a.cxp;
  const unsigned long long *pcw = a.cyw;
  if (active) {
    grp = __ldg(a.groups + g);
    pcx = a.cxp + (size_t)tl * a.N * a.nx + j;
    pcw = a.cyw + (size_t)tl * a.N * a.ngy + grp.y / R;
  }
  double acc[R];
#pragma unroll
  for (int m = 0; m < R; ++m) acc[m] = 0.0;
  // operands of beam i+1 and i+2 in registers (index tables are L2 streams)
  int cx0 = 0, cx1 = 0;
  unsigned long long cw0 = 0, cw1 = 0;
  int2 og0 = make_int2(0, 0), og1 = make_int2(0, 0), ogp = make_int2(0, 0);
  if (a.N > 0) { og0 = __ldg(porg); if (active) { cx0 = __ldg(pcx); cw0 = __ldg(pcw); } }
  if (a.N > 1) { og1 = __ldg(porg + a.nbt); if (active) { cx1 = __ldg(pcx + a.nx); cw1 = __ldg(pcw + a.ngy); } }
  if (threadIdx.x < 32 && SG_TMA_STAGES < a.N) ogp = __ldg(porg + (size_t)SG_TMA_STAGES * a.nbt);
  for (int i = 0; i < a.N; ++i) {
    const int stage = i % SG_TMA_STAGES;
    const int cx = cx0;
    const unsigned long long cw = cw0;
    const int2 og = og0;
    cx0 = cx1; cw0 = cw1; og0 = og1;
    if (i + 2 < a.N) {
      og1 = __ldg(porg + (size_t)(i + 2) * a.nbt);
      if (active) { cx1 = __ldg(pcx + (size_t)(i + 2) * a.nx); cw1 = __ldg(pcw + (size_t)(i + 2) * a.ngy); }
    }
    const double wi = UNIW ? a.w0 : __ldg(a.w + i);
    const double fi = FACTOR ? __ldg(a.f + i) : 1.0;
    mbar_wait(&full_bar[stage], (unsigned)((i / SG_TMA_STAGES) & 1));
    if (active) {
      const double *patch = reinterpret_cast<const double *>(sm_patches + stage * stage_bytes);
      int off = ((int)(unsigned)(cw & 0xffffffffull) - og.y) * a.box_w + (cx - og.x);
      double v[R];
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m > 0) off += nib_sext(cw, m) * a.box_w;
        v[m] = patch[off];
      }
#pragma unroll
      for (int m = 0; m < R; ++m) {
        double term = sg::mul(v[m], wi);
        if (FACTOR) term = sg::mul(term, fi);
        acc[m] = sg::add(acc[m], term);
      }
    }
    __syncthreads();  // every thread is done with this stage: it can be refilled
    if (threadIdx.x < 32 && i + SG_TMA_STAGES < a.N) {
      fill_stage(a, sm_patches + stage * stage_bytes, &full_bar[stage], ogp, box_bytes);
      if (i + SG_TMA_STAGES + 1 < a.N) ogp = __ldg(porg + (size_t)(i + SG_TMA_STAGES + 1) * a.nbt);
    }
  }
  if (active) {
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < grp.z || m >= grp.w) continue;
      double score = a.wsum == 0 ? NAN : sg::div(acc[m], a.wsum);
      long long idx = ((long long)grp.x * a.ny + (grp.y + m)) * a.nx + j;
      a.scores[idx - a.p0] = score;
      if (score == score && beats(score, idx, best_s, best_i)) { best_s = score; best_i = idx; }
    }
  }
  block_argmax(best_s, best_i, a.blk + blockIdx.x);
}

template <int R>
void launch_grid3(slamgpu_ctx *ctx, const GridArgs3 &a, int nblk, size_t shm, bool factor, bool uniw) {
  if (factor) {
    if (uniw) k_score_grid3<R, true, true><<<nblk, 128, shm, ctx->stream>>>(a);
    else k_score_grid3<R, true, false><<<nblk, 128, shm, ctx->stream>>>(a);
  } else {
    if (uniw) k_score_grid3<R, false, true><<<nblk, 128, shm, ctx->stream>>>(a);
    else k_score_grid3<R, false, false><<<nblk, 128, shm, ctx->stream>>>(a);
  }
}

template <int R>
void launch_grid2(slamgpu_ctx *ctx, const GridArgs2 &a, int nblk, bool factor, bool uniw) {
  if (factor) {
    if (uniw) k_score_grid2<R, true, true><<<nblk, 128, 0, ctx->stream>>>(a);
    else k_score_grid2<R, true, false><<<nblk, 128, 0, ctx->stream>>>(a);
  } else {
    if (uniw) k_score_grid2<R, false, true><<<nblk, 128, 0, ctx->stream>>>(a);
    else k_score_grid2<R, false, false><<<nblk, 128, 0, ctx->stream>>>(a);
  }
}

// ------------------------------------------------------------------ window LUTs (max / mean OOPEs on the candidate grid)
// MaxOccupancyObservationPE / MeanOccupancyObservationPE (occupancy_observation_probability.h:29-72) fold the impacts of the
// cells a window covers, walking x outer / y inner (GridRasterizedRectangle, grid_rasterization.h:26-64).  The result depends
// only on the window's first cell and on how many cells it spans per axis, and for a window of fixed size that number takes
// two values (nmin or nmin + 1).  So the four possible folds are tabulated per first cell -- in the reference's operation
// order, hence bit-equal -- and a brute-force evaluation becomes one gather, exactly as for the obstacle OOPE: the existing
// grid kernel runs unchanged on offsets that select table and cell (k_grid_indices, a.win).
struct WlutArgs {
  const double *lut;   // padded impact LUT of the map
  int w, h, pitch;
  int mode, nxmin, nymin, pitch_w, rows_w;
  long long T;
  double *out;         // [vy][vx][rows_w][pitch_w]
};
__global__ void k_build_wlut(WlutArgs a) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= a.pitch_w || py >= a.rows_w) return;
  const int Kx = a.nxmin + 1, Ky = a.nymin + 1;
  const int ix0 = px - Kx, iy0 = py - Ky;  // internal cell the window starts at (may be outside the map)
#pragma unroll
  for (int vy = 0; vy < 2; ++vy)
#pragma unroll
    for (int vx = 0; vx < 2; ++vx) {
      const int nxc = a.nxmin + vx, nyc = a.nymin + vy;
      double acc = 0.0;
      for (int x = 0; x < nxc; ++x)
        for (int y = 0; y < nyc; ++y) {
          const int ix = clampi(ix0 + x, -1, a.w) + SG_LUT_PAD, iy = clampi(iy0 + y, -1, a.h) + SG_LUT_PAD;
          const double impact = __ldg(a.lut + (size_t)iy * a.pitch + ix);
          acc = a.mode == SLAMGPU_OOPE_MAX ? sg::maxd(impact, acc) : sg::add(acc, impact);
        }
      const double v = a.mode == SLAMGPU_OOPE_MAX ? acc : sg::div(acc, (double)(nxc * nyc));
      a.out[(size_t)(vy * 2 + vx) * a.T + (size_t)py * a.pitch_w + px] = v;
    }
}

int sg_map_ensure_wlut(slamgpu_map *m, int oie, int mode, double win_v, double win_h) {
  slamgpu_ctx *ctx = m->ctx;
  SG_TRY(sg_map_ensure_lut(m, oie));
  if (m->wlut_valid && m->wl_oie == oie && m->wl_mode == mode && m->wl_v == win_v && m->wl_h == win_h) return SLAMGPU_OK;
  const int nxmin = (int)std::floor(win_h / m->scale) + 1, nymin = (int)std::floor(win_v / m->scale) + 1;
  const int pitch_w = (m->w + nxmin + 1 + 1 + 1) & ~1, rows_w = m->h + nymin + 1 + 1;
  const size_t T = (size_t)pitch_w * rows_w;
  const size_t need = 4 * T + SG_LUT_SLACK;
  if (need > m->wlut_cap) {
    if (m->d_wlut) cudaFree(m->d_wlut);
    m->d_wlut = nullptr; m->wlut_cap = 0;
    SG_CUDA(ctx, cudaMalloc(&m->d_wlut, need * sizeof(double)));
    m->wlut_cap = need;
  }
  WlutArgs a;
  a.lut = m->d_lut[oie]; a.w = m->w; a.h = m->h; a.pitch = m->pitch;
  a.mode = mode; a.nxmin = nxmin; a.nymin = nymin; a.pitch_w = pitch_w; a.rows_w = rows_w; a.T = (long long)T; a.out = m->d_wlut;
  dim3 blk(32, 8), grd((pitch_w + 31) / 32, (rows_w + 7) / 8);
  k_build_wlut<<<grd, blk, 0, ctx->stream>>>(a);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  m->wl_oie = oie; m->wl_mode = mode; m->wl_v = win_v; m->wl_h = win_h; m->wl_nxmin = nxmin; m->wl_nymin = nymin;
  m->wl_pitch = pitch_w; m->wl_rows = rows_w; m->wl_T = T;
  m->wlut_valid = true;
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ host side
void slice_of(const slamgpu_ctx *ctx, int64_t P, int64_t *p0, int64_t *p1) {
  *p0 = (int64_t)((__int128)P * ctx->rank / ctx->nranks);
  *p1 = (int64_t)((__int128)P * (ctx->rank + 1) / ctx->nranks);
}

MapView make_view(const slamgpu_map *m, int oie) {
  MapView v;
  if (m->pool) (void)sg_map_sync_tiles(const_cast<slamgpu_map *>(m));  // the tile table the kernels read is the current one
  v.lut = m->d_lut[oie]; v.cells = m->d_cells; v.tiles = m->pool ? m->d_tile_ptrs : nullptr; v.tw = m->tw;
  v.w = m->w; v.h = m->h; v.ox = m->ox; v.oy = m->oy; v.pitch = m->pitch; v.stride = m->stride; v.model = m->model;
  v.scale = m->scale; v.unknown_lut = m->unknown_lut[oie];
  memcpy(v.unknown_rec, m->unknown, sizeof v.unknown_rec);
  return v;
}

int check_spe(slamgpu_ctx *ctx, const slamgpu_scan *scan, const slamgpu_spe_params *p) {
  if (!scan || !p) return sg_fail(ctx, SLAMGPU_E_INVALID, "scan/params is NULL");
  if (p->prerotated) SG_TRY(sg_scan_ensure_xy(const_cast<slamgpu_scan *>(scan)));  // pre-rotated scoring reads the points as (x, y)
  if (scan->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "scan belongs to another ctx");
  if (p->oope < 0 || p->oope > SLAMGPU_OOPE_GMAPPING) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad oope %d", p->oope);
  if (p->oie < 0 || p->oie > 1) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad oie %d", p->oie);
  if (p->trig_mode != SLAMGPU_TRIG_DEVICE && p->trig_mode != SLAMGPU_TRIG_HOST)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "bad trig_mode %d", p->trig_mode);
  if (p->oope == SLAMGPU_OOPE_GMAPPING && (p->gm_window < 0 || p->gm_window > 8))
    return sg_fail(ctx, SLAMGPU_E_INVALID, "bad gm_window %d", p->gm_window);
  return SLAMGPU_OK;
}

// host trig table with libm: bit-identical to the reference's std::cos/std::sin calls
void host_trig(const slamgpu_scan *s, const std::vector<double> &thetas, int64_t st_t, int64_t st_i,
               std::vector<double> &rc, std::vector<double> &rs) {
  const int N = s->n;
  const size_t T = thetas.size();
  rc.resize(T * N); rs.resize(T * N);
  for (size_t t = 0; t < T; ++t)
    for (int i = 0; i < N; ++i) {
      double ang = thetas[t] + s->angle[i];
      rc[t * st_t + i * st_i] = s->range[i] * std::cos(ang);
      rs[t * st_t + i * st_i] = s->range[i] * std::sin(ang);
    }
}

int upload(slamgpu_ctx *ctx, DevBuf &dst, const void *src, size_t bytes) {
  if (dst.reserve(std::max<size_t>(bytes, 16)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "device buffer of %zu bytes", bytes);
  if (bytes) SG_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SLAMGPU_OK;
}

int upload_host_trig(slamgpu_ctx *ctx, Candidates &c, const std::vector<double> &thetas, int64_t st_t, int64_t st_i) {
  std::vector<double> rc, rs;
  host_trig(c.scan, thetas, st_t, st_i, rc, rs);
  SG_TRY(upload(ctx, c.trc, rc.data(), rc.size() * sizeof(double)));
  SG_TRY(upload(ctx, c.trs, rs.data(), rs.size() * sizeof(double)));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // rc/rs are stack-lifetime host vectors
  c.trig_is_host = true;
  return SLAMGPU_OK;
}

#define SG_LIST_MAX_TABLE_BYTES (1ull << 30)
#define SG_IDX_SLACK 16  // zeroed beam rows behind the grid index tables (k_score_grid4 prefetches past the last beam)

}  // namespace

static int finish_list_stage(slamgpu_ctx *ctx, const slamgpu_spe_params *p);

// the candidate grid of slamgpu_stage_grid written out as a pose list on the device (OOPEs without a tiled kernel)
__global__ void k_expand_grid(const double *__restrict__ xs, const double *__restrict__ ys, const double *__restrict__ ts, int nx, int ny,
                              long long p0, long long Ploc, double *__restrict__ poses, int *__restrict__ theta_id) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= Ploc) return;
  const long long p = p0 + k;
  const int j = (int)(p % nx);
  const long long row = p / nx;
  const int y = (int)(row % ny), t = (int)(row / ny);
  poses[3 * k] = xs[j]; poses[3 * k + 1] = ys[y]; poses[3 * k + 2] = ts[t];
  theta_id[k] = t;
}

static int stage_grid_as_list(slamgpu_ctx *ctx, slamgpu_scan *scan, const slamgpu_spe_params *p, const double *xs, int32_t nx,
                              const double *ys, int32_t ny, const double *thetas, int32_t nt) {
  Candidates &c = ctx->cand;
  c.kind = -1; c.launched = false;
  c.P = (int64_t)nt * ny * nx; c.spe = *p; c.scan = scan;
  slice_of(ctx, c.P, &c.p0, &c.p1);
  const int64_t Ploc = c.p1 - c.p0;
  if ((unsigned long long)nt * std::max(scan->n, 1) * 16ull > SG_LIST_MAX_TABLE_BYTES)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "%d thetas x %d points exceed the trig table budget; split the call", nt, scan->n);
  c.h_axes.resize((size_t)nx + ny + nt);
  std::copy(xs, xs + nx, c.h_axes.begin());
  std::copy(ys, ys + ny, c.h_axes.begin() + nx);
  std::copy(thetas, thetas + nt, c.h_axes.begin() + nx + ny);
  SG_TRY(upload(ctx, c.d_xs, c.h_axes.data(), c.h_axes.size() * sizeof(double)));
  c.h_thetas.assign(thetas, thetas + nt);
  c.T = nt;
  SG_TRY(upload(ctx, c.d_thetas, c.h_thetas.data(), c.h_thetas.size() * sizeof(double)));
  if (c.poses.reserve(std::max<size_t>((size_t)Ploc * 3, 1) * sizeof(double)) != SLAMGPU_OK ||
      c.theta_id.reserve(std::max<size_t>((size_t)Ploc, 1) * sizeof(int)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "pose list of the candidate grid");
  if (Ploc > 0) {
    const double *d = c.d_xs.as<double>();
    k_expand_grid<<<(unsigned)((Ploc + 255) / 256), 256, 0, ctx->stream>>>(d, d + nx, d + nx + ny, nx, ny, c.p0, Ploc, c.poses.as<double>(),
                                                                         c.theta_id.as<int>());
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
  }
  return finish_list_stage(ctx, p);
}

extern "C" int slamgpu_stage_poses(slamgpu_ctx *ctx, slamgpu_scan *scan, const slamgpu_spe_params *p, const double *poses,
                                   int64_t P) {
  SG_NVTX("K1 stage_poses");
  if (!ctx) return SLAMGPU_E_INVALID;
  SG_TRY(check_spe(ctx, scan, p));
  if (P < 0 || (P > 0 && !poses)) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad pose list");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  Candidates &c = ctx->cand;
  c.kind = -1; c.launched = false;
  c.P = P; c.spe = *p; c.scan = scan;
  slice_of(ctx, P, &c.p0, &c.p1);
  const int64_t Ploc = c.p1 - c.p0;
  const int N = scan->n;
  SG_TRY(upload(ctx, c.poses, poses + 3 * c.p0, (size_t)Ploc * 3 * sizeof(double)));
  c.h_thetas.clear();
  std::vector<int32_t> tid((size_t)Ploc);
  if (!p->prerotated) {
    // distinct thetas of this slice (bit patterns), in order of first appearance
    std::unordered_map<uint64_t, int32_t> seen;
    seen.reserve((size_t)std::min<int64_t>(Ploc, 1 << 16));
    for (int64_t k = 0; k < Ploc; ++k) {
      double th = poses[3 * (c.p0 + k) + 2];
      uint64_t bits;
      memcpy(&bits, &th, 8);
      auto it = seen.find(bits);
      if (it == seen.end()) {
        it = seen.emplace(bits, (int32_t)c.h_thetas.size()).first;
        c.h_thetas.push_back(th);
      }
      tid[k] = it->second;
    }
    if ((unsigned long long)c.h_thetas.size() * std::max(N, 1) * 16ull > SG_LIST_MAX_TABLE_BYTES)
      return sg_fail(ctx, SLAMGPU_E_NOMEM, "%zu distinct thetas x %d points exceed the trig table budget; split the call",
                     c.h_thetas.size(), N);
  }
  c.T = (int32_t)c.h_thetas.size();
  SG_TRY(upload(ctx, c.theta_id, tid.data(), tid.size() * sizeof(int32_t)));
  SG_TRY(upload(ctx, c.d_thetas, c.h_thetas.data(), c.h_thetas.size() * sizeof(double)));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // tid is a local
  return finish_list_stage(ctx, p);
}

// the tail of a list staging: trig tables, score buffers (c.poses / c.theta_id / c.d_thetas / c.h_thetas are in place)
static int finish_list_stage(slamgpu_ctx *ctx, const slamgpu_spe_params *p) {
  Candidates &c = ctx->cand;
  const int N = c.scan->n;
  const int64_t Ploc = c.p1 - c.p0;
  size_t tb = std::max<size_t>((size_t)c.T * N, 1) * sizeof(double);
  if (c.trc.reserve(tb) != SLAMGPU_OK || c.trs.reserve(tb) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "trig table");
  c.trig_is_host = false;
  // the overlap OOPE depends continuously on the point position (not only on its cell), so device
  // trig could move a score by an ulp: it always gets libm trig
  if (p->oope == SLAMGPU_OOPE_OVERLAP) c.spe.trig_mode = SLAMGPU_TRIG_HOST;
  if (!p->prerotated && c.spe.trig_mode == SLAMGPU_TRIG_HOST && c.T > 0) SG_TRY(upload_host_trig(ctx, c, c.h_thetas, 1, c.T));
  int nblk = (int)((Ploc + 127) / 128);
  if (c.scores.reserve(std::max<size_t>(Ploc, 1) * sizeof(double)) != SLAMGPU_OK ||
      c.blk_best.reserve(std::max(nblk, 1) * sizeof(Best)) != SLAMGPU_OK ||
      c.result.reserve(sizeof(Result) * 2) != SLAMGPU_OK ||
      ctx->gather.reserve(sizeof(Result) * std::max(ctx->nranks, 1)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "score buffers");
  c.kind = 0;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_stage_grid(slamgpu_ctx *ctx, slamgpu_scan *scan, const slamgpu_spe_params *p, const double *xs,
                                  int32_t nx, const double *ys, int32_t ny, const double *thetas, int32_t nt) {
  SG_NVTX("K1 stage_grid");
  if (!ctx) return SLAMGPU_E_INVALID;
  SG_TRY(check_spe(ctx, scan, p));
  if (nx <= 0 || ny <= 0 || nt <= 0 || !xs || !ys || !thetas) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad grid axes");
  if (p->prerotated) return sg_fail(ctx, SLAMGPU_E_INVALID, "slamgpu_stage_grid takes polar scans (pre-rotated ones carry one theta)");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  Candidates &c = ctx->cand;
  c.win_mode = 0;
  if (p->oope == SLAMGPU_OOPE_MAX || p->oope == SLAMGPU_OOPE_MEAN) {
    // windows of a finite, non-zero size: scored out of the map's window LUTs by the grid kernel; anything else as a list
    const bool regular = p->win_v > 0 && p->win_h > 0 && std::isfinite(p->win_v) && std::isfinite(p->win_h);
    if (!regular || c.force_list) return stage_grid_as_list(ctx, scan, p, xs, nx, ys, ny, thetas, nt);
    c.win_mode = p->oope;
  } else if (p->oope != SLAMGPU_OOPE_OBSTACLE) {
    // overlap weights and the GMapping OOPE depend on where inside its cell a point falls: no table; the grid is written out
    // as a pose list on the device and scored by the list kernel
    return stage_grid_as_list(ctx, scan, p, xs, nx, ys, ny, thetas, nt);
  }
  c.kind = -1; c.launched = false;
  c.spe = *p; c.scan = scan;
  c.nx = nx; c.ny = ny; c.nt = nt; c.nyp = (ny + SG_GRID_R - 1) / SG_GRID_R * SG_GRID_R;
  c.P = (int64_t)nt * ny * nx;
  {
    static const int env_variant = [] { const char *e = getenv("SLAMGPU_GRID_VARIANT"); return e ? atoi(e) : 0; }();
    int v = c.user_variant ? c.user_variant : (env_variant >= 1 && env_variant <= 5 ? env_variant : 5);
    if (v >= 4) {
      // k_score_grid4 / 5 need ascending y (cell rows of a thread's 8 y then step by 0 or 1) and unit factors
      bool asc = !scan->has_factor;
      for (int k = 1; k < ny && asc; ++k) asc = ys[k] > ys[k - 1];
      if (!asc || c.user_rows == 2 || c.user_rows == 4) v = 2;
    }
    if (v == 5) {  // ... and ascending x (a band's columns then start at its first x)
      bool asc = true;
      for (int j = 1; j < nx && asc; ++j) asc = xs[j] > xs[j - 1];
      if (!asc) v = 4;
    }
    if (v > c.max_variant) v = (v == 5 && c.max_variant == 4) ? 4 : std::min(c.max_variant, 2);  // 5 -> 4 -> 2 -> 1, 3 -> 2
    if (c.force_v1 || c.win_mode) v = 1;  // (window tables are selected per y: the explicit row table carries that)
    if ((size_t)ny * SG_IDX_PAIRS * sizeof(int) > 40 * 1024) v = 1;  // the index kernel stages the rows in smem
    c.grid_variant = v;
    c.grid_v2 = v >= 2;
  }
  {
    // rows per thread: as many as keep the device full (one resident thread per pose column and y-group)
    int64_t r0_, r1_;
    slice_of(ctx, (int64_t)nt * ny, &r0_, &r1_);
    const int64_t want_threads = (int64_t)ctx->sm_count * 768;
    int R = 8;
    while (c.grid_v2 && c.grid_variant < 4 && R > 2 && ((r1_ - r0_ + R - 1) / R) * nx < want_threads) R /= 2;
    if (c.user_rows && c.grid_variant < 4) R = c.user_rows;
    c.grid_R = c.grid_v2 ? R : SG_GRID_R;
    c.ngy = (ny + c.grid_R - 1) / c.grid_R;
    c.ngys = c.grid_variant >= 4 ? (c.ngy + 1) & ~1 : c.ngy;
  }
  if (c.grid_variant == 5 && !c.user_variant) {
    // The cp.async pipeline of v5 hides the memory latency with a handful of warps per SM but is bound by the copy rate
    // (~30 cycles per 512-byte LDGSTS and SM) when the device is full; v4 needs ~28 resident warps per SM to hide its one L2
    // round trip per beam and is then 12 % faster (configs[2] on one GPU: 0.50 vs 0.58 ms; half of it: 0.44 vs 0.30 ms).
    static const int env_variant5 = [] { const char *e = getenv("SLAMGPU_GRID_VARIANT"); return e ? atoi(e) : 0; }();
    int64_t r0_, r1_;
    slice_of(ctx, (int64_t)nt * ny, &r0_, &r1_);
    const int64_t warps5 = ((r1_ - r0_ + 7) / 8) * ((nx + 29) / 30);
    if (env_variant5 != 5 && warps5 > (int64_t)ctx->sm_count * 29) {
      c.grid_variant = 4;
      c.ngys = (c.ngy + 1) & ~1;
    }
  }
  // group / warp tables and the zeroed slack rows depend on the shape of the candidate set only, not on the axis values: a
  // matcher that scores the same window around a new pose every scan re-uses them (saves two uploads and the memsets)
  const int64_t shape_key[8] = {nx, ny, nt, c.grid_variant + 16 * c.win_mode, c.grid_R, scan->n, ctx->rank, ctx->nranks};
  const bool same_shape = c.shape_valid && memcmp(shape_key, c.shape_key, sizeof shape_key) == 0;
  c.shape_valid = false;
  const int GR = c.grid_R;
  c.uniform_w = scan->n > 0;
  for (int i = 1; i < scan->n && c.uniform_w; ++i) c.uniform_w = scan->weight[i] == scan->weight[0];
  c.h_xs.assign(xs, xs + nx); c.h_ys.assign(ys, ys + ny); c.h_ts.assign(thetas, thetas + nt);
  // shard by rows (theta, y): contiguous candidate index ranges, whole rows per rank
  const int64_t rows = (int64_t)nt * ny;
  int64_t r0, r1;
  slice_of(ctx, rows, &r0, &r1);
  c.p0 = r0 * nx; c.p1 = r1 * nx;
  std::vector<int32_t> &groups = c.h_groups;
  if (!same_shape) groups.clear();
  c.t_lo = (int32_t)(r0 / ny);
  c.t_hi = r1 > r0 ? (int32_t)((r1 - 1) / ny) : c.t_lo;
  for (int32_t t = c.t_lo; t <= c.t_hi && r1 > r0 && !same_shape; ++t) {
    int64_t ka = std::max<int64_t>(r0 - (int64_t)t * ny, 0), kb = std::min<int64_t>(r1 - (int64_t)t * ny, ny);
    for (int32_t k0 = (int32_t)(ka / GR * GR); k0 < kb; k0 += GR) {
      int32_t lo = (int32_t)std::max<int64_t>(ka - k0, 0), hi = (int32_t)std::min<int64_t>(kb - k0, GR);
      groups.push_back(t); groups.push_back(k0); groups.push_back(lo); groups.push_back(hi);
    }
  }
  c.n_groups = (int32_t)(groups.size() / 4);
  c.rows_per_group = GR;
  // v3: blocks never straddle a theta (one patch per block and beam); per (theta, block) the y rows it touches
  std::vector<int32_t> blocks3, blk_rows3;
  c.nbt = 0; c.n_blocks3 = 0;
  if (c.grid_variant == 3) {
    const int nt_seg = c.t_hi - c.t_lo + 1;
    std::vector<std::pair<int, int>> seg(nt_seg, std::make_pair(-1, -1));  // [first group, end group) per theta
    for (int g = 0; g < c.n_groups; ++g) {
      const int tl = groups[4 * g] - c.t_lo;
      if (seg[tl].first < 0) seg[tl].first = g;
      seg[tl].second = g + 1;
    }
    for (int tl = 0; tl < nt_seg; ++tl)
      if (seg[tl].first >= 0) c.nbt = std::max<int>(c.nbt, ((seg[tl].second - seg[tl].first) * nx + 127) / 128);
    blk_rows3.assign((size_t)nt_seg * std::max(c.nbt, 1) * 2, 0);
    double max_dy = 0;
    for (int tl = 0; tl < nt_seg; ++tl) {
      for (int b = 0; b < c.nbt; ++b) { blk_rows3[((size_t)tl * c.nbt + b) * 2] = 0; blk_rows3[((size_t)tl * c.nbt + b) * 2 + 1] = -1; }
      if (seg[tl].first < 0) continue;
      const int nthreads = (seg[tl].second - seg[tl].first) * nx;
      for (int b = 0; b * 128 < nthreads; ++b) {
        const int g_first = seg[tl].first + (b * 128) / nx, g_last = seg[tl].first + std::min(b * 128 + 127, nthreads - 1) / nx;
        const int k_lo = groups[4 * g_first + 1], k_hi = std::min(groups[4 * g_last + 1] + GR - 1, ny - 1);
        blk_rows3[((size_t)tl * c.nbt + b) * 2] = k_lo; blk_rows3[((size_t)tl * c.nbt + b) * 2 + 1] = k_hi;
        double lo = ys[k_lo], hi = ys[k_lo];
        for (int k = k_lo; k <= k_hi; ++k) { lo = std::min(lo, ys[k]); hi = std::max(hi, ys[k]); }
        max_dy = std::max(max_dy, hi - lo);
        blocks3.push_back(c.t_lo + tl); blocks3.push_back(seg[tl].first); blocks3.push_back(seg[tl].second); blocks3.push_back(b);
      }
    }
    c.n_blocks3 = (int32_t)(blocks3.size() / 4);
    double xlo = xs[0], xhi = xs[0];
    for (int j = 0; j < nx; ++j) { xlo = std::min(xlo, xs[j]); xhi = std::max(xhi, xs[j]); }
    const double sc = ctx->cand.scan ? 0 : 0;  // (map scale is only known at launch: boxes are sized there)
    (void)sc;
    c.box_w = 0; c.box_h = 0;
    c.stats[6] = 0;
    // remember the extents; launch_staged turns them into cells with the map's scale
    c.h_extent_x = xhi - xlo; c.h_extent_y = max_dy;
  }
  const int N = scan->n;
  const int nt_loc = c.t_hi - c.t_lo + 1;
  if (!same_shape) SG_TRY(upload(ctx, c.groups, groups.data(), groups.size() * sizeof(int32_t)));
  if (c.grid_variant == 3) {
    SG_TRY(upload(ctx, c.blocks, blocks3.data(), blocks3.size() * sizeof(int32_t)));
    SG_TRY(upload(ctx, c.blk_rows, blk_rows3.data(), blk_rows3.size() * sizeof(int32_t)));
  }
  // the three axes travel as one block {xs | ys | thetas}; every source is a member vector, so nothing waits here
  c.h_axes.resize((size_t)nx + ny + nt);
  std::copy(xs, xs + nx, c.h_axes.begin());
  std::copy(ys, ys + ny, c.h_axes.begin() + nx);
  std::copy(thetas, thetas + nt, c.h_axes.begin() + nx + ny);
  SG_TRY(upload(ctx, c.d_xs, c.h_axes.data(), c.h_axes.size() * sizeof(double)));
  c.p_xs = c.d_xs.as<double>(); c.p_ys = c.p_xs + nx; c.p_ts = c.p_ys + ny;
  if (c.grid_variant == 3) SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // blocks3 / blk_rows3 are locals
  size_t tb = std::max<size_t>((size_t)nt * N, 1) * sizeof(double);
  if (c.trc.reserve(tb) != SLAMGPU_OK || c.trs.reserve(tb) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "trig table");
  c.trig_is_host = false;
  if (p->trig_mode == SLAMGPU_TRIG_HOST) SG_TRY(upload_host_trig(ctx, c, c.h_ts, N, 1));
  const int64_t Ploc = c.p1 - c.p0;
  long long threads = (long long)c.n_groups * nx;
  // (v4 packs the leftover columns of several y-groups into one warp: at most one extra warp per y-group)
  int nblk = std::max((int)((threads + 127) / 128), (int)c.n_blocks3) + (c.n_groups + 3) / 4 + 1;
  if (c.grid_variant == 5) nblk = std::max(nblk, (int)((c.n_groups * ((nx + 29) / 30) + 1) / 2));
  if (c.grid_variant == 3 && c.porg.reserve(std::max<size_t>((size_t)nt_loc * N * std::max(c.nbt, 1), 1) * sizeof(int2)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "patch origin table");
  // beam rows of slack behind the index tables: k_score_grid4 prefetches past the last beam without a guard
  const size_t idx_rows = (size_t)nt_loc * N + SG_IDX_SLACK;
  if (c.cxp.reserve((c.grid_variant == 5 ? 16 : idx_rows * nx) * sizeof(int)) != SLAMGPU_OK ||
      c.cyp.reserve(std::max<size_t>(c.grid_v2 ? 1 : (size_t)nt_loc * N * c.nyp, 1) * sizeof(int)) != SLAMGPU_OK ||
      c.cyw.reserve(std::max<size_t>(c.grid_v2 ? idx_rows * c.ngys : 1, 1) * sizeof(unsigned long long)) != SLAMGPU_OK ||
      c.scores.reserve(std::max<size_t>(Ploc, 1) * sizeof(double)) != SLAMGPU_OK ||
      c.blk_best.reserve(std::max(nblk, 1) * sizeof(Best)) != SLAMGPU_OK ||
      c.result.reserve(sizeof(Result) * 2 + 64) != SLAMGPU_OK ||
      ctx->gather.reserve(sizeof(Result) * std::max(ctx->nranks, 1)) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "grid score buffers");
  if (c.grid_variant == 5) {
    // bands of at most 30 consecutive x, balanced: their columns fit the 14-column window of a warp's patch
    c.nb5 = (nx + 29) / 30; c.bw5 = (nx + c.nb5 - 1) / c.nb5;
    if (c.colrec.reserve(idx_rows * c.nb5 * 32) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "column records");
    if (!same_shape) {
      SG_CUDA(ctx, cudaMemsetAsync((char *)c.colrec.p + (idx_rows - SG_IDX_SLACK) * c.nb5 * 32, 0, (size_t)SG_IDX_SLACK * c.nb5 * 32, ctx->stream));
      SG_CUDA(ctx, cudaMemsetAsync(c.cyw.as<unsigned long long>() + (idx_rows - SG_IDX_SLACK) * c.ngys, 0,
                                   SG_IDX_SLACK * (size_t)c.ngys * sizeof(unsigned long long), ctx->stream));
    }
    if (!c.uniform_w) {  // weight pairs {w[i], w[i+1]}: the copy lanes move 16 bytes
      c.h_w2.assign(2 * ((size_t)N + SG_IDX_SLACK), 0.0);
      for (int i = 0; i < N; ++i) { c.h_w2[2 * (size_t)i] = scan->weight[i]; c.h_w2[2 * (size_t)i + 1] = i + 1 < N ? scan->weight[i + 1] : 0.0; }
      SG_TRY(upload(ctx, c.w2, c.h_w2.data(), c.h_w2.size() * sizeof(double)));
    }
  }
  if (c.grid_variant == 4 && !same_shape) {
    // warp table: the full bands of G = 32 / (nx % 32) consecutive y-groups, then ONE warp with the leftover columns of those
    // groups -- slow warps (their indexed branch diverges) so spread evenly over the launch
    std::vector<int32_t> &wt = c.h_wtask;
    wt.clear();
    const int nbf = nx / 32, wr = nx % 32, G = wr ? 32 / wr : c.n_groups + 1;
    for (int g0 = 0; g0 < c.n_groups; g0 += G) {
      const int g1 = std::min(g0 + G, (int)c.n_groups);
      for (int g = g0; g < g1; ++g)
        for (int b = 0; b < nbf; ++b) { wt.push_back(g); wt.push_back(b); }
      if (wr) { wt.push_back(g0); wt.push_back(-(g1 - g0)); }
    }
    c.n_warps4 = (int32_t)(wt.size() / 2);
    SG_TRY(upload(ctx, c.wtask, wt.data(), wt.size() * sizeof(int32_t)));
    SG_CUDA(ctx, cudaMemsetAsync(c.cxp.as<int>() + (idx_rows - SG_IDX_SLACK) * nx, 0, SG_IDX_SLACK * (size_t)nx * sizeof(int), ctx->stream));
    SG_CUDA(ctx, cudaMemsetAsync(c.cyw.as<unsigned long long>() + (idx_rows - SG_IDX_SLACK) * c.ngys, 0, SG_IDX_SLACK * (size_t)c.ngys * sizeof(unsigned long long),
                                 ctx->stream));
  }
  if (c.grid_variant != 3) {  // (v3's block tables depend on the axis values)
    memcpy(c.shape_key, shape_key, sizeof shape_key);
    c.shape_valid = true;
  }
  c.kind = 1;
  return SLAMGPU_OK;
}

namespace {

template <int MODE, bool PREROT, bool GUARD>
void launch_list_f(slamgpu_ctx *ctx, const ListArgs &a, int nblk, bool factor) {
  if (factor) k_score_list<MODE, PREROT, GUARD, true><<<nblk, 128, 0, ctx->stream>>>(a);
  else k_score_list<MODE, PREROT, GUARD, false><<<nblk, 128, 0, ctx->stream>>>(a);
}
template <int MODE>
void launch_list_m(slamgpu_ctx *ctx, const ListArgs &a, int nblk, bool prerot, bool guard, bool factor) {
  if (prerot) launch_list_f<MODE, true, false>(ctx, a, nblk, factor);
  else if (guard) launch_list_f<MODE, false, true>(ctx, a, nblk, factor);
  else launch_list_f<MODE, false, false>(ctx, a, nblk, factor);
}

template <int MODE, bool PREROT, bool GUARD>
void launch_terms_f(slamgpu_ctx *ctx, const PointArgs &a, dim3 grd, bool factor) {
  if (factor) k_point_terms<MODE, PREROT, GUARD, true><<<grd, 128, 0, ctx->stream>>>(a);
  else k_point_terms<MODE, PREROT, GUARD, false><<<grd, 128, 0, ctx->stream>>>(a);
}
template <int MODE>
void launch_terms_m(slamgpu_ctx *ctx, const PointArgs &a, dim3 grd, bool prerot, bool guard, bool factor) {
  if (prerot) launch_terms_f<MODE, true, false>(ctx, a, grd, factor);
  else if (guard) launch_terms_f<MODE, false, true>(ctx, a, grd, factor);
  else launch_terms_f<MODE, false, false>(ctx, a, grd, factor);
}

#define SG_SMALL_MAX_POSES 8192
#define SG_SMALL_MAX_TERMS (4ll << 20)

int launch_staged(slamgpu_ctx *ctx, slamgpu_map *map, double init_score) {
  Candidates &c = ctx->cand;
  if (c.kind < 0 || !c.scan) return sg_fail(ctx, SLAMGPU_E_STATE, "no staged candidate set (stage_* first; re-stage after a scan upload)");
  if (!map || map->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map is NULL or belongs to another ctx");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  const slamgpu_scan *s = c.scan;
  const int N = s->n;
  const int oie = c.spe.oie;
  if (c.spe.oope != SLAMGPU_OOPE_GMAPPING && !c.multi) SG_TRY(sg_map_ensure_lut(map, oie));
  Result *res = c.result.as<Result>();
  SG_CUDA(ctx, cudaMemsetAsync(res, 0, sizeof(Result) * 2, ctx->stream));
  const int64_t Ploc = c.p1 - c.p0;
  const bool device_trig = !c.trig_is_host && !c.spe.prerotated;
  int nblk = 0;
  memset(c.stats, 0, sizeof c.stats);
  c.stats[1] = c.kind == 1 ? c.grid_variant : 0; c.stats[2] = Ploc * N; c.stats[5] = c.grid_R; c.stats[3] = c.p0; c.stats[4] = Ploc;
  if (c.kind == 0) {
    const bool small = Ploc > 0 && Ploc <= SG_SMALL_MAX_POSES && Ploc * (long long)N <= SG_SMALL_MAX_TERMS && N > 0;
    if (!small && device_trig && c.T > 0 && N > 0) {
      long long tot = (long long)c.T * N;
      k_trig_table<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(c.d_thetas.as<double>(), c.T, s->d_range, s->d_angle,
                                                                            N, 1, c.T, c.trc.as<double>(), c.trs.as<double>());
      SG_LAUNCHED(ctx);
    }
    nblk = (int)((Ploc + 127) / 128);
    if (nblk > 0) {
      ListArgs a;
      a.map = make_view(map, oie);
      a.views = c.multi ? c.views.as<MapView>() : nullptr;
      a.view_id = c.multi ? c.view_id.as<int>() : nullptr;
      a.poses = c.poses.as<double>(); a.theta_id = c.theta_id.as<int>();
      a.trc = c.trc.as<double>(); a.trs = c.trs.as<double>(); a.T = c.T;
      a.sx = s->d_x; a.sy = s->d_y; a.w = s->d_w; a.f = s->d_f; a.N = N;
      a.Ploc = Ploc; a.p0 = c.p0; a.wsum = s->wsum; a.win_v = c.spe.win_v; a.win_h = c.spe.win_h;
      a.gm_th = c.spe.gm_fullness_th; a.gm_win = c.spe.gm_window; a.gm_cache = c.spe.gm_cache;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>(); a.result = res;
      const bool pre = c.spe.prerotated != 0, fac = s->has_factor;
      if (small) {
        c.stats[1] = 3;
        const bool gmc = c.spe.oope == SLAMGPU_OOPE_GMAPPING && c.spe.gm_cache != 0;
        size_t tb = (size_t)Ploc * N * sizeof(double), cb = gmc ? (size_t)Ploc * N * sizeof(int2) : 0;
        if (ctx->scratch[7].reserve(tb + cb + 64) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "term buffer");
        PointArgs pa;
        pa.l = a; pa.thetas = c.d_thetas.as<double>(); pa.range = s->d_range; pa.angle = s->d_angle;
        pa.trig_is_table = c.trig_is_host ? 1 : 0;
        pa.terms = ctx->scratch[7].as<double>(); pa.cellids = (int2 *)((char *)ctx->scratch[7].p + tb);
        dim3 grd((N + 127) / 128, (unsigned)Ploc);
        cudaEventRecord(ctx->evk0, ctx->stream);
        switch (c.spe.oope) {
          case SLAMGPU_OOPE_OBSTACLE: launch_terms_m<SLAMGPU_OOPE_OBSTACLE>(ctx, pa, grd, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_MAX: launch_terms_m<SLAMGPU_OOPE_MAX>(ctx, pa, grd, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_MEAN: launch_terms_m<SLAMGPU_OOPE_MEAN>(ctx, pa, grd, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_OVERLAP: launch_terms_m<SLAMGPU_OOPE_OVERLAP>(ctx, pa, grd, pre, device_trig, fac); break;
          default: launch_terms_m<SLAMGPU_OOPE_GMAPPING>(ctx, pa, grd, pre, device_trig, fac); break;
        }
        cudaEventRecord(ctx->evk1, ctx->stream);
        ctx->evk_valid = true;
        if (gmc && c.spe.gm_cache == 2) {
          if (!c.gm_chain) return sg_fail(ctx, SLAMGPU_E_INVALID, "gm_cache == 2 needs the chained entry points (slamgpu_score_poses_chained)");
          ChainArgs ch;
          ch.pred = c.gm_pred.as<int>(); ch.states_in = c.gm_in.as<slamgpu_gm_cache>(); ch.states_out = c.gm_out.as<slamgpu_gm_cache>();
          if (fac) k_pose_sums_chained<true><<<nblk, 128, 0, ctx->stream>>>(pa, ch);
          else k_pose_sums_chained<false><<<nblk, 128, 0, ctx->stream>>>(pa, ch);
        } else if (gmc) {
          if (fac) k_pose_sums<true, true><<<nblk, 128, 0, ctx->stream>>>(pa);
          else k_pose_sums<true, false><<<nblk, 128, 0, ctx->stream>>>(pa);
        } else {
          k_pose_sums<false, false><<<nblk, 128, 0, ctx->stream>>>(pa);
        }
        ctx->launches += 2;
      } else {
        if (c.spe.oope == SLAMGPU_OOPE_GMAPPING && c.spe.gm_cache == 2)
          return sg_fail(ctx, SLAMGPU_E_INVALID, "chained GMapping scoring takes at most %d poses and %lld pose-points a call", SG_SMALL_MAX_POSES, (long long)SG_SMALL_MAX_TERMS);
        cudaEventRecord(ctx->evk0, ctx->stream);
        switch (c.spe.oope) {
          case SLAMGPU_OOPE_OBSTACLE: launch_list_m<SLAMGPU_OOPE_OBSTACLE>(ctx, a, nblk, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_MAX: launch_list_m<SLAMGPU_OOPE_MAX>(ctx, a, nblk, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_MEAN: launch_list_m<SLAMGPU_OOPE_MEAN>(ctx, a, nblk, pre, device_trig, fac); break;
          case SLAMGPU_OOPE_OVERLAP: launch_list_m<SLAMGPU_OOPE_OVERLAP>(ctx, a, nblk, pre, device_trig, fac); break;
          default: launch_list_m<SLAMGPU_OOPE_GMAPPING>(ctx, a, nblk, pre, device_trig, fac); break;
        }
        cudaEventRecord(ctx->evk1, ctx->stream);
        ctx->evk_valid = true;
        SG_LAUNCHED(ctx);
      }
    }
  } else {
    if (c.grid_variant == 3) {
      // size the TMA box from the candidate extents and this map's cell size; too large -> v2
      c.box_w = ((int)std::ceil(c.h_extent_x / map->scale) + 4 + 1) & ~1;  // +1: the patch starts at an even column
      c.box_h = (int)std::ceil(c.h_extent_y / map->scale) + 3;
      const size_t stage_bytes = ((size_t)c.box_w * c.box_h * 8 + 127) & ~(size_t)127;
      if (c.box_w > 256 || c.box_h > 256 || stage_bytes * SG_TMA_STAGES > 26 * 1024 || c.n_blocks3 == 0) {
        const int saved_max = c.max_variant;
        c.max_variant = 2;
        slamgpu_spe_params spe = c.spe;
        std::vector<double> xs = c.h_xs, ys = c.h_ys, ts = c.h_ts;
        int r = slamgpu_stage_grid(ctx, c.scan, &spe, xs.data(), (int32_t)xs.size(), ys.data(), (int32_t)ys.size(), ts.data(),
                                   (int32_t)ts.size());
        c.max_variant = saved_max;
        SG_TRY(r);
        return launch_staged(ctx, map, init_score);
      }
    }
    int dmax = 0;
    if (c.grid_variant >= 4) {
      // distinct cell rows the 8 y of one thread can touch: floor(span / cell) + 2, at most 8 (else v2)
      double span = 0;
      for (int k0 = 0; k0 < c.ny; k0 += 8) span = std::max(span, c.h_ys[std::min(k0 + 7, c.ny - 1)] - c.h_ys[k0]);
      const double cells = std::floor(span / map->scale) + 2.0;
      const size_t lut_elems = (size_t)map->pitch * (map->h + 2 * SG_LUT_PAD + SG_LUT_SLACK_ROWS);
      const size_t idx_elems = ((size_t)(c.t_hi - c.t_lo + 1) * N + SG_IDX_SLACK) * (size_t)std::max(c.nx, std::max(c.ngys, 4 * c.nb5));
      if (!(cells <= 8.0) || lut_elems >= (1ull << 31) || idx_elems >= (1ull << 32) || N <= 0) {
        const int saved_max = c.max_variant;
        c.max_variant = 2;
        slamgpu_spe_params spe = c.spe;
        std::vector<double> xs = c.h_xs, ys = c.h_ys, ts = c.h_ts;
        int r = slamgpu_stage_grid(ctx, c.scan, &spe, xs.data(), (int32_t)xs.size(), ys.data(), (int32_t)ys.size(), ts.data(),
                                   (int32_t)ts.size());
        c.max_variant = saved_max;
        SG_TRY(r);
        return launch_staged(ctx, map, init_score);
      }
      dmax = std::max(2, (int)cells);
      if (dmax == 7) dmax = 8;
      if (c.grid_variant == 5 && dmax > 4) {  // the patch of k_score_grid5 holds 4 rows: more -> v4
        const int saved_max = c.max_variant;
        c.max_variant = 4;
        slamgpu_spe_params spe = c.spe;
        std::vector<double> xs = c.h_xs, ys = c.h_ys, ts = c.h_ts;
        int r = slamgpu_stage_grid(ctx, c.scan, &spe, xs.data(), (int32_t)xs.size(), ys.data(), (int32_t)ys.size(), ts.data(),
                                   (int32_t)ts.size());
        c.max_variant = saved_max;
        SG_TRY(r);
        return launch_staged(ctx, map, init_score);
      }
    }
    const int nt_loc = c.t_hi - c.t_lo + 1;
    // big sets: make the score LUT L2-resident while the trig / index kernels run (side stream)
    static const int env_warm = [] { const char *e = getenv("SLAMGPU_WARM_L2"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool warm = (long long)Ploc * N >= (1ll << 26) && (env_warm >= 0 ? env_warm == 1 : c.warm_l2) && !c.win_mode;
    if (warm) {
      const size_t lut_bytes = (size_t)map->pitch * (map->h + 2 * SG_LUT_PAD) * sizeof(double);
      if (lut_bytes <= (96ull << 20)) {
        SG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        SG_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        k_warm_l2<<<ctx->sm_count * 4, 256, 0, ctx->side>>>(reinterpret_cast<const double2 *>(map->d_lut[oie]), lut_bytes / 16,
                                                            c.result.as<double>() + 8);
        SG_LAUNCHED(ctx);
        SG_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
      }
    }
    long long threads = (long long)c.n_groups * c.nx;
    nblk = (int)((threads + 127) / 128);
    if (nblk > 0 && N > 0) {
      GridIdxArgs ia;
      ia.trc = c.trc.as<double>(); ia.trs = c.trs.as<double>(); ia.xs = c.p_xs; ia.ys = c.p_ys;
      ia.thetas = device_trig ? c.p_ts : nullptr; ia.range = s->d_range; ia.angle = s->d_angle;
      ia.nx = c.nx; ia.ny = c.ny; ia.nyp = c.nyp; ia.N = N; ia.t_lo = c.t_lo; ia.nt_loc = nt_loc;
      ia.w = map->w; ia.h = map->h; ia.ox = map->ox; ia.oy = map->oy; ia.pitch = map->pitch; ia.scale = map->scale;
      ia.guard = device_trig ? 1 : 0;
      ia.cxp = c.cxp.as<int>(); ia.cyp = c.cyp.as<int>(); ia.result = res;
      ia.cyw = c.cyw.as<unsigned long long>(); ia.v2 = c.grid_v2 ? 1 : 0; ia.R = c.grid_R; ia.ngy = c.ngy;
      ia.win = 0;
      if (c.win_mode) {
        SG_TRY(sg_map_ensure_wlut(map, oie, c.win_mode, c.spe.win_v, c.spe.win_h));
        if (4 * map->wl_T >= (1ull << 31)) return sg_fail(ctx, SLAMGPU_E_NOMEM, "window tables too large for 32-bit offsets");
        ia.win = 1; ia.wl_nxmin = map->wl_nxmin; ia.wl_nymin = map->wl_nymin; ia.wl_pitch = map->wl_pitch; ia.wl_T = (long long)map->wl_T;
        ia.win_hh = c.spe.win_h / 2.0; ia.win_hv = c.spe.win_v / 2.0;
      }
      ia.v4 = c.grid_variant >= 4 ? 1 : 0; ia.dmax = dmax; ia.ngys = c.ngys;
      ia.v5 = c.grid_variant == 5 ? 1 : 0; ia.nb = c.nb5; ia.bw = c.bw5; ia.colrec = c.colrec.as<uint4>();
      ia.v3 = c.grid_variant == 3 ? 1 : 0; ia.nbt = c.nbt; ia.box_w = c.box_w; ia.box_h = c.box_h;
      ia.blk_rows = c.blk_rows.as<int2>(); ia.porg = c.porg.as<int2>();
      {
        const size_t shm = c.grid_v2 ? sizeof(int) * SG_IDX_PAIRS * ((size_t)c.ny + (c.grid_variant == 5 ? c.nx : 0)) : 0;
        dim3 grd((unsigned)((N + SG_IDX_PAIRS - 1) / SG_IDX_PAIRS), (unsigned)nt_loc);
        const int mode = (ia.win ? 1 : 0) | (ia.v2 ? 2 : 0) | (ia.v3 ? 4 : 0) | (ia.v4 ? 8 : 0) | (ia.v5 ? 16 : 0);
        switch (mode) {
          case 0: k_grid_indices<0><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          case 1: k_grid_indices<1><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          case 2: k_grid_indices<2><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          case 6: k_grid_indices<6><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          case 10: k_grid_indices<10><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          case 26: k_grid_indices<26><<<grd, SG_IDX_THREADS, shm, ctx->stream>>>(ia); break;
          default: return sg_fail(ctx, SLAMGPU_E_STATE, "grid index kernel: unexpected mode %d", mode);
        }
      }
      SG_LAUNCHED(ctx);
    }
    if (warm) SG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // no-op if the event was not recorded in this call
    if (nblk > 0 && c.grid_variant == 3) {
      GridArgs3 a;
      a.lut = map->d_lut[oie]; a.pitch = map->pitch; a.lut_rows = map->h + 2 * SG_LUT_PAD;
      a.cxp = c.cxp.as<int>(); a.cyw = c.cyw.as<unsigned long long>(); a.porg = c.porg.as<int2>();
      a.groups = c.groups.as<int4>(); a.blocks = c.blocks.as<int4>();
      a.nx = c.nx; a.ny = c.ny; a.ngy = c.ngy; a.N = N; a.t_lo = c.t_lo; a.nbt = c.nbt; a.box_w = c.box_w; a.box_h = c.box_h;
      a.w = s->d_w; a.f = s->d_f; a.w0 = N > 0 ? s->weight[0] : 0.0; a.wsum = s->wsum; a.p0 = c.p0;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>();
      nblk = c.n_blocks3;
      const size_t shm = (((size_t)c.box_w * c.box_h * 8 + 127) & ~(size_t)127) * SG_TMA_STAGES;
      cudaEventRecord(ctx->evk0, ctx->stream);
      if (c.grid_R == 8) launch_grid3<8>(ctx, a, nblk, shm, s->has_factor, c.uniform_w);
      else if (c.grid_R == 4) launch_grid3<4>(ctx, a, nblk, shm, s->has_factor, c.uniform_w);
      else launch_grid3<2>(ctx, a, nblk, shm, s->has_factor, c.uniform_w);
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->evk_valid = true;
      SG_LAUNCHED(ctx);
    } else if (nblk > 0 && c.grid_variant == 5) {
      GridArgs5 a;
      a.lut = map->d_lut[oie]; a.colrec = c.colrec.as<uint4>(); a.cyw = c.cyw.as<uint2>(); a.w2 = c.w2.as<double2>();
      a.groups = c.groups.as<int4>();
      a.n_groups = c.n_groups; a.nx = c.nx; a.ny = c.ny; a.ngy = c.ngy; a.ngys = c.ngys; a.N = N; a.t_lo = c.t_lo; a.pitch = map->pitch;
      a.nb = c.nb5; a.bw = c.bw5; a.w0 = s->weight[0]; a.wsum = s->wsum; a.p0 = c.p0;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>(); a.zero = 0u;
      nblk = (c.n_groups * c.nb5 + 1) / 2;
      cudaEventRecord(ctx->evk0, ctx->stream);
      switch (dmax) {
        case 2: launch_grid5<2>(ctx, a, nblk, c.uniform_w); break;
        case 3: launch_grid5<3>(ctx, a, nblk, c.uniform_w); break;
        default: launch_grid5<4>(ctx, a, nblk, c.uniform_w); break;
      }
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->evk_valid = true;
      SG_LAUNCHED(ctx);
    } else if (nblk > 0 && c.grid_variant == 4) {
      GridArgs4 a;
      a.lut = map->d_lut[oie]; a.cxp = c.cxp.as<unsigned>(); a.cyw = c.cyw.as<uint2>(); a.groups = c.groups.as<int4>();
      a.n_groups = c.n_groups; a.nx = c.nx; a.ny = c.ny; a.ngy = c.ngy; a.ngys = c.ngys; a.N = N; a.t_lo = c.t_lo; a.pitch = map->pitch;
      a.nb_full = c.nx / 32; a.wr = c.nx % 32;
      a.w = s->d_w; a.w0 = s->weight[0]; a.wsum = s->wsum; a.p0 = c.p0;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>();
      a.wtask = c.wtask.as<int2>(); a.n_warps = c.n_warps4;
      nblk = (c.n_warps4 + 3) / 4;
      {
        static const bool env_affine = [] { const char *e = getenv("SLAMGPU_SM_AFFINE"); return !e || atoi(e) != 0; }();
        a.zero = 0u;
        a.sm_cnt = nullptr; a.n_chunks = ctx->sm_count; a.chunk = (nblk + ctx->sm_count - 1) / ctx->sm_count;
        if (env_affine && nblk > ctx->sm_count) {
          if (c.sm_cnt.reserve(sizeof(int) * (size_t)ctx->sm_count) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "task counters");
          SG_CUDA(ctx, cudaMemsetAsync(c.sm_cnt.p, 0, sizeof(int) * (size_t)ctx->sm_count, ctx->stream));
          a.sm_cnt = c.sm_cnt.as<int>();
        }
      }
      cudaEventRecord(ctx->evk0, ctx->stream);
      switch (dmax) {
        case 2: launch_grid4<2>(ctx, a, nblk, c.uniform_w); break;
        case 3: launch_grid4<3>(ctx, a, nblk, c.uniform_w); break;
        case 4: launch_grid4<4>(ctx, a, nblk, c.uniform_w); break;
        case 5: launch_grid4<5>(ctx, a, nblk, c.uniform_w); break;
        case 6: launch_grid4<6>(ctx, a, nblk, c.uniform_w); break;
        default: launch_grid4<8>(ctx, a, nblk, c.uniform_w); break;
      }
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->evk_valid = true;
      SG_LAUNCHED(ctx);
    } else if (nblk > 0 && c.grid_v2) {
      GridArgs2 a;
      a.lut = map->d_lut[oie]; a.cxp = c.cxp.as<int>(); a.cyw = c.cyw.as<unsigned long long>(); a.groups = c.groups.as<int4>();
      a.n_groups = c.n_groups; a.nx = c.nx; a.ny = c.ny; a.ngy = c.ngy; a.N = N; a.t_lo = c.t_lo; a.pitch = map->pitch;
      a.w = s->d_w; a.f = s->d_f; a.w0 = N > 0 ? s->weight[0] : 0.0; a.wsum = s->wsum; a.p0 = c.p0;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>();
      cudaEventRecord(ctx->evk0, ctx->stream);
      if (c.grid_R == 8) launch_grid2<8>(ctx, a, nblk, s->has_factor, c.uniform_w);
      else if (c.grid_R == 4) launch_grid2<4>(ctx, a, nblk, s->has_factor, c.uniform_w);
      else launch_grid2<2>(ctx, a, nblk, s->has_factor, c.uniform_w);
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->evk_valid = true;
      SG_LAUNCHED(ctx);
    } else if (nblk > 0) {
      GridArgs a;
      a.lut = c.win_mode ? map->d_wlut : map->d_lut[oie]; a.cxp = c.cxp.as<int>(); a.cyp = c.cyp.as<int>(); a.groups = c.groups.as<int4>();
      a.n_groups = c.n_groups; a.nx = c.nx; a.ny = c.ny; a.nyp = c.nyp; a.N = N; a.t_lo = c.t_lo;
      a.w = s->d_w; a.f = s->d_f; a.wsum = s->wsum; a.p0 = c.p0;
      a.scores = c.scores.as<double>(); a.blk = c.blk_best.as<Best>();
      cudaEventRecord(ctx->evk0, ctx->stream);
      if (s->has_factor) k_score_grid<true><<<nblk, 128, 0, ctx->stream>>>(a);
      else k_score_grid<false><<<nblk, 128, 0, ctx->stream>>>(a);
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->evk_valid = true;
      SG_LAUNCHED(ctx);
    }
  }
  SG_CUDA(ctx, cudaGetLastError());
  // per-rank best -> res[1] (keeps the guard counter accumulated in res[0].guard)
  Result *local = res + 1;
  const bool p2p = ctx->nranks > 1 && ctx->d_peer_mailbox && !ctx->p2p_broken;
  if (nblk > 0 && ctx->nranks == 1) {
    k_reduce_finalize<<<1, 1024, 0, ctx->stream>>>(c.blk_best.as<Best>(), nblk, res, init_score, local);
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
    c.stats[6] = 0;
    c.launched = true;
    c.init_score = init_score;
    c.last_map = map;
    return SLAMGPU_OK;
  }
  if (nblk > 0 && !p2p) {
    k_reduce_blocks<<<1, 1024, 0, ctx->stream>>>(c.blk_best.as<Best>(), nblk, res);
    SG_LAUNCHED(ctx);
  } else if (nblk <= 0) {
    Result empty{-INFINITY, LLONG_MAX, 0, 0};
    SG_CUDA(ctx, cudaMemcpyAsync(res, &empty, sizeof(Result), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (p2p) {
    const long long deadline = (long long)(ctx->p2p_timeout_ms * 1.0e-3 * ctx->clock_hz);
    k_exchange_finalize<<<1, 1024, 0, ctx->stream>>>(nblk > 0 ? c.blk_best.as<Best>() : nullptr, nblk, res, ctx->d_peer_mailbox, ctx->rank,
                                                      ctx->nranks, ++ctx->p2p_seq, deadline, init_score, local, ctx->d_p2p_status);
    SG_LAUNCHED(ctx);
    SG_CUDA(ctx, cudaGetLastError());
    c.stats[6] = 1;
    c.launched = true;
    c.init_score = init_score;
    c.last_map = map;
    return SLAMGPU_OK;
  }
  c.stats[6] = 0;
  const Result *per_rank = res;
  if (ctx->nranks > 1) {
    std::string err;
    int r = sg_nccl_allgather(ctx->comm, res, ctx->gather.p, sizeof(Result), ctx->stream, &err);
    if (r != SLAMGPU_OK) return sg_fail(ctx, r, "%s", err.c_str());
    per_rank = ctx->gather.as<Result>();
  }
  k_finalize<<<1, 1, 0, ctx->stream>>>(per_rank, ctx->nranks, init_score, local);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  c.launched = true;
  c.init_score = init_score;
  c.last_map = map;
  return SLAMGPU_OK;
}

}  // namespace

extern "C" int slamgpu_score_launch(slamgpu_ctx *ctx, slamgpu_map *map, double init_score) {
  SG_NVTX("K1 score_launch");
  if (!ctx) return SLAMGPU_E_INVALID;
  return launch_staged(ctx, map, init_score);
}

static int fetch_impl(slamgpu_ctx *ctx, slamgpu_map *map, double *out_scores, int64_t *best_idx, double *best_score) {
  Candidates &c = ctx->cand;
  if (!c.launched) return sg_fail(ctx, SLAMGPU_E_STATE, "slamgpu_score_fetch without a launch");
  void *hp;
  SG_TRY(sg_pinned(ctx, sizeof(Result), &hp));
  Result *h = (Result *)hp;
  SG_CUDA(ctx, cudaMemcpyAsync(h, c.result.as<Result>() + 1, sizeof(Result), cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h->idx == LLONG_MIN) {
    ctx->p2p_broken = true;  // a late message must never be taken for a new one: ncclAllGather from here on
    return sg_fail(ctx, SLAMGPU_E_NCCL, "a peer rank did not deliver its result within %.0f ms (peer-memory exchange; ctx option p2p_timeout_ms)",
                   ctx->p2p_timeout_ms);
  }
  if (h->pad > 0 && c.kind == 1 && c.win_mode && map) {
    // some window spans a cell count outside the two tabulated ones (a corner within an ulp of a cell border): as a list
    slamgpu_spe_params spe = c.spe;
    std::vector<double> xs = c.h_xs, ys = c.h_ys, ts = c.h_ts;
    double init = c.init_score;
    c.force_list = true;
    int r = slamgpu_stage_grid(ctx, c.scan, &spe, xs.data(), (int32_t)xs.size(), ys.data(), (int32_t)ys.size(), ts.data(), (int32_t)ts.size());
    c.force_list = false;
    SG_TRY(r);
    SG_TRY(launch_staged(ctx, map, init));
    SG_CUDA(ctx, cudaMemcpyAsync(h, c.result.as<Result>() + 1, sizeof(Result), cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  while (h->pad > 0 && c.kind == 1 && c.grid_variant > 1 && map) {
    // v4: some thread's 8 y do not fall into consecutive cell rows -> v2; v3: some block's cells did not fit its TMA box
    // -> v2; v2: two neighbouring y values are more than 7 cell rows apart, the packed row word cannot hold it -> the
    // explicit row table (v1 kernel)
    const int saved_max = c.max_variant;
    c.max_variant = c.grid_variant == 5 ? 4 : (c.grid_variant >= 3 ? 2 : 1);
    slamgpu_spe_params spe = c.spe;
    std::vector<double> xs = c.h_xs, ys = c.h_ys, ts = c.h_ts;
    double init = c.init_score;
    int r = slamgpu_stage_grid(ctx, c.scan, &spe, xs.data(), (int32_t)xs.size(), ys.data(), (int32_t)ys.size(), ts.data(),
                               (int32_t)ts.size());
    c.max_variant = saved_max;
    SG_TRY(r);
    SG_TRY(launch_staged(ctx, map, init));
    SG_CUDA(ctx, cudaMemcpyAsync(h, c.result.as<Result>() + 1, sizeof(Result), cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (h->guard > 0 && !c.trig_is_host && map) {
    // some world point sits within the guard band of a cell border: redo with libm trig so the
    // cell indices are the reference's by construction (every rank sees the same summed counter)
    int64_t hits = h->guard;
    if (c.kind == 0) SG_TRY(upload_host_trig(ctx, c, c.h_thetas, 1, c.T));
    else SG_TRY(upload_host_trig(ctx, c, c.h_ts, c.scan->n, 1));
    SG_TRY(launch_staged(ctx, map, c.init_score));
    c.stats[0] = hits;
    SG_CUDA(ctx, cudaMemcpyAsync(h, c.result.as<Result>() + 1, sizeof(Result), cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c.trig_is_host = c.spe.trig_mode == SLAMGPU_TRIG_HOST;  // next launch of a DEVICE set uses device trig again
  }
  if (best_idx) *best_idx = h->idx;
  if (best_score) *best_score = h->score;
  if (out_scores && c.p1 > c.p0) {
    SG_CUDA(ctx, cudaMemcpyAsync(out_scores + c.p0, c.scores.p, (size_t)(c.p1 - c.p0) * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SLAMGPU_OK;
}

extern "C" int slamgpu_score_fetch(slamgpu_ctx *ctx, double *out_scores, int64_t *best_idx, double *best_score) {
  SG_NVTX("K1 score_fetch");
  if (!ctx) return SLAMGPU_E_INVALID;
  return fetch_impl(ctx, ctx->cand.last_map, out_scores, best_idx, best_score);
}

extern "C" int slamgpu_score_stats(const slamgpu_ctx *ctx, int64_t stats[8]) {
  if (!ctx || !stats) return SLAMGPU_E_INVALID;
  memcpy(stats, ctx->cand.stats, sizeof ctx->cand.stats);
  return SLAMGPU_OK;
}


#define SG_FUSED_MAX_POSES 1024

// returns 1 if the call was served by the fused small-batch path, 0 if the caller must take the staged path
static int score_small_oneshot(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, slamgpu_scan *scan,
                               const slamgpu_spe_params *p, const double *poses, int64_t P, double init_score, double *out_scores,
                               int64_t *best_idx, double *best_score, int *served) {
  *served = 0;
  const int N = scan->n;
  if (ctx->nranks != 1 || P <= 0 || P > SG_FUSED_MAX_POSES || N <= 0 || P * (int64_t)N > SG_SMALL_MAX_TERMS) return SLAMGPU_OK;
  if (p->trig_mode != SLAMGPU_TRIG_DEVICE && !p->prerotated) return SLAMGPU_OK;
  if (p->oope == SLAMGPU_OOPE_OVERLAP && !p->prerotated) return SLAMGPU_OK;  // always libm trig: staged path
  if (p->oope == SLAMGPU_OOPE_GMAPPING && p->gm_cache) return SLAMGPU_OK;    // cache emulation: staged path
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<MapView> views((size_t)n_maps);
  for (int k = 0; k < n_maps; ++k) {
    if (!maps[k] || maps[k]->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map %d is NULL or belongs to another ctx", k);
    if (p->oope != SLAMGPU_OOPE_GMAPPING) SG_TRY(sg_map_ensure_lut(maps[k], p->oie));
    views[k] = make_view(maps[k], p->oie);
  }
  const bool multi = view_id != nullptr;
  // staging layout (8-byte aligned pieces)
  const size_t o_hdr = 0, o_poses = sizeof(SmallHdr), o_views = o_poses + sizeof(double) * 3 * P;
  const size_t o_vid = o_views + (multi ? sizeof(MapView) * n_maps : 0);
  const size_t in_bytes = ((o_vid + (multi ? sizeof(int32_t) * P : 0)) + 15) & ~(size_t)15;
  const size_t out_bytes = sizeof(SmallHdr) + sizeof(double) * P;
  const size_t terms_bytes = sizeof(double) * (size_t)P * N;
  // device: [in | scores (directly after the header? no: separate out block) | terms]
  DevBuf &db = ctx->scratch[6];
  const size_t d_out = in_bytes, d_terms = (d_out + out_bytes + 15) & ~(size_t)15, d_blk = (d_terms + terms_bytes + 15) & ~(size_t)15;
  if (db.reserve(d_blk + sizeof(Best) * (size_t)(P + 1)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "small-batch buffers");
  void *hp;
  SG_TRY(sg_pinned(ctx, std::max(in_bytes, out_bytes), &hp));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  char *h = (char *)hp;
  memset(h + o_hdr, 0, sizeof(SmallHdr));
  memcpy(h + o_poses, poses, sizeof(double) * 3 * P);
  if (multi) {
    memcpy(h + o_views, views.data(), sizeof(MapView) * n_maps);
    memcpy(h + o_vid, view_id, sizeof(int32_t) * P);
    for (int64_t k = 0; k < P; ++k)
      if (view_id[k] < 0 || view_id[k] >= n_maps) return sg_fail(ctx, SLAMGPU_E_INVALID, "pose %lld: bad particle id", (long long)k);
  }
  char *d = (char *)db.p;
  SG_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  SmallArgs a;
  a.map = views[0];
  a.views = multi ? (const MapView *)(d + o_views) : nullptr;
  a.view_id = multi ? (const int *)(d + o_vid) : nullptr;
  a.poses = (const double *)(d + o_poses);
  a.range = scan->d_range; a.angle = scan->d_angle; a.sx = scan->d_x; a.sy = scan->d_y; a.w = scan->d_w; a.f = scan->d_f;
  a.N = N; a.P = (int)P; a.wsum = scan->wsum; a.win_v = p->win_v; a.win_h = p->win_h; a.gm_th = p->gm_fullness_th;
  a.init_score = init_score; a.gm_win = p->gm_window;
  // the header the kernel updates is the one at the start of the OUT block: copy the zeroed input header there first
  a.hdr = (SmallHdr *)(d + d_out);
  a.scores = (double *)(d + d_out + sizeof(SmallHdr));
  a.terms = (double *)(d + d_terms);
  SG_CUDA(ctx, cudaMemcpyAsync(d + d_out, d, sizeof(SmallHdr), cudaMemcpyDeviceToDevice, ctx->stream));
  dim3 grd((N + 127) / 128, (unsigned)P);
  const bool pre = p->prerotated != 0, fac = scan->has_factor;
  cudaEventRecord(ctx->evk0, ctx->stream);
  switch (p->oope) {
    case SLAMGPU_OOPE_OBSTACLE: launch_small_m<SLAMGPU_OOPE_OBSTACLE>(ctx, a, grd, pre, fac); break;
    case SLAMGPU_OOPE_MAX: launch_small_m<SLAMGPU_OOPE_MAX>(ctx, a, grd, pre, fac); break;
    case SLAMGPU_OOPE_MEAN: launch_small_m<SLAMGPU_OOPE_MEAN>(ctx, a, grd, pre, fac); break;
    case SLAMGPU_OOPE_OVERLAP: launch_small_m<SLAMGPU_OOPE_OVERLAP>(ctx, a, grd, pre, fac); break;
    default: launch_small_m<SLAMGPU_OOPE_GMAPPING>(ctx, a, grd, pre, fac); break;
  }
  cudaEventRecord(ctx->evk1, ctx->stream);
  ctx->evk_valid = true;
  {
    static bool attr_set = false;
    const size_t max_shm = 96 * 1024;
    if (!attr_set) {
      SG_CUDA(ctx, cudaFuncSetAttribute(k_small_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_shm));
      attr_set = true;
    }
    int G = (int)std::min<int64_t>(8, (int64_t)(max_shm / (sizeof(double) * N)));  // one warp of the block per pose
    if (G < 1) return SLAMGPU_OK;  // a scan too long for the staging block: staged path
    const int nb = (int)((P + G - 1) / G);
    k_small_final<<<nb, 256, sizeof(double) * (size_t)G * N, ctx->stream>>>(a, G, (Best *)(d + d_blk));
  }
  ctx->launches += 2;
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaMemcpyAsync(h, d + d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const SmallHdr *rh = (const SmallHdr *)h;
  if (rh->best.guard > 0) return SLAMGPU_OK;  // a point sits on a cell border: the staged path redoes it with libm trig
  if (best_idx) *best_idx = rh->best.idx;
  if (best_score) *best_score = rh->best.score;
  if (out_scores) memcpy(out_scores, h + sizeof(SmallHdr), sizeof(double) * P);
  Candidates &c = ctx->cand;
  c.kind = -1; c.launched = false;
  memset(c.stats, 0, sizeof c.stats);
  c.stats[1] = 4; c.stats[2] = P * N; c.stats[4] = P;
  *served = 1;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_score_poses(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                                   const double *poses, int64_t P, double init_score, double *out_scores, int64_t *best_idx,
                                   double *best_score) {
  SG_NVTX("K1 score_poses");
  if (!ctx) return SLAMGPU_E_INVALID;
  if (map && scan && p && poses && check_spe(ctx, scan, p) == SLAMGPU_OK) {
    int served = 0;
    SG_TRY(score_small_oneshot(ctx, &map, 1, nullptr, scan, p, poses, P, init_score, out_scores, best_idx, best_score, &served));
    if (served) return SLAMGPU_OK;
  }
  SG_TRY(slamgpu_stage_poses(ctx, scan, p, poses, P));
  SG_TRY(launch_staged(ctx, map, init_score));
  return fetch_impl(ctx, map, out_scores, best_idx, best_score);
}

extern "C" int slamgpu_score_grid(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                                  const double *xs, int32_t nx, const double *ys, int32_t ny, const double *thetas, int32_t nt,
                                  double init_score, double *out_scores, int64_t *best_idx, double *best_score) {
  SG_NVTX("K1 score_grid");
  if (!ctx) return SLAMGPU_E_INVALID;
  SG_TRY(slamgpu_stage_grid(ctx, scan, p, xs, nx, ys, ny, thetas, nt));
  SG_TRY(launch_staged(ctx, map, init_score));
  return fetch_impl(ctx, map, out_scores, best_idx, best_score);
}

// K6: one launch scores every particle's candidates against that particle's own map
static int stage_views(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, const slamgpu_spe_params *p, int64_t P) {
  Candidates &c = ctx->cand;
  std::vector<MapView> views(n_maps);
  for (int k = 0; k < n_maps; ++k) {
    if (!maps[k] || maps[k]->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "particle map %d is NULL or belongs to another ctx", k);
    if (p->oope != SLAMGPU_OOPE_GMAPPING) SG_TRY(sg_map_ensure_lut(maps[k], p->oie));
    views[k] = make_view(maps[k], p->oie);
  }
  for (int64_t k = 0; k < P; ++k)
    if (view_id[k] < 0 || view_id[k] >= n_maps) return sg_fail(ctx, SLAMGPU_E_INVALID, "pose %lld: bad particle id", (long long)k);
  SG_TRY(upload(ctx, c.views, views.data(), sizeof(MapView) * n_maps));
  SG_TRY(upload(ctx, c.view_id, view_id + c.p0, sizeof(int32_t) * (size_t)(c.p1 - c.p0)));
  return SLAMGPU_OK;
}

int sg_score_poses_multi(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, slamgpu_scan *scan,
                         const slamgpu_spe_params *p, const double *poses, int64_t P, double *out_scores) {
  if (!ctx || !maps || n_maps <= 0 || (P > 0 && !view_id)) return sg_fail(ctx, SLAMGPU_E_INVALID, "score_multi: bad argument");
  SgLocalScope local_only(ctx);  // every pose of this call is scored here; ranks split the PARTICLES (particles.cu)
  if (scan && p && poses && check_spe(ctx, scan, p) == SLAMGPU_OK) {
    int served = 0;
    SG_TRY(score_small_oneshot(ctx, maps, n_maps, view_id, scan, p, poses, P, -INFINITY, out_scores, nullptr, nullptr, &served));
    if (served) return SLAMGPU_OK;
  }
  SG_TRY(slamgpu_stage_poses(ctx, scan, p, poses, P));
  Candidates &c = ctx->cand;
  SG_TRY(stage_views(ctx, maps, n_maps, view_id, p, P));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  c.multi = true;
  int r = launch_staged(ctx, maps[0], -INFINITY);
  if (r == SLAMGPU_OK) r = fetch_impl(ctx, maps[0], out_scores, nullptr, nullptr);
  c.multi = false;
  return r;
}

int sg_score_chained(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n_maps, const int32_t *view_id, slamgpu_scan *scan,
                     const slamgpu_spe_params *p, const double *poses, int64_t P, const int32_t *pred,
                     const slamgpu_gm_cache *states_in, int n_states, double *out_scores, slamgpu_gm_cache *out_states) {
  if (!ctx || !maps || n_maps <= 0 || !p || !states_in || n_states <= 0 || P < 0) return sg_fail(ctx, SLAMGPU_E_INVALID, "score_chained: bad argument");
  if (p->oope != SLAMGPU_OOPE_GMAPPING) return sg_fail(ctx, SLAMGPU_E_INVALID, "score_chained: only the GMapping OOPE keeps a cache");
  if (P == 0) return SLAMGPU_OK;
  SgLocalScope local_only(ctx);  // a sequence cannot be split over ranks
  slamgpu_spe_params spe = *p;
  spe.gm_cache = 2;
  std::vector<int32_t> chain((size_t)P);
  for (int64_t k = 0; k < P; ++k) {
    chain[k] = pred ? pred[k] : (int32_t)k - 1;
    if (chain[k] >= k || chain[k] < -n_states) return sg_fail(ctx, SLAMGPU_E_INVALID, "score_chained: pose %lld has predecessor %d", (long long)k, chain[k]);
  }
  SG_TRY(slamgpu_stage_poses(ctx, scan, &spe, poses, P));
  Candidates &c = ctx->cand;
  const bool multi = view_id != nullptr;
  if (multi) SG_TRY(stage_views(ctx, maps, n_maps, view_id, &spe, P));
  SG_TRY(upload(ctx, c.gm_pred, chain.data(), sizeof(int32_t) * (size_t)P));
  SG_TRY(upload(ctx, c.gm_in, states_in, sizeof(slamgpu_gm_cache) * (size_t)n_states));
  if (c.gm_out.reserve(sizeof(slamgpu_gm_cache) * (size_t)P) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "cache states");
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  c.multi = multi; c.gm_chain = true;
  int r = launch_staged(ctx, maps[0], -INFINITY);
  if (r == SLAMGPU_OK) r = fetch_impl(ctx, maps[0], out_scores, nullptr, nullptr);
  c.multi = false; c.gm_chain = false;
  if (r == SLAMGPU_OK && out_states) {
    SG_CUDA(ctx, cudaMemcpyAsync(out_states, c.gm_out.p, sizeof(slamgpu_gm_cache) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return r;
}

// One matcher's candidates, scored as the sequence the reference evaluates them in (GMapping OOPE only)
extern "C" int slamgpu_score_poses_chained(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                                           const double *poses, int64_t P, const slamgpu_gm_cache *state_in, double *out_scores,
                                           slamgpu_gm_cache *out_states) {
  SG_NVTX("K1 score_poses_chained");
  if (!ctx || !map || !state_in) return sg_fail(ctx, SLAMGPU_E_INVALID, "score_poses_chained: NULL argument");
  if (map->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map belongs to another ctx");
  return sg_score_chained(ctx, &map, 1, nullptr, scan, p, poses, P, nullptr, state_in, 1, out_scores, out_states);
}

// The whole hill-climbing match of n instances on the device (k_hill_climb).  *served = 0 when the request is outside
// what the kernel covers (the caller then runs the round-by-round path) or when a device-trig border guard fired.
namespace {
template <int MODE>
void launch_hc(slamgpu_ctx *ctx, const HcArgs &a, size_t smem, bool fac) {
  // one matcher: the widest block (latency); many particles: two resident blocks per SM, one wave for <= 2 x SMs
  const int threads = a.n_inst > ctx->sm_count / 2 ? 512 : 1024;
  if (fac) {
    cudaFuncSetAttribute(k_hill_climb<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_hill_climb<MODE, true><<<a.n_inst, threads, smem, ctx->stream>>>(a);
  } else {
    cudaFuncSetAttribute(k_hill_climb<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_hill_climb<MODE, false><<<a.n_inst, threads, smem, ctx->stream>>>(a);
  }
}
}  // namespace

int sg_hill_climb_device(slamgpu_ctx *ctx, slamgpu_map *const *maps, int n, slamgpu_scan *scan, const slamgpu_spe_params *p,
                         const double *init, const uint8_t *active, uint32_t max_failed_rounds, double tr, double rot,
                         double *out8, double *log, int log_cap, slamgpu_gm_cache *gm_states, int *served) {
  *served = 0;
  if (!ctx || !maps || n <= 0 || !scan || !p || !init || !out8) return sg_fail(ctx, SLAMGPU_E_INVALID, "hill_climb: bad argument");
  SG_TRY(check_spe(ctx, scan, p));
  const int N = scan->n;
  const bool chain = p->oope == SLAMGPU_OOPE_GMAPPING && p->gm_cache == 2;
  if (chain && !gm_states) return sg_fail(ctx, SLAMGPU_E_INVALID, "hill_climb: gm_cache == 2 needs the cache states");
  const size_t smem = (size_t)6 * std::max(N, 1) * (sizeof(double) + (chain ? sizeof(int2) : 0));
  if (p->prerotated || p->trig_mode != SLAMGPU_TRIG_DEVICE || p->oope == SLAMGPU_OOPE_OVERLAP || N <= 0 || smem > 200 * 1024 ||
      (p->oope == SLAMGPU_OOPE_GMAPPING && p->gm_cache == 1))
    return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<MapView> views(n);
  for (int k = 0; k < n; ++k) {
    if (!maps[k] || maps[k]->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "hill_climb: map %d is NULL or belongs to another ctx", k);
    if (active && !active[k]) { views[k] = make_view(maps[k], p->oie); continue; }
    if (p->oope != SLAMGPU_OOPE_GMAPPING) SG_TRY(sg_map_ensure_lut(maps[k], p->oie));
    views[k] = make_view(maps[k], p->oie);
  }
  Candidates &c = ctx->cand;
  // staging: {views | init | active} in one upload, {out | log} in one download
  const size_t vb = sizeof(MapView) * n, ib = sizeof(double) * 3 * n, ab = ((size_t)n + 7) & ~(size_t)7;
  const size_t gb = chain ? sizeof(slamgpu_gm_cache) * n : 0;
  std::vector<char> stage(vb + ib + ab + gb, 0);
  memcpy(stage.data(), views.data(), vb);
  memcpy(stage.data() + vb, init, ib);
  if (active) memcpy(stage.data() + vb + ib, active, n);
  if (chain) memcpy(stage.data() + vb + ib + ab, gm_states, gb);
  SG_TRY(upload(ctx, c.views, stage.data(), stage.size()));
  const size_t ob = sizeof(double) * SG_HC_OUT * n, lb = log ? sizeof(double) * 4 * (size_t)log_cap * n : 0;
  if (c.scores.reserve(ob + lb) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "hill_climb output");
  HcArgs a;
  a.map = views[0]; a.views = c.views.as<MapView>(); a.n_inst = n;
  a.init = (const double *)((const char *)c.views.p + vb);
  a.active = active ? (const unsigned char *)c.views.p + vb + ib : nullptr;
  a.max_failed = max_failed_rounds; a.tr = tr; a.rot = rot;
  a.range = scan->d_range; a.angle = scan->d_angle; a.w = scan->d_w; a.f = scan->d_f; a.N = N; a.wsum = scan->wsum;
  a.win_v = p->win_v; a.win_h = p->win_h; a.gm_th = p->gm_fullness_th; a.gm_win = p->gm_window;
  a.out = c.scores.as<double>(); a.log = log ? a.out + SG_HC_OUT * (size_t)n : nullptr; a.log_cap = log_cap;
  a.gm_in = chain ? (const slamgpu_gm_cache *)((const char *)c.views.p + vb + ib + ab) : nullptr;
  cudaEventRecord(ctx->evk0, ctx->stream);
  switch (p->oope) {
    case SLAMGPU_OOPE_OBSTACLE: launch_hc<SLAMGPU_OOPE_OBSTACLE>(ctx, a, smem, scan->has_factor); break;
    case SLAMGPU_OOPE_MAX: launch_hc<SLAMGPU_OOPE_MAX>(ctx, a, smem, scan->has_factor); break;
    case SLAMGPU_OOPE_MEAN: launch_hc<SLAMGPU_OOPE_MEAN>(ctx, a, smem, scan->has_factor); break;
    default: launch_hc<SLAMGPU_OOPE_GMAPPING>(ctx, a, smem, scan->has_factor); break;
  }
  cudaEventRecord(ctx->evk1, ctx->stream);
  ctx->evk_valid = true;
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  std::vector<double> host((ob + lb) / sizeof(double));
  SG_CUDA(ctx, cudaMemcpyAsync(host.data(), c.scores.p, ob + lb, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < n; ++k)
    if (host[SG_HC_OUT * (size_t)k + 5] != 0) return SLAMGPU_OK;  // a point within the device-trig slack of a cell border: host-trig path
  for (int k = 0; k < n; ++k) memcpy(out8 + 8 * (size_t)k, host.data() + SG_HC_OUT * (size_t)k, 8 * sizeof(double));
  if (chain)
    for (int k = 0; k < n; ++k) {
      const double *o = host.data() + SG_HC_OUT * (size_t)k;
      gm_states[k].cx = (int32_t)o[8]; gm_states[k].cy = (int32_t)o[9]; gm_states[k].prob = o[10];
    }
  if (log) memcpy(log, host.data() + SG_HC_OUT * (size_t)n, lb);
  c.stats[0] = 0; c.stats[1] = 5; c.stats[2] = 0;
  for (int k = 0; k < n; ++k) c.stats[2] += (int64_t)host[SG_HC_OUT * (size_t)k + 4] * N;
  *served = 1;
  return SLAMGPU_OK;
}

// A Monte-Carlo match segment on the device (k_monte_carlo): from the enumerator state given, over the noise list given,
// until the budget is spent, the list runs out or an accept resets the dispersion.
namespace {
template <int MODE>
void launch_mc(slamgpu_ctx *ctx, const McArgs &a, size_t smem, bool fac) {
  if (fac) {
    cudaFuncSetAttribute(k_monte_carlo<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_monte_carlo<MODE, true><<<1, 1024, smem, ctx->stream>>>(a);
  } else {
    cudaFuncSetAttribute(k_monte_carlo<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_monte_carlo<MODE, false><<<1, 1024, smem, ctx->stream>>>(a);
  }
}
}  // namespace

extern "C" int slamgpu_match_mc(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *p,
                                const double best_pose[3], double best_prob, int32_t have_best, const double *noise, int32_t K,
                                uint32_t failed, uint32_t poses_nm, uint32_t max_failed, uint32_t max_poses, double out[10],
                                double *log, int32_t log_cap, int32_t *served) {
  SG_NVTX("K1 match_mc");
  if (!ctx || !map || !scan || !p || !best_pose || K < 0 || (K > 0 && !noise) || !out || !served)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "match_mc: bad argument");
  *served = 0;
  if (map->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map belongs to another ctx");
  SG_TRY(check_spe(ctx, scan, p));
  const int N = scan->n;
  if (p->prerotated || p->trig_mode != SLAMGPU_TRIG_DEVICE || p->oope == SLAMGPU_OOPE_OVERLAP || N <= 0 ||
      (p->oope == SLAMGPU_OOPE_GMAPPING && p->gm_cache != 0))
    return SLAMGPU_OK;
  const int batch = (int)std::min<size_t>(32, (size_t)(190 * 1024) / ((size_t)N * sizeof(double)));
  if (batch < 4) return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (p->oope != SLAMGPU_OOPE_GMAPPING) SG_TRY(sg_map_ensure_lut(map, p->oie));
  Candidates &c = ctx->cand;
  const size_t nb = sizeof(double) * 3 * (size_t)std::max(K, 1);
  SG_TRY(upload(ctx, c.poses, noise, sizeof(double) * 3 * (size_t)K));
  if (c.poses.reserve(nb) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "match_mc noise");
  const size_t ob = sizeof(double) * SG_MC_OUT, lb = log && log_cap > 0 ? sizeof(double) * 4 * (size_t)log_cap : 0;
  if (c.scores.reserve(ob + lb) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "match_mc output");
  McArgs a;
  a.map = make_view(map, p->oie);
  a.noise = c.poses.as<double>(); a.K = K;
  a.bx = best_pose[0]; a.by = best_pose[1]; a.bt = best_pose[2]; a.best = best_prob; a.have_best = have_best ? 1 : 0;
  a.failed = failed; a.poses_nm = poses_nm; a.max_failed = max_failed; a.max_poses = max_poses; a.batch = batch;
  a.range = scan->d_range; a.angle = scan->d_angle; a.w = scan->d_w; a.f = scan->d_f; a.N = N; a.wsum = scan->wsum;
  a.win_v = p->win_v; a.win_h = p->win_h; a.gm_th = p->gm_fullness_th; a.gm_win = p->gm_window;
  a.out = c.scores.as<double>(); a.log = lb ? a.out + SG_MC_OUT : nullptr; a.log_cap = lb ? log_cap : 0;
  const size_t smem = (size_t)batch * N * sizeof(double);
  cudaEventRecord(ctx->evk0, ctx->stream);
  switch (p->oope) {
    case SLAMGPU_OOPE_OBSTACLE: launch_mc<SLAMGPU_OOPE_OBSTACLE>(ctx, a, smem, scan->has_factor); break;
    case SLAMGPU_OOPE_MAX: launch_mc<SLAMGPU_OOPE_MAX>(ctx, a, smem, scan->has_factor); break;
    case SLAMGPU_OOPE_MEAN: launch_mc<SLAMGPU_OOPE_MEAN>(ctx, a, smem, scan->has_factor); break;
    default: launch_mc<SLAMGPU_OOPE_GMAPPING>(ctx, a, smem, scan->has_factor); break;
  }
  cudaEventRecord(ctx->evk1, ctx->stream);
  ctx->evk_valid = true;
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  std::vector<double> host((ob + lb) / sizeof(double));
  SG_CUDA(ctx, cudaMemcpyAsync(host.data(), c.scores.p, ob + lb, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (host[8] != 0) return SLAMGPU_OK;  // a point within the device-trig slack of a cell border: the caller's host-trig path
  memcpy(out, host.data(), ob);
  if (lb) memcpy(log, host.data() + SG_MC_OUT, lb);
  c.kind = -1;  // the pose staging buffer was reused
  c.stats[0] = 0; c.stats[1] = 6; c.stats[2] = (int64_t)host[9] * N;
  *served = 1;
  return SLAMGPU_OK;
}

// CUDA loads a kernel's code at its first launch (lazy module loading): tens of milliseconds in the middle of somebody's
// scan.  Touching the kernels once at ctx creation moves that cost to start-up.
#define SG_TOUCH(k) do { cudaFuncAttributes fa_; (void)cudaFuncGetAttributes(&fa_, k); } while (0)
void sg_preload_score() {
  SG_TOUCH((k_score_grid2<8, false, true>)); SG_TOUCH((k_score_grid2<8, false, false>)); SG_TOUCH((k_score_grid2<4, false, true>));
  SG_TOUCH((k_score_grid2<2, false, true>)); SG_TOUCH(k_grid_indices<0>); SG_TOUCH(k_grid_indices<1>); SG_TOUCH(k_grid_indices<2>); SG_TOUCH(k_grid_indices<6>); SG_TOUCH(k_grid_indices<10>); SG_TOUCH(k_grid_indices<26>); SG_TOUCH(k_trig_table); SG_TOUCH(k_reduce_blocks); SG_TOUCH(k_finalize);
  SG_TOUCH((k_score_list<SLAMGPU_OOPE_OBSTACLE, false, true, false>)); SG_TOUCH((k_point_terms<SLAMGPU_OOPE_OBSTACLE, false, true, false>));
  SG_TOUCH((k_pose_sums<false, false>)); SG_TOUCH((k_pose_sums_chained<false>));
  SG_TOUCH((k_small_fused<SLAMGPU_OOPE_OBSTACLE, false, false>)); SG_TOUCH((k_small_fused<SLAMGPU_OOPE_GMAPPING, false, false>));
  SG_TOUCH(k_small_final);
  SG_TOUCH((k_hill_climb<SLAMGPU_OOPE_OBSTACLE, false>)); SG_TOUCH((k_hill_climb<SLAMGPU_OOPE_GMAPPING, false>));
  SG_TOUCH((k_monte_carlo<SLAMGPU_OOPE_OBSTACLE, false>));
  (void)cudaGetLastError();
}
