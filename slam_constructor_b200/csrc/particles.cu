// particles.cu -- K6: the GMapping particle step batched over particles.
//
// Replaces the per-particle loop of GmappingParticleFilter::handle_observation
// (src/slams/gmapping/gmapping_particle_filter.h:70-85) around GmappingWorld::handle_observation
// (src/slams/gmapping/gmapping_world.h:73-101): every particle hill-climbs from its own pose against
// its OWN map (HillClimbingScanMatcher(6, 0.1, 0.1) upstream, src/slams/gmapping/init_gmapping.h:58-60),
// then inserts the scan into its own map.  Here the whole match of every particle runs in ONE launch
// (score.cu: k_hill_climb, one block per particle, the estimator's cell cache carried along); requests
// that kernel does not cover (host trig, overlap OOPE, a border guard hit) advance all particles in
// lock step instead: one K1 launch scores the current round of every particle (6 candidates each,
// each against its particle's map), the accept logic runs on the host.  The scan then goes into all
// maps by one batched insertion (mapping.cu: sg_append_plans).  Resampling moves or copies maps on
// the device; on a distributed ctx the particles are sharded over the ranks.
//
// Each particle owns a dense device map (the reference shares copy-on-write 128x128 tiles between
// particles, src/core/maps/lazy_tiled_grid_map.h:18-118; tile sharing on the device is future work --
// resampling here copies whole maps, device to device).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>

#include "hill_climb.h"
#include "mapping.h"

// On a distributed ctx the PARTICLES are sharded: rank r owns the contiguous range [lo, hi) of the n particles
// (chunks of ceil(n / nranks)), with their maps on its GPU; `maps` has n entries, NULL for the others.  Every call
// takes and returns arrays over all n particles and must be made by every rank with the same arguments; results are
// exchanged with one all-gather per call, cloned maps travel rank to rank in slamgpu_particles_resample.
struct slamgpu_particles {
  slamgpu_ctx *ctx = nullptr;
  std::vector<slamgpu_map *> maps;
  // GROW_TILED particle maps (UnboundedLazyTiledGridMap upstream) share copy-on-write tiles out of this pool; NULL: dense maps
  SgTilePool *pool = nullptr;
  unsigned long long *h_counts = nullptr;  // pinned: per local particle {cells updated, dropped} of the last insertion
  size_t h_counts_cap = 0;
  int64_t resample_bytes = 0, resample_tiles_shared = 0;  // of the last resampling
  double resample_ms = 0;
  int lo = 0, hi = 0, chunk = 0;
  int owner(int i) const { return i / chunk; }
  // per particle: the GMapping OOPE cache its own estimator object would hold (spe.gm_cache == 2)
  std::vector<slamgpu_gm_cache> gm_state;
};

extern "C" int slamgpu_particles_create(slamgpu_ctx *ctx, int32_t n, int32_t w, int32_t h, double scale, int32_t model,
                                        int32_t grow, const double *unknown_rec, slamgpu_particles **out) {
  if (!ctx || !out || n <= 0) return sg_fail(ctx, SLAMGPU_E_INVALID, "particles_create: bad argument");
  *out = nullptr;
  slamgpu_particles *p = new slamgpu_particles();
  p->ctx = ctx;
  p->chunk = (n + ctx->nranks - 1) / ctx->nranks;
  p->lo = std::min(n, p->chunk * ctx->rank);
  p->hi = std::min(n, p->lo + p->chunk);
  p->maps.assign(n, nullptr);
  p->gm_state.assign(n, slamgpu_gm_cache{0, 0, -1.0});
  static const bool env_dense = [] { const char *e = getenv("SLAMGPU_DENSE_PARTICLES"); return e && atoi(e) != 0; }();
  if (grow == SLAMGPU_GROW_TILED && !env_dense) {
    SG_CUDA(ctx, cudaSetDevice(ctx->device));
    int r = sg_pool_create(ctx, model, unknown_rec, &p->pool);
    if (r != SLAMGPU_OK) { slamgpu_particles_destroy(p); return r; }
  }
  for (int i = p->lo; i < p->hi; ++i) {
    int r = p->pool ? sg_map_create_tiled(ctx, p->pool, w, h, scale, model, grow, unknown_rec, &p->maps[i])
                    : slamgpu_map_create(ctx, w, h, scale, model, grow, unknown_rec, &p->maps[i]);
    if (r != SLAMGPU_OK) { slamgpu_particles_destroy(p); return r; }
  }
  *out = p;
  return SLAMGPU_OK;
}

extern "C" void slamgpu_particles_destroy(slamgpu_particles *p) {
  if (!p) return;
  for (slamgpu_map *m : p->maps)
    if (m) slamgpu_map_destroy(m);
  sg_pool_destroy(p->pool);
  if (p->h_counts) cudaFreeHost(p->h_counts);
  delete p;
}

extern "C" int slamgpu_particles_tile_stats(const slamgpu_particles *p, int64_t stats[8]) {
  if (!p || !stats) return SLAMGPU_E_INVALID;
  memset(stats, 0, sizeof(int64_t) * 8);
  stats[0] = p->pool ? 1 : 0;
  if (p->pool) {
    stats[1] = p->pool->tiles_live; stats[2] = p->pool->tiles_cloned;
    stats[3] = (int64_t)(p->pool->tile_doubles * sizeof(double));
    stats[4] = (int64_t)p->pool->chunks.size() * p->pool->tiles_per_chunk * stats[3];
  }
  stats[5] = p->resample_bytes; stats[6] = p->resample_tiles_shared; stats[7] = (int64_t)(p->resample_ms * 1e3);
  return SLAMGPU_OK;
}

extern "C" int slamgpu_particles_count(const slamgpu_particles *p) { return p ? (int)p->maps.size() : SLAMGPU_E_INVALID; }

extern "C" slamgpu_map *slamgpu_particles_map(slamgpu_particles *p, int32_t i) {
  if (!p || i < 0 || i >= (int)p->maps.size()) return nullptr;
  return p->maps[i];
}

extern "C" int slamgpu_particles_score(slamgpu_particles *p, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                                       const double *poses, int32_t per_particle, double *out_scores) {
  SG_NVTX("K6 particles_score");
  if (!p || !spe || per_particle < 0 || (per_particle > 0 && (!poses || !out_scores))) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  const int n = (int)p->maps.size(), nl = p->hi - p->lo;
  const int64_t P = (int64_t)nl * per_particle;
  if (per_particle == 0) return SLAMGPU_OK;
  std::vector<double> all((size_t)p->chunk * ctx->nranks * per_particle, NAN);
  double *mine = all.data() + (size_t)p->chunk * ctx->rank * per_particle;
  if (P > 0) {
    std::vector<int32_t> vid((size_t)P);
    for (int64_t k = 0; k < P; ++k) vid[k] = (int32_t)(k / per_particle);
    SG_TRY(sg_score_poses_multi(ctx, p->maps.data() + p->lo, nl, vid.data(), scan, spe, poses + 3 * (size_t)p->lo * per_particle, P, mine));
  }
  SG_TRY(sg_allgather_host(ctx, all.data(), sizeof(double) * p->chunk * per_particle));
  memcpy(out_scores, all.data(), sizeof(double) * (size_t)n * per_particle);
  return SLAMGPU_OK;
}


extern "C" int slamgpu_particles_match_hc(slamgpu_particles *p, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                                          const double *init_poses, const uint8_t *active, uint32_t max_failed_rounds,
                                          double translation_delta, double rotation_delta, double *out_poses, double *out_probs,
                                          int64_t *out_tested) {
  SG_NVTX("K6 particles_match_hc");
  if (!p || !scan || !spe || !init_poses || !out_poses || !out_probs) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  const int n = (int)p->maps.size(), lo = p->lo, nl = p->hi - p->lo;
  slamgpu_map *const *lmaps = p->maps.data() + lo;  // view ids are local: particle i is view i - lo
  std::vector<char> mine(n, 0);
  for (int i = 0; i < n; ++i) mine[i] = i >= lo && i < p->hi && (!active || active[i]);
  std::vector<HillClimb> hc(n);
  std::vector<double> poses;
  std::vector<int32_t> vid;
  std::vector<double> scores;
  // ---- fast path: every particle's whole match in one launch, one block per particle (score.cu: k_hill_climb)
  int served = 0;
  if (nl > 0) {
    std::vector<double> out8((size_t)8 * nl);
    SG_TRY(sg_hill_climb_device(ctx, lmaps, nl, scan, spe, init_poses + 3 * (size_t)lo, active ? active + lo : nullptr,
                                max_failed_rounds, translation_delta, rotation_delta, out8.data(), nullptr, 0,
                                p->gm_state.data() + lo, &served));
    if (served) {
      for (int i = 0; i < n; ++i) {
        HillClimb &h = hc[i];
        h.bx = init_poses[3 * i]; h.by = init_poses[3 * i + 1]; h.bt = init_poses[3 * i + 2];
        h.best = NAN; h.tested = 0; h.done = true;
        if (!mine[i]) continue;
        const double *o = out8.data() + 8 * (size_t)(i - lo);
        h.bx = o[0]; h.by = o[1]; h.bt = o[2]; h.best = o[3]; h.tested = (int64_t)o[4];
      }
    }
  }
  // ---- general path (the carried GMapping cache, host trig, overlap OOPE, guard hits): all particles advance in lock
  // step, one launch per hill-climbing round
  if (!served) {
  // probability of the initial poses (pose_enumeration_scan_matcher.h:40)
  for (int i = 0; i < n; ++i) {
    if (!mine[i]) continue;
    poses.insert(poses.end(), init_poses + 3 * i, init_poses + 3 * i + 3);
    vid.push_back(i - lo);
  }
  // one batch = consecutive runs of poses per particle.  With the carried GMapping cache every run continues the
  // sequence of its particle's own estimator: first pose from the particle's state, the others from the pose before
  const bool chained = spe->oope == SLAMGPU_OOPE_GMAPPING && spe->gm_cache == 2;
  std::vector<int32_t> pred;
  std::vector<slamgpu_gm_cache> states;
  auto score_batch = [&]() -> int {
    scores.resize(vid.size());
    if (vid.empty()) return SLAMGPU_OK;
    if (!chained) return sg_score_poses_multi(ctx, lmaps, nl, vid.data(), scan, spe, poses.data(), (int64_t)vid.size(), scores.data());
    pred.resize(vid.size()); states.resize(vid.size());
    for (size_t k = 0; k < vid.size(); ++k) pred[k] = k > 0 && vid[k - 1] == vid[k] ? (int32_t)k - 1 : -1 - vid[k];
    SG_TRY(sg_score_chained(ctx, lmaps, nl, vid.data(), scan, spe, poses.data(), (int64_t)vid.size(), pred.data(),
                            p->gm_state.data() + lo, nl, scores.data(), states.data()));
    for (size_t k = 0; k < vid.size(); ++k)
      if (k + 1 == vid.size() || vid[k + 1] != vid[k]) p->gm_state[lo + vid[k]] = states[k];
    return SLAMGPU_OK;
  };
  SG_TRY(score_batch());
  size_t q = 0;
  for (int i = 0; i < n; ++i) {
    HillClimb &h = hc[i];
    h.bx = init_poses[3 * i]; h.by = init_poses[3 * i + 1]; h.bt = init_poses[3 * i + 2];
    h.tr = translation_delta; h.rot = rotation_delta;
    if (!mine[i]) { h.done = true; h.best = NAN; h.tested = 0; continue; }
    h.best = scores[q++];
    h.done = !(0 < max_failed_rounds);
  }
  std::vector<int> who, cnt;
  std::vector<double> cands;  // per entry of `who`: 6 x 3
  for (;;) {
    poses.clear(); vid.clear(); who.clear(); cnt.clear(); cands.clear();
    for (int i = 0; i < n; ++i) {
      if (hc[i].done) continue;
      double c6[6][3];
      int k = hc[i].next_round(max_failed_rounds, c6);
      if (k == 0) { hc[i].done = true; continue; }
      who.push_back(i); cnt.push_back(k);
      cands.insert(cands.end(), &c6[0][0], &c6[0][0] + 18);
      for (int j = 0; j < k; ++j) {
        poses.push_back(c6[j][0]); poses.push_back(c6[j][1]); poses.push_back(c6[j][2]);
        vid.push_back(i - lo);
      }
    }
    if (who.empty()) break;
    SG_TRY(score_batch());
    size_t off = 0;
    for (size_t e = 0; e < who.size(); ++e) {
      double c6[6][3];
      memcpy(c6, cands.data() + 18 * e, sizeof c6);
      hc[who[e]].apply(max_failed_rounds, c6, scores.data() + off, cnt[e]);
      off += cnt[e];
    }
  }
  }  // !served
  if (ctx->nranks > 1) {  // every rank learns every particle's result: 5 x f64 per particle
    std::vector<double> all((size_t)p->chunk * ctx->nranks * 5, 0.0);
    for (int i = lo; i < p->hi; ++i) {
      double *r = all.data() + 5 * (size_t)i;
      r[0] = hc[i].bx; r[1] = hc[i].by; r[2] = hc[i].bt; r[3] = hc[i].best; r[4] = (double)hc[i].tested;
    }
    SG_TRY(sg_allgather_host(ctx, all.data(), sizeof(double) * 5 * p->chunk));
    for (int i = 0; i < n; ++i) {
      const double *r = all.data() + 5 * (size_t)i;
      hc[i].bx = r[0]; hc[i].by = r[1]; hc[i].bt = r[2]; hc[i].best = r[3]; hc[i].tested = (int64_t)r[4];
    }
  }
  for (int i = 0; i < n; ++i) {
    out_poses[3 * i] = hc[i].bx; out_poses[3 * i + 1] = hc[i].by; out_poses[3 * i + 2] = hc[i].bt;
    out_probs[i] = hc[i].best;
    if (out_tested) out_tested[i] = hc[i].tested;
  }
  return SLAMGPU_OK;
}

// Test hook (no GPU, no ctx): the hill-climbing state machine of hill_climb.h -- the one the device kernel and the
// round-by-round path share -- driven by a caller-supplied scoring function, so that CPU-only tests can hold its
// enumeration and accept logic against a sequential reference matcher.
extern "C" int slamgpu_debug_hill_climb(const double init_pose[3], uint32_t max_failed_rounds, double translation_delta,
                                        double rotation_delta, slamgpu_score_fn score, void *user, double out_pose[3],
                                        double *out_prob, int64_t *out_tested) {
  if (!init_pose || !score || !out_pose || !out_prob) return SLAMGPU_E_INVALID;
  HillClimb h;
  h.bx = init_pose[0]; h.by = init_pose[1]; h.bt = init_pose[2];
  h.tr = translation_delta; h.rot = rotation_delta;
  h.best = score(init_pose, user);
  h.done = !(0 < max_failed_rounds);
  while (!h.done) {
    double c6[6][3], s6[6];
    const int k = h.next_round(max_failed_rounds, c6);
    if (k == 0) break;
    for (int j = 0; j < k; ++j) s6[j] = score(c6[j], user);
    h.apply(max_failed_rounds, c6, s6, k);
  }
  out_pose[0] = h.bx; out_pose[1] = h.by; out_pose[2] = h.bt;
  *out_prob = h.best;
  if (out_tested) *out_tested = h.tested;
  return SLAMGPU_OK;
}

// HillClimbingScanMatcher::process_scan for ONE matcher and map: the whole match in one launch when the device kernel
// covers the request, round by round otherwise.  log (optional) receives the candidates in evaluation order.
extern "C" int slamgpu_match_hc(slamgpu_ctx *ctx, slamgpu_map *map, slamgpu_scan *scan, const slamgpu_spe_params *spe,
                                const double init_pose[3], uint32_t max_failed_rounds, double translation_delta,
                                double rotation_delta, double out_pose[3], double *out_prob, int64_t *out_tested,
                                double *log, int32_t log_cap, int32_t *log_count, slamgpu_gm_cache *gm_state) {
  SG_NVTX("K1 match_hc");
  if (!ctx || !map || !scan || !spe || !init_pose || !out_pose || !out_prob) return sg_fail(ctx, SLAMGPU_E_INVALID, "match_hc: NULL argument");
  if (map->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "map belongs to another ctx");
  if (log_count) *log_count = -1;  // -1: no log was produced (round-by-round path)
  int served = 0;
  double out8[8];
  slamgpu_gm_cache fresh{0, 0, -1.0};
  slamgpu_gm_cache *state = gm_state ? gm_state : &fresh;
  SG_TRY(sg_hill_climb_device(ctx, &map, 1, scan, spe, init_pose, nullptr, max_failed_rounds, translation_delta, rotation_delta,
                              out8, log_cap > 0 ? log : nullptr, log_cap > 0 ? log_cap : 0, state, &served));
  if (served) {
    out_pose[0] = out8[0]; out_pose[1] = out8[1]; out_pose[2] = out8[2];
    *out_prob = out8[3];
    if (out_tested) *out_tested = (int64_t)out8[4];
    if (log_count) *log_count = (int32_t)out8[6];
    return SLAMGPU_OK;
  }
  slamgpu_particles one;
  one.ctx = ctx; one.maps.assign(1, map); one.lo = 0; one.hi = 1; one.chunk = 1;
  one.gm_state.assign(1, *state);
  SgLocalScope local_only(ctx);
  int r = slamgpu_particles_match_hc(&one, scan, spe, init_pose, nullptr, max_failed_rounds, translation_delta, rotation_delta, out_pose,
                                     out_prob, out_tested);
  *state = one.gm_state[0];
  return r;
}

// Every rank of a distributed ctx must leave a collective entry point with the same verdict, or the ranks that went on
// block forever in the next all-gather: the local status travels to every rank and the first failure wins.
static int agree_on_status(slamgpu_ctx *ctx, int local_rc) {
  if (ctx->nranks <= 1) return local_rc;
  std::vector<int32_t> all((size_t)ctx->nranks, SLAMGPU_OK);
  all[ctx->rank] = local_rc;
  SG_TRY(sg_allgather_host(ctx, all.data(), sizeof(int32_t)));
  for (int r = 0; r < ctx->nranks; ++r)
    if (all[r] != SLAMGPU_OK) {
      if (local_rc == SLAMGPU_OK) return sg_fail(ctx, all[r], "rank %d failed in this collective call (its own error string says why)", r);
      return local_rc;
    }
  return SLAMGPU_OK;
}

extern "C" int slamgpu_particles_append_scan(slamgpu_particles *p, slamgpu_scan *scan, const double *poses,
                                             const uint8_t *do_update, double scan_quality, int32_t scan_margin,
                                             const slamgpu_estimator *est, double blur, double max_range,
                                             const double *point_quality, int64_t *cells_updated) {
  SG_NVTX("K6 particles_append_scan");
  if (!p || !scan || !poses || !est) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  if (scan->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "scan belongs to another ctx");
  if (scan_margin < 0) return sg_fail(ctx, SLAMGPU_E_INVALID, "negative scan margin");
  const int n = (int)p->maps.size();
  if (cells_updated) for (int i = 0; i < n; ++i) cells_updated[i] = 0;
  if (scan->n == 0) return SLAMGPU_OK;  // on every rank alike: no collective is skipped one-sidedly
  std::vector<int> who;
  for (int i = p->lo; i < p->hi; ++i)
    if (!do_update || do_update[i]) who.push_back(i);
  // per rank: chunk counts + the rank's status, so one all-gather both publishes the counts and agrees on the verdict
  const size_t slot = (size_t)p->chunk + 1;
  std::vector<int64_t> all_counts(slot * ctx->nranks, 0);
  auto count_of = [&](int i) -> int64_t & { return all_counts[(size_t)p->owner(i) * slot + (size_t)(i - p->owner(i) * p->chunk)]; };
  auto publish = [&](int local_rc) -> int {
    all_counts[(size_t)ctx->rank * slot + p->chunk] = local_rc;
    if (ctx->nranks > 1) SG_TRY(sg_allgather_host(ctx, all_counts.data(), sizeof(int64_t) * slot));
    for (int r = 0; r < ctx->nranks; ++r) {
      const int rc_r = (int)all_counts[(size_t)r * slot + p->chunk];
      if (rc_r == SLAMGPU_OK) continue;
      if (local_rc != SLAMGPU_OK) return local_rc;
      return sg_fail(ctx, rc_r, "particles_append_scan failed on rank %d (its own error string says why)", r);
    }
    if (cells_updated) for (int i = 0; i < n; ++i) cells_updated[i] = count_of(i);
    return SLAMGPU_OK;
  };
  if (who.empty()) return publish(SLAMGPU_OK);
  // host beam preparation (libm trig per beam, as the reference computes the end points) on a few threads
  const int m = (int)who.size();
  std::vector<BeamPlan> plans(m);
  std::vector<int> rc(m, SLAMGPU_OK);
  auto prep = [&](int k) {
    rc[k] = sg_prepare_beams(p->maps[who[k]], scan, poses + 3 * who[k], scan_quality, scan_margin, blur, max_range, point_quality, true, &plans[k]);
  };
  const int nthreads = (int)std::min<long long>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u),
                                               ((long long)m * scan->n + 16383) / 16384);
  if (nthreads <= 1) {
    for (int k = 0; k < m; ++k) prep(k);
  } else {
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&] { for (int k; (k = next.fetch_add(1)) < m;) prep(k); });
    for (auto &t : pool) t.join();
  }
  for (int k = 0; k < m; ++k)
    if (rc[k] != SLAMGPU_OK) return publish(sg_fail(ctx, SLAMGPU_E_INVALID, "particle %d: a beam spans more than 2^26 cells", who[k]));
  // batches bounded by the 32-bit (map, cell) key space and 2^30 cell slots
  std::vector<slamgpu_map *> maps;
  if (p->h_counts_cap < (size_t)m * 2) {
    if (p->h_counts) cudaFreeHost(p->h_counts);
    p->h_counts = nullptr; p->h_counts_cap = 0;
    if (cudaMallocHost(&p->h_counts, sizeof(unsigned long long) * 2 * (size_t)p->chunk) != cudaSuccess)
      return publish(sg_fail(ctx, SLAMGPU_E_NOMEM, "particles_append_scan: pinned counters"));
    p->h_counts_cap = 2 * (size_t)p->chunk;
  }
  unsigned long long *h_counts = p->h_counts;
  for (int k0 = 0; k0 < m;) {
    unsigned long long keys = 0;
    long long slots = 0;
    int k1 = k0;
    while (k1 < m) {
      // growth may enlarge a map before the keys are laid out: leave 4x head room
      unsigned long long cells = 4ull * (unsigned long long)p->maps[who[k1]]->w * p->maps[who[k1]]->h;
      if (k1 > k0 && (keys + cells >= 0xF0000000ull || slots + plans[k1].M >= (1ll << 30))) break;
      keys += cells; slots += plans[k1].M; ++k1;
    }
    maps.clear();
    for (int k = k0; k < k1; ++k) maps.push_back(p->maps[who[k]]);
    // deferred: the batch is queued and the loop goes on to prepare the next one while its kernels run; the cell counts
    // of every batch land in one pinned array, read after the one synchronisation below
    const int arc = sg_append_plans(ctx, maps.data(), plans.data() + k0, k1 - k0, est, nullptr, nullptr, h_counts + 2 * (size_t)k0);
    if (arc != SLAMGPU_OK) { cudaStreamSynchronize(ctx->stream); ctx->staged_pending = false; return publish(arc); }
    k0 = k1;
  }
  {
    const cudaError_t se = cudaStreamSynchronize(ctx->stream);
    ctx->staged_pending = false;
    if (se != cudaSuccess) return publish(sg_fail(ctx, SLAMGPU_E_CUDA, "particles_append_scan: %s", cudaGetErrorString(se)));
  }
  for (int k = 0; k < m; ++k) count_of(who[k]) = (int64_t)h_counts[2 * (size_t)k];
  return publish(SLAMGPU_OK);
}

// ParticleFilter::try_resample's copy step (src/core/particle_filter.h:92-98): particle i becomes a
// copy of particle src[i].  A source that survives exactly once is moved, the others are copied on
// the device.
extern "C" int slamgpu_particles_resample(slamgpu_particles *p, const int32_t *src) {
  SG_NVTX("K6 particles_resample");
  if (!p || !src) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = p->ctx;
  const int n = (int)p->maps.size(), lo = p->lo, hi = p->hi;
  for (int i = 0; i < n; ++i)
    if (src[i] < 0 || src[i] >= n) return sg_fail(ctx, SLAMGPU_E_INVALID, "resample: bad source index %d", src[i]);
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  // ---- sources on other ranks: every rank learns every map's geometry, then the cells travel in one NCCL group
  // (sender: the map as it is now; receiver: a staging buffer, so that no map is overwritten while a peer reads it)
  struct Geo { int32_t w, h, ox, oy; };
  std::vector<Geo> geo;
  std::vector<DevBuf> staged(n), send_staged(n);
  auto release_staged = [&]() { for (DevBuf &b : staged) b.release(); for (DevBuf &b : send_staged) b.release(); };
  cudaEvent_t ev_r0 = nullptr, ev_r1 = nullptr;
  cudaEventCreate(&ev_r0); cudaEventCreate(&ev_r1);
  cudaEventRecord(ev_r0, ctx->stream);
  auto leave = [&](int rc) { release_staged(); if (ev_r0) cudaEventDestroy(ev_r0); if (ev_r1) cudaEventDestroy(ev_r1); return rc; };
  p->resample_bytes = 0; p->resample_tiles_shared = 0;
  if (ctx->nranks > 1) {
    geo.assign((size_t)p->chunk * ctx->nranks, Geo{0, 0, 0, 0});
    for (int i = lo; i < hi; ++i) geo[i] = Geo{p->maps[i]->w, p->maps[i]->h, p->maps[i]->ox, p->maps[i]->oy};
    { int grc = sg_allgather_host(ctx, geo.data(), sizeof(Geo) * p->chunk); if (grc != SLAMGPU_OK) return leave(grc); }
    const int stride = slamgpu_model_stride(p->lo < p->hi ? p->maps[lo]->model : 0);
    std::vector<SgXfer> sends, recvs;
    int stage_rc = SLAMGPU_OK;  // a rank that cannot stage still reaches the agreement below: no peer waits in the exchange
    for (int i = 0; i < n && stage_rc == SLAMGPU_OK; ++i) {  // the same order on every rank
      const int q = p->owner(src[i]), r = p->owner(i);
      if (q == r) continue;
      if (ctx->rank == q) {
        slamgpu_map *m = p->maps[src[i]];
        const size_t bytes = (size_t)m->w * m->h * m->stride * sizeof(double);
        void *cells = m->d_cells;
        if (m->pool) {  // a tiled map travels as the dense array it stands for
          if (send_staged[i].reserve(std::max<size_t>(bytes, 16)) != SLAMGPU_OK) { stage_rc = sg_fail(ctx, SLAMGPU_E_NOMEM, "resample: send staging for particle %d", i); break; }
          stage_rc = sg_map_gather_dense(m, send_staged[i].as<double>());
          if (stage_rc != SLAMGPU_OK) break;
          cells = send_staged[i].p;
        }
        sends.push_back(SgXfer{r, cells, bytes});
      } else if (ctx->rank == r) {
        const Geo &g = geo[src[i]];
        const size_t bytes = (size_t)g.w * g.h * stride * sizeof(double);
        if (staged[i].reserve(std::max<size_t>(bytes, 16)) != SLAMGPU_OK) { stage_rc = sg_fail(ctx, SLAMGPU_E_NOMEM, "resample: staging for particle %d", i); break; }
        recvs.push_back(SgXfer{q, staged[i].p, bytes});
      }
    }
    stage_rc = agree_on_status(ctx, stage_rc);
    if (stage_rc != SLAMGPU_OK) return leave(stage_rc);
    std::string err;
    int rc = sg_nccl_exchange(ctx->comm, sends.data(), (int)sends.size(), recvs.data(), (int)recvs.size(), ctx->stream, &err);
    if (rc != SLAMGPU_OK) return leave(sg_fail(ctx, rc, "%s", err.c_str()));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return leave(sg_fail(ctx, SLAMGPU_E_CUDA, "resample: exchange did not complete"));
  }
  // ---- this rank's particles
  std::vector<slamgpu_map *> next(n, nullptr);
  std::vector<char> taken(n, 0);
  auto local = [&](int i) { return i >= lo && i < hi; };
  // keep a surviving particle in place when it can be (no copy at all for the common "i -> i")
  for (int i = lo; i < hi; ++i)
    if (src[i] == i) { next[i] = p->maps[i]; taken[i] = 1; }
  // first other use of a (local) source whose own slot is not reused takes the object itself
  for (int i = lo; i < hi; ++i) {
    if (next[i] || !local(src[i])) continue;
    if (!taken[src[i]]) { next[i] = p->maps[src[i]]; taken[src[i]] = 1; }
  }
  // the rest are copies into the maps nobody kept
  std::vector<slamgpu_map *> spare;
  for (int i = lo; i < hi; ++i)
    if (!taken[i]) spare.push_back(p->maps[i]);
  // local sources first (they are all kept maps), then the staged remote ones
  int copy_rc = SLAMGPU_OK;
  auto copy_ok = [&](int r) { if (r != SLAMGPU_OK && copy_rc == SLAMGPU_OK) copy_rc = r; return r == SLAMGPU_OK; };
  for (int pass = 0; pass < 2 && copy_rc == SLAMGPU_OK; ++pass)
    for (int i = lo; i < hi && copy_rc == SLAMGPU_OK; ++i) {
      if (next[i] || local(src[i]) != (pass == 0)) continue;
      slamgpu_map *to = spare.back();
      spare.pop_back();
      if (pass == 0) {
        slamgpu_map *from = p->maps[src[i]];
        if (to->pool) {
          // copy-on-write: the copy shares every tile of its source (lazy_tiled_grid_map.h:57-71); nothing moves
          if (!copy_ok(sg_map_share_tiles(to, from))) break;
          for (int32_t id : from->tile_ids) p->resample_tiles_shared += id != 0;
        } else {
          if (!copy_ok(sg_map_realloc(to, from->w, from->h))) break;
          to->ox = from->ox; to->oy = from->oy;
          const size_t bytes = (size_t)from->w * from->h * from->stride * sizeof(double);
          if (cudaMemcpyAsync(to->d_cells, from->d_cells, bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { copy_ok(sg_fail(ctx, SLAMGPU_E_CUDA, "resample: copy of particle %d", src[i])); break; }
          p->resample_bytes += (int64_t)bytes;
        }
      } else {
        const Geo &g = geo[src[i]];
        if (!copy_ok(sg_map_realloc(to, g.w, g.h))) break;
        to->ox = g.ox; to->oy = g.oy;
        const size_t bytes = (size_t)g.w * g.h * to->stride * sizeof(double);
        if (to->pool) { if (!copy_ok(sg_map_scatter_dense(to, staged[i].as<double>()))) break; }
        else if (cudaMemcpyAsync(to->d_cells, staged[i].p, bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { copy_ok(sg_fail(ctx, SLAMGPU_E_CUDA, "resample: copy into particle %d", i)); break; }
        p->resample_bytes += (int64_t)bytes;
      }
      sg_map_invalidate_lut(to);
      next[i] = to;
    }
  cudaEventRecord(ev_r1, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) copy_ok(sg_fail(ctx, SLAMGPU_E_CUDA, "resample: copies did not complete"));
  // every rank learns whether every rank's copies went through before the estimator states are exchanged
  copy_rc = agree_on_status(ctx, copy_rc);
  if (copy_rc != SLAMGPU_OK) return leave(copy_rc);
  {
    float ms = 0;
    if (ev_r0 && ev_r1 && cudaEventElapsedTime(&ms, ev_r0, ev_r1) == cudaSuccess) p->resample_ms = ms;
    if (ev_r0) cudaEventDestroy(ev_r0);
    if (ev_r1) cudaEventDestroy(ev_r1);
    ev_r0 = ev_r1 = nullptr;
  }
  release_staged();
  p->maps.swap(next);
  // GMapping OOPE cache: the first particle (in index order) drawn from a source keeps the source's estimator object
  // and with it the cache (particle_filter.h:92-101); every further copy starts with a new estimator, cache empty
  std::vector<slamgpu_gm_cache> all = p->gm_state;
  if (ctx->nranks > 1) {
    all.resize((size_t)p->chunk * ctx->nranks, slamgpu_gm_cache{0, 0, -1.0});
    SG_TRY(sg_allgather_host(ctx, all.data(), sizeof(slamgpu_gm_cache) * p->chunk));
  }
  std::vector<slamgpu_gm_cache> st(n, slamgpu_gm_cache{0, 0, -1.0});
  std::vector<char> drawn(n, 0);
  for (int i = 0; i < n; ++i) {
    if (!drawn[src[i]]) { st[i] = all[src[i]]; drawn[src[i]] = 1; }
  }
  p->gm_state.swap(st);
  return SLAMGPU_OK;
}
