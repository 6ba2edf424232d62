// core.cu -- context, device grid map, scan upload, score-LUT kernels, timers.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dev_math.cuh"
#include "internal.h"

// ------------------------------------------------------------------ errors
static std::string g_last_error;
void sg_set_global_error(const char *msg) { g_last_error = msg; }
int sg_fail(slamgpu_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_last_error = buf;
  return code;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return SLAMGPU_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  size_t want = std::max<size_t>(bytes + bytes / 4, 256);
  if (cudaMalloc(&p, want) != cudaSuccess) {
    if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; (void)cudaGetLastError(); return SLAMGPU_E_NOMEM; }
    want = bytes;
  }
  cap = want;
  return SLAMGPU_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
}

int sg_pinned(slamgpu_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->h_pinned_cap) {
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0;
    size_t want = std::max<size_t>(bytes * 2, 4096);
    SG_CUDA(ctx, cudaMallocHost(&ctx->h_pinned, want));
    ctx->h_pinned_cap = want;
  }
  *out = ctx->h_pinned;
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ context
extern "C" int slamgpu_abi_version(void) { return SLAMGPU_ABI_VERSION; }

extern "C" int slamgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ++ok;
  }
  return ok;
}

static int ctx_create_common(int device, slamgpu_ctx **out) {
  if (!out) return sg_fail(nullptr, SLAMGPU_E_INVALID, "slamgpu_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return sg_fail(nullptr, SLAMGPU_E_NODEVICE, "no CUDA device visible: libslamgpu has no CPU fallback");
  }
  if (device < 0 || device >= n) return sg_fail(nullptr, SLAMGPU_E_INVALID, "device %d out of range (%d visible)", device, n);
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess)
    return sg_fail(nullptr, SLAMGPU_E_CUDA, "cudaGetDeviceProperties(%d) failed", device);
  if (pr.major != 10)
    return sg_fail(nullptr, SLAMGPU_E_NODEVICE, "device %d is sm_%d%d; libslamgpu is built for sm_100a only", device,
                   pr.major, pr.minor);
  slamgpu_ctx *ctx = new slamgpu_ctx();
  ctx->device = device;
  ctx->sm_count = pr.multiProcessorCount;
  {
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) == cudaSuccess && khz > 0) ctx->clock_hz = 1e3 * khz;
  }
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->evk0) != cudaSuccess || cudaEventCreate(&ctx->evk1) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    int r = sg_fail(nullptr, SLAMGPU_E_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return r;
  }
  // kernels loaded now rather than at their first launch in the middle of a scan (SLAMGPU_LAZY_KERNELS=1 skips)
  {
    const char *lazy = getenv("SLAMGPU_LAZY_KERNELS");
    if (!(lazy && lazy[0] == '1')) { sg_preload_score(); sg_preload_mapping(); sg_preload_pyramid(); }
  }
  // pinned staging sized up front: growing it later (cudaFreeHost + cudaMallocHost) can stall a scan for ~100 ms
  void *hp;
  if (sg_pinned(ctx, (size_t)4 << 20, &hp) != SLAMGPU_OK) (void)cudaGetLastError();
  *out = ctx;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_ctx_create(int device, slamgpu_ctx **out) { return ctx_create_common(device, out); }

extern "C" int slamgpu_nccl_unique_id(void *id128) {
  std::string err;
  int r = sg_nccl_unique_id(id128, &err);
  if (r != SLAMGPU_OK) return sg_fail(nullptr, r, "%s", err.c_str());
  return r;
}

extern "C" int slamgpu_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id128, slamgpu_ctx **out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return sg_fail(nullptr, SLAMGPU_E_INVALID, "bad rank %d / %d", rank, nranks);
  SG_TRY(ctx_create_common(device, out));
  slamgpu_ctx *ctx = *out;
  ctx->rank = rank; ctx->nranks = nranks;
  if (nranks > 1) {
    if (!nccl_id128) { slamgpu_ctx_destroy(ctx); *out = nullptr; return sg_fail(nullptr, SLAMGPU_E_INVALID, "nccl id is NULL"); }
    std::string err;
    int r = sg_nccl_init(nranks, rank, nccl_id128, &ctx->comm, &err);
    if (r != SLAMGPU_OK) { slamgpu_ctx_destroy(ctx); *out = nullptr; return sg_fail(nullptr, r, "%s", err.c_str()); }
    sg_p2p_setup(ctx);  // optional: without it the per-rank results travel by ncclAllGather
  }
  return SLAMGPU_OK;
}

// Mailboxes for the fused result exchange: each rank allocates one, the CUDA IPC handles are all-gathered over NCCL,
// every rank maps every peer's mailbox (NVLink peer access).  All ranks must agree, so the outcome is all-gathered too;
// any failure (no peer access, IPC not permitted in this container, SLAMGPU_NO_P2P=1) leaves the NCCL path in place.
void sg_p2p_setup(slamgpu_ctx *ctx) {
  const int n = ctx->nranks;
  if (n < 2 || n > 64) return;
  const char *off = getenv("SLAMGPU_NO_P2P");
  int ok = !(off && off[0] == '1');
  const size_t box_bytes = (size_t)2 * n * 64;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok && (cudaMalloc(&ctx->mailbox, box_bytes) != cudaSuccess || cudaMemset(ctx->mailbox, 0, box_bytes) != cudaSuccess ||
             cudaIpcGetMemHandle(&mine, ctx->mailbox) != cudaSuccess))
    ok = 0;
  (void)cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  std::vector<cudaIpcMemHandle_t> all(n);
  all[ctx->rank] = mine;
  if (sg_allgather_host(ctx, all.data(), sizeof(cudaIpcMemHandle_t)) != SLAMGPU_OK) ok = 0;
  std::vector<int> oks(n, 0);
  oks[ctx->rank] = ok;
  if (sg_allgather_host(ctx, oks.data(), sizeof(int)) != SLAMGPU_OK) return;
  for (int r = 0; r < n; ++r) ok &= oks[r];
  std::vector<void *> peers(n, nullptr);
  if (ok) {
    for (int r = 0; r < n && ok; ++r) {
      if (r == ctx->rank) { peers[r] = ctx->mailbox; continue; }
      if (cudaIpcOpenMemHandle(&ctx->peer_mapped[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; (void)cudaGetLastError(); }
      peers[r] = ctx->peer_mapped[r];
    }
  }
  if (ok && (cudaMalloc(&ctx->d_peer_mailbox, sizeof(void *) * n) != cudaSuccess ||
             cudaMemcpy(ctx->d_peer_mailbox, peers.data(), sizeof(void *) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
             cudaMalloc(&ctx->d_p2p_status, sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_p2p_status, 0, sizeof(int)) != cudaSuccess))
    ok = 0;
  oks.assign(n, 0);
  oks[ctx->rank] = ok;
  if (sg_allgather_host(ctx, oks.data(), sizeof(int)) != SLAMGPU_OK) ok = 0;
  for (int r = 0; r < n; ++r) ok &= oks[r];
  if (!ok) sg_p2p_teardown(ctx);
}

void sg_p2p_teardown(slamgpu_ctx *ctx) {
  for (int r = 0; r < 64; ++r)
    if (ctx->peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->peer_mapped[r]); ctx->peer_mapped[r] = nullptr; }
  if (ctx->d_peer_mailbox) { cudaFree(ctx->d_peer_mailbox); ctx->d_peer_mailbox = nullptr; }
  if (ctx->d_p2p_status) { cudaFree(ctx->d_p2p_status); ctx->d_p2p_status = nullptr; }
  if (ctx->mailbox) { cudaFree(ctx->mailbox); ctx->mailbox = nullptr; }
  (void)cudaGetLastError();
}

extern "C" void slamgpu_ctx_destroy(slamgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  sg_p2p_teardown(ctx);
  if (ctx->comm) sg_nccl_destroy(ctx->comm);
  Candidates &c = ctx->cand;
  DevBuf *bufs[] = {&c.poses, &c.theta_id, &c.d_thetas, &c.d_xs, &c.d_ys, &c.groups, &c.cxp, &c.cyp, &c.cyw, &c.sm_cnt, &c.wtask, &c.colrec, &c.w2, &c.views, &c.view_id, &c.blocks, &c.blk_rows, &c.porg, &c.trc, &c.trs,
                    &c.scores, &c.blk_best, &c.result, &c.gm_pred, &c.gm_in, &c.gm_out, &ctx->flush, &ctx->gather};
  for (DevBuf *b : bufs) b->release();
  for (DevBuf &b : ctx->scratch) b.release();
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->evk0);
  cudaEventDestroy(ctx->evk1);
  if (ctx->ev_staged) cudaEventDestroy(ctx->ev_staged);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_join);
  cudaStreamDestroy(ctx->side);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *slamgpu_last_error(const slamgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

extern "C" int slamgpu_sync(slamgpu_ctx *ctx) {
  if (!ctx) return SLAMGPU_E_INVALID;
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}
extern "C" int slamgpu_timer_begin(slamgpu_ctx *ctx) {
  if (!ctx) return SLAMGPU_E_INVALID;
  SG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return SLAMGPU_OK;
}
extern "C" int slamgpu_timer_end(slamgpu_ctx *ctx, float *ms) {
  if (!ctx || !ms) return SLAMGPU_E_INVALID;
  SG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  SG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  SG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return SLAMGPU_OK;
}
extern "C" int slamgpu_last_kernel_ms(slamgpu_ctx *ctx, float *ms) {
  if (!ctx || !ms) return SLAMGPU_E_INVALID;
  if (!ctx->evk_valid) return sg_fail(ctx, SLAMGPU_E_STATE, "no dominant-kernel timing recorded yet");
  SG_CUDA(ctx, cudaEventSynchronize(ctx->evk1));
  SG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->evk0, ctx->evk1));
  return SLAMGPU_OK;
}
extern "C" int slamgpu_ctx_set_option(slamgpu_ctx *ctx, const char *name, int64_t value) {
  if (!ctx || !name) return SLAMGPU_E_INVALID;
  if (strcmp(name, "grid_variant") == 0) {
    if (value < 0 || value > 5) return sg_fail(ctx, SLAMGPU_E_INVALID, "grid_variant must be 0 (automatic) or 1..5");
    ctx->cand.user_variant = (int)value;
    return SLAMGPU_OK;
  }
  if (strcmp(name, "grid_rows") == 0) {  // rows per thread of the v2 grid kernel: 0 = automatic, or 2 / 4 / 8
    if (value != 0 && value != 2 && value != 4 && value != 8) return sg_fail(ctx, SLAMGPU_E_INVALID, "grid_rows must be 0, 2, 4 or 8");
    ctx->cand.user_rows = (int)value;
    return SLAMGPU_OK;
  }
  if (strcmp(name, "p2p_timeout_ms") == 0) {  // how long a rank waits for its peers' results in the fused exchange
    if (value < 1) return sg_fail(ctx, SLAMGPU_E_INVALID, "p2p_timeout_ms must be positive");
    ctx->p2p_timeout_ms = (double)value;
    return SLAMGPU_OK;
  }
  if (strcmp(name, "warm_l2") == 0) {  // experiment: 1 = stream the score LUT through L2 (side stream) before big grid launches
    ctx->cand.warm_l2 = value != 0;
    return SLAMGPU_OK;
  }
  return sg_fail(ctx, SLAMGPU_E_INVALID, "unknown option '%s'", name);
}
extern "C" int64_t slamgpu_launch_count(const slamgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

__global__ void k_flush_l2(float4 *p, size_t n, float v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = make_float4(v, v, v, v);
}
extern "C" int slamgpu_flush_l2(slamgpu_ctx *ctx) {
  if (!ctx) return SLAMGPU_E_INVALID;
  const size_t bytes = 256u << 20;  // > 126 MB L2
  if (ctx->flush.reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "flush buffer");
  k_flush_l2<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->flush.as<float4>(), bytes / 16, 1.0f);
  SG_CUDA(ctx, cudaGetLastError());
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ gather micro-benchmark (roofline context)
// Uniformly random 8-byte loads over a table the size of a score LUT: the rate at which the memory system serves
// gathers that share NOTHING (one 32-byte sector each).  K1's gathers share sectors between neighbouring candidates,
// so it may exceed this rate; bench.py reports both.
namespace {
__global__ void __launch_bounds__(256) k_gather_probe(const double *__restrict__ table, unsigned long long mask, int loads, double *sink) {
  unsigned long long x = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < loads; k += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
      acc[u] += __ldg(table + (x & mask));
    }
  }
  double t = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) t += acc[u];
  if (t == 123.456) sink[0] = t;  // keeps the loads alive
}
}  // namespace

extern "C" int slamgpu_probe_gather(slamgpu_ctx *ctx, int64_t table_bytes, int32_t loads_per_thread, double *gathers_per_s) {
  if (!ctx || !gathers_per_s || table_bytes < 4096 || loads_per_thread < 8) return sg_fail(ctx, SLAMGPU_E_INVALID, "probe_gather: bad argument");
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long n = 1;
  while (n * 2 * sizeof(double) <= (unsigned long long)table_bytes) n *= 2;  // power of two entries
  if (ctx->scratch[4].reserve(n * sizeof(double) + 64) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "probe table");
  SG_CUDA(ctx, cudaMemsetAsync(ctx->scratch[4].p, 0, n * sizeof(double) + 64, ctx->stream));
  const int blocks = ctx->sm_count * 8, loads = (loads_per_thread + 7) & ~7;
  double *table = ctx->scratch[4].as<double>();
  k_gather_probe<<<blocks, 256, 0, ctx->stream>>>(table, n - 1, loads, table + n);  // warm-up: the table becomes L2 resident
  SG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  k_gather_probe<<<blocks, 256, 0, ctx->stream>>>(table, n - 1, loads, table + n);
  SG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->launches += 2;
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  SG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *gathers_per_s = (double)blocks * 256.0 * loads / (ms * 1e-3);
  return SLAMGPU_OK;
}

int sg_allgather16(slamgpu_ctx *ctx, const void *d_send16, void *d_recv) {
  std::string err;
  int r = sg_nccl_allgather(ctx->comm, d_send16, d_recv, 16, ctx->stream, &err);
  if (r != SLAMGPU_OK) return sg_fail(ctx, r, "%s", err.c_str());
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ test hook
namespace {
__global__ void k_debug_div(const double *a, const double *b, int n, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sg::div_chain(a[i], b[i]);
}
}  // namespace

extern "C" int slamgpu_debug_div(slamgpu_ctx *ctx, int32_t n, const double *a, const double *b, double *out) {
  if (!ctx || n < 0 || (n > 0 && (!a || !b || !out))) return sg_fail(ctx, SLAMGPU_E_INVALID, "debug_div: bad argument");
  if (n == 0) return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(double) * (size_t)n;
  if (ctx->scratch[0].reserve(bytes) != SLAMGPU_OK || ctx->scratch[1].reserve(bytes) != SLAMGPU_OK || ctx->scratch[2].reserve(bytes) != SLAMGPU_OK)
    return sg_fail(ctx, SLAMGPU_E_NOMEM, "debug_div buffers");
  SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0].p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[1].p, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_debug_div<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch[0].as<double>(), ctx->scratch[1].as<double>(), n, ctx->scratch[2].as<double>());
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaMemcpyAsync(out, ctx->scratch[2].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

int sg_allgather_host(slamgpu_ctx *ctx, void *host, size_t chunk_bytes) {
  if (ctx->nranks <= 1 || chunk_bytes == 0) return SLAMGPU_OK;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t total = chunk_bytes * ctx->nranks, mine = chunk_bytes * ctx->rank;
  if (ctx->gather.reserve(total) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "all-gather staging");
  char *d = ctx->gather.as<char>();
  SG_CUDA(ctx, cudaMemcpyAsync(d + mine, (char *)host + mine, chunk_bytes, cudaMemcpyHostToDevice, ctx->stream));
  std::string err;
  int r = sg_nccl_allgather(ctx->comm, d + mine, d, chunk_bytes, ctx->stream, &err);  // in place
  if (r != SLAMGPU_OK) return sg_fail(ctx, r, "%s", err.c_str());
  SG_CUDA(ctx, cudaMemcpyAsync(host, d, total, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ map
extern "C" int slamgpu_model_stride(int model) { return sg::model_stride(model); }

// prototypes: test/core/mock_grid_cell.h:10-11, naive_grid_cells.h:8,27, tbm_grid_cells.h:10,
// slams/gmapping/gmapping_grid_cell.h:14
extern "C" void slamgpu_default_unknown(int model, double *r) {
  memset(r, 0, sizeof(double) * SLAMGPU_MAX_STRIDE);
  switch (model) {
    case SLAMGPU_CELL_LWW:
    case SLAMGPU_CELL_AFFINE:
    case SLAMGPU_CELL_MEAN: r[0] = 0.5; break;
    case SLAMGPU_CELL_TBM_CONSISTENT:
    case SLAMGPU_CELL_TBM_UNKNOWN_EVEN:
    case SLAMGPU_CELL_CREDIBILIST: r[0] = 0.5; r[1] = 1; r[2] = 1; break;
    case SLAMGPU_CELL_GMAPPING: r[0] = -1; break;
  }
}

struct RecParam { double v[SLAMGPU_MAX_STRIDE]; };

__global__ void k_fill_cells(double *cells, size_t n_cells, int stride, RecParam rec) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  size_t n = n_cells * stride;
  for (; i < n; i += st) cells[i] = rec.v[i % stride];
}

static int map_reset_tiles(slamgpu_map *m, int32_t w, int32_t h);
int sg_map_realloc(slamgpu_map *m, int32_t w, int32_t h) {
  slamgpu_ctx *ctx = m->ctx;
  if (m->pool) return map_reset_tiles(m, w, h);
  size_t need = (size_t)w * h * m->stride;
  if (need > m->cells_cap) {
    if (m->d_cells) cudaFree(m->d_cells);
    m->d_cells = nullptr; m->cells_cap = 0;
    SG_CUDA(ctx, cudaMalloc(&m->d_cells, std::max<size_t>(need, 1) * sizeof(double)));
    m->cells_cap = need;
  }
  m->w = w; m->h = h;
  m->pitch = (w + 2 * SG_LUT_PAD + 1) & ~1;
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}

void sg_map_invalidate_lut(slamgpu_map *m) { m->lut_valid[0] = m->lut_valid[1] = false; m->wlut_valid = false; }


// ------------------------------------------------------------------ copy-on-write tiles (see internal.h: SgTilePool)
struct TileCopy { const double *src; double *dst; };
__global__ void k_copy_tiles(const TileCopy *__restrict__ jobs, size_t tile_doubles) {
  const TileCopy j = jobs[blockIdx.y];
  const double2 *s = reinterpret_cast<const double2 *>(j.src);
  double2 *d = reinterpret_cast<double2 *>(j.dst);
  const size_t n2 = tile_doubles / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// dense [h][w][stride] <-> tiles
__global__ void k_tiles_gather(double *const *__restrict__ tiles, int tw, int w, int h, int stride, double *__restrict__ dense) {
  const size_t n = (size_t)w * h * stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = i / stride;
    const int k = (int)(i - cell * stride), y = (int)(cell / w), x = (int)(cell - (size_t)y * w);
    const double *t = tiles[(y >> SG_TILE_BITS) * tw + (x >> SG_TILE_BITS)];
    dense[i] = t[((size_t)(y & (SG_TILE - 1)) * SG_TILE + (x & (SG_TILE - 1))) * stride + k];
  }
}
__global__ void k_tiles_scatter(double *const *__restrict__ tiles, int tw, int w, int h, int stride, const double *__restrict__ dense) {
  const size_t n = (size_t)w * h * stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = i / stride;
    const int k = (int)(i - cell * stride), y = (int)(cell / w), x = (int)(cell - (size_t)y * w);
    double *t = tiles[(y >> SG_TILE_BITS) * tw + (x >> SG_TILE_BITS)];
    t[((size_t)(y & (SG_TILE - 1)) * SG_TILE + (x & (SG_TILE - 1))) * stride + k] = dense[i];
  }
}

static int pool_alloc(SgTilePool *pool, int32_t *id) {
  slamgpu_ctx *ctx = pool->ctx;
  if (pool->free_ids.empty()) {
    double *chunk = nullptr;
    SG_CUDA(ctx, cudaMalloc(&chunk, pool->tile_doubles * sizeof(double) * pool->tiles_per_chunk));
    const int32_t base = (int32_t)pool->chunks.size() * pool->tiles_per_chunk;
    pool->chunks.push_back(chunk);
    pool->refcnt.resize((size_t)base + pool->tiles_per_chunk, 0);
    for (int k = pool->tiles_per_chunk - 1; k >= 0; --k) pool->free_ids.push_back(base + k);
  }
  *id = pool->free_ids.back();
  pool->free_ids.pop_back();
  pool->refcnt[*id] = 1;
  ++pool->tiles_live;
  return SLAMGPU_OK;
}
static void pool_decref(SgTilePool *pool, int32_t id) {
  if (id == 0) return;  // the shared unknown tile lives as long as the pool
  if (--pool->refcnt[id] == 0) { pool->free_ids.push_back(id); --pool->tiles_live; }
}

int sg_pool_create(slamgpu_ctx *ctx, int model, const double *unknown_rec, SgTilePool **out) {
  SgTilePool *pool = new SgTilePool();
  pool->ctx = ctx; pool->stride = sg::model_stride(model);
  pool->tile_doubles = (size_t)SG_TILE * SG_TILE * pool->stride;
  int32_t id0 = -1;
  int r = pool_alloc(pool, &id0);  // tile 0: every cell unknown
  if (r != SLAMGPU_OK) { sg_pool_destroy(pool); return r; }
  RecParam rp;
  memset(rp.v, 0, sizeof rp.v);
  if (unknown_rec) memcpy(rp.v, unknown_rec, sizeof(double) * pool->stride);
  else slamgpu_default_unknown(model, rp.v);
  k_fill_cells<<<ctx->sm_count, 256, 0, ctx->stream>>>(pool->ptr(0), (size_t)SG_TILE * SG_TILE, pool->stride, rp);
  SG_LAUNCHED(ctx);
  pool->refcnt[0] = 1 << 30;
  *out = pool;
  return SLAMGPU_OK;
}
void sg_pool_destroy(SgTilePool *pool) {
  if (!pool) return;
  cudaStreamSynchronize(pool->ctx->stream);
  for (double *c : pool->chunks) cudaFree(c);
  pool->jobs.release();
  delete pool;
}

static void map_set_tile(slamgpu_map *m, size_t slot, int32_t id) {
  m->tile_ids[slot] = id;
  m->h_tile_ptrs[slot] = m->pool->ptr(id);
  m->tiles_dirty = true;
}
// new geometry, every tile unknown (the old tiles are released)
static int map_reset_tiles(slamgpu_map *m, int32_t w, int32_t h) {
  for (int32_t id : m->tile_ids) pool_decref(m->pool, id);
  m->w = w; m->h = h;
  m->tw = (w + SG_TILE - 1) >> SG_TILE_BITS; m->th = (h + SG_TILE - 1) >> SG_TILE_BITS;
  m->tile_ids.assign((size_t)m->tw * m->th, 0);
  m->h_tile_ptrs.assign((size_t)m->tw * m->th, m->pool->ptr(0));
  m->tiles_dirty = true;
  m->pitch = (w + 2 * SG_LUT_PAD + 1) & ~1;
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}
int sg_map_sync_tiles(slamgpu_map *m) {
  if (!m->pool || !m->tiles_dirty) return SLAMGPU_OK;
  slamgpu_ctx *ctx = m->ctx;
  const size_t n = m->h_tile_ptrs.size();
  if (n > m->d_tile_cap) {
    if (m->d_tile_ptrs) { SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(m->d_tile_ptrs); }
    m->d_tile_ptrs = nullptr; m->d_tile_cap = 0;
    SG_CUDA(ctx, cudaMalloc(&m->d_tile_ptrs, std::max<size_t>(n, 1) * sizeof(double *)));
    m->d_tile_cap = n;
  }
  if (n) SG_CUDA(ctx, cudaMemcpyAsync(m->d_tile_ptrs, m->h_tile_ptrs.data(), n * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
  m->tiles_dirty = false;
  return SLAMGPU_OK;
}
int sg_map_make_writable(slamgpu_map *m, int x0, int y0, int x1, int y1, std::vector<int32_t> *copies) {
  if (!m->pool || m->w <= 0 || m->h <= 0) return SLAMGPU_OK;
  x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, m->w - 1); y1 = std::min(y1, m->h - 1);
  for (int ty = y0 >> SG_TILE_BITS; ty <= (y1 >> SG_TILE_BITS); ++ty)
    for (int tx = x0 >> SG_TILE_BITS; tx <= (x1 >> SG_TILE_BITS); ++tx) {
      const size_t slot = (size_t)ty * m->tw + tx;
      const int32_t id = m->tile_ids[slot];
      if (id != 0 && m->pool->refcnt[id] == 1) continue;  // already private
      int32_t nid;
      SG_TRY(pool_alloc(m->pool, &nid));
      copies->push_back(id); copies->push_back(nid);
      pool_decref(m->pool, id);
      map_set_tile(m, slot, nid);
      ++m->pool->tiles_cloned;
    }
  return SLAMGPU_OK;
}
int sg_pool_run_copies(SgTilePool *pool, const std::vector<int32_t> &copies) {
  if (copies.empty()) return SLAMGPU_OK;
  slamgpu_ctx *ctx = pool->ctx;
  const size_t n = copies.size() / 2;
  std::vector<TileCopy> jobs(n);
  for (size_t k = 0; k < n; ++k) jobs[k] = TileCopy{pool->ptr(copies[2 * k]), pool->ptr(copies[2 * k + 1])};
  // (a source tile freed by the same batch keeps its content until something writes to the id again, which only happens
  // after these copies on the stream)
  DevBuf &buf = pool->jobs;
  if (buf.reserve(n * sizeof(TileCopy)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "tile copy list");
  SG_CUDA(ctx, cudaMemcpyAsync(buf.p, jobs.data(), n * sizeof(TileCopy), cudaMemcpyHostToDevice, ctx->stream));
  for (size_t k0 = 0; k0 < n; k0 += 65535) {
    const unsigned cnt = (unsigned)std::min<size_t>(65535, n - k0);
    k_copy_tiles<<<dim3(8, cnt), 256, 0, ctx->stream>>>(buf.as<TileCopy>() + k0, pool->tile_doubles);
    SG_LAUNCHED(ctx);
  }
  SG_CUDA(ctx, cudaGetLastError());
  return SLAMGPU_OK;
}
int sg_map_share_tiles(slamgpu_map *to, const slamgpu_map *from) {
  if (!to->pool || to->pool != from->pool) return sg_fail(to->ctx, SLAMGPU_E_STATE, "tile sharing needs two maps of one pool");
  for (int32_t id : from->tile_ids)
    if (id != 0) ++to->pool->refcnt[id];
  for (int32_t id : to->tile_ids) pool_decref(to->pool, id);
  to->w = from->w; to->h = from->h; to->ox = from->ox; to->oy = from->oy; to->tw = from->tw; to->th = from->th;
  to->tile_ids = from->tile_ids; to->h_tile_ptrs = from->h_tile_ptrs; to->tiles_dirty = true;
  to->pitch = from->pitch;
  sg_map_invalidate_lut(to);
  return SLAMGPU_OK;
}
int sg_map_gather_dense(slamgpu_map *m, double *d_dst) {
  slamgpu_ctx *ctx = m->ctx;
  SG_TRY(sg_map_sync_tiles(m));
  if ((size_t)m->w * m->h == 0) return SLAMGPU_OK;
  k_tiles_gather<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->d_tile_ptrs, m->tw, m->w, m->h, m->stride, d_dst);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  return SLAMGPU_OK;
}
int sg_map_scatter_dense(slamgpu_map *m, const double *d_src) {
  slamgpu_ctx *ctx = m->ctx;
  std::vector<int32_t> copies;
  SG_TRY(sg_map_make_writable(m, 0, 0, m->w - 1, m->h - 1, &copies));  // (no need to run the copies: every cell is overwritten...
  // ... except the cells of edge tiles past w / h, which nothing ever reads)
  SG_TRY(sg_map_sync_tiles(m));
  if ((size_t)m->w * m->h == 0) return SLAMGPU_OK;
  k_tiles_scatter<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->d_tile_ptrs, m->tw, m->w, m->h, m->stride, d_src);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  sg_map_invalidate_lut(m);
  return SLAMGPU_OK;
}
int sg_map_create_tiled(slamgpu_ctx *ctx, SgTilePool *pool, int32_t w, int32_t h, double scale, int32_t model, int32_t grow,
                        const double *unknown_rec, slamgpu_map **out) {
  if (!ctx || !pool || !out || w < 0 || h < 0 || !(scale > 0)) return sg_fail(ctx, SLAMGPU_E_INVALID, "sg_map_create_tiled: bad arguments");
  slamgpu_map *m = new slamgpu_map();
  m->ctx = ctx; m->scale = scale; m->model = model; m->grow = grow; m->pool = pool;
  m->stride = sg::model_stride(model);
  m->ox = w / 2; m->oy = h / 2;
  if (unknown_rec) memcpy(m->unknown, unknown_rec, sizeof(double) * m->stride);
  else slamgpu_default_unknown(model, m->unknown);
  map_reset_tiles(m, w, h);
  *out = m;
  return SLAMGPU_OK;
}
// dense staging of a tiled map for the slow paths (download, LUT build): scratch[5] holds [h][w][stride]
static int map_dense_view(slamgpu_map *m, const double **cells) {
  if (!m->pool) { *cells = m->d_cells; return SLAMGPU_OK; }
  slamgpu_ctx *ctx = m->ctx;
  const size_t bytes = std::max<size_t>((size_t)m->w * m->h * m->stride, 1) * sizeof(double);
  if (ctx->scratch[5].reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "dense staging of a tiled map (%zu bytes)", bytes);
  SG_TRY(sg_map_gather_dense(m, ctx->scratch[5].as<double>()));
  *cells = ctx->scratch[5].as<double>();
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_create(slamgpu_ctx *ctx, int32_t w, int32_t h, double scale, int32_t model, int32_t grow,
                                  const double *unknown_rec, slamgpu_map **out) {
  if (!ctx || !out) return SLAMGPU_E_INVALID;
  *out = nullptr;
  if (w < 0 || h < 0 || !(scale > 0) || model < 0 || model >= SLAMGPU_CELL_MODELS || grow < 0 || grow > SLAMGPU_GROW_TILED)
    return sg_fail(ctx, SLAMGPU_E_INVALID, "slamgpu_map_create: bad arguments (w=%d h=%d scale=%g model=%d grow=%d)", w, h,
                   scale, model, grow);
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  slamgpu_map *m = new slamgpu_map();
  m->ctx = ctx; m->scale = scale; m->model = model; m->grow = grow;
  m->stride = sg::model_stride(model);
  m->ox = w / 2; m->oy = h / 2;  // regular_squares_grid.h:120-122
  if (unknown_rec) memcpy(m->unknown, unknown_rec, sizeof(double) * m->stride);
  else slamgpu_default_unknown(model, m->unknown);
  int r = sg_map_realloc(m, w, h);
  if (r != SLAMGPU_OK) { slamgpu_map_destroy(m); return r; }
  RecParam rp;
  memcpy(rp.v, m->unknown, sizeof rp.v);
  if ((size_t)w * h > 0) {
    k_fill_cells<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(m->d_cells, (size_t)w * h, m->stride, rp);
    SG_LAUNCHED(ctx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { slamgpu_map_destroy(m); return sg_fail(ctx, SLAMGPU_E_CUDA, "k_fill_cells: %s", cudaGetErrorString(e)); }
  }
  *out = m;
  return SLAMGPU_OK;
}

extern "C" void slamgpu_map_destroy(slamgpu_map *m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  if (m->d_cells) cudaFree(m->d_cells);
  if (m->pool) {
    for (int32_t id : m->tile_ids) pool_decref(m->pool, id);
    if (m->d_tile_ptrs) cudaFree(m->d_tile_ptrs);
  }
  for (int i = 0; i < 2; ++i)
    if (m->d_lut[i]) cudaFree(m->d_lut[i]);
  if (m->d_wlut) cudaFree(m->d_wlut);
  delete m;
}

extern "C" int slamgpu_map_info(const slamgpu_map *m, int32_t *w, int32_t *h, double *scale, int32_t *ox, int32_t *oy,
                                int32_t *stride) {
  if (!m) return SLAMGPU_E_INVALID;
  if (w) *w = m->w;
  if (h) *h = m->h;
  if (scale) *scale = m->scale;
  if (ox) *ox = m->ox;
  if (oy) *oy = m->oy;
  if (stride) *stride = m->stride;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_upload(slamgpu_map *m, const double *cells, int32_t w, int32_t h, int32_t ox, int32_t oy) {
  if (!m || !cells || w < 0 || h < 0) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  SG_TRY(sg_map_realloc(m, w, h));
  m->ox = ox; m->oy = oy;
  size_t bytes = (size_t)w * h * m->stride * sizeof(double);
  if (m->pool) {
    if (ctx->scratch[5].reserve(std::max<size_t>(bytes, 16)) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "upload staging");
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[5].p, cells, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SG_TRY(sg_map_scatter_dense(m, ctx->scratch[5].as<double>()));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMGPU_OK;
  }
  SG_CUDA(ctx, cudaMemcpyAsync(m->d_cells, cells, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_download(slamgpu_map *m, double *cells) {
  if (!m || !cells) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  size_t bytes = (size_t)m->w * m->h * m->stride * sizeof(double);
  const double *src = nullptr;
  SG_TRY(map_dense_view(m, &src));
  SG_CUDA(ctx, cudaMemcpyAsync(cells, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_read_cell(slamgpu_map *m, int32_t x, int32_t y, double *rec) {
  if (!m || !rec) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  int ix = x + m->ox, iy = y + m->oy;
  if (ix < 0 || ix >= m->w || iy < 0 || iy >= m->h) {  // plain_grid_map.h:69-73: the unknown cell outside
    memcpy(rec, m->unknown, sizeof(double) * m->stride);
    return SLAMGPU_OK;
  }
  const double *cell = m->pool ? m->pool->ptr(m->tile_ids[(size_t)(iy >> SG_TILE_BITS) * m->tw + (ix >> SG_TILE_BITS)]) +
                                   ((size_t)(iy & (SG_TILE - 1)) * SG_TILE + (ix & (SG_TILE - 1))) * m->stride
                             : m->d_cells + ((size_t)iy * m->w + ix) * m->stride;
  SG_CUDA(ctx, cudaMemcpyAsync(rec, cell, sizeof(double) * m->stride, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ score LUT
__global__ void k_build_lut(const double *__restrict__ cells, int w, int h, int stride, int model, int oie,
                            double *__restrict__ lut, int pitch, double unknown_value) {
  int px = blockIdx.x * blockDim.x + threadIdx.x;  // padded coordinates
  int py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= pitch || py >= h + 2 * SG_LUT_PAD) return;
  int x = px - SG_LUT_PAD, y = py - SG_LUT_PAD;
  double v = unknown_value;
  if (x >= 0 && x < w && y >= 0 && y < h) {
    double r[SLAMGPU_MAX_STRIDE];
    const double *src = cells + ((size_t)y * w + x) * stride;
    for (int k = 0; k < stride; ++k) r[k] = src[k];
    v = sg::cell_impact(model, oie, r, 0.0, 0.0);
  }
  lut[(size_t)py * pitch + px] = v;
}

__global__ void k_unknown_impact(int model, int oie, RecParam rec, double *out) {
  *out = sg::cell_impact(model, oie, rec.v, 0.0, 0.0);
}

int sg_map_ensure_lut(slamgpu_map *m, int oie) {
  slamgpu_ctx *ctx = m->ctx;
  if (oie < 0 || oie > 1) return sg_fail(ctx, SLAMGPU_E_INVALID, "bad oie %d", oie);
  if (m->lut_valid[oie]) return SLAMGPU_OK;
  // SG_LUT_SLACK_ROWS rows + SG_LUT_SLACK doubles past the padded LUT: k_score_grid4 loads up to 7 rows below the
  // one it needs, a staged patch row (v3) may run past the last row; the values are never used
  const size_t rows_used = (size_t)m->pitch * (m->h + 2 * SG_LUT_PAD);
  size_t need = rows_used + (size_t)m->pitch * SG_LUT_SLACK_ROWS + SG_LUT_SLACK;
  if (need > m->lut_cap[oie]) {
    if (m->d_lut[oie]) cudaFree(m->d_lut[oie]);
    m->d_lut[oie] = nullptr; m->lut_cap[oie] = 0;
    SG_CUDA(ctx, cudaMalloc(&m->d_lut[oie], need * sizeof(double)));
    SG_CUDA(ctx, cudaMemsetAsync(m->d_lut[oie] + rows_used, 0, (need - rows_used) * sizeof(double), ctx->stream));
    m->lut_cap[oie] = need;
  }
  // the unknown cell's impact, computed by the same device code as every other cell
  double *d_tmp = nullptr;
  SG_TRY(ctx->scratch[7].reserve(64));
  d_tmp = ctx->scratch[7].as<double>();
  RecParam rp;
  memcpy(rp.v, m->unknown, sizeof rp.v);
  k_unknown_impact<<<1, 1, 0, ctx->stream>>>(m->model, oie, rp, d_tmp);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaMemcpyAsync(&m->unknown_lut[oie], d_tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dim3 blk(32, 8), grd((m->pitch + 31) / 32, (m->h + 2 * SG_LUT_PAD + 7) / 8);
  const double *dense = nullptr;
  SG_TRY(map_dense_view(m, &dense));
  k_build_lut<<<grd, blk, 0, ctx->stream>>>(dense, m->w, m->h, m->stride, m->model, oie, m->d_lut[oie], m->pitch,
                                             m->unknown_lut[oie]);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  m->lut_valid[oie] = true;
  return SLAMGPU_OK;
}

extern "C" int slamgpu_map_lut_download(slamgpu_map *m, int32_t oie, double *lut, double *unknown_value) {
  if (!m) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  SG_TRY(sg_map_ensure_lut(m, oie));
  if (lut && m->w > 0 && m->h > 0)
    SG_CUDA(ctx, cudaMemcpy2DAsync(lut, (size_t)m->w * sizeof(double),
                                   m->d_lut[oie] + (size_t)SG_LUT_PAD * m->pitch + SG_LUT_PAD,
                                   (size_t)m->pitch * sizeof(double), (size_t)m->w * sizeof(double), m->h,
                                   cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (unknown_value) *unknown_value = m->unknown_lut[oie];
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ scan
extern "C" int slamgpu_scan_create(slamgpu_ctx *ctx, slamgpu_scan **out) {
  if (!ctx || !out) return SLAMGPU_E_INVALID;
  slamgpu_scan *s = new slamgpu_scan();
  s->ctx = ctx;
  *out = s;
  return SLAMGPU_OK;
}
extern "C" void slamgpu_scan_destroy(slamgpu_scan *s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->ctx->cand.scan == s) { s->ctx->cand.scan = nullptr; s->ctx->cand.kind = -1; }
  s->d.release();
  delete s;
}

extern "C" int slamgpu_scan_upload(slamgpu_scan *s, int32_t n, int32_t cartesian, const double *a, const double *b,
                                   const uint8_t *occ, const double *factor, const double *weight) {
  if (!s || n < 0 || (n > 0 && (!a || !b))) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = s->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  s->n = n; s->cartesian = cartesian; s->has_factor = factor != nullptr;
  s->range.resize(n); s->angle.resize(n); s->x.resize(n); s->y.resize(n); s->weight.resize(n); s->factor.resize(n);
  s->occ.resize(n);
  // ScanPoint2D accessors, src/core/states/sensor_data.h:47-70: the representation that was
  // not given is derived with libm on the host (bit-identical to the reference's own calls)
  for (int i = 0; i < n; ++i) {
    if (cartesian) {
      s->x[i] = a[i]; s->y[i] = b[i];
      s->range[i] = std::sqrt(a[i] * a[i] + b[i] * b[i]);
      s->angle[i] = std::atan2(b[i], a[i]);
    } else {
      s->range[i] = a[i]; s->angle[i] = b[i];
      s->x[i] = NAN; s->y[i] = NAN;  // derived on demand (sg_scan_ensure_xy): 2 libm calls per point nobody may need
    }
    s->weight[i] = weight ? weight[i] : 1.0 / n;
    s->factor[i] = factor ? factor[i] : 1.0;
    s->occ[i] = occ ? occ[i] : 1;
  }
  s->xy_valid = cartesian != 0;
  double ws = 0;
  for (int i = 0; i < n; ++i) ws += s->weight[i];  // weighted_mean_point_probability_spe.h:125
  s->wsum = ws;
  size_t nn = std::max(n, 1);
  size_t bytes = nn * 6 * sizeof(double) + nn;
  if (s->d.reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "scan buffer");
  void *hp;
  SG_TRY(sg_pinned(ctx, bytes, &hp));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging may still be in flight
  double *h = (double *)hp;
  memcpy(h + 0 * nn, s->range.data(), n * sizeof(double));
  memcpy(h + 1 * nn, s->angle.data(), n * sizeof(double));
  memcpy(h + 2 * nn, s->x.data(), n * sizeof(double));
  memcpy(h + 3 * nn, s->y.data(), n * sizeof(double));
  memcpy(h + 4 * nn, s->weight.data(), n * sizeof(double));
  memcpy(h + 5 * nn, s->factor.data(), n * sizeof(double));
  memcpy(h + 6 * nn, s->occ.data(), n);
  double *d = s->d.as<double>();
  s->d_range = d; s->d_angle = d + nn; s->d_x = d + 2 * nn; s->d_y = d + 3 * nn; s->d_w = d + 4 * nn; s->d_f = d + 5 * nn;
  s->d_occ = (uint8_t *)(d + 6 * nn);
  SG_CUDA(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->cand.scan == s) ctx->cand.kind = -1;  // staged candidates depend on the scan
  return SLAMGPU_OK;
}

// x = r*cos(a), y = r*sin(a) of a polar scan (ScanPoint2D::x / y, sensor_data.h:47-70), for the paths that read points
int sg_scan_ensure_xy(slamgpu_scan *s) {
  if (!s || s->xy_valid) return SLAMGPU_OK;
  slamgpu_ctx *ctx = s->ctx;
  const int n = s->n;
  for (int i = 0; i < n; ++i) {
    s->x[i] = s->range[i] * std::cos(s->angle[i]);
    s->y[i] = s->range[i] * std::sin(s->angle[i]);
  }
  s->xy_valid = true;
  if (n > 0) {
    SG_CUDA(ctx, cudaSetDevice(ctx->device));
    SG_CUDA(ctx, cudaMemcpyAsync(s->d_x, s->x.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(s->d_y, s->y.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  return SLAMGPU_OK;
}

// Many Cartesian copies of one scan at once (the pre-rotated scans of the multi-resolution matcher): only what the
// window kernels read is filled in -- x, y, weight, factor 1, occupied -- so no atan2/sqrt per point, one staging
// buffer, one synchronisation.  xs / ys hold `count` rows of n values.
int sg_scans_upload_xy(slamgpu_ctx *ctx, slamgpu_scan *const *scans, int count, int32_t n, const double *xs, const double *ys,
                       const double *weight) {
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t nn = std::max(n, 1);
  const size_t bytes = nn * 6 * sizeof(double) + ((nn + 7) & ~(size_t)7);
  void *hp;
  SG_TRY(sg_pinned(ctx, bytes * count, &hp));
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging may still be in flight
  std::vector<double> w(nn, 0.0);
  for (int i = 0; i < n; ++i) w[i] = weight ? weight[i] : 1.0 / n;
  double ws = 0;
  for (int i = 0; i < n; ++i) ws += w[i];  // weighted_mean_point_probability_spe.h:125
  for (int k = 0; k < count; ++k) {
    slamgpu_scan *s = scans[k];
    if (!s || s->ctx != ctx) return sg_fail(ctx, SLAMGPU_E_INVALID, "scan %d is NULL or belongs to another ctx", k);
    s->n = n; s->cartesian = 1; s->has_factor = false; s->wsum = ws;
    s->x.assign(xs + (size_t)k * n, xs + (size_t)(k + 1) * n); s->y.assign(ys + (size_t)k * n, ys + (size_t)(k + 1) * n);
    s->range.assign(n, NAN); s->angle.assign(n, NAN);  // not derived: these copies are only ever read as points
    s->weight.assign(w.begin(), w.begin() + n); s->factor.assign(n, 1.0); s->occ.assign(n, 1);
    if (s->d.reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "scan buffer");
    double *h = (double *)((char *)hp + bytes * k);
    for (size_t i = 0; i < 2 * nn; ++i) h[i] = NAN;
    memcpy(h + 2 * nn, s->x.data(), n * sizeof(double));
    memcpy(h + 3 * nn, s->y.data(), n * sizeof(double));
    memcpy(h + 4 * nn, w.data(), n * sizeof(double));
    for (size_t i = 0; i < nn; ++i) h[5 * nn + i] = 1.0;
    memset(h + 6 * nn, 1, nn);
    double *d = s->d.as<double>();
    s->d_range = d; s->d_angle = d + nn; s->d_x = d + 2 * nn; s->d_y = d + 3 * nn; s->d_w = d + 4 * nn; s->d_f = d + 5 * nn;
    s->d_occ = (uint8_t *)(d + 6 * nn);
    SG_CUDA(ctx, cudaMemcpyAsync(d, h, nn * 6 * sizeof(double) + nn, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->cand.scan == s) ctx->cand.kind = -1;
  }
  return SLAMGPU_OK;
}

// a score-only snapshot of a host-side map: the caller evaluated its own ObservationImpactEstimator per
// cell (any cell class works), the kernels only ever gather this LUT
__global__ void k_pad_lut(const double *__restrict__ src, int w, int h, double *__restrict__ lut, int pitch, double unknown_value) {
  int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= pitch || py >= h + 2 * SG_LUT_PAD) return;
  int x = px - SG_LUT_PAD, y = py - SG_LUT_PAD;
  lut[(size_t)py * pitch + px] = (x >= 0 && x < w && y >= 0 && y < h) ? src[(size_t)y * w + x] : unknown_value;
}

extern "C" int slamgpu_map_upload_lut(slamgpu_map *m, int32_t oie, const double *lut, double unknown_value, int32_t w, int32_t h,
                                      int32_t ox, int32_t oy) {
  if (!m || !lut || w < 0 || h < 0 || oie < 0 || oie > 1) return SLAMGPU_E_INVALID;
  slamgpu_ctx *ctx = m->ctx;
  SG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (w != m->w || h != m->h) {
    SG_TRY(sg_map_realloc(m, w, h));
    // cell records are not part of a score-only snapshot: keep them at the unknown prototype
    RecParam rp;
    memcpy(rp.v, m->unknown, sizeof rp.v);
    if ((size_t)w * h > 0) {
      k_fill_cells<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(m->d_cells, (size_t)w * h, m->stride, rp);
      SG_LAUNCHED(ctx);
    }
  }
  m->ox = ox; m->oy = oy;
  // SG_LUT_SLACK_ROWS rows + SG_LUT_SLACK doubles past the padded LUT: k_score_grid4 loads up to 7 rows below the
  // one it needs, a staged patch row (v3) may run past the last row; the values are never used
  const size_t rows_used = (size_t)m->pitch * (m->h + 2 * SG_LUT_PAD);
  size_t need = rows_used + (size_t)m->pitch * SG_LUT_SLACK_ROWS + SG_LUT_SLACK;
  if (need > m->lut_cap[oie]) {
    if (m->d_lut[oie]) cudaFree(m->d_lut[oie]);
    m->d_lut[oie] = nullptr; m->lut_cap[oie] = 0;
    SG_CUDA(ctx, cudaMalloc(&m->d_lut[oie], need * sizeof(double)));
    SG_CUDA(ctx, cudaMemsetAsync(m->d_lut[oie] + rows_used, 0, (need - rows_used) * sizeof(double), ctx->stream));
    m->lut_cap[oie] = need;
  }
  size_t bytes = std::max<size_t>((size_t)w * h, 1) * sizeof(double);
  if (ctx->scratch[7].reserve(bytes) != SLAMGPU_OK) return sg_fail(ctx, SLAMGPU_E_NOMEM, "LUT staging");
  SG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[7].p, lut, (size_t)w * h * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grd((m->pitch + 31) / 32, (m->h + 2 * SG_LUT_PAD + 7) / 8);
  k_pad_lut<<<grd, blk, 0, ctx->stream>>>(ctx->scratch[7].as<double>(), w, h, m->d_lut[oie], m->pitch, unknown_value);
  SG_LAUNCHED(ctx);
  SG_CUDA(ctx, cudaGetLastError());
  SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->unknown_lut[oie] = unknown_value;
  m->lut_valid[oie] = true;
  m->lut_valid[1 - oie] = false;
  return SLAMGPU_OK;
}

// ------------------------------------------------------------------ per-scan host preparation
// (once per scan, O(n), libm; the reference does the same work on the host before its hot loops)
namespace {
bool h_less(double a, double b) { return a < b + 2.220446049250313e-16; }

// AngleHistogram (src/core/features/angle_histogram.h:9-102, 20 bins): occupancy of the bin each
// point's direction-to-previous-point falls into; the first point reports n
void angle_histogram_values(int n, const double *range, const double *angle, std::vector<uint32_t> &values) {
  const int NB = 20;
  unsigned hist[NB] = {0};
  const double step = (180 * M_PI / 180) / NB;
  std::vector<int> bin(std::max(n, 1), 0);
  for (int i = 1; i < n; ++i) {
    const double bx = range[i - 1] * std::cos(angle[i - 1]), by = range[i - 1] * std::sin(angle[i - 1]);
    const double x = range[i] * std::cos(angle[i]), y = range[i] * std::sin(angle[i]);
    const double d_x = x - bx, d_y = y - by;
    double a = 0;
    if (d_y != 0) {
      const double d_d = std::sqrt(d_x * d_x + d_y * d_y);
      a = std::acos(d_x / d_d);
      if (d_y < 0 && d_x != 0) a = M_PI - a;
    }
    bin[i] = (int)(size_t)std::floor(a / step);
    hist[bin[i]]++;
  }
  values.resize(n);
  for (int i = 0; i < n; ++i) values[i] = i == 0 ? (uint32_t)n : hist[bin[i]];
}
}  // namespace

extern "C" int slamgpu_scan_filter(const slamgpu_map *m, int32_t n, const double *range, const double *angle, const uint8_t *occ,
                                   const double pose[3], uint32_t skip_rate, double max_range, int32_t *keep_idx) {
  if (!m || n < 0 || (n > 0 && (!range || !angle || !keep_idx)) || !pose) return SLAMGPU_E_INVALID;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (skip_rate && (unsigned)i % skip_rate) continue;
    const double wx = pose[0] + range[i] * std::cos(pose[2] + angle[i]);
    const double wy = pose[1] + range[i] * std::sin(pose[2] + angle[i]);
    const int cx = (int)std::floor(wx / m->scale), cy = (int)std::floor(wy / m->scale);
    bool has_cell = true;  // unbounded maps own every cell (plain_grid_map.h:77)
    if (m->grow == SLAMGPU_GROW_NONE) {
      const int ix = cx + m->ox, iy = cy + m->oy;
      has_cell = ix >= 0 && ix < m->w && iy >= 0 && iy < m->h;
    }
    const bool skip = (occ && !occ[i]) || !has_cell || (h_less(0.0, max_range) && h_less(max_range, range[i]));
    if (!skip) keep_idx[k++] = i;
  }
  return k;
}

extern "C" int slamgpu_point_weights(int32_t kind, int32_t n, const double *range, const double *angle, double *out_w) {
  if (n < 0 || (n > 0 && (!range || !angle || !out_w))) return SLAMGPU_E_INVALID;
  if (kind == SLAMGPU_SPW_EVEN) {
    const double c = 1.0 / n;
    for (int i = 0; i < n; ++i) out_w[i] = c;
  } else if (kind == SLAMGPU_SPW_VINY) {
    for (int i = 0; i < n; ++i) {
      const double a = angle[i];
      double w = std::fabs(std::sin(a)) + std::fabs(std::cos(a));
      if (0.9 < std::fabs(std::cos(a))) w = 3;
      else if (0.8 < std::fabs(std::cos(a))) w = 2;
      out_w[i] = w * std::sqrt(range[i]);
    }
  } else if (kind == SLAMGPU_SPW_AHR) {
    std::vector<uint32_t> v;
    angle_histogram_values(n, range, angle, v);
    for (int i = 0; i < n; ++i) out_w[i] = 1.0 / v[i];
  } else {
    return SLAMGPU_E_INVALID;
  }
  return SLAMGPU_OK;
}

extern "C" int slamgpu_mapping_quality(int32_t kind, int32_t n, const double *range, const double *angle, double *out_q) {
  if (n < 0 || (n > 0 && (!range || !angle || !out_q))) return SLAMGPU_E_INVALID;
  if (kind == SLAMGPU_OMQE_IDLE) {
    for (int i = 0; i < n; ++i) out_q[i] = 1.0;
  } else if (kind == SLAMGPU_OMQE_AHR) {
    std::vector<uint32_t> v;
    angle_histogram_values(n, range, angle, v);
    for (int i = 0; i < n; ++i) out_q[i] = 1.0 / v[i];
  } else {
    return SLAMGPU_E_INVALID;
  }
  return SLAMGPU_OK;
}
