// tma_rate_probe.cu -- how fast can an SM issue small 2-D tensor-map loads?  Every warp keeps RING boxes of BW x BH doubles in
// flight (one elected lane issues, completion on a per-slot mbarrier) and walks the map like the grid kernel does (a few cells
// per step).  Prints cycles per box and SM for several warps-per-SM counts.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RING, bool TEST = false>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int iters, int box_bytes, int W, int H, double *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar[32][RING];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned stage_bytes = (box_bytes + 127) & ~127;
  unsigned char *my = sm + (size_t)w * RING * stage_bytes;
  if (lane == 0) for (int s = 0; s < RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[w][s])), "r"(1));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  int x = ((blockIdx.x * 37 + w * 101) % (W - 64)) & ~1, y = (blockIdx.x * 53 + w * 17) % (H - 16);
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters + RING; ++i) {
    const int s = i % RING;
    if (i >= RING) {  // consume the box issued RING iterations ago
      unsigned ok = 0;
      const unsigned par = ((i / RING) - 1) & 1;
      if (TEST) { while (!ok) asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[w][s])), "r"(par) : "memory"); }
      else while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[w][s])), "r"(par) : "memory");
      acc += ((const double *)(my + s * stage_bytes))[lane % (box_bytes / 8)];
      __syncwarp();
    }
    if (i < iters && lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[w][s])), "r"(box_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(my + s * stage_bytes)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(&bar[w][s])), "r"(x), "r"(y) : "memory");
    }
    x += 2; if (x > W - 64) x -= W - 64;  // the box start has to be 16-byte aligned: even columns only
    y += (i & 3) == 0; if (y > H - 16) y -= H - 16;
  }
  long long t1 = clock64();
  if (acc == 1.2345e-300) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int RING, int NB>
__global__ void k2(const __grid_constant__ CUtensorMap tmap, int iters, int box_bytes, int W, int H, double *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar[32][RING];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned stage_bytes = (box_bytes + 127) & ~127;
  unsigned char *my = sm + (size_t)w * RING * stage_bytes;
  if (lane == 0) for (int s = 0; s < RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[w][s])), "r"(1));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  int x = ((blockIdx.x * 37 + w * 101) % (W - 64)) & ~1, y = (blockIdx.x * 53 + w * 17) % (H - 16);
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters + RING; i += NB) {
    if (i >= RING) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int s = (i + b) % RING;
        unsigned ok = 0;
        const unsigned par = (((i + b) / RING) - 1) & 1;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[w][s])), "r"(par) : "memory");
        acc += ((const double *)(my + s * stage_bytes))[lane];
      }
      __syncwarp();
    }
    if (i < iters && lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int s = (i + b) % RING;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[w][s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(my + s * stage_bytes)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(&bar[w][s])), "r"(x + 2 * b), "r"(y) : "memory");
      }
    }
    x += 2 * NB; if (x > W - 64) x -= W - 64;
    y += (i & 3) == 0; if (y > H - 16) y -= H - 16;
  }
  long long t1 = clock64();
  if (acc == 1.2345e-300) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  const int W = 2002, H = 2002;
  double *d, *o; long long *c;
  cudaMalloc(&d, sizeof(double) * (size_t)W * (H + 8)); cudaMemset(d, 0, sizeof(double) * (size_t)W * (H + 8));
  cudaMalloc(&o, 64); cudaMalloc(&c, 8);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*E)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  const int iters = 2000;
  struct B { int bw, bh; } boxes[] = {{14, 4}, {16, 16}, {44, 10}};
  for (const B &b : boxes) {
    CUtensorMap tm;
    cuuint64_t gd[2] = {(cuuint64_t)W, (cuuint64_t)H}, gs[1] = {(cuuint64_t)W * 8};
    cuuint32_t box[2] = {(cuuint32_t)b.bw, (cuuint32_t)b.bh}, es[2] = {1, 1};
    CUresult r = ((E)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int box_bytes = b.bw * b.bh * 8, stage = (box_bytes + 127) & ~127;
    auto run = [&](auto kern, int ring, int warps, int blocks_per_sm) {
      if ((size_t)warps * ring * stage * blocks_per_sm > 200 * 1024) return true;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * ring * stage);
      kern<<<148 * blocks_per_sm, warps * 32, warps * ring * stage>>>(tm, iters, box_bytes, W, H, o, c);
      cudaError_t e = cudaDeviceSynchronize();
      long long cy = 0; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
      const int wsm = warps * blocks_per_sm;
      printf("box %2dx%-2d (%4d B) %2d warps/SM, %d boxes in flight per warp: %s  %.1f cycles per box and warp, %.1f cycles per box and SM, %.1f B/cycle/SM\n",
             b.bw, b.bh, box_bytes, wsm, ring, cudaGetErrorString(e), (double)cy / iters, (double)cy / iters / wsm, box_bytes * (double)wsm * iters / cy);
      return e == cudaSuccess;
    };
    for (int warps : {1, 8, 16, 32}) {
      if (!run(k<2>, 2, warps, 1) || !run(k<4>, 4, warps, 1) || !run(k<8>, 8, warps, 1)) return 1;
    }
    if (!run(k<2>, 2, 32, 2) || !run(k<4>, 4, 32, 2)) return 1;
    printf("-- the same with mbarrier.test_wait (no suspend) in place of try_wait\n");
    for (int warps : {1, 8, 32}) if (!run(k<4, true>, 4, warps, 1) || !run(k<8, true>, 8, warps, 1)) return 1;
    if (!run(k<4, true>, 4, 32, 2)) return 1;
    printf("-- NB boxes issued back to back per loop trip (cycles are per box)\n");
    for (int warps : {1, 8, 32}) if (!run(k2<8, 1>, 8, warps, 1) || !run(k2<8, 2>, 8, warps, 1) || !run(k2<8, 4>, 8, warps, 1)) return 1;
  }
  return 0;
}
