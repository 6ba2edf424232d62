#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdlib>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap_param, const CUtensorMap *tmap_glob, int c0, int c1, int bw, int bh, double *out) {
  const CUtensorMap *tmp = tmap_glob ? tmap_glob : &tmap_param;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    printf("smem base %u align %u\n", smem_u32(sm), smem_u32(sm) & 127);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * bh * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sm)), "l"(tmp), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  const double *p = (const double *)sm;
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = p[i];
}
int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 202, H = 202, bw = argc > 2 ? atoi(argv[2]) : 12, bh = argc > 3 ? atoi(argv[3]) : 9;
  std::vector<double> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = i;
  double *d, *o;
  cudaMalloc(&d, sizeof(double) * W * H); cudaMalloc(&o, sizeof(double) * bw * bh);
  cudaMemcpy(d, h.data(), sizeof(double) * W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*E)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  // mode 0: FLOAT64, 1: UINT64, 2: FLOAT32 with doubled inner extent, 3: UINT8 with 8x inner extent
  const int mul = mode == 2 ? 2 : (mode == 3 ? 8 : 1);
  CUtensorMapDataType ty = mode == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : mode == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : mode == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  cuuint64_t gd[2] = {(cuuint64_t)W * mul, H}, gs[1] = {W * 8}; cuuint32_t box[2] = {(cuuint32_t)bw * mul, (cuuint32_t)bh}, es[2] = {1, 1};
  for (int dt = 0; dt < 1; ++dt) {
    CUresult r = ((E)fn)(&tm, ty, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d box %dx%d encode -> %d\n", mode, bw, bh, (int)r);
    CUtensorMap *dtm = nullptr; if (argc > 4) { cudaMalloc(&dtm, sizeof(CUtensorMap)); cudaMemcpy(dtm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice); }
    k<<<1, 128, bw * bh * 8 + 128>>>(tm, dtm, 5 * mul, 7, bw, bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> r2(bw * bh);
    cudaMemcpy(r2.data(), o, sizeof(double) * bw * bh, cudaMemcpyDeviceToHost);
    printf("got %g %g %g (expect %d %d %d)\n", r2[0], r2[1], r2[bw], 7 * W + 5, 7 * W + 6, 8 * W + 5);
  }
  return 0;
}
