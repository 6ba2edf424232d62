// tma_probe2.cu -- canonical 2-D tensor-map load (cp.async.bulk.tensor.2d, SASS UTMALDG), the CUTLASS way: the tensor map
// is a __grid_constant__ kernel parameter whose address goes to the instruction unchanged; one thread issues the copy.
// Prints, for a few element types / box shapes, whether the patch arrived.  (tma_probe.cu selected between a parameter and a
// global copy of the map at run time, which makes nvcc spill the parameter to local memory -- TMA cannot read that.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, int n8, double *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  const double *p = (const double *)sm;
  for (int i = threadIdx.x; i < n8; i += blockDim.x) out[i] = p[i];
}
int main() {
  const int W = 2048, H = 256;  // doubles per row, rows (row pitch 16 KB: a multiple of 16 bytes)
  std::vector<double> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *o;
  cudaMalloc(&d, sizeof(double) * W * H); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), sizeof(double) * W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*E)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  struct Case { const char *name; CUtensorMapDataType ty; int esz; int bw8, bh; };  // bw8: box width in doubles
  Case cases[] = {{"f64 16x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 16, 4}, {"f64 14x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 14, 4},
                  {"u64 16x4", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 16, 4}, {"f32 32x4 (=16 doubles)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 16, 4},
                  {"u8 128x4 (=16 doubles)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 16, 4}, {"f64 44x10", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 44, 10}};
  for (const Case &c : cases) {
    const int mul = 8 / c.esz;
    CUtensorMap tm;
    cuuint64_t gd[2] = {(cuuint64_t)W * mul, (cuuint64_t)H}, gs[1] = {(cuuint64_t)W * 8};
    cuuint32_t box[2] = {(cuuint32_t)(c.bw8 * mul), (cuuint32_t)c.bh}, es[2] = {1, 1};
    CUresult r = ((E)fn)(&tm, c.ty, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int c0 = 6, c1 = 7, n8 = c.bw8 * c.bh;
    cudaMemset(o, 0, 65536);
    k<<<1, 128, n8 * 8 + 256>>>(tm, c0 * mul, c1, n8 * 8, n8, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> got(n8);
    if (e == cudaSuccess) cudaMemcpy(got.data(), o, n8 * 8, cudaMemcpyDeviceToHost);
    bool okv = e == cudaSuccess;
    for (int y = 0; y < c.bh && okv; ++y) for (int x = 0; x < c.bw8; ++x) okv &= got[y * c.bw8 + x] == (double)((c1 + y) * W + c0 + x);
    printf("%-24s encode %d kernel '%s' patch %s\n", c.name, (int)r, cudaGetErrorString(e), okv ? "CORRECT" : "wrong/none");
    if (e != cudaSuccess) { printf("(context lost: stopping)\n"); return 1; }
  }
  return 0;
}
