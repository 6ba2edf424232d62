#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const double *src, int W, int c0, int c1, int bw, int bh, double *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * bh * 8) : "memory");
  __syncthreads();
  if (threadIdx.x < bh) {
    const double *g = src + (size_t)(c1 + threadIdx.x) * W + c0;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sm + threadIdx.x * bw * 8)), "l"(g), "r"(bw * 8), "r"(smem_u32(&bar)) : "memory");
  }
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  const double *p = (const double *)sm;
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = p[i];
}
int main() {
  const int W = 202, H = 202, bw = 12, bh = 9;
  std::vector<double> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = i;
  double *d, *o;
  cudaMalloc(&d, sizeof(double) * W * H); cudaMalloc(&o, sizeof(double) * bw * bh);
  cudaMemcpy(d, h.data(), sizeof(double) * W * H, cudaMemcpyHostToDevice);
  k<<<1, 128, bw * bh * 8 + 128>>>(d, W, 6, 7, bw, bh, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("bulk kernel -> %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double> r2(bw * bh);
  cudaMemcpy(r2.data(), o, sizeof(double) * bw * bh, cudaMemcpyDeviceToHost);
  printf("got %g %g %g (expect %d %d %d)\n", r2[0], r2[1], r2[bw], 7 * W + 6, 7 * W + 7, 8 * W + 6);
  return 0;
}
