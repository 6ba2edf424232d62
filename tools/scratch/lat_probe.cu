// latency probe: dependent-load chains through L1 / L2 / DRAM, and a batch-of-independent-loads test (one warp)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__global__ void chase(const unsigned *p, int n, unsigned *out, long long *cyc, int cg) {
  unsigned k = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) k = cg ? __ldcg(p + k) : __ldg(p + k);
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = t1 - t0; *out = k; }
}
// one warp issues B independent 8-byte gathers (lane stride `stride` doubles apart in B rows), then uses them all; repeated
template <int B>
__global__ void batch(const double *p, size_t pitch, int iters, unsigned step, double *out, long long *cyc) {
  double acc = 0; unsigned off = threadIdx.x / 3;  // ~13 distinct cells per warp like the score kernel
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double v[B];
#pragma unroll
    for (int b = 0; b < B; ++b) v[b] = __ldg(p + (size_t)b * pitch + off);
#pragma unroll
    for (int b = 0; b < B; ++b) acc += v[b];
    off += step;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = t1 - t0; *out = acc; }
}
int main() {
  unsigned *d; unsigned *o; long long *c; double *dd, *od;
  cudaMalloc(&o, 8); cudaMalloc(&c, 8); cudaMalloc(&od, 8);
  for (size_t bytes : {16u << 10, 8u << 20, 512u << 20}) {
    size_t n = bytes / 4; std::vector<unsigned> h(n);
    // random cycle with stride >= 128 B
    size_t stride = 37 * 32 + 32; for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i + stride) % n);
    cudaMalloc(&d, bytes); cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    for (int cg = 0; cg < 2; ++cg) {
      chase<<<1, 1>>>(d, 2000, o, c, cg); chase<<<1, 1>>>(d, 20000, o, c, cg);
      long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
      printf("chase %8zu KB %s: %.1f cycles/load\n", bytes >> 10, cg ? "ld.cg" : "ld.nc", cy / 20000.0);
    }
    cudaFree(d);
  }
  size_t pitch = 2002, rows = 2002; cudaMalloc(&dd, pitch * (rows + 8) * 8); cudaMemset(dd, 0, pitch * (rows + 8) * 8);
  for (unsigned step : {0u, 3u, 64u, 2002u * 3}) {
    long long cy;
    batch<1><<<1, 32>>>(dd, pitch, 1000, step, od, c); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("batch B=1 step %5u: %.1f cycles/iter\n", step, cy / 1000.0);
    batch<4><<<1, 32>>>(dd, pitch, 1000, step, od, c); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("batch B=4 step %5u: %.1f cycles/iter\n", step, cy / 1000.0);
    batch<8><<<1, 32>>>(dd, pitch, 1000, step % 64, od, c); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("batch B=8 step %5u: %.1f cycles/iter\n", step % 64, cy / 1000.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
