// pf_probe.cu -- does prefetch.global.L1 (SASS CCTL.E.PF1) bring a line into L1 on this part?  One warp walks random 8-byte
// cells of an L2-resident 32 MB table; the load of cell i+1 is timed after ~1000 cycles of dependent arithmetic, with and
// without a prefetch of its line issued before the arithmetic.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__global__ void warm(const double *t, size_t n, double *o) {
  double a = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += t[i];
  if (a == 1.234e-300) *o = a;
}
template <int MODE>  // 0 none, 1 prefetch.L1, 2 prefetch.L2
__global__ void k(const double *t, const unsigned *idx, int n, long long *cyc, double *o) {
  const int lane = threadIdx.x;
  double acc = 0, x = 1.0 + lane;
  long long tot = 0;
  for (int i = 0; i + 1 < n; ++i) {
    const double *next = t + idx[(i + 1) * 32 + lane];
    if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(next));
    if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(next));
    for (int k = 0; k < 120; ++k) x = x * 1.0000001 + 1e-9;  // ~1000 cycles of dependent FP64
    long long t0 = clock64();
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(next) : "memory");
    acc += v;
    long long t1 = clock64();
    tot += t1 - t0;
  }
  if (lane == 0) *cyc = tot / (n - 1);
  if (acc + x == 1.234e-300) *o = acc;
}
int main() {
  const size_t n = (32u << 20) / 8;
  double *t, *o; unsigned *idx; long long *cyc;
  cudaMalloc(&t, n * 8); cudaMemset(t, 0, n * 8); cudaMalloc(&o, 64); cudaMalloc(&cyc, 8);
  const int steps = 2000;
  std::vector<unsigned> h(steps * 32);
  srand(3);
  for (auto &v : h) v = (unsigned)(((size_t)rand() * 7919u + rand()) % n);
  cudaMalloc(&idx, h.size() * 4); cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const char *names[3] = {"no prefetch", "prefetch.global.L1 (CCTL.E.PF1)", "prefetch.global.L2 (CCTL.E.PF2)"};
  for (int mode = 0; mode < 3; ++mode) {
    warm<<<592, 256>>>(t, n, o);
    if (mode == 0) k<0><<<1, 32>>>(t, idx, steps, cyc, o);
    if (mode == 1) k<1><<<1, 32>>>(t, idx, steps, cyc, o);
    if (mode == 2) k<2><<<1, 32>>>(t, idx, steps, cyc, o);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-36s %s: %lld cycles from load issue to use (32 random lines per warp load)\n", names[mode], cudaGetErrorString(e), c);
  }
  return 0;
}
