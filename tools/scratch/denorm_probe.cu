// latency probe: dependent chains of FP64 ops with normal vs subnormal operands (B200)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../slam_constructor_b200/csrc/dev_math.cuh"

template <int OP>
__global__ void chain(double x0, double y, int n, double *out, long long *cycles) {
  double x = x0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = __dmul_rn(x, y);
    if (OP == 1) x = __dadd_rn(x, y);
    if (OP == 2) x = __ddiv_rn(x, y);
    if (OP == 3) x = sg::div(x, y);
    if (OP == 4) x = sg::div_chain(x, y);
  }
  long long t1 = clock64();
  out[0] = x; cycles[0] = t1 - t0;
}

template <int OP>
void run(const char *name, double x0, double y) {
  double *out; long long *cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int n = 4096;
  chain<OP><<<1, 1>>>(x0, y, n, out, cyc);
  chain<OP><<<1, 1>>>(x0, y, n, out, cyc);
  long long c; double r;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8.1f cycles/op   (result %g)\n", name, (double)c / n, r);
}

int main() {
  run<0>("dmul normal x*0.9999999", 1.5, 0.99999999);
  run<0>("dmul subnormal 5e-324*0.99", 5e-324, 0.99);
  run<0>("dmul subnormal 1e-310*0.9999999", 1e-310, 0.99999999);
  run<1>("dadd normal", 1.5, 1e-9);
  run<1>("dadd subnormal 5e-324 + 5e-324", 5e-324, 5e-324);
  run<2>("__ddiv_rn normal x/1.0000001", 1.5, 1.0000001);
  run<2>("__ddiv_rn subnormal 1e-310/1.0000001", 1e-310, 1.0000001);
  run<2>("__ddiv_rn 5e-324/0.99997", 5e-324, 0.99997);
  run<3>("sg::div 1e-310/1.0000001", 1e-310, 1.0000001);
  run<3>("sg::div 5e-324/0.99997", 5e-324, 0.99997);
  run<4>("sg::div_chain 1e-310/1.0000001", 1e-310, 1.0000001);
  run<4>("sg::div_chain 5e-324/0.99997", 5e-324, 0.99997);
  run<3>("sg::div normal x/1.0000001", 1.5, 1.0000001);
  return 0;
}
