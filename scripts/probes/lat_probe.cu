// Dependent-issue latency of the FP64 and integer instructions the ensemble loop chains (development probe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat_probe lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b, int ia) {
  double v = a + threadIdx.x * 1e-9;
  int iv = ia + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (KIND == 0) v = __dadd_rn(v, b);
      if (KIND == 1) v = __dmul_rn(v, b);
      if (KIND == 2) v = fma(v, b, a);
      if (KIND == 3) v = (v < b) ? __dadd_rn(v, a) : v;              // DSETP + predicated op
      if (KIND == 4) iv = __dp4a(iv, 0x01010101, iv);
      if (KIND == 5) iv = (iv ^ (iv << 3)) + ia;                      // LOP3/IADD chain
      if (KIND == 6) v = __hiloint2double(__double2hiint(v), __double2loint(v) + 1) - b;  // int update of the low word, then DADD
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + iv;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DADD", "DMUL", "DFMA", "DSETP+@DADD", "IDP.4A", "LOP3+IADD", "IADD(lo)+DADD"};
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int k = 0; k < 7; ++k) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        switch (k) {
          case 0: chain<0><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1e-9, 3); break;
          case 1: chain<1><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1.0000001, 3); break;
          case 2: chain<2><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 0.999, 3); break;
          case 3: chain<3><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1e300, 3); break;
          case 4: chain<4><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1.0, 3); break;
          case 5: chain<5><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1.0, 3); break;
          case 6: chain<6><<<148, 128 * warps>>>(out, cyc, iters, 1.0, 1.0, 3); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps/SMSP=%d %-14s %.2f clk per dependent step\n", warps, names[k], (double)c / (iters * 16.0));
    }
  }
  return 0;
}
