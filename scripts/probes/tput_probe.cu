// Issue throughput of the instruction kinds the ensemble pass is made of, alone and in pairs (development probe):
// which pipes are half rate on this part, and which pairs of pipes overlap.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tput_probe tput_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(x) x(0) x(1) x(2) x(3) x(4) x(5) x(6) x(7)

template <int KIND>
__global__ void tput(double* out, long long* cycles, int iters, double b, int ib) {
  double v0 = 1.0 + threadIdx.x * 1e-9, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#define DADD(n) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v##n) : "d"(b));
#define DMUL(n) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(v##n) : "d"(b));
#define DFMA(n) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(v##n) : "d"(b));
#define DSETP_IADD(n) asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(i##n) : "d"(v##n), "d"(b));
#define ISETP_IADD(n) asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(i##n) : "r"(i##n), "r"(ib));
#define IADD(n) asm volatile("add.s32 %0, %0, %1;" : "+r"(i##n) : "r"(ib));
#define LOP3(n) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i##n) : "r"(ib), "r"(i0));
#define SHF(n) asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(i##n) : "r"(ib));
#define IMAD(n) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i##n) : "r"(ib));
#define IDP(n) asm volatile("dp4a.s32.s32 %0, %0, %1, %0;" : "+r"(i##n) : "r"(ib));
#define SELP(n) asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 0; selp.s32 %0, %0, 7, p;}" : "+r"(i##n) : "r"(ib));
#define DSETP_SEL(n) asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, %0, 7, p;}" : "+r"(i##n) : "d"(v##n), "d"(b));
    if (KIND == 0) { REP8(DADD) REP8(DADD) }
    if (KIND == 1) { REP8(DMUL) REP8(DMUL) }
    if (KIND == 2) { REP8(DFMA) REP8(DFMA) }
    if (KIND == 3) { REP8(DSETP_IADD) REP8(DSETP_IADD) }
    if (KIND == 4) { REP8(ISETP_IADD) REP8(ISETP_IADD) }
    if (KIND == 5) { REP8(IADD) REP8(IADD) }
    if (KIND == 6) { REP8(LOP3) REP8(LOP3) }
    if (KIND == 7) { REP8(SHF) REP8(SHF) }
    if (KIND == 8) { REP8(IMAD) REP8(IMAD) }
    if (KIND == 9) { REP8(IDP) REP8(IDP) }
    if (KIND == 10) { REP8(DSETP_SEL) REP8(DSETP_SEL) }
#define DADD_LOP3(n) DADD(n) LOP3(n)
#define DADD_IMAD(n) DADD(n) IMAD(n)
#define LOP3_IMAD(n) LOP3(n) IMAD(n)
#define DADD_LOP3_IMAD(n) DADD(n) LOP3(n) IMAD(n)
    if (KIND == 11) { REP8(DADD_LOP3) }
    if (KIND == 12) { REP8(DADD_IMAD) }
    if (KIND == 13) { REP8(LOP3_IMAD) }
    if (KIND == 14) { REP8(DADD_LOP3_IMAD) }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7 + i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DADD", "DMUL", "DFMA", "DSETP+@IADD", "ISETP+@IADD", "IADD", "LOP3", "SHF", "IMAD", "IDP.4A", "DSETP+SEL",
                         "DADD+LOP3", "DADD+IMAD", "LOP3+IMAD", "DADD+LOP3+IMAD"};
  const int per_iter[] = {16, 16, 16, 32, 32, 16, 16, 16, 16, 16, 32, 16, 16, 16, 24};
  const int warps = 8;  // per scheduler
  for (int k = 0; k < 15; ++k) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
      switch (k) {
#define CASE(K) case K: tput<K><<<148, 128 * warps>>>(out, cyc, iters, 1.0000001, 3); break;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-16s %.2f clk per warp instruction per scheduler (%d warps)\n", names[k], (double)c / ((double)iters * per_iter[k] * warps), warps);
  }
  return 0;
}
