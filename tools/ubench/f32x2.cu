// Microbenchmark: issue rate of sm_100's packed binary32 instructions (FMUL2 / FADD2) against scalar FMUL / FADD.
// Every thread runs ILP independent multiply (or add) chains; 148 SMs x 8 CTAs x 256 threads.  Prints warp
// instructions per clock per SM.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8;
constexpr int ITERS = 4096;

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, float c) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(1.0f + threadIdx.x * 1e-6f + i, 2.0f + i);
  const float2 cc = make_float2(c, c);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) { a[i].x = __fmul_rn(a[i].x, c); }                       // ILP scalar FMUL
      if (KIND == 1) { a[i] = __fmul2_rn(a[i], cc); }                         // ILP FMUL2
      if (KIND == 2) { a[i].x = __fmul_rn(a[i].x, c); a[i].y = __fmul_rn(a[i].y, c); }  // 2 ILP scalar FMUL (same math as 1)
      if (KIND == 3) { a[i].x = __fadd_rn(a[i].x, c); }
      if (KIND == 4) { a[i] = __fadd2_rn(a[i], cc); }
      if (KIND == 5) { a[i] = __fmul2_rn(a[i], cc); a[i].x = fminf(a[i].x, a[i].y); }   // FMUL2 + FMNMX mix
      if (KIND == 6) { a[i].x = __fmul_rn(a[i].x, c); a[i].y = __fmul_rn(a[i].y, c); a[i].x = fminf(a[i].x, a[i].y); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
  if (s == 12345.678f) out[0] = s;
}

template <int KIND>
void run(const char* name, double inst_per_iter_per_chain) {
  float* d;
  cudaMalloc(&d, 4);
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<KIND><<<grid, 256>>>(d, 1.0000001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<KIND><<<grid, 256>>>(d, 1.0000001f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)grid * 8;
  const double inst = warps * ITERS * ILP * inst_per_iter_per_chain;
  const double clocks = ms * 1e-3 * khz * 1e3;
  printf("%-28s %8.3f ms   %.3f warp-inst/clk/SM (nominal clock %d MHz)\n", name, ms, inst / clocks / sms, khz / 1000);
  cudaFree(d);
}

int main() {
  run<0>("FMUL  (8 chains)", 1);
  run<1>("FMUL2 (8 chains)", 1);
  run<2>("2x FMUL (16 chains)", 2);
  run<3>("FADD  (8 chains)", 1);
  run<4>("FADD2 (8 chains)", 1);
  run<5>("FMUL2 + FMNMX", 2);
  run<6>("2x FMUL + FMNMX", 3);
  return 0;
}
