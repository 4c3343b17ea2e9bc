// Microbenchmark: issue rate of FFMA, FHFMA (fma.rn.f32.f16), HFMA2 and F2F.f32.f16 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rates fma_rates.cu && ./fma_rates
#include <cstdio>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters, unsigned seed) {
  float acc[8];
  unsigned short a = (unsigned short)(0x3c00 + (threadIdx.x & 7)), b = (unsigned short)(0x3800 + (seed & 3));
  __half2 ha = __floats2half2_rn(1.0f + threadIdx.x * 1e-3f, 0.5f), hb = __floats2half2_rn(0.999f, 1.001f);
  __half2 hacc[8];
  for (int i = 0; i < 8; ++i) { acc[i] = i; hacc[i] = __floats2half2_rn(float(i), 1.f); }
  float fa = 1.0001f + seed, fb = 0.9999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) acc[u] = fmaf(fa, fb, acc[u]);
      if (MODE == 1) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[u]) : "h"(a), "h"(b));
      if (MODE == 2) hacc[u] = __hfma2(ha, hb, hacc[u]);
      if (MODE == 3) { unsigned short h = (unsigned short)(a + u + it); float f; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h)); acc[u] += f; }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i] + __low2float(hacc[i]) + __high2float(hacc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 8, 1024>>>(d, 100, 1);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 1024>>>(d, iters, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = double(148) * 8 * 1024 * iters * 8;
  printf("%-28s %8.3f ms  %7.1f Gop/s  (%.1f ops/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(d);
}
int main() { run<0>("FFMA"); run<1>("FHFMA fma.rn.f32.f16"); run<2>("HFMA2 (2 MAC each)"); run<3>("cvt.f32.f16 + FADD"); return 0; }
