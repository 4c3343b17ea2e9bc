// Microbenchmark: does the FHFMA stream of the 5x5 depthwise kernel (half-register selects, 8 window registers, 25 weight
// registers, 56 accumulators per thread) issue back to back?  Same FMA block as dwconv_persist_kernel<5,1,7>, no memory.
//   MODE 0: FHFMA, lo and hi halves of packed registers (the kernel's form)
//   MODE 1: FHFMA, lo halves only (operands pre-split into separate registers)
//   MODE 2: FFMA on fp32 copies of the same operands
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_pattern fhfma_pattern.cu && ./fhfma_pattern
#include <cstdio>
#include <cstdint>
constexpr int K = 5, R = 7, S = 4, WIN = S + K - 1;
__device__ __forceinline__ void fh(uint16_t x, uint16_t w, float& a) { asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a) : "h"(x), "h"(w)); }
template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) k(float* out, const uint32_t* in, int iters) {
  uint32_t wq[K * K];
  for (int t = 0; t < K * K; ++t) wq[t] = in[t * 128 + threadIdx.x];
  float acc[R][S][2];
  for (int r = 0; r < R; ++r) for (int s = 0; s < S; ++s) acc[r][s][0] = acc[r][s][1] = 0.f;
  uint32_t win[WIN];
  for (int j = 0; j < WIN; ++j) win[j] = in[(32 + j) * 128 + threadIdx.x];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int iy = 0; iy < R; ++iy) {
#pragma unroll
      for (int j = 0; j < WIN; ++j) win[j] = win[j] * 0x9E3779B1u + it;  // a fresh window per input row (1 IMAD each)
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int r = iy + K / 2 - ky;
        if (r < 0 || r >= R) continue;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const uint32_t x = win[s + kx], w = wq[ky * K + kx];
            if (MODE == 0) { fh(uint16_t(x & 0xffff), uint16_t(w & 0xffff), acc[r][s][0]); fh(uint16_t(x >> 16), uint16_t(w >> 16), acc[r][s][1]); }
            if (MODE == 1) { fh(uint16_t(x & 0xffff), uint16_t(w & 0xffff), acc[r][s][0]); fh(uint16_t(x & 0xffff), uint16_t(w & 0xffff), acc[r][s][1]); }
            if (MODE == 2) { acc[r][s][0] = fmaf(__uint_as_float(x), __uint_as_float(w), acc[r][s][0]); acc[r][s][1] = fmaf(__uint_as_float(x), __uint_as_float(w), acc[r][s][1]); }
          }
      }
    }
  }
  float sum = 0;
  for (int r = 0; r < R; ++r) for (int s = 0; s < S; ++s) sum += acc[r][s][0] + acc[r][s][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
template <int MODE, int MINB> void run(const char* name) {
  float* d; uint32_t* in;
  cudaMalloc(&d, 148 * 8 * 128 * 4); cudaMalloc(&in, 64 * 128 * 4); cudaMemset(in, 0x3c, 64 * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  k<MODE, MINB><<<148 * MINB, 128>>>(d, in, 10);
  cudaEventRecord(e0); k<MODE, MINB><<<148 * MINB, 128>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int taps = 0;
  for (int iy = 0; iy < R; ++iy) for (int ky = 0; ky < K; ++ky) { int r = iy + K / 2 - ky; if (r >= 0 && r < R) ++taps; }
  const double fma = double(148) * MINB * 128 * iters * taps * S * K * 2;
  printf("%-44s %8.3f ms  %6.1f FMA/clk/SM at 1.965 GHz (%d FMA + %d IMAD per iteration and thread)\n", name, ms,
         fma / ms / 1e6 / 148 / 1.965, taps * S * K * 2, R * WIN);
  cudaFree(d); cudaFree(in);
}
int main() {
  run<0, 4>("FHFMA lo+hi halves (kernel form), 4 CTA/SM");
  run<0, 3>("FHFMA lo+hi halves, 3 CTA/SM (<=168 regs)");
  run<0, 2>("FHFMA lo+hi halves, 2 CTA/SM (<=255 regs)");
  run<0, 5>("FHFMA lo+hi halves, 5 CTA/SM (<=96 regs)");
  run<1, 4>("FHFMA lo halves only, 4 CTA/SM");
  run<2, 4>("FFMA fp32 operands, 4 CTA/SM");
  return 0;
}
