// Microbenchmark: throughput of the legacy warp-level tensor path (mma.sync m16n8k16 / m16n8k8, fp16 in, fp32 acc) and of
// ldmatrix.x4 on sm_100a, per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rates mma_sync_rates.cu
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ __align__(16) uint32_t sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0x3c003c00u + i;
  __syncthreads();
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x38003800u + threadIdx.x, b1 = 0x34003400u;
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 1024;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      if (MODE == 1)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
      if (MODE == 2) {
        uint32_t r0, r1, r2, r3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr + u * 512));
        a[0] ^= r0; a[1] ^= r1; a[2] ^= r2; a[3] ^= r3;
      }
      if (MODE == 3) {  // the depthwise pattern: one ldmatrix.x4 feeding five mma
        uint32_t r[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr + u * 512));
#pragma unroll
        for (int v = 0; v < 5; ++v)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[(u + v) & 7][0]), "+f"(c[(u + v) & 7][1]), "+f"(c[(u + v) & 7][2]), "+f"(c[(u + v) & 7][3])
                       : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(b0 + v), "r"(b1));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + a[0] + a[1] + a[2] + a[3];
}
template <int MODE> void run(const char* name, double macs_per_instr, int per_iter) {
  float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 4, 256>>>(d, 100);
  cudaEventRecord(e0); k<MODE><<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = double(148) * 4 * 8 * iters * per_iter;  // warp instructions
  printf("%-34s %8.3f ms  %7.2f warp-instr/clk/SM  %8.1f dense MAC/clk/SM (1.9 GHz)\n", name, ms,
         instr / (ms * 1e-3) / 148 / 1.9e9, instr * macs_per_instr / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(d);
}
int main() {
  run<0>("mma.sync m16n8k16 f16->f32", 2048, 8);
  run<1>("mma.sync m16n8k8  f16->f32", 1024, 8);
  run<2>("ldmatrix.x4", 0, 8);
  run<3>("ldmatrix.x4 + 5 x m16n8k16", 2048 * 5.0 / 6.0, 8 * 6);
  return 0;
}
