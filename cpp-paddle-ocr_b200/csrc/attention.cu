// SVTR neck self-attention on tensor cores (mma.sync m16n8k16, fp16 operands, fp32 accumulate) for sm_100a.
//
// Replaces the reshape2 / transpose2 / matmul_v2 / softmax / matmul_v2 chain of the recognizer's two mixing blocks
// (8 heads x 15 dims, T = W/8 <= ~130 tokens).  The contraction sizes (15 and T) are far too small for tcgen05
// (M = 128 tiles); one warp-level MMA per 16 queries x 8 keys is the right granularity.
//
// One CTA per sequence, one warp per head.  The sequence's packed qkv rows ([3][heads][d] halves per token) are
// read ONCE with coalesced 16-byte loads and re-laid in shared memory with d padded 15 -> 16 (Q, K token-major;
// V transposed, so that every MMA fragment is an aligned 32-bit shared-memory load).  Per 16-query tile the warp
// streams the keys in blocks of 32 with an online soft-max (running max / sum, flash-attention style): S = Q.K^T in
// registers, P = exp2(S - m) repacked in place as the A operand of P.V.  The number of key blocks depends only on
// the row's own valid length, so a row of a ragged batch is bit-identical to the same row in a dense batch.
#include "kernels.h"
#include "pdl.h"

#include <cfloat>
#include <cstdlib>

namespace b200ocr {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kHeadsMax = 8;
constexpr int kTok = 264;  // halves per token in shared memory: Q[8][16], K[8][16] + 8 pad (fragment loads hit 32 banks)

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// smem: QK [Tp][kTok] halves (Q then K per token; the Q slots are reused for the output), Vt [8][16][Tp + 8]
__global__ void __launch_bounds__(kAttnThreads)
attention_mma_kernel(TV qkv, TV out, int heads, int hd, float scale_log2e, const int* __restrict__ vw, int Tp) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* QK = reinterpret_cast<__half*>(attn_smem);
  const int vpitch = Tp + 8;                 // +8 halves: rows of Vt land on different banks
  __half* Vt = QK + size_t(Tp) * kTok;
  const int Tfull = qkv.h * qkv.w;
  const int n = blockIdx.x;
  const int Tv = vw ? min(vw[n], Tfull) : Tfull;
  const int C = heads * hd;                  // 120
  const __half* base = qkv.p + long(n) * Tfull * qkv.pitch;

  // ---- zero fill (pad dims, pad tokens), then scatter the valid rows
  {
    uint4* z = reinterpret_cast<uint4*>(attn_smem);
    const int nz = int((size_t(Tp) * kTok + size_t(kHeadsMax) * 16 * vpitch) * 2 / 16);
    for (int i = threadIdx.x; i < nz; i += kAttnThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    // a thread keeps ONE 16-byte slot of the token row (45 slots for 8 x 15): where its 8 halves go is computed once
    const int vec_per_row = (3 * C) / 8;
    const int tpp = kAttnThreads / vec_per_row;            // tokens per pass
    const int v = threadIdx.x % vec_per_row, t0 = threadIdx.x / vec_per_row;
    int dbase[8], dstep[8];                                 // destination (in halves from attn_smem) = dbase + t * dstep
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      const int which = c / C, r = c - which * C;
      const int head = r / hd, d = r - head * hd;
      if (which < 2) { dbase[e] = which * 128 + head * 16 + d; dstep[e] = kTok; }
      else { dbase[e] = Tp * kTok + (head * 16 + d) * vpitch; dstep[e] = 1; }
    }
    if (t0 < tpp)
      for (int t = t0; t < Tv; t += tpp) {
        const uint4 raw = *reinterpret_cast<const uint4*>(base + long(t) * qkv.pitch + v * 8);
        const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) QK[dbase[e] + t * dstep[e]] = h[e];
      }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  if (warp < heads) {
    const int head = warp;
    const __half* Vh = Vt + size_t(head) * 16 * vpitch;
    const int nkb = (Tv + 31) >> 5;          // key blocks of 32: a function of the row's own length only
    for (int q0 = 0; q0 < Tv; q0 += 16) {
      uint32_t qa[4];
      {
        const __half* q = QK + size_t(q0 + g) * kTok + head * 16 + 2 * t4;
        qa[0] = *reinterpret_cast<const uint32_t*>(q);
        qa[1] = *reinterpret_cast<const uint32_t*>(q + 8 * kTok);
        qa[2] = *reinterpret_cast<const uint32_t*>(q + 8);
        qa[3] = *reinterpret_cast<const uint32_t*>(q + 8 * kTok + 8);
      }
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g + 8
      float o[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int j0 = kb * 32;
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
          const __half* k = QK + size_t(j0 + nt * 8 + g) * kTok + 128 + head * 16 + 2 * t4;
          mma16816(s[nt], qa, *reinterpret_cast<const uint32_t*>(k), *reinterpret_cast<const uint32_t*>(k + 8));
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = j0 + nt * 8 + 2 * t4;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool ok = key + e < Tv;
            s[nt][e] = ok ? s[nt][e] * scale_log2e : -INFINITY;
            s[nt][2 + e] = ok ? s[nt][2 + e] * scale_log2e : -INFINITY;
            bm0 = fmaxf(bm0, s[nt][e]);
            bm1 = fmaxf(bm1, s[nt][2 + e]);
          }
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);  // finite: every block holds at least one valid key
        const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pa[2][4];  // P as A fragments: k-step ks covers keys j0 + 16*ks .. +15 (score tiles 2ks, 2ks+1)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = fast_exp2(s[nt][0] - mn0), p1 = fast_exp2(s[nt][1] - mn0);
          const float p2 = fast_exp2(s[nt][2] - mn1), p3 = fast_exp2(s[nt][3] - mn1);
          rs0 += p0 + p1;
          rs1 += p2 + p3;
          pa[nt >> 1][(nt & 1) * 2] = pack_half2(p0, p1);
          pa[nt >> 1][(nt & 1) * 2 + 1] = pack_half2(p2, p3);
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
          o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const __half* v = Vh + size_t(dt * 8 + g) * vpitch + j0 + ks * 16 + 2 * t4;
            mma16816(o[dt], pa[ks], *reinterpret_cast<const uint32_t*>(v), *reinterpret_cast<const uint32_t*>(v + 8));
          }
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      __syncwarp();  // every lane has read this tile's Q fragments: the Q slots now take the output
#pragma unroll
      for (int dt = 0; dt < 2; ++dt) {
        __half* d0 = QK + size_t(q0 + g) * kTok + head * 16 + dt * 8 + 2 * t4;
        *reinterpret_cast<uint32_t*>(d0) = pack_half2(o[dt][0] * i0, o[dt][1] * i0);
        *reinterpret_cast<uint32_t*>(d0 + 8 * kTok) = pack_half2(o[dt][2] * i1, o[dt][3] * i1);
      }
    }
  }
  __syncthreads();
  // ---- write back: [T][C] halves, 16-byte vectors; rows beyond the valid length are zero
  {
    const int vec_per_row = C / 8;  // 15
    const int tpp = kAttnThreads / vec_per_row;
    const int v = threadIdx.x % vec_per_row, t0 = threadIdx.x / vec_per_row;
    int sbase[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      const int head = c / hd, d = c - head * hd;
      sbase[e] = head * 16 + d;
    }
    __half* obase = out.p + long(n) * Tfull * out.pitch + v * 8;
    if (t0 < tpp)
      for (int t = t0; t < Tfull; t += tpp) {
        uint4 pk = make_uint4(0, 0, 0, 0);
        if (t < Tv) {
          __half* h = reinterpret_cast<__half*>(&pk);
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = QK[size_t(t) * kTok + sbase[e]];
        }
        *reinterpret_cast<uint4*>(obase + long(t) * out.pitch) = pk;
      }
  }
}

}  // namespace

bool launch_attention_mma(const TV& qkv, const TV& out, int heads, int hd, float scale, cudaStream_t s, const int* vw) {
  static const bool disabled = getenv("B200OCR_OLD_ATTENTION") != nullptr;
  if (disabled) return false;
  const int T = qkv.h * qkv.w, C = heads * hd;
  if (heads > kHeadsMax || hd > 16 || hd < 1 || (3 * C) % 8 || C % 8 || qkv.c != 3 * C || out.c != C) return false;
  if (qkv.pitch % 8 || out.pitch % 8 || (reinterpret_cast<uintptr_t>(qkv.p) & 15) || (reinterpret_cast<uintptr_t>(out.p) & 15)) return false;
  const int Tp = (T + 31) & ~31;
  const size_t smem = (size_t(Tp) * kTok + size_t(kHeadsMax) * 16 * (Tp + 8)) * 2;
  if (smem > 200 * 1024) return false;  // T > ~250 tokens: the CUDA-core kernel takes it
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev < 64 && smem > configured[dev]) {
    cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured[dev] = 200 * 1024;
  }
  launch_k(attention_mma_kernel, dim3(qkv.n), dim3(kAttnThreads), smem, s, qkv, out, heads, hd, scale * 1.4426950408889634f, vw, Tp);
  return true;
}

}  // namespace b200ocr
