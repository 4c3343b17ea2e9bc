// Narrow 1x1 convolutions (C_in <= 48, C_out <= 64 channels by default) on mma.sync tensor cores.
//
// The first blocks of the three backbones are 1x1 convolutions with 8-64 input channels (rec 16->32, 32->64; det 16->32,
// 32->48, 48->48; most of the classifier).  Their operands are 16-128 bytes per pixel: the tcgen05 path pays a TMA box,
// an mbarrier round trip and a TMEM epilogue per 128-pixel tile for 2-8 KB of data and ran them at 10-45 % of the HBM
// rate.  Here a warp owns 16-pixel tiles: the A fragments of m16n8k8 are read straight from the NHWC rows (a pixel's
// 8-channel group is 16 contiguous bytes = the four threads of a fragment row), the filter block sits in shared memory
// once per CTA (row stride padded by 8 halves: conflict-free fragment reads), accumulators stay in registers, and the
// epilogue (bias, activation, post-affine, residual, ragged-width zeroing) is applied on the C fragments and stored as
// half2.  No shared-memory staging of activations, no barriers in the loop: the kernel is a stream over the pixels.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "pdl.h"

namespace b200ocr {
namespace {

constexpr int kPwThreads = 256;
constexpr int kPwWarps = kPwThreads / 32;
constexpr int kPwNChunk = 4;  // 8-channel output tiles accumulated together per pass over the A fragments

__device__ __forceinline__ void mma_m16n8k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

// ACT >= 0: the activation known at compile time (0 none, 1 relu, 2 hard-swish); ACT < 0: the rarer ones, by value
template <int ACT>
__device__ __forceinline__ float pw_act(float v, int act, float a, float b) {
  if (ACT == 0) return v;
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));
    case 3: return v / (1.f + __expf(-v));
    case 4: return __saturatef(fmaf(v, a, b));
    case 5: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// accumulator (u, j, i) with a run-time j: a select chain instead of a dynamically indexed register array
template <int TILES>
__device__ __forceinline__ float cj(const float (&c)[TILES][kPwNChunk][4], int u, int j, int i) {
  float v = c[u][0][i];
#pragma unroll
  for (int q = 1; q < kPwNChunk; ++q) v = j == q ? c[u][q][i] : v;
  return v;
}

// K8 = input channel groups of 8 (in.pitch / 8), TILES = 16-pixel tiles a warp loads before it multiplies
template <int K8, int TILES, int ACT>
__global__ void __launch_bounds__(kPwThreads)
pwconv_mma_kernel(const TV in, const TV out, const __half* __restrict__ w, const int w_stride,
                  const float* __restrict__ bias, const Epi e, const int* __restrict__ vw, const long npix) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t pw_smem[];
  constexpr int WS = K8 * 8 + 8;  // shared filter row stride (halves): rows 4 banks apart -> conflict-free b fragments
  const int nt = out.pitch >> 3;  // 8-channel output tiles
  __half* sw = reinterpret_cast<__half*>(pw_smem);
  float* sb = reinterpret_cast<float*>(pw_smem + size_t(nt) * 8 * WS * 2);
  for (int i = threadIdx.x; i < nt * 8 * K8; i += kPwThreads) {
    const int row = i / K8, kc = i - row * K8;
    *reinterpret_cast<uint4*>(sw + row * WS + kc * 8) = *reinterpret_cast<const uint4*>(w + long(row) * w_stride + kc * 8);
  }
  for (int i = threadIdx.x; i < nt * 8; i += kPwThreads) sb[i] = bias[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long ntile = (npix + 15) >> 4;
  const long hw = long(out.h) * out.w;
  const bool affine = e.s2 != 1.f || e.t2 != 0.f;
  for (long tile0 = (long(blockIdx.x) * kPwWarps + warp) * TILES; tile0 < ntile;
       tile0 += long(gridDim.x) * kPwWarps * TILES) {
    // A fragments of TILES tiles: row g and row g + 8 of each, every 8-channel group (all loads before any use)
    uint32_t a[TILES][K8][2];
#pragma unroll
    for (int u = 0; u < TILES; ++u) {
      const long p_lo = (tile0 + u) * 16 + g, p_hi = p_lo + 8;
      const __half* r_lo = in.p + p_lo * in.pitch + 2 * t;
      const __half* r_hi = in.p + p_hi * in.pitch + 2 * t;
#pragma unroll
      for (int k = 0; k < K8; ++k) {
        a[u][k][0] = p_lo < npix ? __ldg(reinterpret_cast<const unsigned*>(r_lo + 8 * k)) : 0u;
        a[u][k][1] = p_hi < npix ? __ldg(reinterpret_cast<const unsigned*>(r_hi + 8 * k)) : 0u;
      }
    }
    // per fragment row: output pointer (at this thread's channel pair of tile 0), in-range flag, and for ragged batches
    // whether the pixel lies beyond its row's valid width (zero output)
    __half* orow[TILES][2];
    const __half* rrow[TILES][2];
    bool live[TILES][2], dead[TILES][2];
#pragma unroll
    for (int u = 0; u < TILES; ++u)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const long p = (tile0 + u) * 16 + g + 8 * h2;
        live[u][h2] = p < npix;
        dead[u][h2] = vw && p < npix && int(p % out.w) >= vw[p / hw];
        orow[u][h2] = out.p + p * out.pitch + 2 * t;
        rrow[u][h2] = e.res ? e.res + p * e.res_pitch + 2 * t : nullptr;
      }
    const bool plain = !vw && !e.res && (tile0 + TILES) * 16 <= npix;
    for (int nc = 0; nc < nt; nc += kPwNChunk) {
      float c[TILES][kPwNChunk][4];
#pragma unroll
      for (int u = 0; u < TILES; ++u)
#pragma unroll
        for (int j = 0; j < kPwNChunk; ++j) c[u][j][0] = c[u][j][1] = c[u][j][2] = c[u][j][3] = 0.f;
#pragma unroll
      for (int k = 0; k < K8; ++k)
#pragma unroll
        for (int j = 0; j < kPwNChunk; ++j)
          if (nc + j < nt) {
            const uint32_t b = *reinterpret_cast<const uint32_t*>(sw + ((nc + j) * 8 + g) * WS + 8 * k + 2 * t);
#pragma unroll
            for (int u = 0; u < TILES; ++u) mma_m16n8k8(c[u][j], a[u][k][0], a[u][k][1], b);
          }
      // Fast path (warp-uniform): every pixel of the warp's tiles exists, no ragged rows, no residual, and the chunk's
      // channels are all logical ones -> straight-line bias / activation / affine / pack / store.
      if (plain && (nc + kPwNChunk) * 8 <= out.c) {
#pragma unroll
        for (int j = 0; j < kPwNChunk; ++j) {
          const int ch = (nc + j) * 8 + 2 * t;
          const float b0 = sb[ch], b1 = sb[ch + 1];
#pragma unroll
          for (int u = 0; u < TILES; ++u)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              float v0 = pw_act<ACT>(c[u][j][2 * h2] + b0, e.act, e.a, e.b);
              float v1 = pw_act<ACT>(c[u][j][2 * h2 + 1] + b1, e.act, e.a, e.b);
              if (affine) { v0 = fmaf(e.s2, v0, e.t2); v1 = fmaf(e.s2, v1, e.t2); }
              *reinterpret_cast<__half2*>(orow[u][h2] + (nc + j) * 8) = __floats2half2_rn(v0, v1);
            }
        }
        continue;
      }
#pragma unroll 1
      for (int j = 0; j < kPwNChunk; ++j) {
        if (nc + j >= nt) break;
        const int ch = (nc + j) * 8 + 2 * t;
        const float b0 = sb[ch], b1 = sb[ch + 1];
        const bool k0 = ch < out.c, k1 = ch + 1 < out.c;  // channels beyond the logical count stay zero
#pragma unroll
        for (int u = 0; u < TILES; ++u)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (!live[u][h2]) continue;
            float v0 = pw_act<ACT>(cj(c, u, j, 2 * h2) + b0, e.act, e.a, e.b);
            float v1 = pw_act<ACT>(cj(c, u, j, 2 * h2 + 1) + b1, e.act, e.a, e.b);
            if (affine) { v0 = fmaf(e.s2, v0, e.t2); v1 = fmaf(e.s2, v1, e.t2); }
            if (e.res) {
              const float2 r = __half22float2(*reinterpret_cast<const __half2*>(rrow[u][h2] + (nc + j) * 8));
              v0 += r.x; v1 += r.y;
            }
            if (!k0 || dead[u][h2]) v0 = 0.f;
            if (!k1 || dead[u][h2]) v1 = 0.f;
            *reinterpret_cast<__half2*>(orow[u][h2] + (nc + j) * 8) = __floats2half2_rn(v0, v1);
          }
      }
    }
  }
}

template <int K8, int TILES>
void pw_launch(const TV& in, const TV& out, const __half* w, int w_stride, const float* bias, const Epi& e,
               cudaStream_t s, const int* vw) {
  const long npix = long(out.n) * out.h * out.w;
  const int nt = out.pitch / 8;
  const size_t smem = size_t(nt) * 8 * (K8 * 8 + 8) * 2 + size_t(nt) * 8 * 4;
  const long ntile = (npix + 15) / 16;
  const long want = (ntile + long(kPwWarps) * TILES - 1) / (long(kPwWarps) * TILES);
  const int grid = int(std::max<long>(1, std::min<long>(want, 148 * 2)));
  if (e.act == 0) launch_k(pwconv_mma_kernel<K8, TILES, 0>, dim3(grid), dim3(kPwThreads), smem, s, in, out, w, w_stride, bias, e, vw, npix);
  else if (e.act == 1) launch_k(pwconv_mma_kernel<K8, TILES, 1>, dim3(grid), dim3(kPwThreads), smem, s, in, out, w, w_stride, bias, e, vw, npix);
  else if (e.act == 2) launch_k(pwconv_mma_kernel<K8, TILES, 2>, dim3(grid), dim3(kPwThreads), smem, s, in, out, w, w_stride, bias, e, vw, npix);
  else launch_k(pwconv_mma_kernel<K8, TILES, -1>, dim3(grid), dim3(kPwThreads), smem, s, in, out, w, w_stride, bias, e, vw, npix);
}

}  // namespace

bool launch_pwconv_mma(const TV& in, const TV& out, const __half* w, const float* bias, const ConvGeom& g, const Epi& e,
                       cudaStream_t s, const int* vw) {
  static const bool disabled = getenv("B200OCR_NO_PWCONV") != nullptr;
  // measured on B200 (profiles/r01_notes.md): ahead of the tcgen05 path up to 48 input and 64 output channels; wider
  // outputs need several passes over the A fragments and fall behind it
  static const int max_cin = getenv("B200OCR_PWCONV_MAX_CIN") ? atoi(getenv("B200OCR_PWCONV_MAX_CIN")) : 48;
  static const int max_cout = getenv("B200OCR_PWCONV_MAX_COUT") ? atoi(getenv("B200OCR_PWCONV_MAX_COUT")) : 64;
  if (disabled || in.pitch > max_cin || out.pitch > max_cout) return false;
  if (g.kh != 1 || g.kw != 1 || g.sh != 1 || g.sw != 1 || g.ph != 0 || g.pw != 0) return false;
  if (in.n != out.n || in.h != out.h || in.w != out.w) return false;
  if (in.pitch % 8 || out.pitch % 8 || in.pitch > 64 || out.pitch > 256 || g.cin_pad < in.pitch || g.cin_pad % 8) return false;
  if (g.cout_pad < out.pitch) return false;
  if ((reinterpret_cast<uintptr_t>(in.p) | reinterpret_cast<uintptr_t>(out.p) | reinterpret_cast<uintptr_t>(w)) & 15) return false;
  if (e.res && ((reinterpret_cast<uintptr_t>(e.res) & 3) || e.res_pitch % 2)) return false;
  switch (in.pitch / 8) {
    case 1: pw_launch<1, 4>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 2: pw_launch<2, 4>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 3: pw_launch<3, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 4: pw_launch<4, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 5: pw_launch<5, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 6: pw_launch<6, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 7: pw_launch<7, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    case 8: pw_launch<8, 2>(in, out, w, g.cin_pad, bias, e, s, vw); break;
    default: return false;
  }
  return true;
}

}  // namespace b200ocr
