// CUDA-core kernels of the forward passes: stem / odd-shape direct conv, depthwise conv,
// SE block, FPN glue, pooling, LayerNorm, attention, DB head tail, cls head, CTC head (CUDA-core
// variant).  All activations are NHWC fp16 with 16-byte (8-channel) vector access; accumulation
// is fp32.  HBM/L2-bound kernels: one thread per (pixel, 8-channel group), coalesced along C.
#include "kernels.h"
#include "pdl.h"
#include "resize.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace b200ocr {

namespace {

constexpr int kThreads = 256;


struct H8 {
  uint4 u;
  __device__ __forceinline__ void to_float(float* f) const {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float* f) {
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
};

__device__ __forceinline__ H8 ld8(const __half* p) {
  H8 r;
  r.u = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ void st8(__half* p, const H8& v) { *reinterpret_cast<uint4*>(p) = v.u; }

template <int ACT>
__device__ __forceinline__ float act_t(float v, float a, float b) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));
  if (ACT == 3) return v / (1.f + __expf(-v));
  if (ACT == 4) return __saturatef(fmaf(v, a, b));
  if (ACT == 5) return 1.f / (1.f + __expf(-v));
  return v;
}
template <int ACT>
__device__ __forceinline__ void epilogue8_act(float* acc, const float* bias, int c0, const Epi& e) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = fmaf(e.s2, act_t<ACT>(acc[i] + bias[c0 + i], e.a, e.b), e.t2);
}

// Shared epilogue for 8 consecutive channels starting at c0 of pixel `pix`.  The activation is selected ONCE per
// call (a per-element switch compiles to an indirect branch per value).
__device__ __forceinline__ void epilogue8(float* acc, const float* bias, int c0, int cout, const Epi& e,
                                          const TV& out, long pix, bool masked = false) {
  if (masked) {  // ragged batch: beyond this row's valid width the tensor is zero
    H8 z;
    z.u = make_uint4(0, 0, 0, 0);
    st8(out.p + pix * out.pitch + c0, z);
    return;
  }
  switch (e.act) {
    case 1: epilogue8_act<1>(acc, bias, c0, e); break;
    case 2: epilogue8_act<2>(acc, bias, c0, e); break;
    case 3: epilogue8_act<3>(acc, bias, c0, e); break;
    case 4: epilogue8_act<4>(acc, bias, c0, e); break;
    case 5: epilogue8_act<5>(acc, bias, c0, e); break;
    default: epilogue8_act<0>(acc, bias, c0, e); break;
  }
  if (e.res) {
    float r[8];
    ld8(e.res + pix * e.res_pitch + c0).to_float(r);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += r[i];
  }
  if (c0 + 8 > cout) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i >= cout) acc[i] = 0.f;
  }
  H8 o;
  o.from_float(acc);
  st8(out.p + pix * out.pitch + c0, o);
}

// ---------------------------------------------------------------- direct conv
__global__ void __launch_bounds__(kThreads)
conv_simt_kernel(TV in, TV out, const __half* __restrict__ w, const float* __restrict__ bias,
                 ConvGeom g, Epi e, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (out.c + 7) >> 3;
  const long npix = long(out.n) * out.h * out.w;
  const long total = npix * cgs;
  const int taps = g.kh * g.kw;
  const int cin8 = (in.c + 7) >> 3;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    // consecutive threads -> consecutive pixels of one channel group (weights broadcast)
    const int cg = int(t / npix);
    const long pix = t - long(cg) * npix;
    const int ox = int(pix % out.w);
    const int oy = int((pix / out.w) % out.h);
    const int n = int(pix / (long(out.w) * out.h));
    if (vw && ox >= vw[n]) { epilogue8(nullptr, bias, cg * 8, out.c, e, out, pix, true); continue; }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < g.kh; ++ky) {
      const int iy = oy * g.sh - g.ph + ky;
      if (iy < 0 || iy >= in.h) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int ix = ox * g.sw - g.pw + kx;
        if (ix < 0 || ix >= in.w) continue;
        const __half* ip = in.p + ((long(n) * in.h + iy) * in.w + ix) * in.pitch;
        const __half* wp = w + (long(cg) * 8 * taps + (ky * g.kw + kx)) * g.cin_pad;
        for (int c8 = 0; c8 < cin8; ++c8) {
          float x[8];
          ld8(ip + c8 * 8).to_float(x);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            float wv[8];
            ld8(wp + long(o) * taps * g.cin_pad + c8 * 8).to_float(wv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[o] = fmaf(x[i], wv[i], acc[o]);
          }
        }
      }
    }
    epilogue8(acc, bias, cg * 8, out.c, e, out, pix);
  }
}

// ---------------------------------------------------------------- depthwise conv
template <int KH, int KW>
__global__ void __launch_bounds__(kThreads)
dwconv_kernel(TV in, TV out, const float* __restrict__ wb, ConvGeom g, Epi e, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (out.c + 7) >> 3;
  const int cp = g.cout_pad;
  const long total = long(out.n) * out.h * out.w * cgs;
  const float* bias = wb + long(KH * KW) * cp;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    const int ox = int(pix % out.w);
    const int oy = int((pix / out.w) % out.h);
    const int n = int(pix / (long(out.w) * out.h));
    if (vw && ox >= vw[n]) { epilogue8(nullptr, bias, cg * 8, out.c, e, out, pix, true); continue; }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const int iy = oy * g.sh - g.ph + ky;
      if (iy < 0 || iy >= in.h) continue;
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const int ix = ox * g.sw - g.pw + kx;
        if (ix < 0 || ix >= in.w) continue;
        float x[8];
        ld8(in.p + ((long(n) * in.h + iy) * in.w + ix) * in.pitch + cg * 8).to_float(x);
        const float4* wp = reinterpret_cast<const float4*>(wb + long(ky * KW + kx) * cp + cg * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        acc[0] = fmaf(x[0], w0.x, acc[0]); acc[1] = fmaf(x[1], w0.y, acc[1]);
        acc[2] = fmaf(x[2], w0.z, acc[2]); acc[3] = fmaf(x[3], w0.w, acc[3]);
        acc[4] = fmaf(x[4], w1.x, acc[4]); acc[5] = fmaf(x[5], w1.y, acc[5]);
        acc[6] = fmaf(x[6], w1.z, acc[6]); acc[7] = fmaf(x[7], w1.w, acc[7]);
      }
    }
    epilogue8(acc, bias, cg * 8, out.c, e, out, pix);
  }
}

// fp16 x fp16 + fp32 -> fp32 in one instruction (sm_100 FHFMA): the product of two halves is exact in fp32, so this
// equals converting both operands and issuing an FFMA -- without the conversions.
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
  return c;
}
__device__ __forceinline__ void fhfma8(const uint4& x, const uint4& w, float* acc) {
  const unsigned xs[4] = {x.x, x.y, x.z, x.w}, ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fhfma((unsigned short)(xs[i] & 0xffffu), (unsigned short)(ws[i] & 0xffffu), acc[2 * i]);
    acc[2 * i + 1] = fhfma((unsigned short)(xs[i] >> 16), (unsigned short)(ws[i] >> 16), acc[2 * i + 1]);
  }
}

// Depthwise conv, register-blocked along x: one thread owns S consecutive output pixels of one 8-channel group.
// Per filter row it loads the KW weight vectors (fp16, 16 B each) once and every input column once (a
// (S-1)*SW+KW wide window) instead of KH*KW input + weight loads per output pixel; all loads of a filter row are
// unconditional (clamped address, zeroed afterwards when out of range) so that they issue back to back, and the
// multiply-accumulates take the fp16 operands directly (FHFMA), fp32 accumulation.
template <int KH, int KW, int SW, int S>
__global__ void __launch_bounds__(kThreads)
dwconv_strip_kernel(TV in, TV out, const float* __restrict__ wb, const __half* __restrict__ wh, ConvGeom g, Epi e,
                    const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (out.c + 7) >> 3;
  const int cp = g.cout_pad;
  const int strips = (out.w + S - 1) / S;
  const long total = long(out.n) * out.h * strips * cgs;
  const float* bias = wb + long(KH * KW) * cp;
  constexpr int WIN = (S - 1) * SW + KW;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    long r = t / cgs;
    const int sx = int(r % strips); r /= strips;
    const int oy = int(r % out.h);
    const int n = int(r / out.h);
    const int ox0 = sx * S;
    const int ix0 = ox0 * SW - g.pw;
    const bool interior = ix0 >= 0 && ix0 + WIN <= in.w;
    float acc[S][8];
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[s_][i] = 0.f;
    const __half* img = in.p + long(n) * in.h * in.w * in.pitch + cg * 8;
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const int iy = oy * g.sh - g.ph + ky;
      if (iy < 0 || iy >= in.h) continue;
      const __half* row = img + long(iy) * in.w * in.pitch;
      uint4 raw[WIN];
      if (interior) {
#pragma unroll
        for (int j = 0; j < WIN; ++j) raw[j] = *reinterpret_cast<const uint4*>(row + long(ix0 + j) * in.pitch);
      } else {
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
          const int ix = ix0 + j;
          raw[j] = *reinterpret_cast<const uint4*>(row + long(min(max(ix, 0), in.w - 1)) * in.pitch);
          if (ix < 0 || ix >= in.w) raw[j] = make_uint4(0, 0, 0, 0);
        }
      }
      uint4 w[KW];
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) w[kx] = __ldg(reinterpret_cast<const uint4*>(wh + long(ky * KW + kx) * cp + cg * 8));
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) {
          const int kx = j - s_ * SW;  // compile-time after unrolling
          if (kx >= 0 && kx < KW) fhfma8(raw[j], w[kx], acc[s_]);
        }
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_) {
      const int ox = ox0 + s_;
      if (ox >= out.w) break;
      const long pix = (long(n) * out.h + oy) * out.w + ox;
      epilogue8(acc[s_], bias, cg * 8, out.c, e, out, pix, vw && ox >= vw[n]);
    }
  }
}

// fp32-weight variant (used by the detector, whose thresholded output is the most precision-sensitive):
// Depthwise conv, register-blocked along x: one thread owns S consecutive output pixels of one 8-channel group.
// Per filter row it loads the KW weight vectors once and every input column once (a (S-1)*SW+KW wide window),
// instead of KH*KW input + weight loads per output pixel.  All loads of a filter row are unconditional (clamped
// address, zeroed afterwards when out of range) so that they are issued back to back instead of one per branch.
template <int KH, int KW, int SW, int S>
__global__ void __launch_bounds__(kThreads)
dwconv_strip_f32w_kernel(TV in, TV out, const float* __restrict__ wb, ConvGeom g, Epi e, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (out.c + 7) >> 3;
  const int cp = g.cout_pad;
  const int strips = (out.w + S - 1) / S;
  const long total = long(out.n) * out.h * strips * cgs;
  const float* bias = wb + long(KH * KW) * cp;
  constexpr int WIN = (S - 1) * SW + KW;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    long r = t / cgs;
    const int sx = int(r % strips); r /= strips;
    const int oy = int(r % out.h);
    const int n = int(r / out.h);
    const int ox0 = sx * S;
    const int ix0 = ox0 * SW - g.pw;
    float acc[S][8];
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[s_][i] = 0.f;
    const __half* img = in.p + long(n) * in.h * in.w * in.pitch + cg * 8;
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const int iy = oy * g.sh - g.ph + ky;
      const bool yok = iy >= 0 && iy < in.h;
      const __half* row = img + long(min(max(iy, 0), in.h - 1)) * in.w * in.pitch;
      H8 raw[WIN];
#pragma unroll
      for (int j = 0; j < WIN; ++j) raw[j] = ld8(row + long(min(max(ix0 + j, 0), in.w - 1)) * in.pitch);
      float w[KW][8];
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(wb + long(ky * KW + kx) * cp + cg * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        w[kx][0] = w0.x; w[kx][1] = w0.y; w[kx][2] = w0.z; w[kx][3] = w0.w;
        w[kx][4] = w1.x; w[kx][5] = w1.y; w[kx][6] = w1.z; w[kx][7] = w1.w;
      }
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        const int ix = ix0 + j;
        const bool ok = yok && ix >= 0 && ix < in.w;
        float x[8];
        raw[j].to_float(x);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = ok ? x[i] : 0.f;
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) {
          const int kx = j - s_ * SW;  // compile-time after unrolling
          if (kx >= 0 && kx < KW) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[s_][i] = fmaf(x[i], w[kx][i], acc[s_][i]);
          }
        }
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_) {
      const int ox = ox0 + s_;
      if (ox >= out.w) break;
      const long pix = (long(n) * out.h + oy) * out.w + ox;
      epilogue8(acc[s_], bias, cg * 8, out.c, e, out, pix, vw && ox >= vw[n]);
    }
  }
}

// First layer: 3 input channels (pitch 8), 3x3, any stride; all COUT channels of one output pixel per thread,
// filter in shared memory as fp32 [tap][ci][co].
template <int COUT>
__global__ void __launch_bounds__(kThreads)
stem_conv_kernel(TV in, TV out, const __half* __restrict__ w, const float* __restrict__ bias, ConvGeom g, Epi e,
                 const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sw[9 * 3 * COUT];
  __shared__ float sb[COUT];
  for (int i = threadIdx.x; i < 9 * 3 * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % 3, tap = i / (3 * COUT);
    sw[i] = __half2float(w[(long(co) * 9 + tap) * g.cin_pad + ci]);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const long npix = long(out.n) * out.h * out.w;
  for (long pix = blockIdx.x * long(blockDim.x) + threadIdx.x; pix < npix; pix += long(gridDim.x) * blockDim.x) {
    const int ox = int(pix % out.w);
    const int oy = int((pix / out.w) % out.h);
    const int n = int(pix / (long(out.w) * out.h));
    float acc[COUT];
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[i] = 0.f;
    const bool masked = vw && ox >= vw[n];
    if (!masked) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * g.sh - g.ph + ky;
        if (iy < 0 || iy >= in.h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * g.sw - g.pw + kx;
          if (ix < 0 || ix >= in.w) continue;
          const uint2 raw = *reinterpret_cast<const uint2*>(in.p + ((long(n) * in.h + iy) * in.w + ix) * in.pitch);
          const float2 ab = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float c = __half2float(*reinterpret_cast<const __half*>(&raw.y));
          const float* wp = sw + (ky * 3 + kx) * 3 * COUT;
#pragma unroll
          for (int co = 0; co < COUT; ++co)
            acc[co] = fmaf(ab.x, wp[co], fmaf(ab.y, wp[COUT + co], fmaf(c, wp[2 * COUT + co], acc[co])));
        }
      }
    }
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 8) epilogue8(acc + c0, sb, c0, out.c, e, out, pix, masked);
  }
}

template <int COUT, int ACT, int PX>
__device__ __forceinline__ void stem_store(float (*acc)[COUT], const float* sb, const Epi& e, const TV& out, long pix0, int ox0,
                                           int vwn) {
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + p;
    if (ox >= out.w) break;
    __half* op = out.p + (pix0 + p) * out.pitch;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 8) {
      H8 o;
      if (ox >= vwn) {
        o.u = make_uint4(0, 0, 0, 0);
      } else {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = fmaf(e.s2, act_t<ACT>(acc[p][c0 + i] + sb[c0 + i], e.a, e.b), e.t2);
          if (c0 + i >= out.c) v[i] = 0.f;
        }
        o.from_float(v);
      }
      st8(op + c0, o);
    }
  }
}

// First layer, stride 2: one thread = PX consecutive output pixels x all COUT channels.  Per filter row the thread
// loads its 2*PX+1 input pixels once (8 B each: 3 of the 8 channel slots are real) and every filter tap's COUT weights
// come from shared memory as broadcast 16-byte loads shared by the 4 pixels: 27 * COUT / 4 shared loads per
// 4 * 27 * COUT multiply-adds (the one-pixel kernel above issues one shared load per multiply-add).
template <int COUT, int PX>
__global__ void __launch_bounds__(128, PX == 4 ? 3 : 4)
stem_conv_s2x4_kernel(TV in, TV out, const __half* __restrict__ w, const float* __restrict__ bias, ConvGeom g, Epi e,
                      const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float sw[9 * 3 * COUT];
  __shared__ float sb[COUT];
  for (int i = threadIdx.x; i < 9 * 3 * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % 3, tap = i / (3 * COUT);
    sw[i] = __half2float(w[(long(co) * 9 + tap) * g.cin_pad + ci]);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int strips = (out.w + PX - 1) / PX;
  const long total = long(out.n) * out.h * strips;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int sx = int(t % strips);
    const int oy = int((t / strips) % out.h);
    const int n = int(t / (long(strips) * out.h));
    const int ox0 = sx * PX;
    const int ix0 = ox0 * 2 - g.pw;
    float acc[PX][COUT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int i = 0; i < COUT; ++i) acc[p][i] = 0.f;
    // all 27 input pixels are requested before the first multiply-add (one round trip to memory per thread)
    constexpr int WINX = 2 * PX + 1;
    uint2 raw[3][WINX];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - g.ph + ky;
      const bool y_ok = iy >= 0 && iy < in.h;
      const __half* row = in.p + (long(n) * in.h + (y_ok ? iy : 0)) * in.w * in.pitch;
#pragma unroll
      for (int j = 0; j < WINX; ++j) {
        const int ix = ix0 + j;
        raw[ky][j] = make_uint2(0u, 0u);
        if (y_ok && ix >= 0 && ix < in.w) raw[ky][j] = *reinterpret_cast<const uint2*>(row + long(ix) * in.pitch);
      }
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float x[WINX][3];
#pragma unroll
      for (int j = 0; j < WINX; ++j) {
        const float2 ab = __half22float2(*reinterpret_cast<const __half2*>(&raw[ky][j].x));
        x[j][0] = ab.x; x[j][1] = ab.y;
        x[j][2] = __half2float(*reinterpret_cast<const __half*>(&raw[ky][j].y));
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * COUT);
#pragma unroll
          for (int q = 0; q < COUT / 4; ++q) {
            const float4 wv = wp[q];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float xv = x[2 * p + kx][ci];
              acc[p][4 * q] = fmaf(xv, wv.x, acc[p][4 * q]);
              acc[p][4 * q + 1] = fmaf(xv, wv.y, acc[p][4 * q + 1]);
              acc[p][4 * q + 2] = fmaf(xv, wv.z, acc[p][4 * q + 2]);
              acc[p][4 * q + 3] = fmaf(xv, wv.w, acc[p][4 * q + 3]);
            }
          }
        }
    }
    const int vwn = vw ? vw[n] : out.w;
    const long pix0 = (long(n) * out.h + oy) * out.w + ox0;
    switch (e.act) {  // one activation dispatch per thread, not per value
      case 1: stem_store<COUT, 1, PX>(acc, sb, e, out, pix0, ox0, vwn); break;
      case 2: stem_store<COUT, 2, PX>(acc, sb, e, out, pix0, ox0, vwn); break;
      default: stem_store<COUT, 0, PX>(acc, sb, e, out, pix0, ox0, vwn); break;
    }
  }
}

// First layer fused with the stage's pre-processing: the CTA builds the resized + normalised fp16 input tile of its
// 32 x TH output pixels in shared memory straight from the 8-bit source (cv::resize INTER_LINEAR + Normalize, the
// arithmetic of preproc.cu: resize.cuh) and convolves from there, so the [n, h, w, 8] fp16 network input -- written
// once and read once at 16 B per pixel for 6 B of data -- never exists in HBM (reference: ResizeImgType0 / CrnnResizeImg /
// ClsResizeImg + Normalize + Permute feeding predictor->Run, src/preprocess_op.cpp:19-137, src/ocr_det.cpp:93-120).
// Values and summation order are those of det/crop_preprocess followed by stem_conv_s2x4_kernel<COUT, 2>: the output is
// bit-identical (tests/test_env_paths_gpu.py::test_fused_stem_equals_unfused).
// The axis coefficients of the tile's 65 columns and 2*TH+1 rows are computed once per CTA (they cost two double
// divisions each), not once per pixel.
constexpr int kFsTW = 32;                    // output columns per tile
constexpr int kFsCols = 2 * kFsTW + 1;       // input columns per tile
constexpr int kFsPlane = kFsCols / 4 + 1;    // tile columns are stored de-interleaved by (column mod 4): conflict-free reads
template <int COUT, int KIND>                // KIND 1: DetPreItem (whole images), 2: CropItem (ROIs, right-padded)
__global__ void __launch_bounds__(128, 4)
fused_stem_kernel(const void* __restrict__ items_, NormParams np, float pad_value, int dh, int dw, TV out,
                  const __half* __restrict__ w, const float* __restrict__ bias, ConvGeom g, Epi e,
                  const int* __restrict__ vw, int th, int tiles_x, int tiles_y) {
  pdl_trigger();
  __shared__ __align__(16) float sw[9 * 3 * COUT];
  __shared__ float sb[COUT];
  __shared__ uint2 tile[17][4 * kFsPlane];
  __shared__ resize::Coef cxs[kFsCols], cys[17];
  // the filter is a constant: fetched while the previous kernel drains
  for (int i = threadIdx.x; i < 9 * 3 * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % 3, tap = i / (3 * COUT);
    sw[i] = __half2float(w[(long(co) * 9 + tap) * g.cin_pad + ci]);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  pdl_wait();
  const int tx = blockIdx.x % tiles_x;
  const int ty = (blockIdx.x / tiles_x) % tiles_y;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int ox_t = tx * kFsTW, oy_t = ty * th;
  const int vwn = vw ? vw[n] : out.w;
  const int r = threadIdx.x >> 4, strip = threadIdx.x & 15;
  const int oy = oy_t + r, ox0 = ox_t + strip * 2;
  const bool active = r < th && oy < out.h && ox0 < out.w;
  if (ox_t >= vwn) {  // a tile beyond this row's width (ragged batch): zeros, like stem_store's mask
    if (active) {
      for (int p = 0; p < 2 && ox0 + p < out.w; ++p)
        for (int c0 = 0; c0 < COUT; c0 += 8) {
          H8 o;
          o.u = make_uint4(0, 0, 0, 0);
          st8(out.p + ((long(n) * out.h + oy) * out.w + ox0 + p) * out.pitch + c0, o);
        }
    }
    return;
  }
  resize::Src src;
  int rw, pad_w;   // resized width of this item; columns [rw, pad_w) hold the pad value, [pad_w, dw) zero
  if (KIND == 1) {
    const DetPreItem it = static_cast<const DetPreItem*>(items_)[n];
    src = resize::Src{it.src, it.w, it.h, it.stride};
    rw = dw; pad_w = dw;
  } else {
    const CropItem it = static_cast<const CropItem*>(items_)[n];
    src = resize::Src{it.img + long(it.y) * it.stride + long(it.x) * 3, it.w, it.h, it.stride};
    rw = it.resize_w; pad_w = it.pad_w;
  }
  const int mode = (src.w == rw && src.h == dh) ? 0 : (src.w == 2 * rw && src.h == 2 * dh) ? 1 : 2;  // copy / 2x2 area / bilinear
  const int rows = 2 * th + 1;
  const int ix_t = ox_t * 2 - g.pw, iy_t = oy_t * 2 - g.ph;
  if (mode == 2) {
    if (threadIdx.x < kFsCols) {
      const int ix = ix_t + threadIdx.x;
      if (ix >= 0 && ix < rw) cxs[threadIdx.x] = resize::coef_x(ix, 1.0 / (double(rw) / double(src.w)), src.w);
    } else if (threadIdx.x < kFsCols + rows) {
      const int k = threadIdx.x - kFsCols;
      const int iy = iy_t + k;
      if (iy >= 0 && iy < dh) cys[k] = resize::coef_y(iy, 1.0 / (double(dh) / double(src.h)));
    }
    __syncthreads();
  }
  const __half2 pad2 = __floats2half2_rn(pad_value, pad_value);
  const __half2 pad1 = __floats2half2_rn(pad_value, 0.f);
  for (int i = threadIdx.x; i < rows * kFsCols; i += blockDim.x) {
    const int k = i / kFsCols, j = i - k * kFsCols;
    const int iy = iy_t + k, ix = ix_t + j;
    uint2 v = make_uint2(0u, 0u);   // the convolution's zero padding, and the ragged filler beyond pad_w
    if (iy >= 0 && iy < dh && ix >= 0 && ix < dw) {
      if (ix >= rw) {
        if (ix < pad_w) { v.x = *reinterpret_cast<const uint32_t*>(&pad2); v.y = *reinterpret_cast<const uint32_t*>(&pad1); }
      } else {
        int px[3];
        if (mode == 0) {
          const uint8_t* q = src.p + iy * src.stride + ix * 3;
          px[0] = q[0]; px[1] = q[1]; px[2] = q[2];
        } else if (mode == 1) {
          const uint8_t* q0 = src.p + (2 * iy) * src.stride + (2 * ix) * 3;
          const uint8_t* q1 = q0 + src.stride;
#pragma unroll
          for (int c = 0; c < 3; ++c) px[c] = (int(q0[c]) + int(q0[c + 3]) + int(q1[c]) + int(q1[c + 3]) + 2) >> 2;
        } else {
          resize::resize_px_coef(src, cxs[j], cys[k], px);
        }
        const __half2 ab = __floats2half2_rn(resize::norm1(px[0], np.scale[0], np.shift[0]),
                                             resize::norm1(px[1], np.scale[1], np.shift[1]));
        const __half2 c0 = __floats2half2_rn(resize::norm1(px[2], np.scale[2], np.shift[2]), 0.f);
        v.x = *reinterpret_cast<const uint32_t*>(&ab);
        v.y = *reinterpret_cast<const uint32_t*>(&c0);
      }
    }
    tile[k][(j & 3) * kFsPlane + (j >> 2)] = v;
  }
  __syncthreads();
  if (!active) return;
  float acc[2][COUT];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[p][i] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    float x[5][3];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const uint2 raw = tile[2 * r + ky][(j & 3) * kFsPlane + strip + (j >> 2)];
      const float2 ab = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      x[j][0] = ab.x; x[j][1] = ab.y;
      x[j][2] = __half2float(*reinterpret_cast<const __half*>(&raw.y));
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * COUT);
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
          const float4 wv = wp[q];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const float xv = x[2 * p + kx][ci];
            acc[p][4 * q] = fmaf(xv, wv.x, acc[p][4 * q]);
            acc[p][4 * q + 1] = fmaf(xv, wv.y, acc[p][4 * q + 1]);
            acc[p][4 * q + 2] = fmaf(xv, wv.z, acc[p][4 * q + 2]);
            acc[p][4 * q + 3] = fmaf(xv, wv.w, acc[p][4 * q + 3]);
          }
        }
      }
  }
  const long pix0 = (long(n) * out.h + oy) * out.w + ox0;
  switch (e.act) {
    case 1: stem_store<COUT, 1, 2>(acc, sb, e, out, pix0, ox0, vwn); break;
    case 2: stem_store<COUT, 2, 2>(acc, sb, e, out, pix0, ox0, vwn); break;
    default: stem_store<COUT, 0, 2>(acc, sb, e, out, pix0, ox0, vwn); break;
  }
}

// ---------------------------------------------------------------- SE block
// partial[n][split][cp] = sum over the split's pixels (fp32, fixed order -> deterministic)
// Two ways of cutting a sample into splits:
//  * cbs > 0 (recognizer, whose batches may be ragged): split = (row range, block of gap_col_block() columns).  A lane owns
//    the columns x == x0 + lane (mod lanes) of its block and adds them in (row, column) order, so the order in which
//    its accumulator sees the valid pixels of a text line does not depend on how wide the surrounding tensor is: zero
//    columns add exactly 0, and so do the all-zero blocks a wider tensor appends to the sum over splits.  Pooled
//    values are bit-identical to a dense run of the line's own width.
//  * cbs == 0 (detector / classifier, always dense): split = a contiguous range of the sample's pixels, so that every
//    lane has a full ring of loads in flight whatever the aspect ratio.
// The pixels are staged through shared memory by cp.async: each thread keeps kGapStages 16-byte copies in flight into
// ring slots of its own (bytes in flight do not cost registers, and no block barrier sits in the loop).
constexpr int kGapColBlockDefault = 256;
constexpr int kGapStages = 8;  // 16-byte cp.async copies in flight per thread (its own ring slots: no block barrier)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// sm: kGapStages * kThreads * 16 bytes (ring), reused for the cross-lane reduction
__device__ __forceinline__ void gap_partial_body(TV in, float* __restrict__ partial, int splits, int cbs, int cbw,
                                                 float* sm) {
  const int cgs = (in.c + 7) >> 3;
  const int cp = cgs * 8;
  const int lanes = max(1, kThreads / cgs);
  const int split = blockIdx.x, n = blockIdx.y;
  const int hw = in.h * in.w;
  const int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
  // One accumulator per channel, fed in (row, column) order: the loads are decoupled from the adds by the ring, so
  // a single fixed-order chain costs nothing, and interleaved zero columns leave it unchanged.
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane < lanes) {
    int y0 = 0, y1 = 1, xa, xb;  // rows [y0, y1); this lane's pixels of a row: xa + lane + k * lanes < xb
    long row_pix = 0;
    if (cbs > 0) {
      const int rs = splits / cbs, rsplit = split / cbs, cb = split - rsplit * cbs;
      const int rows_per = (in.h + rs - 1) / rs;
      y0 = rsplit * rows_per;
      y1 = min(in.h, y0 + rows_per);
      xa = cb * cbw;
      xb = min(in.w, xa + cbw);
      row_pix = in.w;
    } else {
      const int chunk = (hw + splits - 1) / splits;
      xa = split * chunk;
      xb = min(hw, xa + chunk);
    }
    const int k_per_row = xb - xa > lane ? (xb - xa - lane + lanes - 1) / lanes : 0;
    const int total = k_per_row * max(0, y1 - y0);
    const __half* base = in.p + (long(n) * hw + xa + lane) * in.pitch + cg * 8;
    uint4* ring = reinterpret_cast<uint4*>(sm) + threadIdx.x;  // slot s at ring[s * kThreads]
    int iy = y0, ik = 0, issued = 0;
    auto issue = [&](int slot) {
      cp_async16(ring + slot * kThreads, base + (long(iy) * row_pix + long(ik) * lanes) * in.pitch);
      if (++ik == k_per_row) { ik = 0; ++iy; }
      ++issued;
    };
#pragma unroll
    for (int s = 0; s < kGapStages; ++s) {  // one group per slot, empty past the end: the wait count stays uniform
      if (s < total) issue(s);
      cp_async_commit();
    }
    for (int j = 0; j < total; ++j) {
      cp_async_wait<kGapStages - 1>();
      const int slot = j % kGapStages;
      H8 v;
      v.u = ring[slot * kThreads];
      float x[8];
      v.to_float(x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x[i];
      if (issued < total) issue(slot);
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // every ring slot has been consumed: the buffer becomes the reduction scratch
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[lane * cp + cg * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cp; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * cp + c];
    partial[(long(n) * splits + split) * cp + c] = s;
  }
}

__global__ void __launch_bounds__(kThreads)
gap_partial_kernel(TV in, float* __restrict__ partial, int splits, int cbs, int cbw) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  gap_partial_body(in, partial, splits, cbs, cbw, sm);
}

// blk: w1[cmid][c], b1[cmid], w2[c][cmid], b2[c]
// SE gate of kSeSamples consecutive samples starting at n0, computed by one block: the two small matrices (c x cmid
// floats each) are read once per block; one thread owns one output neuron of all its samples (no reductions).
// blk: w1t[c][cmid], b1[cmid], w2t[cmid][c], b2[c].  `partial` may have been written by other blocks of the SAME
// kernel (fused pool + gate): it is read through L2 (__ldcg).
constexpr int kSeMaxParts = 16;  // partitions of a contraction in the vector path of se_fc_body

template <int kSeSamples>
__device__ __forceinline__ void se_fc_body(float* sm, const float* partial, int splits, float inv_hw0, int n0, int n_total,
                                           int c, int cmid, const float* __restrict__ blk, float slope, float offset,
                                           float* __restrict__ gate, const int* __restrict__ vw_in, int h) {
  const int cp = (c + 7) / 8 * 8;
  float* pooled = sm;                      // [kSeSamples][cp]
  float* hidden = sm + kSeSamples * cp;    // [kSeSamples][cmid]
  const int ns = min(kSeSamples, n_total - n0);
  for (int t = threadIdx.x; t < kSeSamples * cp; t += blockDim.x) {
    const int sidx = t / cp, i = t - sidx * cp;
    float v = 0.f;
    if (sidx < ns) {
      const int n = n0 + sidx;
      const float inv_hw = vw_in ? 1.f / float(h * vw_in[n]) : inv_hw0;
      float s = 0.f;
      const float* pp = partial + long(n) * splits * cp + i;
      int k = 0;
      for (; k + 8 <= splits; k += 8) {  // eight loads in flight, added in split order
        float q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = __ldcg(pp + long(k + u) * cp);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += q[u];
      }
      for (; k < splits; ++k) s += __ldcg(pp + long(k) * cp);
      v = s * inv_hw;
    }
    pooled[t] = v;
  }
  __syncthreads();
  const float* w1t = blk;
  const float* b1 = w1t + long(cmid) * c;
  const float* w2t = b1 + cmid;
  const float* b2 = w2t + long(c) * cmid;
  if ((c & 3) == 0 && (cmid & 3) == 0) {
    // Vector path (every SE block of det and rec).  The gate is on the critical path of the fused pool + gate kernel
    // (the block that pooled a sample's last split computes it alone), so the two mat-vecs are cut for LATENCY:
    // thread = (four neighbouring neurons, one partition of the contraction), eight 16-byte weight loads in flight,
    // partitions added in index order afterwards.
    float* scratch = hidden + kSeSamples * cmid;  // [P][kSeSamples][cmid] then [P][kSeSamples][c]
    {
      const int quads = cmid >> 2;
      const int parts = max(1, min(min(int(blockDim.x) / quads, kSeMaxParts), c));
      const int per = (c + parts - 1) / parts;
      for (int t = threadIdx.x; t < quads * parts; t += blockDim.x) {
        const int pt = t / quads, mq = t - pt * quads;
        const int i0 = pt * per, i1 = min(c, i0 + per);
        float4 a[kSeSamples];
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wp = reinterpret_cast<const float4*>(w1t) + mq;
        int i = i0;
        for (; i + 8 <= i1; i += 8) {
          float4 w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __ldg(wp + long(i + u) * quads);
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < kSeSamples; ++q) {
              const float x = pooled[q * cp + i + u];
              a[q].x = fmaf(w[u].x, x, a[q].x); a[q].y = fmaf(w[u].y, x, a[q].y);
              a[q].z = fmaf(w[u].z, x, a[q].z); a[q].w = fmaf(w[u].w, x, a[q].w);
            }
        }
        for (; i < i1; ++i) {
          const float4 w = __ldg(wp + long(i) * quads);
#pragma unroll
          for (int q = 0; q < kSeSamples; ++q) {
            const float x = pooled[q * cp + i];
            a[q].x = fmaf(w.x, x, a[q].x); a[q].y = fmaf(w.y, x, a[q].y);
            a[q].z = fmaf(w.z, x, a[q].z); a[q].w = fmaf(w.w, x, a[q].w);
          }
        }
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q)
          reinterpret_cast<float4*>(scratch + (long(pt) * kSeSamples + q) * cmid)[mq] = a[q];
      }
      __syncthreads();
      for (int t = threadIdx.x; t < kSeSamples * cmid; t += blockDim.x) {
        const int q = t / cmid, m = t - q * cmid;
        float v = b1[m];
        for (int pt = 0; pt < parts; ++pt) v += scratch[(long(pt) * kSeSamples + q) * cmid + m];
        hidden[t] = fmaxf(v, 0.f);
      }
      __syncthreads();
    }
    {
      const int quads = c >> 2;
      const int parts = max(1, min(min(int(blockDim.x) / quads, kSeMaxParts), cmid));
      const int per = (cmid + parts - 1) / parts;
      for (int t = threadIdx.x; t < quads * parts; t += blockDim.x) {
        const int pt = t / quads, iq = t - pt * quads;
        const int m0 = pt * per, m1 = min(cmid, m0 + per);
        float4 a[kSeSamples];
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wp = reinterpret_cast<const float4*>(w2t) + iq;
        int m = m0;
        for (; m + 8 <= m1; m += 8) {
          float4 w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __ldg(wp + long(m + u) * quads);
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < kSeSamples; ++q) {
              const float x = hidden[q * cmid + m + u];
              a[q].x = fmaf(w[u].x, x, a[q].x); a[q].y = fmaf(w[u].y, x, a[q].y);
              a[q].z = fmaf(w[u].z, x, a[q].z); a[q].w = fmaf(w[u].w, x, a[q].w);
            }
        }
        for (; m < m1; ++m) {
          const float4 w = __ldg(wp + long(m) * quads);
#pragma unroll
          for (int q = 0; q < kSeSamples; ++q) {
            const float x = hidden[q * cmid + m];
            a[q].x = fmaf(w.x, x, a[q].x); a[q].y = fmaf(w.y, x, a[q].y);
            a[q].z = fmaf(w.z, x, a[q].z); a[q].w = fmaf(w.w, x, a[q].w);
          }
        }
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q)
          reinterpret_cast<float4*>(scratch + (long(pt) * kSeSamples + q) * c)[iq] = a[q];
      }
      __syncthreads();
      for (int t = threadIdx.x; t < kSeSamples * cp; t += blockDim.x) {
        const int q = t / cp, i = t - q * cp;
        float v = 0.f;
        if (i < c) {
          v = b2[i];
          for (int pt = 0; pt < parts; ++pt) v += scratch[(long(pt) * kSeSamples + q) * c + i];
          v = __saturatef(fmaf(v, slope, offset));
        }
        if (q < ns) gate[long(n0 + q) * cp + i] = v;
      }
    }
    return;
  }
  // Scalar path (channel counts that are not multiples of four: some classifier blocks).
  // fc1: thread = (half of the input channels, hidden neuron); 8 independent weight loads in flight per thread
  float* part = hidden + kSeSamples * cmid;  // [2][kSeSamples][cmid] partial sums
  for (int t = threadIdx.x; t < 2 * cmid; t += blockDim.x) {
    const int half = t / cmid, m = t - half * cmid;
    const int i0 = half * (c / 2), i1 = half ? c : c / 2;
    float a[kSeSamples];
#pragma unroll
    for (int q = 0; q < kSeSamples; ++q) a[q] = 0.f;
    int i = i0;
    for (; i + 8 <= i1; i += 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldg(w1t + long(i + u) * cmid + m);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) a[q] = fmaf(w[u], pooled[q * cp + i + u], a[q]);
    }
    for (; i < i1; ++i) {
      const float w = __ldg(w1t + long(i) * cmid + m);
#pragma unroll
      for (int q = 0; q < kSeSamples; ++q) a[q] = fmaf(w, pooled[q * cp + i], a[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeSamples; ++q) part[(half * kSeSamples + q) * cmid + m] = a[q];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kSeSamples * cmid; t += blockDim.x) {
    const int m = t % cmid;
    hidden[t] = fmaxf(part[t] + part[kSeSamples * cmid + t] + b1[m], 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cp; i += blockDim.x) {
    float a[kSeSamples];
#pragma unroll
    for (int q = 0; q < kSeSamples; ++q) a[q] = 0.f;
    if (i < c) {
      const float bb = b2[i];
#pragma unroll
      for (int q = 0; q < kSeSamples; ++q) a[q] = bb;
      int m = 0;
      for (; m + 8 <= cmid; m += 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(w2t + long(m + u) * c + i);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < kSeSamples; ++q) a[q] = fmaf(w[u], hidden[q * cmid + m + u], a[q]);
      }
      for (; m < cmid; ++m) {
        const float w = __ldg(w2t + long(m) * c + i);
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) a[q] = fmaf(w, hidden[q * cmid + m], a[q]);
      }
#pragma unroll
      for (int q = 0; q < kSeSamples; ++q) a[q] = __saturatef(fmaf(a[q], slope, offset));
    }
#pragma unroll
    for (int q = 0; q < kSeSamples; ++q)
      if (q < ns) gate[long(n0 + q) * cp + i] = a[q];
  }
}

template <int kSeSamples>
__global__ void __launch_bounds__(kThreads)
se_fc_kernel(const float* __restrict__ partial, int splits, float inv_hw0, int n_total, int c, int cmid,
             const float* __restrict__ blk, float slope, float offset, float* __restrict__ gate,
             const int* __restrict__ vw_in, int h) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  se_fc_body<kSeSamples>(sm, partial, splits, inv_hw0, blockIdx.x * kSeSamples, n_total, c, cmid, blk, slope, offset, gate,
                         vw_in, h);
}

// Global average pool fused with the SE gate: the block that finishes a sample's LAST row split (ticket counter) turns
// the sample's partial sums into its gate right away -- one launch less per SE block, and the gates of the first
// samples are computed while the later samples are still being pooled.  The sums are added in split order by that
// one block, so the result does not depend on which block came last.
__global__ void __launch_bounds__(kThreads)
gap_se_kernel(TV in, float* __restrict__ partial, int splits, int cbs, int cbw, SeFuse fc) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  __shared__ int s_last;
  gap_partial_body(in, partial, splits, cbs, cbw, sm);
  // Release / acquire through the ticket, by ONE thread: the block barrier orders every thread's partial sums before
  // thread 0's gpu-scope acq_rel atomic (cumulativity), and the last block's threads read them after the second
  // barrier, through L2 (__ldcg).  A __threadfence() by all 256 threads was the kernel's top stall (ncu: ERRBAR).
  __syncthreads();
  const int n = blockIdx.y;
  if (threadIdx.x == 0) {
    int ticket;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(fc.counters + n) : "memory");
    s_last = ticket == splits - 1;
    if (s_last) fc.counters[n] = 0;  // ready for the next launch (every other block of this sample is past its atomic)
  }
  __syncthreads();
  if (!s_last) return;
  se_fc_body<1>(sm, partial, splits, fc.inv_hw, n, in.n, fc.c, fc.cmid, fc.blk, fc.slope, fc.offset, fc.gate, fc.vw_in, fc.h);
  if (!fc.sc_out) return;
  // Gate applied by the same block (x * gate [+ x], the arithmetic of scale_kernel): one launch less per SE block, and
  // the map is still in L2 from the pooling pass.  Blocks of different samples work in parallel; nothing waits on
  // another block, so any number of these kernels can share the GPU.
  __syncthreads();  // the gate was written by this block's threads
  const int cgs = (in.c + 7) >> 3;
  const int cp = cgs * 8;
  for (int i = threadIdx.x; i < cp; i += blockDim.x) sm[i] = fc.gate[long(n) * cp + i];
  __syncthreads();
  const long hw = long(in.h) * in.w;
  const long total = hw * cgs;
  const __half* src = in.p + long(n) * hw * in.pitch;
  __half* dst = fc.sc_out + long(n) * hw * fc.sc_out_pitch;
  for (long t = threadIdx.x; t < total; t += blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    float x[8];
    ld8(src + pix * in.pitch + cg * 8).to_float(x);
    const float* g = sm + cg * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fc.sc_add_x ? fmaf(x[i], g[i], x[i]) : x[i] * g[i];
    H8 o;
    o.from_float(x);
    st8(dst + pix * fc.sc_out_pitch + cg * 8, o);
  }
}

__global__ void __launch_bounds__(kThreads)
scale_kernel(TV in, const float* __restrict__ gate, int add_x, TV out) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (in.c + 7) >> 3;
  const int cp = cgs * 8;
  const long hw = long(in.h) * in.w;
  const long total = long(in.n) * hw * cgs;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    const int n = int(pix / hw);
    float x[8];
    ld8(in.p + pix * in.pitch + cg * 8).to_float(x);
    const float4* gp = reinterpret_cast<const float4*>(gate + long(n) * cp + cg * 8);
    const float4 g0 = gp[0], g1 = gp[1];
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = add_x ? fmaf(x[i], g[i], x[i]) : x[i] * g[i];
    H8 o;
    o.from_float(x);
    st8(out.p + pix * out.pitch + cg * 8, o);
  }
}

// ---------------------------------------------------------------- glue
__global__ void __launch_bounds__(kThreads) upadd_kernel(TV a, TV b, TV out) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (a.c + 7) >> 3;
  const long total = long(a.n) * a.h * a.w * cgs;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    const int x = int(pix % a.w), y = int((pix / a.w) % a.h), n = int(pix / (long(a.w) * a.h));
    float u[8], v[8];
    ld8(a.p + pix * a.pitch + cg * 8).to_float(u);
    ld8(b.p + ((long(n) * b.h + (y >> 1)) * b.w + (x >> 1)) * b.pitch + cg * 8).to_float(v);
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] += v[i];
    H8 o;
    o.from_float(u);
    st8(out.p + pix * out.pitch + cg * 8, o);
  }
}

struct UpCatArgs {
  TV in[4];
  int shift[4];
  int coff[4];
  int nin;
};

__global__ void __launch_bounds__(kThreads) upcat_kernel(UpCatArgs a, TV out) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (out.c + 7) >> 3;
  const long total = long(out.n) * out.h * out.w * cgs;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    const int x = int(pix % out.w), y = int((pix / out.w) % out.h), n = int(pix / (long(out.w) * out.h));
    const int c0 = cg * 8;
    int k = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (j < a.nin && c0 >= a.coff[j]) k = j;
    const TV& s = a.in[k];
    const int sh = a.shift[k];
    H8 v = ld8(s.p + ((long(n) * s.h + (y >> sh)) * s.w + (x >> sh)) * s.pitch + (c0 - a.coff[k]));
    st8(out.p + pix * out.pitch + c0, v);
  }
}

__global__ void __launch_bounds__(kThreads) add_kernel(TV a, TV b, TV out) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (a.c + 7) >> 3;
  const long total = long(a.n) * a.h * a.w * cgs;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    float u[8], v[8];
    ld8(a.p + pix * a.pitch + cg * 8).to_float(u);
    ld8(b.p + pix * b.pitch + cg * 8).to_float(v);
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] += v[i];
    H8 o;
    o.from_float(u);
    st8(out.p + pix * out.pitch + cg * 8, o);
  }
}

// windows are clipped to the input; avg divides by the clipped size (Paddle exclusive=true)
__global__ void __launch_bounds__(kThreads)
pool_kernel(TV in, TV out, int kh, int kw, int sh, int sw, int is_max, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const int cgs = (in.c + 7) >> 3;
  const long total = long(out.n) * out.h * out.w * cgs;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int cg = int(t % cgs);
    const long pix = t / cgs;
    const int ox = int(pix % out.w), oy = int((pix / out.w) % out.h), n = int(pix / (long(out.w) * out.h));
    const int y0 = oy * sh, y1 = min(y0 + kh, in.h), x0 = ox * sw, x1 = min(x0 + kw, in.w);
    if (vw && ox >= vw[n]) {
      H8 z;
      z.u = make_uint4(0, 0, 0, 0);
      st8(out.p + pix * out.pitch + cg * 8, z);
      continue;
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = is_max ? -FLT_MAX : 0.f;
    if (kh <= 3 && kw <= 2) {
      // windows of up to 3 x 2 (the recognizer's pool): all loads first, then the same y-outer / x-inner order of adds
      H8 v[6];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          if (y0 + dy < y1 && x0 + dx < x1)
            v[dy * 2 + dx] = ld8(in.p + ((long(n) * in.h + y0 + dy) * in.w + x0 + dx) * in.pitch + cg * 8);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          if (y0 + dy < y1 && x0 + dx < x1) {
            float f[8];
            v[dy * 2 + dx].to_float(f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = is_max ? fmaxf(acc[i], f[i]) : acc[i] + f[i];
          }
    } else {
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          float v[8];
          ld8(in.p + ((long(n) * in.h + y) * in.w + x) * in.pitch + cg * 8).to_float(v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = is_max ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
        }
    }
    if (!is_max) {
      const float inv = 1.f / float((y1 - y0) * (x1 - x0));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= inv;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (cg * 8 + i >= in.c) acc[i] = 0.f;
    H8 o;
    o.from_float(acc);
    st8(out.p + pix * out.pitch + cg * 8, o);
  }
}

// ---------------------------------------------------------------- SVTR neck
// LPT lanes (16 for C <= 128, else 32; 8 channels per lane, C <= 256) own one token, and a lane group normalises
// kLnTok consecutive tokens at a time: all their 16-byte loads are issued before the first reduction, so a warp has
// 4-8 tokens in flight instead of one load followed by two shuffle chains.
constexpr int kLnTok = 4;

template <int LPT>
__global__ void __launch_bounds__(kThreads)
layernorm_kernel(TV in, TV out, const float* __restrict__ gb, float eps, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  const long rows = long(in.n) * in.h * in.w;
  const int lane = threadIdx.x % LPT;
  const long group = (blockIdx.x * long(blockDim.x) + threadIdx.x) / LPT;
  const long row0 = group * kLnTok;
  if (row0 >= rows) return;  // uniform over the lane group (and over the warp's shuffles: masks below are per group)
  const unsigned gmask = LPT == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
  const int c0 = lane * 8;
  const bool have = c0 < in.c;
  const long hw = long(in.h) * in.w;
  H8 raw[kLnTok];
  bool live[kLnTok], pad[kLnTok];
#pragma unroll
  for (int t = 0; t < kLnTok; ++t) {
    const long row = row0 + t;
    live[t] = row < rows;
    pad[t] = live[t] && vw && int(row % in.w) >= vw[row / hw];
    raw[t].u = make_uint4(0, 0, 0, 0);
    if (live[t] && !pad[t] && have) raw[t] = ld8(in.p + row * in.pitch + c0);
  }
  float g[8], bta[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    g[i] = c < in.c ? gb[c] : 0.f;
    bta[i] = c < in.c ? gb[in.c + c] : 0.f;
  }
  const float inv_c = 1.f / float(in.c);
#pragma unroll
  for (int t = 0; t < kLnTok; ++t) {
    if (!live[t]) continue;
    const long row = row0 + t;
    float x[8];
    raw[t].to_float(x);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (c0 + i < in.c) ? x[i] : 0.f;
#pragma unroll
    for (int o = LPT / 2; o; o >>= 1) s += __shfl_xor_sync(gmask, s, o);
    const float mean = s * inv_c;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = (c0 + i < in.c) ? x[i] - mean : 0.f;
      v = fmaf(d, d, v);
    }
#pragma unroll
    for (int o = LPT / 2; o; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    const float rstd = rsqrtf(v * inv_c + eps);
    if (have) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (c0 + i < in.c && !pad[t]) ? (x[i] - mean) * rstd * g[i] + bta[i] : 0.f;
      H8 o;
      o.from_float(x);
      st8(out.p + row * out.pitch + c0, o);
    }
  }
}

// qkv row layout (Paddle reshape [N,T,3,heads,d]): [which(3)][head][d].  One block per (n, head).
// smem: K[T][d], V[T][d], P[warps][T]
__global__ void __launch_bounds__(128)
attention_kernel(TV qkv, TV out, int heads, int hd, float scale, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int Tfull = qkv.h * qkv.w;
  const int n = blockIdx.x / heads, head = blockIdx.x % heads;
  const int T = vw ? vw[n] : Tfull;  // valid tokens of this sequence
  float* K = sm;
  float* V = K + Tfull * hd;
  float* P = V + Tfull * hd;
  const __half* base = qkv.p + long(n) * Tfull * qkv.pitch;
  const int C = heads * hd;
  for (int i = T * hd + threadIdx.x; i < Tfull * hd; i += blockDim.x)  // queries beyond the valid length -> zeros
    out.p[(long(n) * Tfull + i / hd) * out.pitch + head * hd + i % hd] = __float2half_rn(0.f);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    K[i] = __half2float(base[long(t) * qkv.pitch + C + head * hd + d]);
    V[i] = __half2float(base[long(t) * qkv.pitch + 2 * C + head * hd + d]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* p = P + warp * T;
  for (int tq = warp; tq < T; tq += nwarps) {
    float q[32];
#pragma unroll
    for (int d = 0; d < 32; ++d)
      q[d] = d < hd ? __half2float(base[long(tq) * qkv.pitch + head * hd + d]) * scale : 0.f;
    float mx = -FLT_MAX;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d)
        if (d < hd) s = fmaf(q[d], K[j * hd + d], s);
      p[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = __expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    // lane d accumulates output dim d (hd <= 32)
    float acc = 0.f;
    if (lane < hd)
      for (int j = 0; j < T; ++j) acc = fmaf(p[j], V[j * hd + lane], acc);
    if (lane < hd)
      out.p[(long(n) * Tfull + tq) * out.pitch + head * hd + lane] = __float2half_rn(acc / sum);
    __syncwarp();
  }
}

// ---------------------------------------------------------------- heads
// blk: w1[q][ci][cm], b1[cm], w2[cm][4], b2.  One thread per 1/4-scale pixel -> 4x4 output block.
template <int CIN, int CMID>
__global__ void __launch_bounds__(128, 4)
dbhead_kernel(TV in, const float* __restrict__ blk, float* __restrict__ prob,
              uint8_t* __restrict__ bitmap, int thresh_u8) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sw1[4 * CIN * CMID];
  __shared__ float sb1[CMID];
  __shared__ float sw2[CMID * 4];
  __shared__ float sb2;
  for (int i = threadIdx.x; i < 4 * CIN * CMID; i += blockDim.x) sw1[i] = blk[i];
  for (int i = threadIdx.x; i < CMID; i += blockDim.x) sb1[i] = blk[4 * CIN * CMID + i];
  for (int i = threadIdx.x; i < CMID * 4; i += blockDim.x) sw2[i] = blk[4 * CIN * CMID + CMID + i];
  if (threadIdx.x == 0) sb2 = blk[4 * CIN * CMID + CMID + CMID * 4];
  __syncthreads();
  const long npix = long(in.n) * in.h * in.w;
  const long pix = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (pix >= npix) return;
  const int x = int(pix % in.w), y = int((pix / in.w) % in.h), n = int(pix / (long(in.w) * in.h));
  float xin[CIN];
#pragma unroll
  for (int c8 = 0; c8 < CIN / 8; ++c8) ld8(in.p + pix * in.pitch + c8 * 8).to_float(xin + c8 * 8);
  const int OW = in.w * 4, OH = in.h * 4;
  float pv[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float mid[CMID];
#pragma unroll
    for (int cm = 0; cm < CMID; ++cm) mid[cm] = sb1[cm];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float xv = xin[ci];
#pragma unroll
      for (int cm = 0; cm < CMID; ++cm) mid[cm] = fmaf(xv, sw1[(q * CIN + ci) * CMID + cm], mid[cm]);
    }
    float o[4] = {sb2, sb2, sb2, sb2};
#pragma unroll
    for (int cm = 0; cm < CMID; ++cm) {
      const float m = fmaxf(mid[cm], 0.f);
#pragma unroll
      for (int r = 0; r < 4; ++r) o[r] = fmaf(m, sw2[cm * 4 + r], o[r]);
    }
    const int dy = q >> 1, dx = q & 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) pv[dy * 2 + (r >> 1)][dx * 2 + (r & 1)] = 1.f / (1.f + __expf(-o[r]));
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long o = (long(n) * OH + (y * 4 + r)) * OW + x * 4;
    *reinterpret_cast<float4*>(prob + o) = make_float4(pv[r][0], pv[r][1], pv[r][2], pv[r][3]);
    if (bitmap) {
      uchar4 b;
      // reference src/ocr_det.cpp:146,151-154: cbuf = (uchar)(p*255); bit = cbuf > floor(thresh*255)
      b.x = (int((unsigned char)(pv[r][0] * 255.f)) > thresh_u8) ? 255 : 0;
      b.y = (int((unsigned char)(pv[r][1] * 255.f)) > thresh_u8) ? 255 : 0;
      b.z = (int((unsigned char)(pv[r][2] * 255.f)) > thresh_u8) ? 255 : 0;
      b.w = (int((unsigned char)(pv[r][3] * 255.f)) > thresh_u8) ? 255 : 0;
      *reinterpret_cast<uchar4*>(bitmap + o) = b;
    }
  }
}

// blk: w[cin][cout], b[cout]; one warp per image
__global__ void fc_softmax_kernel(const float* __restrict__ partial, int splits, float inv_hw, int cin,
                                  int cout, const float* __restrict__ blk, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x, lane = threadIdx.x;
  const int cp = (cin + 7) / 8 * 8;
  float logit[8];
  for (int o = 0; o < cout; ++o) {
    float s = 0.f;
    for (int i = lane; i < cin; i += 32) {
      float p = 0.f;
      for (int k = 0; k < splits; ++k) p += partial[(long(n) * splits + k) * cp + i];
      s = fmaf(p * inv_hw, blk[long(i) * cout + o], s);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    logit[o] = s + blk[long(cin) * cout + o];
  }
  if (lane == 0) {
    float mx = logit[0];
    for (int o = 1; o < cout; ++o) mx = fmaxf(mx, logit[o]);
    float sum = 0.f;
    for (int o = 0; o < cout; ++o) { logit[o] = expf(logit[o] - mx); sum += logit[o]; }
    for (int o = 0; o < cout; ++o) out[long(n) * cout + o] = logit[o] / sum;
  }
}

// CUDA-core CTC head: 8 tokens per block share each weight row; per token running (max, argmax, sum).
__global__ void __launch_bounds__(kThreads)
ctc_head_simt_kernel(TV feat, const __half* __restrict__ w, const float* __restrict__ bias, int cin_pad,
                     int ncls_pad, int* __restrict__ idx, float* __restrict__ prob, const int* __restrict__ vw) {
  pdl_trigger();
  pdl_wait();
  constexpr int ROWS = 8;
  extern __shared__ float sm[];  // feat[ROWS][cin_pad]
  const long rows = long(feat.n) * feat.h * feat.w;
  const long r0 = long(blockIdx.x) * ROWS;
  for (int i = threadIdx.x; i < ROWS * cin_pad; i += blockDim.x) {
    const int r = i / cin_pad, c = i % cin_pad;
    sm[i] = (r0 + r < rows && c < feat.c) ? __half2float(feat.p[(r0 + r) * feat.pitch + c]) : 0.f;
  }
  __syncthreads();
  float mx[ROWS], sum[ROWS];
  int am[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) { mx[r] = -FLT_MAX; sum[r] = 0.f; am[r] = 0; }
  for (int o = threadIdx.x; o < ncls_pad; o += blockDim.x) {
    float acc[ROWS];
    const float b = bias[o];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = b;
    const __half* wp = w + long(o) * cin_pad;
    for (int c8 = 0; c8 < cin_pad; c8 += 8) {
      float wv[8];
      ld8(wp + c8).to_float(wv);
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[r] = fmaf(wv[i], sm[r * cin_pad + c8 + i], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float v = acc[r];
      if (v > mx[r]) { sum[r] = sum[r] * __expf(mx[r] - v) + 1.f; mx[r] = v; am[r] = o; }
      else sum[r] += __expf(v - mx[r]);
    }
  }
  // block reduce (max, first argmax, rescaled sum)
  __shared__ float smx[kThreads / 32][ROWS], ssum[kThreads / 32][ROWS];
  __shared__ int sam[kThreads / 32][ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float m = mx[r], s = sum[r];
    int a = am[r];
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, d);
      const float s2 = __shfl_xor_sync(0xffffffffu, s, d);
      const int a2 = __shfl_xor_sync(0xffffffffu, a, d);
      const float nm = fmaxf(m, m2);
      s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
      if (m2 > m || (m2 == m && a2 < a)) a = a2;
      m = nm;
    }
    if (lane == 0) { smx[warp][r] = m; ssum[warp][r] = s; sam[warp][r] = a; }
  }
  __syncthreads();
  if (threadIdx.x < ROWS && r0 + threadIdx.x < rows) {
    const int r = threadIdx.x;
    float m = smx[0][r], s = ssum[0][r];
    int a = sam[0][r];
    for (int k = 1; k < kThreads / 32; ++k) {
      const float m2 = smx[k][r], s2 = ssum[k][r];
      const int a2 = sam[k][r];
      const float nm = fmaxf(m, m2);
      s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
      if (m2 > m || (m2 == m && a2 < a)) a = a2;
      m = nm;
    }
    const long row = r0 + r;
    const bool pad = vw && int(row % feat.w) >= vw[row / (long(feat.h) * feat.w)];
    idx[row] = pad ? 0 : a;  // beyond the valid length: blank, which the CTC collapse skips
    prob[row] = pad ? 0.f : 1.f / s;
  }
}

__global__ void __launch_bounds__(kThreads) nhwc_to_nchw_f32_kernel(TV in, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long total = long(in.n) * in.c * in.h * in.w;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int x = int(t % in.w);
    const int y = int((t / in.w) % in.h);
    const int c = int((t / (long(in.w) * in.h)) % in.c);
    const int n = int(t / (long(in.w) * in.h * in.c));
    out[t] = __half2float(in.p[((long(n) * in.h + y) * in.w + x) * in.pitch + c]);
  }
}

__global__ void __launch_bounds__(kThreads)
nchw3_to_input_kernel(const float* __restrict__ in, int n, int h, int w, __half* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long total = long(n) * h * w;
  const long hw = long(h) * w;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const long img = t / hw, p = t - img * hw;
    float f[8] = {in[(img * 3 + 0) * hw + p], in[(img * 3 + 1) * hw + p], in[(img * 3 + 2) * hw + p], 0, 0, 0, 0, 0};
    H8 o;
    o.from_float(f);
    st8(out + t * 8, o);
  }
}

inline int grid_for(long total, int threads = kThreads) {
  long b = (total + threads - 1) / threads;
  const long cap = 148L * 16;  // grid-stride: a few waves over the 148 SMs
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

bool fused_stem_eligible(const TV& in, const TV& out, const ConvGeom& g, const Epi& e) {
  static const bool off = getenv("B200OCR_FUSED_STEM") && atoi(getenv("B200OCR_FUSED_STEM")) == 0;
  return !off && in.c == 3 && g.kh == 3 && g.kw == 3 && g.sh == 2 && g.sw == 2 && g.ph == 1 && g.pw == 1 && e.res == nullptr &&
         (out.c == 8 || out.c == 16) && e.act >= 0 && e.act <= 2 && out.h == (in.h + 1) / 2 && out.w == (in.w + 1) / 2;
}

void launch_fused_stem(const StemSource& src, const TV& in, const TV& out, const __half* w, const float* bias,
                       const ConvGeom& g, const Epi& e, cudaStream_t s, const int* vw) {
  // rows per tile: 8, or 7 when that divides the output height (the recognizer at rec_img_h 28: 14 rows)
  const int th = (out.h % 8 != 0 && out.h % 7 == 0) ? 7 : 8;
  const int tiles_x = (out.w + kFsTW - 1) / kFsTW, tiles_y = (out.h + th - 1) / th;
  const dim3 grid(unsigned(long(out.n) * tiles_x * tiles_y));
  auto go = [&](auto kern) {
    launch_k(kern, grid, dim3(128), 0, s, src.items, src.np, src.pad_value, in.h, in.w, out, w, bias, g, e, vw, th, tiles_x, tiles_y);
  };
  if (src.kind == 1) { if (out.c == 16) go(fused_stem_kernel<16, 1>); else go(fused_stem_kernel<8, 1>); }
  else { if (out.c == 16) go(fused_stem_kernel<16, 2>); else go(fused_stem_kernel<8, 2>); }
}

void launch_conv_simt(const TV& in, const TV& out, const __half* w, const float* bias,
                      const ConvGeom& g, const Epi& e, cudaStream_t s, const int* vw) {
  if (in.c == 3 && in.pitch == 8 && g.kh == 3 && g.kw == 3 && e.res == nullptr && (out.c == 8 || out.c == 16)) {
    if (g.sh == 2 && g.sw == 2 && out.w >= 8 && e.act >= 0 && e.act <= 2 && !getenv("B200OCR_OLD_STEM")) {
      static const int px = getenv("B200OCR_STEM_PX") ? atoi(getenv("B200OCR_STEM_PX")) : 2;
      const int sg = grid_for(long(out.n) * out.h * ((out.w + px - 1) / px), 128);
      if (out.c == 16 && px == 4) launch_k(stem_conv_s2x4_kernel<16, 4>, dim3(sg), dim3(128), 0, s, in, out, w, bias, g, e, vw);
      else if (out.c == 16) launch_k(stem_conv_s2x4_kernel<16, 2>, dim3(sg), dim3(128), 0, s, in, out, w, bias, g, e, vw);
      else if (px == 4) launch_k(stem_conv_s2x4_kernel<8, 4>, dim3(sg), dim3(128), 0, s, in, out, w, bias, g, e, vw);
      else launch_k(stem_conv_s2x4_kernel<8, 2>, dim3(sg), dim3(128), 0, s, in, out, w, bias, g, e, vw);
      return;
    }
    const int grid = grid_for(long(out.n) * out.h * out.w);
    if (out.c == 16) launch_k(stem_conv_kernel<16>, dim3(grid), dim3(kThreads), 0, s, in, out, w, bias, g, e, vw);
    else launch_k(stem_conv_kernel<8>, dim3(grid), dim3(kThreads), 0, s, in, out, w, bias, g, e, vw);
    return;
  }
  const long total = long(out.n) * out.h * out.w * ((out.c + 7) / 8);
  launch_k(conv_simt_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, out, w, bias, g, e, vw);
}

void launch_dwconv(const TV& in, const TV& out, const float* wb, const __half* wh, const ConvGeom& g, const Epi& e,
                   cudaStream_t s, const int* vw) {
  if (launch_dwconv_tile(in, out, wb, wh, g, e, s, vw)) return;
  {
    // register-blocked strips along x whenever the row is long enough to fill them
    constexpr int S = 4;
    const long st = long(out.n) * out.h * ((out.w + S - 1) / S) * ((out.c + 7) / 8);
    const int sg = grid_for(st);
    if (out.w >= 2 * S && e.res == nullptr && wh == nullptr) {
      if (g.kh == 3 && g.kw == 3 && g.sw == 1) { launch_k(dwconv_strip_f32w_kernel<3, 3, 1, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, g, e, vw); return; }
      if (g.kh == 3 && g.kw == 3 && g.sw == 2) { launch_k(dwconv_strip_f32w_kernel<3, 3, 2, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, g, e, vw); return; }
      if (g.kh == 5 && g.kw == 5 && g.sw == 1) { launch_k(dwconv_strip_f32w_kernel<5, 5, 1, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, g, e, vw); return; }
      if (g.kh == 5 && g.kw == 5 && g.sw == 2) { launch_k(dwconv_strip_f32w_kernel<5, 5, 2, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, g, e, vw); return; }
    }
    if (out.w >= 2 * S && e.res == nullptr && wh != nullptr) {
      if (g.kh == 3 && g.kw == 3 && g.sw == 1) { launch_k(dwconv_strip_kernel<3, 3, 1, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, wh, g, e, vw); return; }
      if (g.kh == 3 && g.kw == 3 && g.sw == 2) { launch_k(dwconv_strip_kernel<3, 3, 2, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, wh, g, e, vw); return; }
      if (g.kh == 5 && g.kw == 5 && g.sw == 1) { launch_k(dwconv_strip_kernel<5, 5, 1, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, wh, g, e, vw); return; }
      if (g.kh == 5 && g.kw == 5 && g.sw == 2) { launch_k(dwconv_strip_kernel<5, 5, 2, S>, dim3(sg), dim3(kThreads), 0, s, in, out, wb, wh, g, e, vw); return; }
    }
  }
  const long total = long(out.n) * out.h * out.w * ((out.c + 7) / 8);
  const int grid = grid_for(total);
  if (g.kh == 3 && g.kw == 3) launch_k(dwconv_kernel<3, 3>, dim3(grid), dim3(kThreads), 0, s, in, out, wb, g, e, vw);
  else if (g.kh == 5 && g.kw == 5) launch_k(dwconv_kernel<5, 5>, dim3(grid), dim3(kThreads), 0, s, in, out, wb, g, e, vw);
  else throw std::runtime_error("depthwise convolution: only 3x3 and 5x5 filters are implemented");
}

// pooled [S][cp] + hidden [S][cmid] + scratch: [2][S][cmid] (scalar path) or [parts][S][cmid | c] with
// parts * cmid, parts * c <= 4 * kThreads (vector path)
static int gap_col_block() {
  static const int v = [] {
    const char* e = getenv("B200OCR_GAP_COLBLOCK");
    return e ? std::max(8, atoi(e)) : kGapColBlockDefault;
  }();
  return v;
}

static size_t se_fc_smem_floats(int S, int cp, int cmid) {
  return size_t(S) * (size_t(cp) + cmid + std::max(2 * cmid, 4 * kThreads));
}

int gap_splits(int n, int h, int w, int c, bool ragged_safe) {
  if (ragged_safe) {
    // row ranges x fixed-width column blocks: depends on the height and on the number of column blocks only, and a
    // wider (ragged) tensor only appends all-zero blocks, so every text line is reduced in the order a dense batch of
    // its own width would use (see gap_partial_body)
    const int rs = h < 1 ? 1 : (h > 8 ? 8 : h);
    return rs * ((w + gap_col_block() - 1) / gap_col_block());
  }
  // dense: splits of >= 8 pixels per lane (two rounds of the unrolled loop), at most 32 per sample; a function of the
  // sample's shape only, never of the batch size, so that batching leaves every pooled value unchanged
  (void)n;
  const int cgs = (c + 7) / 8;
  const int lanes = kThreads / cgs > 0 ? kThreads / cgs : 1;
  return std::max(1, std::min(h * w / (lanes * 8), 32));
}

void launch_gap_partial(const TV& in, float* partial, int splits, bool ragged_safe, cudaStream_t s, const SeFuse* fuse) {
  const int cgs = (in.c + 7) / 8;
  const int lanes = kThreads / cgs > 0 ? kThreads / cgs : 1;
  const int cbw = gap_col_block();
  const int cbs = ragged_safe ? (in.w + cbw - 1) / cbw : 0;
  size_t smem = std::max(size_t(lanes) * cgs * 8 * sizeof(float), size_t(kGapStages) * kThreads * 16);
  if (fuse) {
    smem = std::max(smem, se_fc_smem_floats(1, cgs * 8, fuse->cmid) * sizeof(float));
    launch_k(gap_se_kernel, dim3(dim3(splits, in.n)), dim3(kThreads), smem, s, in, partial, splits, cbs, cbw, *fuse);
    return;
  }
  launch_k(gap_partial_kernel, dim3(dim3(splits, in.n)), dim3(kThreads), smem, s, in, partial, splits, cbs, cbw);
}

void launch_se_fc(const float* partial, int splits, int hw, int n, int c, int cmid, const float* blk,
                  float slope, float offset, float* gate, cudaStream_t s, const int* vw_in, int h) {
  const int cp = (c + 7) / 8 * 8;
  // samples per block: share the weight reads between samples once there are enough samples to fill the SMs
  const int S = n >= 4 * 148 ? 4 : (n >= 2 * 148 ? 2 : 1);
  const size_t smem = se_fc_smem_floats(S, cp, cmid) * sizeof(float);
  const float inv = 1.f / float(hw);
  if (S == 4) launch_k(se_fc_kernel<4>, dim3((n + 3) / 4), dim3(kThreads), smem, s, partial, splits, inv, n, c, cmid, blk, slope, offset, gate, vw_in, h);
  else if (S == 2) launch_k(se_fc_kernel<2>, dim3((n + 1) / 2), dim3(kThreads), smem, s, partial, splits, inv, n, c, cmid, blk, slope, offset, gate, vw_in, h);
  else launch_k(se_fc_kernel<1>, dim3(n), dim3(kThreads), smem, s, partial, splits, inv, n, c, cmid, blk, slope, offset, gate, vw_in, h);
}

void launch_scale(const TV& in, const float* gate, bool add_x, const TV& out, cudaStream_t s) {
  const long total = long(in.n) * in.h * in.w * ((in.c + 7) / 8);
  launch_k(scale_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, gate, add_x ? 1 : 0, out);
}

void launch_upadd(const TV& a, const TV& b, const TV& out, cudaStream_t s) {
  const long total = long(a.n) * a.h * a.w * ((a.c + 7) / 8);
  launch_k(upadd_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, a, b, out);
}

void launch_upcat(const TV in[4], const int shift[4], int nin, const TV& out, cudaStream_t s) {
  UpCatArgs a;
  int off = 0;
  for (int i = 0; i < 4; ++i) {
    a.in[i] = i < nin ? in[i] : TV();
    a.shift[i] = i < nin ? shift[i] : 0;
    a.coff[i] = off;
    if (i < nin) off += in[i].c;
  }
  a.nin = nin;
  const long total = long(out.n) * out.h * out.w * ((out.c + 7) / 8);
  launch_k(upcat_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, a, out);
}

void launch_add(const TV& a, const TV& b, const TV& out, cudaStream_t s) {
  const long total = long(a.n) * a.h * a.w * ((a.c + 7) / 8);
  launch_k(add_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, a, b, out);
}

void launch_pool(const TV& in, const TV& out, int kh, int kw, int sh, int sw, bool is_max, cudaStream_t s,
                 const int* vw) {
  const long total = long(out.n) * out.h * out.w * ((out.c + 7) / 8);
  launch_k(pool_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, out, kh, kw, sh, sw, is_max ? 1 : 0, vw);
}

void launch_layernorm(const TV& in, const TV& out, const float* gb, float eps, cudaStream_t s, const int* vw) {
  const long rows = long(in.n) * in.h * in.w;
  const long groups = (rows + kLnTok - 1) / kLnTok;
  if (in.c <= 128) launch_k(layernorm_kernel<16>, dim3(int((groups * 16 + kThreads - 1) / kThreads)), dim3(kThreads), 0, s, in, out, gb, eps, vw);
  else launch_k(layernorm_kernel<32>, dim3(int((groups * 32 + kThreads - 1) / kThreads)), dim3(kThreads), 0, s, in, out, gb, eps, vw);
}

void launch_attention(const TV& qkv, const TV& out, int heads, int hd, float scale, cudaStream_t s, const int* vw) {
  if (launch_attention_mma(qkv, out, heads, hd, scale, s, vw)) return;
  const int T = qkv.h * qkv.w;
  const size_t smem = (size_t(2) * T * hd + size_t(4) * T) * sizeof(float);
  if (smem > 227 * 1024)
    throw std::runtime_error("attention: a text line of " + std::to_string(T) +
                             " tokens does not fit in shared memory (limit ~1700 tokens = a crop ~13600 px wide "
                             "at rec height 48)");
  static size_t configured[64] = {0};  // cudaFuncSetAttribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && (dev >= 64 || smem > configured[dev])) {
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      throw std::runtime_error("attention: cannot opt in to " + std::to_string(smem) + " bytes of shared memory");
    if (dev < 64) configured[dev] = smem;
  }
  launch_k(attention_kernel, dim3(qkv.n * heads), dim3(128), smem, s, qkv, out, heads, hd, scale, vw);
  if (cudaPeekAtLastError() != cudaSuccess)
    throw std::runtime_error(std::string("attention launch: ") + cudaGetErrorString(cudaGetLastError()));
}

void launch_dbhead(const TV& in, const float* blk, int cmid, float* prob, uint8_t* bitmap, int thresh_u8,
                   cudaStream_t s) {
  const long npix = long(in.n) * in.h * in.w;
  if (in.c != 24 || cmid != 24) throw std::runtime_error("DB head: only the 24 -> 24 -> 1 head of the shipped det graph is implemented");
  launch_k(dbhead_kernel<24, 24>, dim3(int((npix + 127) / 128)), dim3(128), 0, s, in, blk, prob, bitmap, thresh_u8);
}

void launch_fc_softmax(const float* partial, int splits, int hw, int n, int cin, int cout, const float* blk,
                       float* out, cudaStream_t s) {
  if (cout > 8) throw std::runtime_error("cls head: at most 8 classes");
  launch_k(fc_softmax_kernel, dim3(n), dim3(32), 0, s, partial, splits, 1.f / float(hw), cin, cout, blk, out);
}

void launch_ctc_head_simt(const TV& feat, const __half* w, const float* bias, int cin_pad, int ncls,
                          int ncls_pad, int* idx, float* prob, cudaStream_t s, const int* vw) {
  (void)ncls;
  const long rows = long(feat.n) * feat.h * feat.w;
  launch_k(ctc_head_simt_kernel, dim3(int((rows + 7) / 8)), dim3(kThreads), size_t(8) * cin_pad * sizeof(float), s, 
      feat, w, bias, cin_pad, ncls_pad, idx, prob, vw);
}

void launch_nchw3_to_input(const float* in, int n, int h, int w, __half* out, cudaStream_t s) {
  launch_k(nchw3_to_input_kernel, dim3(grid_for(long(n) * h * w)), dim3(kThreads), 0, s, in, n, h, w, out);
}

void launch_nhwc_to_nchw_f32(const TV& in, float* out, cudaStream_t s) {
  const long total = long(in.n) * in.c * in.h * in.w;
  launch_k(nhwc_to_nchw_f32_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, out);
}

}  // namespace b200ocr
