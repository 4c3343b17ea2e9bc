// cv::resize(INTER_LINEAR) on 8-bit BGR pixels + Normalize, per output pixel, shared by the stand-alone pre-processing
// kernels (preproc.cu) and the stem convolution that pre-processes on the fly (kernels_simt.cu: fused_stem_kernel).
//
// cv::resize(INTER_LINEAR) on 8-bit data is fixed point: 11-bit horizontal coefficients, the vertical pass computes
// (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, an exact 2x2 shrink goes through the INTER_AREA box average,
// equal sizes copy (reference src/preprocess_op.cpp:57-118 calls cv::resize with INTER_LINEAR).  Bit-identical to
// cv2.resize (tests/test_preproc_gpu.py).  Every floating-point step is an explicit round-to-nearest intrinsic, so the
// result does not depend on the translation unit's -fmad setting.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace b200ocr {
namespace resize {

struct Coef {
  int s;       // source index of the first tap
  int c0, c1;  // 11-bit weights
};

// x axis: indices outside the image collapse onto the edge pixel with weight 1 (cv::resize xofs/alpha set-up)
__device__ __forceinline__ Coef coef_x(int d, double scale, int sn) {
  float f = float(__dsub_rn(__dmul_rn(__dadd_rn(double(d), 0.5), scale), 0.5));
  int s = int(floorf(f));
  f = __fsub_rn(f, float(s));
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  Coef c;
  c.s = s;
  c.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}
// y axis: weights are kept, row indices are clamped when read
__device__ __forceinline__ Coef coef_y(int d, double scale) {
  float f = float(__dsub_rn(__dmul_rn(__dadd_rn(double(d), 0.5), scale), 0.5));
  int s = int(floorf(f));
  f = __fsub_rn(f, float(s));
  Coef c;
  c.s = s;
  c.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}

struct Src {
  const uint8_t* p;  // top-left pixel of the (cropped) source
  int w, h;
  long stride;  // bytes per row
};

// The general (bilinear) case with the two axis coefficients already computed.
__device__ __forceinline__ void resize_px_coef(const Src& s, const Coef& cx, const Coef& cy, int out[3]) {
  const int x1 = min(cx.s + 1, s.w - 1);
  const int y0 = min(max(cy.s, 0), s.h - 1), y1 = min(max(cy.s + 1, 0), s.h - 1);
  const uint8_t* r0 = s.p + y0 * s.stride;
  const uint8_t* r1 = s.p + y1 * s.stride;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int S0 = int(r0[cx.s * 3 + c]) * cx.c0 + int(r0[x1 * 3 + c]) * cx.c1;
    const int S1 = int(r1[cx.s * 3 + c]) * cx.c0 + int(r1[x1 * 3 + c]) * cx.c1;
    int v = (((cy.c0 * (S0 >> 4)) >> 16) + ((cy.c1 * (S1 >> 4)) >> 16) + 2) >> 2;
    out[c] = min(max(v, 0), 255);
  }
}

// One output pixel (3 channels) of cv::resize(src -> dw x dh, INTER_LINEAR) on CV_8UC3.
__device__ __forceinline__ void resize_px(const Src& s, int dw, int dh, int dx, int dy, int out[3]) {
  if (s.w == dw && s.h == dh) {
    const uint8_t* q = s.p + dy * s.stride + dx * 3;
    out[0] = q[0]; out[1] = q[1]; out[2] = q[2];
    return;
  }
  if (s.w == 2 * dw && s.h == 2 * dh) {  // INTER_LINEAR with an exact 2x2 shrink runs as INTER_AREA
    const uint8_t* q0 = s.p + (2 * dy) * s.stride + (2 * dx) * 3;
    const uint8_t* q1 = q0 + s.stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = (int(q0[c]) + int(q0[c + 3]) + int(q1[c]) + int(q1[c + 3]) + 2) >> 2;
    return;
  }
  const double scale_x = 1.0 / (double(dw) / double(s.w));
  const double scale_y = 1.0 / (double(dh) / double(s.h));
  const Coef cx = coef_x(dx, scale_x, s.w);
  const Coef cy = coef_y(dy, scale_y);
  resize_px_coef(s, cx, cy, out);
}

// Normalize::Run: f = u8 * (1/255.f); f * scale + shift, two fp32 roundings, no contraction.
__device__ __forceinline__ float norm1(int v, float scale, float shift) {
  return __fadd_rn(__fmul_rn(__fmul_rn(float(v), 0.0039215688593685627f /* (float)(1/255.) */), scale), shift);
}

}  // namespace resize
}  // namespace b200ocr
