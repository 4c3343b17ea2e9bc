// Kernel-level test entry points (declared in include/b200ocr.h, "kernel-level entry points"): one hot kernel on
// host tensors, so that parity tests can sweep shapes the shipped graphs do not contain (odd widths, partial
// channel chunks, every stride, ragged rows).  Host fp32 NCHW in, host fp32 NCHW out; the device side is exactly
// what the networks run (NHWC fp16 activations, fp32 accumulation).
#include <cstring>
#include <vector>

#include "../../include/b200ocr.h"
#include "capi_util.h"
#include "engine.h"
#include "kernels.h"
#include "plan.h"

using namespace b200ocr;

namespace {
struct DevMem {
  void* p = nullptr;
  explicit DevMem(size_t bytes) { cuda_check(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc"); }
  ~DevMem() { cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};
// fp32 NCHW (host) -> fp16 NHWC with channel pitch `pitch` (host), pad channels zero
std::vector<uint16_t> to_nhwc_f16(const float* x, int n, int c, int h, int w, int pitch) {
  std::vector<uint16_t> o(size_t(n) * h * w * pitch, 0);
  for (int in = 0; in < n; ++in)
    for (int ic = 0; ic < c; ++ic)
      for (int y = 0; y < h; ++y)
        for (int xx = 0; xx < w; ++xx)
          o[((size_t(in) * h + y) * w + xx) * pitch + ic] = f32_to_f16_bits(x[((size_t(in) * c + ic) * h + y) * w + xx]);
  return o;
}
void from_nhwc_f16(const std::vector<uint16_t>& v, int n, int c, int h, int w, int pitch, float* out) {
  for (int in = 0; in < n; ++in)
    for (int ic = 0; ic < c; ++ic)
      for (int y = 0; y < h; ++y)
        for (int xx = 0; xx < w; ++xx)
          out[((size_t(in) * c + ic) * h + y) * w + xx] = f16_bits_to_f32(v[((size_t(in) * h + y) * w + xx) * pitch + ic]);
}
}  // namespace

extern "C" {

int b200ocr_kernel_dwconv(int device, const float* x, int n, int c, int h, int w, const float* filt, const float* bias,
                          int k, int sh, int sw, int act, float post_scale, float post_shift, int fp16_weights,
                          const int* out_widths, float* out, int* out_h, int* out_w) {
  return capi_guard([&] {
    if (!x || !filt || !bias || !out || n < 1 || c < 1 || h < 1 || w < 1) throw std::invalid_argument("bad argument");
    if (k != 3 && k != 5) throw std::invalid_argument("k must be 3 or 5");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    const int pad = k / 2, pitch = round_up(c, 8), cp = pitch, taps = k * k;
    const int oh = (h + 2 * pad - k) / sh + 1, ow = (w + 2 * pad - k) / sw + 1;
    if (out_h) *out_h = oh;
    if (out_w) *out_w = ow;
    std::vector<uint16_t> hx = to_nhwc_f16(x, n, c, h, w, pitch);
    std::vector<float> wb(size_t(taps + 1) * cp, 0.f);      // [taps][cp] + [cp] bias, as plan.cpp packs it
    std::vector<uint16_t> wh(size_t(taps) * cp, 0);
    for (int ic = 0; ic < c; ++ic) {
      for (int t = 0; t < taps; ++t) {
        wb[size_t(t) * cp + ic] = filt[size_t(ic) * taps + t];
        wh[size_t(t) * cp + ic] = f32_to_f16_bits(filt[size_t(ic) * taps + t]);
      }
      wb[size_t(taps) * cp + ic] = bias[ic];
    }
    DevMem dx(hx.size() * 2), dy(size_t(n) * oh * ow * pitch * 2), dwb(wb.size() * 4), dwh(wh.size() * 2), dvw(size_t(n) * 4);
    cuda_check(cudaMemcpy(dx.p, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(dwb.p, wb.data(), wb.size() * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(dwh.p, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice), "upload");
    if (out_widths) cuda_check(cudaMemcpy(dvw.p, out_widths, size_t(n) * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemset(dy.p, 0xff, size_t(n) * oh * ow * pitch * 2), "memset");  // NaN pattern: every element must be written
    TV in, o;
    in.p = dx.as<__half>(); in.n = n; in.h = h; in.w = w; in.c = c; in.pitch = pitch;
    o.p = dy.as<__half>(); o.n = n; o.h = oh; o.w = ow; o.c = c; o.pitch = pitch;
    ConvGeom g;
    g.kh = g.kw = k; g.sh = sh; g.sw = sw; g.ph = g.pw = pad; g.cin_pad = g.cout_pad = cp;
    Epi e;
    e.act = act; e.s2 = post_scale; e.t2 = post_shift;
    launch_dwconv(in, o, dwb.as<float>(), fp16_weights ? dwh.as<__half>() : nullptr, g, e, nullptr,
                  out_widths ? dvw.as<int>() : nullptr);
    cuda_check(cudaGetLastError(), "dwconv launch");
    cuda_check(cudaDeviceSynchronize(), "dwconv");
    std::vector<uint16_t> hy(size_t(n) * oh * ow * pitch);
    cuda_check(cudaMemcpy(hy.data(), dy.p, hy.size() * 2, cudaMemcpyDeviceToHost), "download");
    for (size_t i = 0; i < hy.size(); ++i)   // pad channels must come back as zeros (consumers read whole 8-channel groups)
      if (int(i % pitch) >= c && hy[i] != 0) throw std::runtime_error("dwconv left a non-zero pad channel");
    from_nhwc_f16(hy, n, c, oh, ow, pitch, out);
  });
}

int b200ocr_kernel_attention(int device, const float* qkv, int n, int t, int heads, int head_dim, float scale,
                             const int* valid, float* out) {
  return capi_guard([&] {
    if (!qkv || !out || n < 1 || t < 1 || heads < 1 || head_dim < 1) throw std::invalid_argument("bad argument");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    const int C = heads * head_dim, c3 = 3 * C, p3 = round_up(c3, 8), p1 = round_up(C, 8);
    // qkv host layout: [n][t][3*C] (token-major, what the fused qkv linear writes)
    std::vector<uint16_t> hq(size_t(n) * t * p3, 0);
    for (size_t r = 0; r < size_t(n) * t; ++r)
      for (int ch = 0; ch < c3; ++ch) hq[r * p3 + ch] = f32_to_f16_bits(qkv[r * c3 + ch]);
    DevMem dq(hq.size() * 2), dy(size_t(n) * t * p1 * 2), dvw(size_t(n) * 4);
    cuda_check(cudaMemcpy(dq.p, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice), "upload");
    if (valid) cuda_check(cudaMemcpy(dvw.p, valid, size_t(n) * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemset(dy.p, 0xff, size_t(n) * t * p1 * 2), "memset");
    TV in, o;
    in.p = dq.as<__half>(); in.n = n; in.h = 1; in.w = t; in.c = c3; in.pitch = p3;
    o.p = dy.as<__half>(); o.n = n; o.h = 1; o.w = t; o.c = C; o.pitch = p1;
    launch_attention(in, o, heads, head_dim, scale, nullptr, valid ? dvw.as<int>() : nullptr);
    cuda_check(cudaGetLastError(), "attention launch");
    cuda_check(cudaDeviceSynchronize(), "attention");
    std::vector<uint16_t> hy(size_t(n) * t * p1);
    cuda_check(cudaMemcpy(hy.data(), dy.p, hy.size() * 2, cudaMemcpyDeviceToHost), "download");
    for (size_t r = 0; r < size_t(n) * t; ++r)
      for (int ch = 0; ch < C; ++ch) out[r * C + ch] = f16_bits_to_f32(hy[r * p1 + ch]);
  });
}

int b200ocr_kernel_conv(int device, const float* x, int n, int cin, int h, int w, const float* filt, const float* bias,
                        int cout, int kh, int kw, int act, float post_scale, float post_shift, const float* residual,
                        const int* out_widths, int force_simt, float* out) {
  return capi_guard([&] {
    if (!x || !filt || !bias || !out || n < 1 || cin < 1 || cout < 1 || h < 1 || w < 1) throw std::invalid_argument("bad argument");
    if (kh < 1 || kw < 1 || !(kh & 1) || !(kw & 1)) throw std::invalid_argument("odd filter sizes only (same padding, stride 1)");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    const int ip = round_up(cin, 8), op = round_up(cout, 8), taps = kh * kw;
    const int cin_pad = round_up(cin, 64), cout_pad = round_up(cout, 16);
    std::vector<uint16_t> hx = to_nhwc_f16(x, n, cin, h, w, ip);
    // dense filter as plan.cpp packs it: fp16 [cout_pad][taps][cin_pad]; filt is [cout][cin][kh][kw]
    std::vector<uint16_t> hw(size_t(cout_pad) * taps * cin_pad, 0);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < taps; ++t)
          hw[(size_t(co) * taps + t) * cin_pad + ci] = f32_to_f16_bits(filt[(size_t(co) * cin + ci) * taps + t]);
    std::vector<float> hb(cout_pad, 0.f);
    for (int co = 0; co < cout; ++co) hb[co] = bias[co];
    std::vector<uint16_t> hr;
    if (residual) hr = to_nhwc_f16(residual, n, cout, h, w, op);
    DevMem dx(hx.size() * 2), dy(size_t(n) * h * w * op * 2), dw(hw.size() * 2), db(hb.size() * 4), dr(hr.size() * 2 + 16),
        dvw(size_t(n) * 4);
    cuda_check(cudaMemcpy(dx.p, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(dw.p, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(db.p, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice), "upload");
    if (residual) cuda_check(cudaMemcpy(dr.p, hr.data(), hr.size() * 2, cudaMemcpyHostToDevice), "upload");
    if (out_widths) cuda_check(cudaMemcpy(dvw.p, out_widths, size_t(n) * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemset(dy.p, 0xff, size_t(n) * h * w * op * 2), "memset");
    TV in, o;
    in.p = dx.as<__half>(); in.n = n; in.h = h; in.w = w; in.c = cin; in.pitch = ip;
    o.p = dy.as<__half>(); o.n = n; o.h = h; o.w = w; o.c = cout; o.pitch = op;
    ConvGeom g;
    g.kh = kh; g.kw = kw; g.sh = g.sw = 1; g.ph = kh / 2; g.pw = kw / 2; g.cin_pad = cin_pad; g.cout_pad = cout_pad;
    Epi e;
    e.act = act; e.s2 = post_scale; e.t2 = post_shift;
    if (residual) { e.res = dr.as<__half>(); e.res_pitch = op; }
    const int* vw = out_widths ? dvw.as<int>() : nullptr;
    // force_simt: 0 = the engine's choice (mma.sync stream for narrow 1x1, else tcgen05, else CUDA cores), 1 = CUDA cores,
    // 2 = skip the narrow-1x1 kernel (tcgen05 where eligible)
    if (!force_simt && launch_pwconv_mma(in, o, dw.as<__half>(), db.as<float>(), g, e, nullptr, vw)) {
      cuda_check(cudaDeviceSynchronize(), "pwconv_mma");
    } else if (force_simt != 1 && conv_tc_eligible(in, o, g)) {
      ConvTcPlan plan = make_conv_tc_plan(in, o, dw.as<__half>(), g);
      launch_conv_tc(plan, db.as<float>(), e, nullptr, vw);
      cuda_check(cudaDeviceSynchronize(), "conv_tc");
      free_conv_tc_plan(&plan);
    } else {
      launch_conv_simt(in, o, dw.as<__half>(), db.as<float>(), g, e, nullptr, vw);
      cuda_check(cudaDeviceSynchronize(), "conv_simt");
    }
    cuda_check(cudaGetLastError(), "conv launch");
    std::vector<uint16_t> hy(size_t(n) * h * w * op);
    cuda_check(cudaMemcpy(hy.data(), dy.p, hy.size() * 2, cudaMemcpyDeviceToHost), "download");
    for (size_t i = 0; i < hy.size(); ++i)
      if (int(i % op) >= cout && hy[i] != 0) throw std::runtime_error("conv left a non-zero pad channel");
    from_nhwc_f16(hy, n, cout, h, w, op, out);
  });
}

int b200ocr_kernel_ctc_head(int device, const float* feat, int n, int t, int cin, const float* w, const float* bias,
                            int ncls, int force_simt, int32_t* idx, float* prob, int32_t* collapsed, int32_t* lens,
                            float* scores) {
  return capi_guard([&] {
    if (!feat || !w || !bias || !idx || !prob || n < 1 || t < 1 || cin < 1 || ncls < 2) throw std::invalid_argument("bad argument");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    // exactly what plan.cpp builds for the rec head: fp16 [ncls_pad16][cin_pad64] K-major rows, padded classes biased
    // to -30000, and the same bias pre-multiplied by log2(e) for the tcgen05 kernel
    const int pitch = round_up(cin, 8), cin_pad = round_up(cin, 64), ncls_pad = round_up(ncls, 16);
    std::vector<uint16_t> hf(size_t(n) * t * pitch, 0);
    for (size_t r = 0; r < size_t(n) * t; ++r)
      for (int c = 0; c < cin; ++c) hf[r * pitch + c] = f32_to_f16_bits(feat[r * cin + c]);
    std::vector<uint16_t> hw(size_t(ncls_pad) * cin_pad, 0);
    for (int co = 0; co < ncls; ++co)
      for (int ci = 0; ci < cin; ++ci) hw[size_t(co) * cin_pad + ci] = f32_to_f16_bits(w[size_t(ci) * ncls + co]);
    std::vector<float> hb(ncls_pad, -30000.f), hb2(ncls_pad);
    for (int co = 0; co < ncls; ++co) hb[co] = bias[co];
    for (int co = 0; co < ncls_pad; ++co) hb2[co] = hb[co] * 1.4426950408889634f;
    const size_t rows = size_t(n) * t;
    DevMem df(hf.size() * 2), dw(hw.size() * 2), db(hb.size() * 4), db2(hb2.size() * 4), di(rows * 4), dp(rows * 4),
        dc(rows * 4), dl(size_t(n) * 4), ds(size_t(n) * 4);
    cuda_check(cudaMemcpy(df.p, hf.data(), hf.size() * 2, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(dw.p, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(db.p, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(db2.p, hb2.data(), hb2.size() * 4, cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemset(di.p, 0xff, rows * 4), "memset");
    cuda_check(cudaMemset(dp.p, 0xff, rows * 4), "memset");
    TV f;
    f.p = df.as<__half>(); f.n = n; f.h = 1; f.w = t; f.c = cin; f.pitch = pitch;
    if (!force_simt && ctc_tc_eligible(f, cin_pad))
      launch_ctc_head_tc(f, dw.as<__half>(), db2.as<float>(), cin_pad, ncls, ncls_pad, di.as<int>(), dp.as<float>(), nullptr);
    else if (force_simt)
      launch_ctc_head_simt(f, dw.as<__half>(), db.as<float>(), cin_pad, ncls, ncls_pad, di.as<int>(), dp.as<float>(), nullptr);
    else
      throw std::runtime_error("shape not eligible for the tcgen05 CTC head");
    cuda_check(cudaGetLastError(), "ctc head launch");
    launch_ctc_collapse(di.as<int>(), dp.as<float>(), n, t, dc.as<int>(), dl.as<int>(), ds.as<float>(), nullptr);
    cuda_check(cudaDeviceSynchronize(), "ctc head");
    cuda_check(cudaMemcpy(idx, di.p, rows * 4, cudaMemcpyDeviceToHost), "download");
    cuda_check(cudaMemcpy(prob, dp.p, rows * 4, cudaMemcpyDeviceToHost), "download");
    if (collapsed) cuda_check(cudaMemcpy(collapsed, dc.p, rows * 4, cudaMemcpyDeviceToHost), "download");
    if (lens) cuda_check(cudaMemcpy(lens, dl.p, size_t(n) * 4, cudaMemcpyDeviceToHost), "download");
    if (scores) cuda_check(cudaMemcpy(scores, ds.p, size_t(n) * 4, cudaMemcpyDeviceToHost), "download");
  });
}

}  // extern "C"
