// Launchers for the sm_100a kernels of the det/cls/rec forward passes.
// Everything here runs on the GPU; there is no host fallback for any of it.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace b200ocr {

// NHWC fp16 view: element (n,y,x,ch) lives at p[((n*h + y)*w + x)*pitch + ch].
struct TV {
  __half* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0, pitch = 0;
};

// y = s2 * act(acc + bias[c]) + t2 (+ residual)
struct Epi {
  int act = 0;
  float a = 0.f, b = 0.f, s2 = 1.f, t2 = 0.f;
  const __half* res = nullptr;
  int res_pitch = 0;
};

// Ragged batches: row n of a [N,h,w] tensor may only be `vw[n]` columns wide (the recognizer's text lines have
// different padded widths).  A kernel that is given `vw` (device int[N], valid width of its OUTPUT tensor) writes
// zeros at x >= vw[n], so that the next convolution sees exactly the zero padding it would see at the edge of a
// tensor that is vw[n] wide.  nullptr = dense batch.
struct ConvGeom {
  int kh = 1, kw = 1, sh = 1, sw = 1, ph = 0, pw = 0;
  int cin_pad = 0;  // weight row stride per tap
  int cout_pad = 0;
};

// ---- convolution family -----------------------------------------------------
// CUDA-core direct convolution (stems with C_in=3 and shapes the tensor-core path does not take).
void launch_conv_simt(const TV& in, const TV& out, const __half* w, const float* bias,
                      const ConvGeom& g, const Epi& e, cudaStream_t s, const int* vw = nullptr);
// tcgen05 implicit-GEMM convolution (stride 1; 1x1, 1x3, 3x3): TMA -> smem -> UMMA -> TMEM -> epilogue.
// Returns false if the shape is not eligible (caller then uses the CUDA-core kernel).
// Tensor maps are encoded once per (layer, shape) by make_conv_tc_plan.
struct ConvTcPlanImpl;
struct ConvTcPlan { ConvTcPlanImpl* impl = nullptr; };
bool conv_tc_eligible(const TV& in, const TV& out, const ConvGeom& g);
ConvTcPlan make_conv_tc_plan(const TV& in, const TV& out, const __half* w, const ConvGeom& g);
void free_conv_tc_plan(ConvTcPlan* p);
void launch_conv_tc(const ConvTcPlan& p, const float* bias, const Epi& e, cudaStream_t s, const int* vw = nullptr);
// Narrow 1x1 convolutions (<= 64 input channels) on mma.sync, streaming over the pixels (pwconv.cu); false = shape not
// covered, nothing launched.  w: fp16 [cout_pad][cin_pad] as plan.cpp packs it.
bool launch_pwconv_mma(const TV& in, const TV& out, const __half* w, const float* bias, const ConvGeom& g, const Epi& e,
                       cudaStream_t s, const int* vw = nullptr);
// w_bias: fp32 [taps][cp] weights + [cp] bias; w_half: the same weights as fp16 [taps][cp] (FHFMA kernel), or
// nullptr to multiply with the fp32 weights (conversion + FFMA: slower, one rounding less)
void launch_dwconv(const TV& in, const TV& out, const float* w_bias, const __half* w_half, const ConvGeom& g,
                   const Epi& e, cudaStream_t s, const int* vw = nullptr);
// shared-memory tiled variant (dwconv.cu): 3x3 / 5x5, strides 1|2; false = shape not covered, nothing launched
bool launch_dwconv_tile(const TV& in, const TV& out, const float* w_bias, const __half* w_half, const ConvGeom& g,
                        const Epi& e, cudaStream_t s, const int* vw);

// ---- SE block ---------------------------------------------------------------
int gap_splits(int n, int h, int w, int c, bool ragged_safe);
// SE gate fused into the pooling kernel (the block that finishes a sample's last split computes its gate):
// counters = int[n] device scratch, zero before the first use (the kernel leaves it zero)
struct SeFuse {
  const float* blk = nullptr;  // w1t[c][cmid], b1[cmid], w2t[cmid][c], b2[c]
  float* gate = nullptr;
  int* counters = nullptr;
  int c = 0, cmid = 0, h = 0;
  float slope = 0.f, offset = 0.f, inv_hw = 0.f;
  const int* vw_in = nullptr;
  // also apply the gate (the SE block's elementwise_mul [+ add]): the block that computes a sample's gate multiplies the
  // sample's map with it right away -- nullptr = the Scale layer runs as its own kernel
  __half* sc_out = nullptr;
  int sc_out_pitch = 0, sc_add_x = 0;
};
void launch_gap_partial(const TV& in, float* partial, int splits, bool ragged_safe, cudaStream_t s,
                        const SeFuse* fuse = nullptr);
// vw_in: valid width of the pooled tensor per row (mean over h * vw_in[n] pixels), h its height
void launch_se_fc(const float* partial, int splits, int hw, int n, int c, int cmid, const float* blk,
                  float slope, float offset, float* gate, cudaStream_t s, const int* vw_in = nullptr, int h = 0);
void launch_scale(const TV& in, const float* gate, bool add_x, const TV& out, cudaStream_t s);

// ---- glue ---------------------------------------------------------------------
void launch_upadd(const TV& a, const TV& b_half_res, const TV& out, cudaStream_t s);
void launch_upcat(const TV in[4], const int shift[4], int nin, const TV& out, cudaStream_t s);
void launch_add(const TV& a, const TV& b, const TV& out, cudaStream_t s);
void launch_pool(const TV& in, const TV& out, int kh, int kw, int sh, int sw, bool is_max, cudaStream_t s,
                 const int* vw = nullptr);

// ---- SVTR neck ----------------------------------------------------------------
void launch_layernorm(const TV& in, const TV& out, const float* gamma_beta, float eps, cudaStream_t s,
                      const int* vw = nullptr);
// vw: number of valid tokens per sequence (keys beyond it are ignored, queries beyond it produce zeros)
void launch_attention(const TV& qkv, const TV& out, int heads, int head_dim, float scale, cudaStream_t s,
                      const int* vw = nullptr);

// tensor-core variant (attention.cu); false = shape not covered, nothing launched
bool launch_attention_mma(const TV& qkv, const TV& out, int heads, int head_dim, float scale, cudaStream_t s,
                          const int* vw);

// ---- heads ----------------------------------------------------------------------
// DB head tail: deconv2x2+BN+relu -> deconv2x2 -> sigmoid, plus cbuf=(u8)(p*255) > thresh bitmap.
void launch_dbhead(const TV& in, const float* blk, int cmid, float* prob, uint8_t* bitmap,
                   int thresh_u8, cudaStream_t s);
void launch_fc_softmax(const float* partial, int splits, int hw, int n, int cin, int cout,
                       const float* blk, float* out, cudaStream_t s);
// CTC head: logits = feat . W^T + b; per (n,t): argmax index and softmax probability of the max.
void launch_ctc_head_simt(const TV& feat, const __half* w, const float* bias, int cin_pad, int ncls,
                          int ncls_pad, int* idx, float* prob, cudaStream_t s, const int* vw = nullptr);
bool ctc_tc_eligible(const TV& feat, int cin_pad);
void launch_ctc_head_tc(const TV& feat, const __half* w, const float* bias, int cin_pad, int ncls,
                        int ncls_pad, int* idx, float* prob, cudaStream_t s, const int* vw = nullptr);

// ---- pre-processing (preproc.cu) ---------------------------------------------------
struct NormParams { float scale[3], shift[3]; };  // y = (u8 * (1/255.f)) * scale + shift, per BGR channel
NormParams make_norm(const float mean[3], const float scale[3]);
struct DetPreItem { const uint8_t* src; int w, h; long stride; };          // one source image (device memory)
// ROI of a device image; columns [resize_w, pad_w) hold the pad value, columns >= pad_w (ragged batches) zero
struct CropItem { const uint8_t* img; long stride; int x, y, w, h, resize_w, pad_w; };
// [n] images -> [n, dh, dw] network input; every image is resized to the same dh x dw
void launch_det_preprocess(const DetPreItem* items_dev, int n, int dh, int dw, const NormParams& np, __half* out,
                           cudaStream_t s);
// [n] ROIs -> [n, dh, dw]; columns >= resize_w hold pad_value (rec: -1 = normalised u8 zero; cls: 0)
void launch_crop_preprocess(const CropItem* items_dev, int n, int dh, int dw, const NormParams& np, float pad_value,
                            __half* out, cudaStream_t s);
// The 8-bit sources a network's stem convolution (3 -> 8/16, 3x3, stride 2) can pre-process on the fly instead of
// reading the network input tensor (kernels_simt.cu: fused_stem_kernel); `items` is device memory.
struct StemSource {
  int kind = 0;               // 0: none; 1: DetPreItem[n], every image resized to the input's h x w; 2: CropItem[n]
  const void* items = nullptr;
  NormParams np;
  float pad_value = 0.f;      // CropItem columns [resize_w, pad_w)
};
bool fused_stem_eligible(const TV& in, const TV& out, const ConvGeom& g, const Epi& e);
void launch_fused_stem(const StemSource& src, const TV& in, const TV& out, const __half* w, const float* bias,
                       const ConvGeom& g, const Epi& e, cudaStream_t s, const int* vw);
void launch_rotate180_if(uint8_t* img, long stride, int x0, int y0, int w, int h, const int* label_dev, cudaStream_t s);
void launch_resize_u8(const uint8_t* src, int sw, int sh, long stride, int dw, int dh, uint8_t* out, cudaStream_t s);

// ---- perspective crop (warp.cu): Utility::GetRotateCropImage, reference src/utility.cpp:137-190
// box = 4 points (x,y) tl,tr,br,bl in image pixels.  out: rotate_crop_dims() rows x cols x 3 (device memory).
void rotate_crop_dims(const int box[8], int* out_rows, int* out_cols, int* crop_w, int* crop_h);
void launch_rotate_crop(const uint8_t* img, int rows, int cols, long stride, const int box[8], uint8_t* out, cudaStream_t s);

// ---- DB post-process (dbpost.cu) ------------------------------------------------------
struct DbPostParams {
  int n, h, w;            // batch of bitmaps / probability maps
  float box_thresh, unclip_ratio;
  int max_candidates;     // 1000 in the reference
  int score_slow = 0;     // det_db_score_mode "slow": PolygonScoreAcc (mean over the filled contour polygon)
};
struct DbImageInfo { float ratio_h, ratio_w; int src_h, src_w; };
struct DbBox { int valid; int pts[8]; float score; int start; };  // start = start pixel (y*w+x) of the contour
size_t dbpost_workspace_bytes(const DbPostParams& p);
// prob [n,h,w] fp32, bitmap [n,h,w] u8 (0/255), info [n] (device).  Writes counts[n] and boxes[n][max_candidates]
// (device), contour order = the reference's (cv::findContours RETR_LIST order).
void launch_dbpost(const DbPostParams& p, const float* prob, const uint8_t* bitmap, const DbImageInfo* info_dev,
                   void* workspace, int* counts_dev, DbBox* boxes_dev, cudaStream_t s);
// cbuf = (uchar)(p * 255); bit = cbuf > thresh_u8 ? 255 : 0   (reference src/ocr_det.cpp:143-154)
void launch_threshold(const float* prob, long n, int thresh_u8, uint8_t* bitmap, cudaStream_t s);
void launch_dilate2x2(const uint8_t* in, uint8_t* out, int n, int h, int w, cudaStream_t s);
// debug: contour start pixels + bounding boxes of the last launch_dbpost are left in the workspace; see dbpost.cu

// ---- CTC collapse (ctc.cu) -------------------------------------------------------------
// idx/prob [n, T] from the CTC head -> out_idx [n, T] (collapsed label ids), out_len [n], out_score [n]
void launch_ctc_collapse(const int* idx, const float* prob, int n, int T, int* out_idx, int* out_len,
                         float* out_score, cudaStream_t s);

// debug / test helpers
void launch_nhwc_to_nchw_f32(const TV& in, float* out, cudaStream_t s);
// fp32 NCHW (3 channels) -> network input NHWC fp16, channel pitch 8, channels 3..7 zero
void launch_nchw3_to_input(const float* in, int n, int h, int w, __half* out, cudaStream_t s);

}  // namespace b200ocr
