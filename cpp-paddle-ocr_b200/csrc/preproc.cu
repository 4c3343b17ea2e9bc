// Pre-processing kernels: one fused pass per stage from 8-bit BGR pixels to the network input
// (NHWC fp16, channel pitch 8, channels 3..7 zero).
//
//   det_preprocess   = ResizeImgType0 (cv::resize, reference src/preprocess_op.cpp:57-93) + Normalize (:40-55)
//                      + Permute (:19-26)
//   crop_preprocess  = ROI crop (src/ocr_worker.cpp:244-259) + CrnnResizeImg (:95-118) / ClsResizeImg (:120-137)
//                      + right padding + Normalize + PermuteBatch (:28-38), batched, variable width
//   rotate180        = cv::rotate(ROTATE_180) in place on an ROI (src/ocr_worker.cpp:277-281)
//
// cv::resize(INTER_LINEAR) on 8-bit data is fixed point: 11-bit horizontal coefficients, the vertical
// pass computes (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, an exact 2x2 shrink goes through the
// INTER_AREA box average, equal sizes copy.  resize_px() restates that per output pixel and is bit-identical
// to cv2.resize (tests/test_preproc_gpu.py).  HBM-bound: every source byte is read once (neighbouring
// threads share rows through L1), every output pixel is one 16-byte store.
#include "kernels.h"
#include "pdl.h"
#include "resize.cuh"

namespace b200ocr {

namespace {

constexpr int kThreads = 256;

using namespace resize;

__device__ __forceinline__ void store_px(__half* out, long pix, float a, float b, float c) {
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
  h[0] = __floats2half2_rn(a, b);
  h[1] = __floats2half2_rn(c, 0.f);
  h[2] = __floats2half2_rn(0.f, 0.f);
  h[3] = h[2];
  *reinterpret_cast<uint4*>(out + pix * 8) = o;
}

__global__ void __launch_bounds__(kThreads)
det_preprocess_kernel(const DetPreItem* __restrict__ items, int n, int dh, int dw, NormParams np,
                      __half* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long per = long(dh) * dw;
  const long total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int img = int(t / per);
    const int r = int(t - long(img) * per);
    const int dy = r / dw, dx = r - dy * dw;
    const DetPreItem it = items[img];
    Src s{it.src, it.w, it.h, it.stride};
    int v[3];
    resize_px(s, dw, dh, dx, dy, v);
    store_px(out, t, norm1(v[0], np.scale[0], np.shift[0]), norm1(v[1], np.scale[1], np.shift[1]),
             norm1(v[2], np.scale[2], np.shift[2]));
  }
}

__global__ void __launch_bounds__(kThreads)
crop_preprocess_kernel(const CropItem* __restrict__ items, int n, int dh, int dw, NormParams np, float pad_value,
                       __half* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long per = long(dh) * dw;
  const long total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int b = int(t / per);
    const int r = int(t - long(b) * per);
    const int dy = r / dw, dx = r - dy * dw;
    const CropItem it = items[b];
    if (dx >= it.resize_w) {
      const float v = dx < it.pad_w ? pad_value : 0.f;  // beyond this row's own batch width: ragged filler
      store_px(out, t, v, v, v);
      continue;
    }
    Src s{it.img + long(it.y) * it.stride + long(it.x) * 3, it.w, it.h, it.stride};
    int v[3];
    resize_px(s, it.resize_w, dh, dx, dy, v);
    store_px(out, t, norm1(v[0], np.scale[0], np.shift[0]), norm1(v[1], np.scale[1], np.shift[1]),
             norm1(v[2], np.scale[2], np.shift[2]));
  }
}

// In-place 180 degree rotation of one ROI, executed only when *label == 1.
__global__ void __launch_bounds__(kThreads)
rotate180_kernel(uint8_t* __restrict__ img, long stride, int x0, int y0, int w, int h, const int* __restrict__ label) {
  pdl_trigger();
  pdl_wait();
  if (*label != 1) return;
  const long total = long(w) * h;
  const long half = total / 2;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < half; t += long(gridDim.x) * blockDim.x) {
    const long u = total - 1 - t;
    const int ay = int(t / w), ax = int(t - long(ay) * w);
    const int by = int(u / w), bx = int(u - long(by) * w);
    uint8_t* a = img + long(y0 + ay) * stride + long(x0 + ax) * 3;
    uint8_t* b = img + long(y0 + by) * stride + long(x0 + bx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const uint8_t v = a[c]; a[c] = b[c]; b[c] = v; }
  }
}

// Plain resized 8-bit image (test hook for the fixed-point resize; also used by the warp path tests).
__global__ void __launch_bounds__(kThreads)
resize_u8_kernel(const uint8_t* __restrict__ src, int sw, int sh, long stride, int dw, int dh, uint8_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long total = long(dw) * dh;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int dy = int(t / dw), dx = int(t - long(dy) * dw);
    Src s{src, sw, sh, stride};
    int v[3];
    resize_px(s, dw, dh, dx, dy, v);
    out[t * 3] = uint8_t(v[0]); out[t * 3 + 1] = uint8_t(v[1]); out[t * 3 + 2] = uint8_t(v[2]);
  }
}

inline int grid_for(long total) {
  long b = (total + kThreads - 1) / kThreads;
  const long cap = 148L * 8;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

NormParams make_norm(const float mean[3], const float scale[3]) {
  NormParams p;
  for (int i = 0; i < 3; ++i) {
    p.scale[i] = scale[i];                                       // (float)(1.0 * scale[i])
    p.shift[i] = float((0.0 - double(mean[i])) * double(scale[i]));  // (float)((0.0 - mean[i]) * scale[i])
  }
  return p;
}

void launch_det_preprocess(const DetPreItem* items_dev, int n, int dh, int dw, const NormParams& np, __half* out,
                           cudaStream_t s) {
  launch_k(det_preprocess_kernel, dim3(grid_for(long(n) * dh * dw)), dim3(kThreads), 0, s, items_dev, n, dh, dw, np, out);
}

void launch_crop_preprocess(const CropItem* items_dev, int n, int dh, int dw, const NormParams& np, float pad_value,
                            __half* out, cudaStream_t s) {
  launch_k(crop_preprocess_kernel, dim3(grid_for(long(n) * dh * dw)), dim3(kThreads), 0, s, items_dev, n, dh, dw, np, pad_value, out);
}

void launch_rotate180_if(uint8_t* img, long stride, int x0, int y0, int w, int h, const int* label_dev,
                         cudaStream_t s) {
  launch_k(rotate180_kernel, dim3(grid_for(long(w) * h / 2 + 1)), dim3(kThreads), 0, s, img, stride, x0, y0, w, h, label_dev);
}

void launch_resize_u8(const uint8_t* src, int sw, int sh, long stride, int dw, int dh, uint8_t* out, cudaStream_t s) {
  launch_k(resize_u8_kernel, dim3(grid_for(long(dw) * dh)), dim3(kThreads), 0, s, src, sw, sh, stride, dw, dh, out);
}

}  // namespace b200ocr
