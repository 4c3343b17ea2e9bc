// Perspective crop of one text box: Utility::GetRotateCropImage (reference src/utility.cpp:137-190) =
// bounding-box crop + cv::getPerspectiveTransform + cv::warpPerspective + optional transpose/flip.
//
// The reference passes cv::BORDER_REPLICATE (== 1) in the *flags* argument of warpPerspective, so what actually
// runs is INTER_LINEAR with the default BORDER_CONSTANT (black) border; that call is what is reproduced.
// cv::warpPerspective on 8-bit data is fixed point: source coordinates are rounded to 1/32 pixel
// (X = round(32 * x'), sx = X >> 5, ax = X & 31) and the four taps are blended with integer weights
// (32-ay)(32-ax)*32 ... that sum to 2^15, then (sum + 2^14) >> 15.  One thread per output pixel.
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "kernels.h"

namespace b200ocr {

namespace {

struct WarpArgs {
  const uint8_t* src;  // top-left pixel of the cropped region
  int sw, sh;
  long stride;
  double m[9];         // inverse map: destination pixel -> source coordinates
  int dw, dh;          // size of the warped image (before the optional transpose + flip)
  int bw0;             // cv::WarpPerspectiveInvoker's block width for this output size
  int rot;             // 1: out = flip(transpose(warped), 0)
  uint8_t* out;        // rot ? [dw][dh][3] : [dh][dw][3]
};

__global__ void __launch_bounds__(256) warp_perspective_kernel(WarpArgs a) {
  const long total = long(a.dw) * a.dh;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int y = int(t / a.dw), x = int(t - long(y) * a.dw);
    // cv::WarpPerspectiveInvoker (imgproc/src/imgwarp.cpp) evaluates the map per block of bw0 columns: the row terms at
    // the block's first column, then one more product per pixel -- the same double operations in the same order here,
    // with explicit round-to-nearest multiplies / adds (no FMA contraction: the CPU code has none)
    const int xb = (x / a.bw0) * a.bw0, x1 = x - xb;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(a.m[0], double(xb)), __dmul_rn(a.m[1], double(y))), a.m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(a.m[3], double(xb)), __dmul_rn(a.m[4], double(y))), a.m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(a.m[6], double(xb)), __dmul_rn(a.m[7], double(y))), a.m[8]);
    double W = __dadd_rn(W0, __dmul_rn(a.m[6], double(x1)));
    W = W != 0. ? __ddiv_rn(32., W) : 0.;
    const double fX = fmax(-2147483648., fmin(2147483647., __dmul_rn(__dadd_rn(X0, __dmul_rn(a.m[0], double(x1))), W)));
    const double fY = fmax(-2147483648., fmin(2147483647., __dmul_rn(__dadd_rn(Y0, __dmul_rn(a.m[3], double(x1))), W)));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));  // saturate_cast<short>
    sy = max(-32768, min(32767, sy));
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
    int v[3] = {0, 0, 0};
    if (!(sx >= a.sw || sx + 1 < 0 || sy >= a.sh || sy + 1 < 0)) {
      const bool x0 = sx >= 0, x1 = sx + 1 < a.sw, y0 = sy >= 0, y1 = sy + 1 < a.sh;
      const uint8_t* p = a.src + long(sy) * a.stride + long(sx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int s00 = (x0 && y0) ? p[c] : 0;
        const int s01 = (x1 && y0) ? p[3 + c] : 0;
        const int s10 = (x0 && y1) ? p[a.stride + c] : 0;
        const int s11 = (x1 && y1) ? p[a.stride + 3 + c] : 0;
        v[c] = (s00 * w00 + s01 * w01 + s10 * w10 + s11 * w11 + (1 << 14)) >> 15;
      }
    }
    long o;
    if (a.rot) o = (long(a.dw - 1 - x) * a.dh + y) * 3;  // F[i][j] = warped[j][dw-1-i]
    else o = t * 3;
    a.out[o] = uint8_t(v[0]); a.out[o + 1] = uint8_t(v[1]); a.out[o + 2] = uint8_t(v[2]);
  }
}

// cv::getPerspectiveTransform: 8x8 system, LU with partial pivoting (cv::solve DECOMP_LU), double precision
void perspective_transform(const double src[8], const double dst[8], double m[9]) {
  double a[8][9];
  for (int i = 0; i < 4; ++i) {
    const double sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
    double r0[9] = {sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx, dx};
    double r1[9] = {0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy, dy};
    memcpy(a[i], r0, sizeof r0);
    memcpy(a[i + 4], r1, sizeof r1);
  }
  for (int i = 0; i < 8; ++i) {
    int k = i;
    for (int j = i + 1; j < 8; ++j)
      if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
    if (std::fabs(a[k][i]) < 2.220446049250313e-16 * 100) { for (int q = 0; q < 9; ++q) m[q] = 0; return; }
    if (k != i) for (int q = i; q < 9; ++q) std::swap(a[i][q], a[k][q]);
    const double d = -1 / a[i][i];
    for (int j = i + 1; j < 8; ++j) {
      const double alpha = a[j][i] * d;
      for (int q = i + 1; q < 9; ++q) a[j][q] += alpha * a[i][q];
    }
  }
  double x[8];
  for (int i = 7; i >= 0; --i) {
    double s = a[i][8];
    for (int q = i + 1; q < 8; ++q) s -= a[i][q] * x[q];
    x[i] = s / a[i][i];
  }
  for (int i = 0; i < 8; ++i) m[i] = x[i];
  m[8] = 1.;
}

// cv::invert of a 3x3 (closed form, double)
bool invert3(const double s[9], double d[9]) {
  const double det = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
  if (det == 0) return false;
  const double id = 1. / det;
  d[0] = (s[4] * s[8] - s[5] * s[7]) * id; d[1] = (s[2] * s[7] - s[1] * s[8]) * id; d[2] = (s[1] * s[5] - s[2] * s[4]) * id;
  d[3] = (s[5] * s[6] - s[3] * s[8]) * id; d[4] = (s[0] * s[8] - s[2] * s[6]) * id; d[5] = (s[2] * s[3] - s[0] * s[5]) * id;
  d[6] = (s[3] * s[7] - s[4] * s[6]) * id; d[7] = (s[1] * s[6] - s[0] * s[7]) * id; d[8] = (s[0] * s[4] - s[1] * s[3]) * id;
  return true;
}

}  // namespace

void rotate_crop_dims(const int box[8], int* out_rows, int* out_cols, int* crop_w, int* crop_h) {
  // img_crop_width = int(|p0 - p1|), img_crop_height = int(|p0 - p3|)   (utility.cpp:160-163)
  const int w = int(std::sqrt(std::pow(double(box[0] - box[2]), 2) + std::pow(double(box[1] - box[3]), 2)));
  const int h = int(std::sqrt(std::pow(double(box[0] - box[6]), 2) + std::pow(double(box[1] - box[7]), 2)));
  *crop_w = w; *crop_h = h;
  if (float(h) >= float(w) * 1.5f) { *out_rows = w; *out_cols = h; }  // transpose + flip
  else { *out_rows = h; *out_cols = w; }
}

void launch_rotate_crop(const uint8_t* img, int rows, int cols, long stride, const int box[8], uint8_t* out, cudaStream_t s) {
  int xs[4] = {box[0], box[2], box[4], box[6]}, ys[4] = {box[1], box[3], box[5], box[7]};
  int left = xs[0], right = xs[0], top = ys[0], bottom = ys[0];
  for (int i = 1; i < 4; ++i) {
    left = std::min(left, xs[i]); right = std::max(right, xs[i]);
    top = std::min(top, ys[i]); bottom = std::max(bottom, ys[i]);
  }
  if (left < 0 || top < 0 || right > cols || bottom > rows || right <= left || bottom <= top)
    throw std::invalid_argument("rotate_crop: the box's bounding rectangle must be a non-empty part of the image");
  int orows, ocols, cw, ch;
  rotate_crop_dims(box, &orows, &ocols, &cw, &ch);
  if (cw < 1 || ch < 1) throw std::invalid_argument("rotate_crop: degenerate box");
  double src[8], dst[8] = {0, 0, double(cw), 0, double(cw), double(ch), 0, double(ch)};
  for (int i = 0; i < 4; ++i) { src[2 * i] = float(xs[i] - left); src[2 * i + 1] = float(ys[i] - top); }
  double m[9], inv[9];
  perspective_transform(src, dst, m);
  WarpArgs a;
  if (!invert3(m, inv)) memset(inv, 0, sizeof inv);  // cv::invert leaves zeros for a singular matrix
  memcpy(a.m, inv, sizeof inv);
  a.src = img + long(top) * stride + long(left) * 3;
  a.sw = right - left; a.sh = bottom - top; a.stride = stride;
  a.dw = cw; a.dh = ch;
  {  // BLOCK_SZ = 32: bh0 = min(16, height); bw0 = min(1024 / bh0, width)
    const int bh0 = std::min(16, ch);
    a.bw0 = std::max(1, std::min(1024 / bh0, cw));
  }
  a.rot = float(ch) >= float(cw) * 1.5f;
  a.out = out;
  const long total = long(cw) * ch;
  int grid = int((total + 255) / 256);
  grid = grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid);
  warp_perspective_kernel<<<grid, 256, 0, s>>>(a);
}

}  // namespace b200ocr
