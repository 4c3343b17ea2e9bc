// TEST-ONLY host build of csrc/geom.h (the per-box geometry the dbpost kernels run on the device), so
// that the CPU suite can compare that arithmetic with OpenCV / the reference Clipper without a GPU.
// Not linked into libb200ocr.so; the product never runs this code on the host.
#include <vector>
#include <algorithm>
#include "../../cpp-paddle-ocr_b200/csrc/geom.h"

using namespace b200ocr::geom;

extern "C" {

int geomtest_hull(const float* xy, int n, float* out_xy, int cap) {
  std::vector<P2> p(n);
  for (int i = 0; i < n; ++i) p[i] = P2{xy[2 * i], xy[2 * i + 1]};
  std::stable_sort(p.begin(), p.end(), [](const P2& a, const P2& b) { return a.y < b.y || (a.y == b.y && a.x < b.x); });
  p.erase(std::unique(p.begin(), p.end(), [](const P2& a, const P2& b) { return a.x == b.x && a.y == b.y; }), p.end());
  std::vector<P2> h(cap);
  int k = convex_hull_sorted_yx([&](int i) { return p[i]; }, int(p.size()), h.data(), cap);
  for (int i = 0; i < k; ++i) { out_xy[2 * i] = h[i].x; out_xy[2 * i + 1] = h[i].y; }
  return k;
}

// xy: the contour as cv::findContours returns it (xy[0..1] = its start point); outer: 1 = outer border, 0 = hole border
void geomtest_min_area_rect(const float* xy, int n, int outer, float out[5]) {
  std::vector<float> hxy(2 * (n + 1));
  int k = geomtest_hull(xy, n, hxy.data(), n + 1);
  std::vector<P2> h(k);
  for (int i = 0; i < k; ++i) h[i] = P2{hxy[2 * i], hxy[2 * i + 1]};
  hull_order_like_cv(h.data(), k, outer != 0, P2{xy[0], xy[1]});
  std::vector<float> vx(k + 1), vy(k + 1), il(k + 1);
  RotRect r = min_area_rect_hull(h.data(), k, vx.data(), vy.data(), il.data());
  out[0] = r.cx; out[1] = r.cy; out[2] = r.w; out[3] = r.h; out[4] = r.angle;
}

float geomtest_mini_box(const float rr[5], float out[8]) {
  RotRect r; r.cx = rr[0]; r.cy = rr[1]; r.w = rr[2]; r.h = rr[3]; r.angle = rr[4];
  P2 b[4];
  float ssid = mini_box(r, b);
  for (int i = 0; i < 4; ++i) { out[2 * i] = b[i].x; out[2 * i + 1] = b[i].y; }
  return ssid;
}

void geomtest_quad_mask(const int* x, const int* y, int w, int h, unsigned char* mask) {
  QuadMask q;
  q.init(x, y, w, h);
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) mask[r * w + c] = q.inside(c, r) ? 1 : 0;
}

float geomtest_unclip_distance(const float box[8], float ratio) {
  P2 b[4];
  for (int i = 0; i < 4; ++i) b[i] = P2{box[2 * i], box[2 * i + 1]};
  return unclip_distance(b, ratio);
}

int geomtest_offset_round(const long long* qx, const long long* qy, double delta, float* ox, float* oy, int cap) {
  return offset_round(qx, qy, delta, ox, oy, cap);
}

int geomtest_finish_box(const float clip[8], int width, int height, float ratio_w, float ratio_h, int src_w, int src_h,
                        int out[8]) {
  P2 c[4];
  for (int i = 0; i < 4; ++i) c[i] = P2{clip[2 * i], clip[2 * i + 1]};
  return finish_box(c, width, height, ratio_w, ratio_h, src_w, src_h, out) ? 1 : 0;
}
}
