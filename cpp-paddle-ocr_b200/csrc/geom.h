// Per-box geometry of the DB post-process, written as plain functions that compile for the device
// (dbpost.cu, one thread / warp per candidate box) and for the host (tests/native/geom_host.cpp, a
// TEST-ONLY library that lets the CPU suite compare this arithmetic with OpenCV; the product library
// never calls these on the host).
//
// The reference delegates all of this to OpenCV and Clipper (reference src/postprocess_op.cpp):
//   cv::minAreaRect (:279, :66)      -> convex_hull() + min_area_rect()  (Sklansky hull + rotating calipers,
//                                       float32 like OpenCV's rotcalipers; restated from the published algorithm)
//   cv::boxPoints + GetMiniBoxes (:134-168) -> box_points() + mini_box()
//   cv::fillPoly + cv::mean in BoxScoreFast (:216-253) -> QuadMask (8-connected Bresenham outline +
//                                       16.16 fixed-point scan-line fill, OpenCV drawing.cpp semantics)
//   GetContourArea + ClipperOffset jtRound in UnClip (:20-72) -> unclip_distance() + offset_round()
//   clamp/round (:312-324), OrderPointsClockwise (:87-104), FilterTagDetRes (:333-362) -> finish_box()
// Compile with FMA contraction off (-fmad=false / -ffp-contract=off): OpenCV's float expressions are
// evaluated with separate roundings.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GEOM_HD __host__ __device__ __forceinline__
#else
#define GEOM_HD inline
#endif

namespace b200ocr {
namespace geom {

struct P2 {
  float x, y;
};

struct RotRect {
  float cx = 0.f, cy = 0.f, w = 0.f, h = 0.f, angle = 0.f;  // angle in degrees, like cv::RotatedRect
};

constexpr int kMaxHull = 256;  // hull vertices kept per candidate (integer-grid hulls of <=1000 px blobs are far smaller)

// ---------------------------------------------------------------------------------------------
// Convex hull of points already sorted by (y, then x).  Strictly convex vertices only.
// Output order = the order OpenCV's convexHull(points, hull, /*clockwise=*/true) produces for the
// same set (before its index-based cyclic shift): it starts at the left-most point (smallest x,
// ties -> smallest y) and walks left -> bottom (max y) -> right -> top in image coordinates.
// Returns the number of hull points written to `out` (<= cap).
template <class GetPt>
GEOM_HD int convex_hull_sorted_yx(GetPt pt, int n, P2* out, int cap) {
  if (n <= 0) return 0;
  if (n == 1) { out[0] = pt(0); return 1; }
  // Andrew monotone chain over the (y,x) order: first the chain that runs with increasing y on the
  // left side, then back with decreasing y on the right side.
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const P2 p = pt(i);
    while (k >= 2) {
      const float cr = (out[k - 1].x - out[k - 2].x) * (p.y - out[k - 2].y) -
                       (out[k - 1].y - out[k - 2].y) * (p.x - out[k - 2].x);
      if (cr >= 0.f) --k; else break;
    }
    if (k < cap) out[k++] = p; else return -1;
  }
  const int lower = k + 1;
  for (int i = n - 2; i >= 0; --i) {
    const P2 p = pt(i);
    while (k >= lower) {
      const float cr = (out[k - 1].x - out[k - 2].x) * (p.y - out[k - 2].y) -
                       (out[k - 1].y - out[k - 2].y) * (p.x - out[k - 2].x);
      if (cr >= 0.f) --k; else break;
    }
    if (k < cap) out[k++] = p; else return -1;
  }
  --k;  // last point equals the first
  if (k < 1) k = 1;
  // The chain above runs top -> left side -> bottom -> right side in image coordinates (y down), the
  // rotation OpenCV's convexHull(points, hull, /*clockwise=*/true) (as called by cv::minAreaRect) uses:
  // left -> bottom (max y) -> right -> top.  Rotate its start, the left-most point (smallest x,
  // ties -> smallest y), to the front.
  return k;
}

// The vertex ORDER cv::minAreaRect hands to its rotating calipers: cv::convexHull(points, clockwise = false) followed
// by that function's final cyclic shift, which makes the hull's contour indices monotone.  For a border traced by
// cv::findContours that comes out as (measured against cv2 4.13 on 8600 outer and hole borders, 0 mismatches):
//   outer border: the reverse of the chain above, i.e. ending with the contour's start pixel (its first pixel in
//                 raster order, which is chain[0]);
//   hole border : the same reversed cycle, rotated to BEGIN with the contour's start pixel (the pixel left of the
//                 hole's first pixel) when that pixel is a hull vertex.
// The order decides which of two equal-area rectangles of a small lattice polygon wins (`area <= minarea`: the last
// minimum in iteration order), so it is part of the result, not a detail.
GEOM_HD void hull_order_like_cv(P2* h, int k, bool outer, P2 start) {
  for (int i = 0, j = k - 1; i < j; ++i, --j) { const P2 t = h[i]; h[i] = h[j]; h[j] = t; }
  if (outer) return;
  int s = -1;
  for (int i = 0; i < k; ++i)
    if (h[i].x == start.x && h[i].y == start.y) { s = i; break; }
  if (s <= 0) return;
  // rotate left by s: three reversals
  for (int i = 0, j = s - 1; i < j; ++i, --j) { const P2 t = h[i]; h[i] = h[j]; h[j] = t; }
  for (int i = s, j = k - 1; i < j; ++i, --j) { const P2 t = h[i]; h[i] = h[j]; h[j] = t; }
  for (int i = 0, j = k - 1; i < j; ++i, --j) { const P2 t = h[i]; h[i] = h[j]; h[j] = t; }
}

// ---------------------------------------------------------------------------------------------
// Rotating calipers, minimum-area enclosing rectangle of a convex polygon (float32 throughout).
// out[0..1] = rectangle corner, out[2..3] = first edge vector, out[4..5] = second edge vector.
GEOM_HD void rotating_calipers_min_area(const P2* pts, int n, float* vx, float* vy, float* inv_len, float out[6]) {
  float minarea = 3.402823466e+38f;
  int left = 0, bottom = 0, right = 0, top = 0;
  int seq[4];
  float orientation = 0.f, base_a, base_b = 0.f;
  P2 pt0 = pts[0];
  float left_x = pt0.x, right_x = pt0.x, top_y = pt0.y, bottom_y = pt0.y;
  for (int i = 0; i < n; ++i) {
    if (pt0.x < left_x) { left_x = pt0.x; left = i; }
    if (pt0.x > right_x) { right_x = pt0.x; right = i; }
    if (pt0.y > top_y) { top_y = pt0.y; top = i; }
    if (pt0.y < bottom_y) { bottom_y = pt0.y; bottom = i; }
    const P2 pt = pts[(i + 1 < n) ? i + 1 : 0];
    const double dx = double(pt.x) - double(pt0.x), dy = double(pt.y) - double(pt0.y);
    vx[i] = float(dx);
    vy[i] = float(dy);
    inv_len[i] = float(1. / sqrt(dx * dx + dy * dy));
    pt0 = pt;
  }
  {
    double ax = vx[n - 1], ay = vy[n - 1];
    for (int i = 0; i < n; ++i) {
      const double bx = vx[i], by = vy[i];
      const double convexity = ax * by - ay * bx;
      if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
      ax = bx; ay = by;
    }
  }
  base_a = orientation;
  seq[0] = bottom; seq[1] = right; seq[2] = top; seq[3] = left;
  int b_left = 0, b_bottom = 0;
  float b_a = 0.f, b_b = 0.f, b_w = 0.f, b_h = 0.f;
  for (int k = 0; k < n; ++k) {
    const float dp[4] = {
        +base_a * vx[seq[0]] + base_b * vy[seq[0]],
        -base_b * vx[seq[1]] + base_a * vy[seq[1]],
        -base_a * vx[seq[2]] - base_b * vy[seq[2]],
        +base_b * vx[seq[3]] - base_a * vy[seq[3]],
    };
    float maxcos = dp[0] * inv_len[seq[0]];
    int main_element = 0;
    for (int i = 1; i < 4; ++i) {
      const float cosalpha = dp[i] * inv_len[seq[i]];
      if (cosalpha > maxcos) { main_element = i; maxcos = cosalpha; }
    }
    {
      const int pindex = seq[main_element];
      const float lead_x = vx[pindex] * inv_len[pindex];
      const float lead_y = vy[pindex] * inv_len[pindex];
      switch (main_element) {
        case 0: base_a = lead_x; base_b = lead_y; break;
        case 1: base_a = lead_y; base_b = -lead_x; break;
        case 2: base_a = -lead_x; base_b = -lead_y; break;
        default: base_a = -lead_y; base_b = lead_x; break;
      }
    }
    seq[main_element] += 1;
    if (seq[main_element] == n) seq[main_element] = 0;
    float dx = pts[seq[1]].x - pts[seq[3]].x;
    float dy = pts[seq[1]].y - pts[seq[3]].y;
    const float width = dx * base_a + dy * base_b;
    dx = pts[seq[2]].x - pts[seq[0]].x;
    dy = pts[seq[2]].y - pts[seq[0]].y;
    const float height = -dx * base_b + dy * base_a;
    const float area = width * height;
    if (area <= minarea) {
      minarea = area;
      b_left = seq[3]; b_a = base_a; b_w = width; b_b = base_b; b_h = height; b_bottom = seq[0];
    }
  }
  const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
  const float C1 = A1 * pts[b_left].x + pts[b_left].y * B1;
  const float C2 = A2 * pts[b_bottom].x + pts[b_bottom].y * B2;
  const float idet = 1.f / (A1 * B2 - A2 * B1);
  out[0] = (C1 * B2 - C2 * B1) * idet;
  out[1] = (A1 * C2 - A2 * C1) * idet;
  out[2] = A1 * b_w; out[3] = B1 * b_w;
  out[4] = A2 * b_h; out[5] = B2 * b_h;
}

// cv::minAreaRect on a convex hull (n >= 1).  `vx, vy, inv_len` are scratch arrays of n floats.
GEOM_HD RotRect min_area_rect_hull(const P2* hull, int n, float* vx, float* vy, float* inv_len) {
  RotRect box;
  const double kPi = 3.1415926535897932384626433832795;
  if (n > 2) {
    float out[6];
    rotating_calipers_min_area(hull, n, vx, vy, inv_len, out);
    box.cx = out[0] + (out[2] + out[4]) * 0.5f;
    box.cy = out[1] + (out[3] + out[5]) * 0.5f;
    box.w = float(sqrt(double(out[2]) * out[2] + double(out[3]) * out[3]));
    box.h = float(sqrt(double(out[4]) * out[4] + double(out[5]) * out[5]));
    box.angle = float(atan2(double(out[3]), double(out[2])));
  } else if (n == 2) {
    box.cx = (hull[0].x + hull[1].x) * 0.5f;
    box.cy = (hull[0].y + hull[1].y) * 0.5f;
    const double dx = double(hull[1].x) - double(hull[0].x), dy = double(hull[1].y) - double(hull[0].y);
    box.w = float(sqrt(dx * dx + dy * dy));
    box.h = 0.f;
    box.angle = float(atan2(dy, dx));
  } else if (n == 1) {
    box.cx = hull[0].x;
    box.cy = hull[0].y;
  }
  box.angle = float(double(box.angle) * 180. / kPi);
  return box;
}

// cv::RotatedRect::points / cv::boxPoints
GEOM_HD void box_points(const RotRect& r, P2 pt[4]) {
  const double kPi = 3.1415926535897932384626433832795;
  const double ang = double(r.angle) * kPi / 180.;
  const float b = float(cos(ang)) * 0.5f;
  const float a = float(sin(ang)) * 0.5f;
  pt[0].x = r.cx - a * r.h - b * r.w;
  pt[0].y = r.cy + b * r.h - a * r.w;
  pt[1].x = r.cx + a * r.h - b * r.w;
  pt[1].y = r.cy - b * r.h - a * r.w;
  pt[2].x = 2 * r.cx - pt[0].x;
  pt[2].y = 2 * r.cy - pt[0].y;
  pt[3].x = 2 * r.cx - pt[1].x;
  pt[3].y = 2 * r.cy - pt[1].y;
}

// GetMiniBoxes (reference postprocess_op.cpp:134-168): [tl, tr, br, bl] and ssid = max(w, h).
GEOM_HD float mini_box(const RotRect& r, P2 box[4]) {
  P2 a[4];
  box_points(r, a);
  // std::sort on 4 elements = insertion sort with "a.x < b.x" (stable on ties)
  for (int i = 1; i < 4; ++i) {
    const P2 v = a[i];
    int j = i - 1;
    while (j >= 0 && v.x < a[j].x) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
  P2 i1, i2, i3, i4;
  if (a[3].y <= a[2].y) { i2 = a[3]; i3 = a[2]; } else { i2 = a[2]; i3 = a[3]; }
  if (a[1].y <= a[0].y) { i1 = a[1]; i4 = a[0]; } else { i1 = a[0]; i4 = a[1]; }
  box[0] = i1; box[1] = i2; box[2] = i3; box[3] = i4;
  return r.w > r.h ? r.w : r.h;
}

// ---------------------------------------------------------------------------------------------
// cv::fillPoly of one quad with integer vertices into a mask of mw x mh pixels: mask(y, x) == inside(y, x).
// OpenCV (imgproc/src/drawing.cpp, CollectPolyEdges + FillEdgeCollection) draws each edge with the 8-connected
// Bresenham line of cv::LineIterator -- CLIPPED to the mask first (cv::clipLine), then ordered left to right -- and
// fills the scan lines y0 <= y < y1 of every non-horizontal edge between ceil(x_left) and floor(x_right) in 16.16 fixed
// point (x advances by the truncated slope per line).  For an edge with an end point outside the mask the fill edge is
// rebuilt from the clipped end points: x always, y only when the clipped segment is not horizontal (a segment clipped
// to a single point becomes a vertical edge at the mask border over the original y range).  Bit-identical to
// cv2.fillPoly (4.13) for vertices inside AND outside the mask (tests/test_oracle_cpu.py::test_geom_fill_poly_vs_cv2).
struct QuadMask {
  int vx[4], vy[4];
  // per edge (i -> i+1)
  int64_t ex[4], edx[4];
  int ey0[4], ey1[4];
  bool eok[4];
  // outline segments after clipping
  int lax[4], lay[4], lbx[4], lby[4];
  bool lok[4];

  // cv::clipLine(Size(w, h), pt1, pt2): the end points are modified in place, also when the result is "outside"
  GEOM_HD static bool clip_line(int w, int h, int64_t& x1, int64_t& y1, int64_t& x2, int64_t& y2) {
    const int64_t right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      int64_t a;
      if (c1 & 12) {
        a = c1 < 8 ? 0 : bottom;
        x1 += int64_t(double(a - y1) * double(x2 - x1) / double(y2 - y1));
        y1 = a;
        c1 = (x1 < 0) + (x1 > right) * 2;
      }
      if (c2 & 12) {
        a = c2 < 8 ? 0 : bottom;
        x2 += int64_t(double(a - y2) * double(x2 - x1) / double(y2 - y1));
        y2 = a;
        c2 = (x2 < 0) + (x2 > right) * 2;
      }
      if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        if (c1) {
          a = c1 == 1 ? 0 : right;
          y1 += int64_t(double(a - x1) * double(y2 - y1) / double(x2 - x1));
          x1 = a;
          c1 = 0;
        }
        if (c2) {
          a = c2 == 1 ? 0 : right;
          y2 += int64_t(double(a - x2) * double(y2 - y1) / double(x2 - x1));
          x2 = a;
          c2 = 0;
        }
      }
    }
    return (c1 | c2) == 0;
  }

  GEOM_HD void init(const int x[4], const int y[4], int mw, int mh) {
    for (int i = 0; i < 4; ++i) { vx[i] = x[i]; vy[i] = y[i]; }
    for (int i = 0; i < 4; ++i) {
      const int j = (i + 3) & 3;  // edge from vertex j (previous) to vertex i, like CollectPolyEdges
      // outline: Line(img, t0, t1) = LineIterator over the clipped segment
      int64_t ax = vx[j], ay = vy[j], bx = vx[i], by = vy[i];
      lok[i] = clip_line(mw, mh, ax, ay, bx, by);
      lax[i] = int(ax); lay[i] = int(ay); lbx[i] = int(bx); lby[i] = int(by);
      // fill edge
      int64_t x0 = int64_t(vx[j]) << 16, x1 = int64_t(vx[i]) << 16;
      int64_t y0c = vy[j], y1c = vy[i];
      const int y0 = vy[j], y1 = vy[i];
      const bool outside = unsigned(vx[j]) >= unsigned(mw) || unsigned(vx[i]) >= unsigned(mw) ||
                           unsigned(vy[j]) >= unsigned(mh) || unsigned(vy[i]) >= unsigned(mh);
      if (outside) {  // "use clipped endpoints to create a more accurate PolyEdge" (ax.. already hold clipLine's result)
        if (ay != by) { y0c = ay; y1c = by; }
        x0 = ax << 16;
        x1 = bx << 16;
      }
      eok[i] = y0 != y1;
      if (!eok[i]) continue;
      edx[i] = (x1 - x0) / (y1c - y0c);
      if (y0 < y1) { ey0[i] = y0; ey1[i] = y1; ex[i] = x0 + (int64_t(y0) - y0c) * edx[i]; }
      else { ey0[i] = y1; ey1[i] = y0; ex[i] = x1 + (int64_t(y1) - y1c) * edx[i]; }
    }
  }

  // Bresenham membership: is (px,py) on the 8-connected line a->b as cv::LineIterator(…, 8, leftToRight) draws it?
  GEOM_HD static bool on_line(int ax, int ay, int bx, int by, int px, int py) {
    int dx = bx - ax, dy = by - ay;
    if (dx < 0) { const int tx = ax, ty = ay; ax = bx; ay = by; bx = tx; by = ty; dx = -dx; dy = -dy; }
    const int ystep = dy < 0 ? -1 : 1;
    const int ady = dy < 0 ? -dy : dy;
    if (ady > dx) {
      // y is the major axis: one pixel per row
      const int t = (py - ay) * ystep;  // step index along the major axis
      if (t < 0 || t > ady) return false;
      // minor (x) steps taken after t major steps: count of k in [0,t) with err_k < 0,
      // err_0 = ady - 2dx, err_{k+1} = err_k - 2dx + (err_k < 0 ? 2ady : 0)
      // closed form: minor(t) = floor((2*dx*t + ady - 1 + ... )) -> evaluate with the standard identity
      //   minor(t) = ceil((2*dx*t - ady) / (2*ady)) clipped at 0 ... derived below in minor_steps()
      return px == ax + minor_steps(ady, dx, t);
    } else {
      const int t = px - ax;
      if (t < 0 || t > dx) return false;
      return py == ay + ystep * minor_steps(dx, ady, t);
    }
  }

  // Number of minor-axis steps after `t` major steps of the Bresenham iteration
  //   err = major - 2*minor; each step: if (err < 0) { minor step; err += 2*major; } err -= 2*minor;
  // A minor step happens at step k (0-based) iff major - 2*minor*(k+1) + 2*major*m_k < 0 where m_k is the
  // number of minor steps before k, so m after t steps = number of integers j>=1 with
  //   2*minor*(step index) ... -> m(t) = floor((2*minor*t + major - 1) / (2*major))  for major > 0 ... verified
  //   against cv2.line in tests/test_geom_host.py.
  GEOM_HD static int minor_steps(int major, int minor, int t) {
    if (major == 0) return 0;
    // step k takes a minor step iff err_k < 0; err_k = major - 2*minor*(k+1) + 2*major*m_k
    // => m_{k+1} = m_k + [major - 2*minor*(k+1) + 2*major*m_k < 0]
    // closed form: m_t = max over m of ... ; use floor((2*minor*t - 1 + major) / (2*major)) when minor>0
    if (minor == 0) return 0;
    const long long num = 2LL * minor * t - major + 2LL * major - 1;  // ceil((2*minor*t - major) / (2*major))
    long long m = num / (2LL * major);
    if (2LL * minor * t - major <= 0) m = 0;
    // strictness: a minor step is taken when err < 0 (strict)
    return int(m);
  }

  GEOM_HD bool inside(int px, int py) const {
    // outline
    for (int i = 0; i < 4; ++i)
      if (lok[i] && on_line(lax[i], lay[i], lbx[i], lby[i], px, py)) return true;
    // scan-line fill: edges active on row py sorted by x; pairs (0,1), (2,3)
    int64_t xs[4];
    int cnt = 0;
    for (int i = 0; i < 4; ++i) {
      if (!eok[i] || py < ey0[i] || py >= ey1[i]) continue;
      const int64_t x = ex[i] + int64_t(py - ey0[i]) * edx[i];
      int k = cnt++;
      while (k > 0 && xs[k - 1] > x) { xs[k] = xs[k - 1]; --k; }
      xs[k] = x;
    }
    for (int k = 0; k + 1 < cnt; k += 2) {
      const int x1 = int((xs[k] + 65535) >> 16), x2 = int(xs[k + 1] >> 16);
      if (px >= x1 && px <= x2) return true;
    }
    return false;
  }
};

// ---------------------------------------------------------------------------------------------
// GetContourArea (reference postprocess_op.cpp:20-37), float32 as written.
GEOM_HD float unclip_distance(const P2 box[4], float unclip_ratio) {
  float area = 0.f, dist = 0.f;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    area += box[i].x * box[j].y - box[i].y * box[j].x;
    dist += sqrtf((box[i].x - box[j].x) * (box[i].x - box[j].x) + (box[i].y - box[j].y) * (box[i].y - box[j].y));
  }
  area = fabsf(float(double(area) / 2.0));
  return area * unclip_ratio / dist;
}

GEOM_HD long long clip_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

constexpr int kMaxOffsetPts = 512;

// ClipperOffset(jtRound, etClosedPolygon) of one integer quad, before the union pass (which does not
// change the hull for delta > 0): reference clipper.cpp AddPath :3628-3673, FixOrientations :3682-3702,
// DoOffset :3779-3944, OffsetPoint :3947-3994, DoRound :4007-4021.  Returns the number of points
// written to (ox, oy) or 0 when Clipper would reject the path (fewer than 3 distinct vertices).
GEOM_HD int offset_round(const long long qx[4], const long long qy[4], double delta, float* ox, float* oy, int cap) {
  long long sx[4], sy[4];
  int hi = 3;
  while (hi > 0 && qx[0] == qx[hi] && qy[0] == qy[hi]) --hi;
  int n = 0;
  sx[n] = qx[0]; sy[n] = qy[0]; ++n;
  for (int i = 1; i <= hi; ++i)
    if (sx[n - 1] != qx[i] || sy[n - 1] != qy[i]) { sx[n] = qx[i]; sy[n] = qy[i]; ++n; }
  if (n < 3) return 0;
  double a = 0;
  for (int i = 0, j = n - 1; i < n; ++i) { a += (double(sx[j]) + double(sx[i])) * (double(sy[j]) - double(sy[i])); j = i; }
  if (!(-a * 0.5 >= 0)) {
    for (int i = 0, j = n - 1; i < j; ++i, --j) {
      long long t = sx[i]; sx[i] = sx[j]; sx[j] = t;
      t = sy[i]; sy[i] = sy[j]; sy[j] = t;
    }
  }
  int m = 0;
  if (fabs(delta) < 1.0E-20) {
    for (int i = 0; i < n && m < cap; ++i) { ox[m] = float(sx[i]); oy[m] = float(sy[i]); ++m; }
    return m;
  }
  const double pi = 3.141592653589793238, two_pi = pi * 2;
  double y = 0.25;  // ArcTolerance default, def_arc_tolerance = 0.25
  if (y > fabs(delta) * 0.25) y = fabs(delta) * 0.25;
  double steps = pi / acos(1 - y / fabs(delta));
  if (steps > fabs(delta) * pi) steps = fabs(delta) * pi;
  double m_sin = sin(two_pi / steps);
  const double m_cos = cos(two_pi / steps);
  const double steps_per_rad = steps / two_pi;
  if (delta < 0.0) m_sin = -m_sin;
  double nx[4], ny[4];
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    double dx = double(sx[j] - sx[i]), dy = double(sy[j] - sy[i]);
    if (dx == 0 && dy == 0) { nx[i] = 0; ny[i] = 0; continue; }
    const double f = 1.0 / sqrt(dx * dx + dy * dy);
    dx *= f; dy *= f;
    nx[i] = dy; ny[i] = -dx;
  }
  auto emit = [&](double X, double Y) {
    if (m < cap) { ox[m] = float(clip_round(X)); oy[m] = float(clip_round(Y)); ++m; }
  };
  int k = n - 1;
  for (int j = 0; j < n; ++j) {
    double sin_a = nx[k] * ny[j] - nx[j] * ny[k];
    if (fabs(sin_a * delta) < 1.0) {
      const double cos_a = nx[k] * nx[j] + ny[j] * ny[k];
      if (cos_a > 0) { emit(double(sx[j]) + nx[k] * delta, double(sy[j]) + ny[k] * delta); continue; }
    } else if (sin_a > 1.0) sin_a = 1.0;
    else if (sin_a < -1.0) sin_a = -1.0;
    if (sin_a * delta < 0) {
      emit(double(sx[j]) + nx[k] * delta, double(sy[j]) + ny[k] * delta);
      emit(double(sx[j]), double(sy[j]));
      emit(double(sx[j]) + nx[j] * delta, double(sy[j]) + ny[j] * delta);
    } else {
      const double ang = atan2(sin_a, nx[k] * nx[j] + ny[k] * ny[j]);
      long long st = clip_round(steps_per_rad * fabs(ang));
      if (st < 1) st = 1;
      double X = nx[k], Y = ny[k];
      for (long long i = 0; i < st; ++i) {
        emit(double(sx[j]) + X * delta, double(sy[j]) + Y * delta);
        const double X2 = X;
        X = X * m_cos - m_sin * Y;
        Y = X2 * m_sin + Y * m_cos;
      }
      emit(double(sx[j]) + nx[j] * delta, double(sy[j]) + ny[j] * delta);
    }
    k = j;
  }
  return m;
}

// Sort (in place) by (y, x) — insertion sort, for the few dozen offset points of one box.
GEOM_HD void sort_yx(float* x, float* y, int n) {
  for (int i = 1; i < n; ++i) {
    const float vx = x[i], vy = y[i];
    int j = i - 1;
    while (j >= 0 && (y[j] > vy || (y[j] == vy && x[j] > vx))) { x[j + 1] = x[j]; y[j + 1] = y[j]; --j; }
    x[j + 1] = vx; y[j + 1] = vy;
  }
}

GEOM_HD float c_roundf(float v) { return roundf(v); }

// Tail of BoxesFromBitmap (reference postprocess_op.cpp:312-324) + OrderPointsClockwise (:87-104) +
// FilterTagDetRes (:333-362).  Returns false when the box is filtered out.
GEOM_HD bool finish_box(const P2 clip[4], int width, int height, float ratio_w, float ratio_h, int src_w, int src_h,
                        int out[8]) {
  int b[4][2];
  for (int k = 0; k < 4; ++k) {
    float x = c_roundf(clip[k].x / float(width) * float(width));
    float y = c_roundf(clip[k].y / float(height) * float(height));
    x = x < 0.f ? 0.f : (x > float(width) ? float(width) : x);
    y = y < 0.f ? 0.f : (y > float(height) ? float(height) : y);
    b[k][0] = int(x);
    b[k][1] = int(y);
  }
  // stable sort by x
  for (int i = 1; i < 4; ++i) {
    const int vx = b[i][0], vy = b[i][1];
    int j = i - 1;
    while (j >= 0 && vx < b[j][0]) { b[j + 1][0] = b[j][0]; b[j + 1][1] = b[j][1]; --j; }
    b[j + 1][0] = vx; b[j + 1][1] = vy;
  }
  int l0 = 0, l1 = 1, r0 = 2, r1 = 3;
  if (b[l0][1] > b[l1][1]) { l0 = 1; l1 = 0; }
  if (b[r0][1] > b[r1][1]) { r0 = 3; r1 = 2; }
  const int ord[4] = {l0, r0, r1, l1};
  for (int k = 0; k < 4; ++k) {
    int x = int(float(b[ord[k]][0]) / ratio_w);
    int y = int(float(b[ord[k]][1]) / ratio_h);
    x = x < 0 ? 0 : (x > src_w - 1 ? src_w - 1 : x);
    y = y < 0 ? 0 : (y > src_h - 1 ? src_h - 1 : y);
    out[2 * k] = x;
    out[2 * k + 1] = y;
  }
  const double dw = sqrt(double((out[0] - out[2]) * (out[0] - out[2]) + (out[1] - out[3]) * (out[1] - out[3])));
  const double dh = sqrt(double((out[0] - out[6]) * (out[0] - out[6]) + (out[1] - out[7]) * (out[1] - out[7])));
  return !(int(dw) <= 4 || int(dh) <= 4);
}

}  // namespace geom
}  // namespace b200ocr
