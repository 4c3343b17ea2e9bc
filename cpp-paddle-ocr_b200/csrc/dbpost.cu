// DB post-process on the GPU: thresholded bitmap -> text boxes, for a whole batch of images at once.
//
// Replaces DBPostProcessor::BoxesFromBitmap + FilterTagDetRes (reference src/postprocess_op.cpp:255-362),
// i.e. cv::findContours(RETR_LIST, CHAIN_APPROX_SIMPLE) + cv::minAreaRect + GetMiniBoxes + BoxScoreFast +
// UnClip + clamp/round + OrderPointsClockwise + size filter.
//
// The reference follows borders sequentially (Suzuki).  A border is a pair (8-connected foreground
// component C, 4-connected background component B) that touch; its pixels are the pixels of C that have a
// 4-neighbour in B.  That set-based statement is what runs here, data-parallel over pixels:
//   1. ccl_*       one union-find labelling for both classes at once (foreground 8-connected, background
//                  4-connected); every component's label is its first pixel in raster order.
//   2. mark        every border pixel finds its border's *start pixel* -- the pixel where the reference's
//                  raster scan would have started following it: the first pixel of C for C's outer border,
//                  the pixel left of B's first pixel for the border of hole B -- and flags it / grows that
//                  border's bounding box.
//   3. list        start pixels in descending raster order = the reference's contour order
//                  (cv::findContours returns RETR_LIST contours last-found first); the first
//                  `max_candidates` (1000) are kept, like postprocess_op.cpp:271-272.
//   4. boxes       one CTA per candidate: row extremes of the border -> convex hull -> rotating calipers
//                  (= cv::minAreaRect of the contour, which only depends on the hull) -> mini box -> masked
//                  mean of the probability map -> unclip -> final integer box, all with geom.h.
// Verified against cv2.findContours on random bitmaps (same count, order, start pixels, pixel sets) in
// tests/test_dbpost_gpu.py.  Integer / index work is bit-exact; box vertices agree within 1 px.
#include "kernels.h"
#include "pdl.h"
#include "geom.h"

#include <algorithm>
#include <cstdio>

namespace b200ocr {

namespace {

constexpr int kFrame = -1;

// ---------------------------------------------------------------- 1. labelling
__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) { i = p; p = L[i]; }
  return i;
}
// The same with path halving (the intermediate pointer jumping of ECL-CC, Jaiganesh & Burtscher, HPDC 2018): every
// element the walk passes is re-pointed at its grandparent.  Parents only ever move towards the root of the same set, so
// the plain stores race benignly with each other and with uf_union's atomicMin (which only ever LINKS at a root; a
// non-root never becomes a root again).  Without it the frame-connected background -- one run per row, each linked to
// the row above -- is a chain as long as the map is high, walked in full by every union and every pixel of the flatten
// pass.  The labels that come out are the same: the smallest pixel index of the component.
__device__ __forceinline__ int uf_find_halving(int* L, int i) {
  int cur = L[i];
  if (cur != i) {
    int prev = i, next;
    while (cur > (next = L[cur])) { L[prev] = next; prev = cur; cur = next; }
  }
  return cur;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find_halving(L, a);
    b = uf_find_halving(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256)
ccl_init_kernel(int* __restrict__ L, int* __restrict__ aux, long total) {
  pdl_trigger();
  pdl_wait();
  // aux: per pixel, zeroed: border-touch flag of background roots / contour flag of start pixels
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    L[t] = int(t);  // global index; images never merge because neighbours are taken inside one image
    aux[t] = 0;
  }
}

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int n, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int r = int(t % per);
    const int y = r / w, x = r - y * w;
    const bool fg = bm[t] != 0;
    if (x > 0 && (bm[t - 1] != 0) == fg) uf_union(L, int(t), int(t - 1));
    if (y > 0) {
      if ((bm[t - w] != 0) == fg) uf_union(L, int(t), int(t - w));
      if (fg) {
        if (x > 0 && bm[t - w - 1] != 0) uf_union(L, int(t), int(t - w - 1));
        if (x + 1 < w && bm[t - w + 1] != 0) uf_union(L, int(t), int(t - w + 1));
      }
    }
  }
}

// Run-based variant for w % 32 == 0 (every detection map: its sides are multiples of 32): a warp owns 32
// consecutive pixels of one row, a ballot finds the horizontal runs, and every pixel starts out pointing at the first
// pixel of its run inside the warp's segment -- no atomics.  Unions are then only issued where runs meet: at segment
// boundaries, at the first pixel of every vertical overlap with a run of the row above, and for the two diagonal
// contacts of the 8-connected foreground that are not already implied by a horizontal or vertical contact.
__global__ void __launch_bounds__(256)
ccl_runs_init_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int* __restrict__ aux, long total) {
  pdl_trigger();
  pdl_wait();
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(0xffffffffu, bm[t] != 0);
    const unsigned diff = (mask ^ (mask << 1)) | 1u;       // bit i: pixel i starts a run inside this segment
    const unsigned upto = diff & (0xffffffffu >> (31 - lane));
    const int start = 31 - __clz(upto);
    L[t] = int(t - lane + start);
    aux[t] = 0;
  }
}

__global__ void __launch_bounds__(256)
ccl_runs_merge_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int n, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int r = int(t % per);
    const int y = r / w, x = r - y * w;
    const bool fg = bm[t] != 0;
    const bool left_same = x > 0 && (bm[t - 1] != 0) == fg;
    if ((x & 31) == 0 && left_same) uf_union(L, int(t), int(t - 1));  // runs continue across the segment boundary
    if (y == 0) continue;
    const bool up_same = (bm[t - w] != 0) == fg;
    if (up_same) {
      // first pixel of this overlap: the pair (left, up-left) does not already carry the same contact
      const bool carried = left_same && (bm[t - w - 1] != 0) == fg;
      if (!carried) uf_union(L, int(t), int(t - w));
    } else if (fg) {
      // 8-connectivity: diagonal contacts that no horizontal / vertical contact implies
      if (x > 0 && bm[t - w - 1] != 0 && bm[t - 1] == 0) uf_union(L, int(t), int(t - w - 1));
      if (x + 1 < w && bm[t - w + 1] != 0 && bm[t + 1] == 0) uf_union(L, int(t), int(t - w + 1));
    }
  }
}

// flatten + flag background components that touch the image frame (they are the "outside")
__global__ void __launch_bounds__(256)
ccl_flatten_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int* __restrict__ touch, int n, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    // read-only walk: a halving store of another thread could land after this pixel's final label and leave it non-flat
    const int root = uf_find(L, int(t));
    L[t] = root;
    if (bm[t] == 0) {
      const int r = int(t % per);
      const int y = r / w, x = r - y * w;
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) touch[root] = 1;
    }
  }
}

// background component id as the reference sees it: everything connected to the frame is one component
__device__ __forceinline__ int canon_bg(const int* L, const int* touch, long q) {
  const int b = L[q];
  return touch[b] ? kFrame : b;
}

// the background component that surrounds foreground component C (C = global index of its first pixel)
__device__ __forceinline__ int outer_bg(const int* L, const int* touch, int C, int w) {
  if ((C % w) == 0) return kFrame;
  return canon_bg(L, touch, long(C) - 1);
}

struct Slot {  // per-pixel arrays, meaningful at contour start pixels only
  int* flag;
  int* x0; int* y0; int* x1; int* y1;
  // "slow" score mode only (nullptr otherwise).  Fixed-point (Q32) sums of the probability map and pixel counts:
  //   at the first pixel of a component (foreground or hole): over the component AND everything nested inside it,
  //   at the start pixel of a hole border (the pixel left of the hole's first pixel): over that border's pixels.
  // The two kinds of index never coincide (a hole's first pixel lies below the first row of the component around it).
  unsigned long long* sum;
  int* cnt;
  // start pixels appended by mark_kernel (first setter of a flag appends): cand[img][kCandCap], cand_n[img]
  int* cand;
  int* cand_n;
};
constexpr int kCandCap = 4096;

__global__ void __launch_bounds__(256)
slot_init_kernel(Slot s, long total) {
  pdl_trigger();
  pdl_wait();
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    s.flag[t] = 0;
    s.x0[t] = 0x7fffffff; s.y0[t] = 0x7fffffff; s.x1[t] = -1; s.y1[t] = -1;
    if (s.sum) { s.sum[t] = 0ull; s.cnt[t] = 0; }
  }
}

// Q32 fixed point: exact for p >= 2^-9, truncated below (error < 2^-32 per pixel); integer sums make the score
// independent of the order in which the atomics arrive
__device__ __forceinline__ unsigned long long prob_q32(float p) {
  return (unsigned long long)(double(fminf(fmaxf(p, 0.f), 1.f)) * 4294967296.0);
}

// ---------------------------------------------------------------- "slow" score: sums over nested regions
// PolygonScoreAcc (reference src/postprocess_op.cpp:170-214) fills the contour polygon and averages the probability
// map under it.  A traced border only has horizontal, vertical and diagonal unit steps, so cv::fillPoly of it is the
// border's pixels plus everything the border encloses.  Set-based: the components form a containment tree
// (a foreground component's parent is the background component left of its first pixel; a hole's parent is the
// foreground component left of its first pixel; the frame-connected background is the root).  The polygon of C's
// outer border covers C and all its descendants; the polygon of hole B's border covers B, its descendants and the
// border pixels themselves.  Every pixel adds its value to all its ancestors (depth is 1-3 on text maps).
__global__ void __launch_bounds__(256)
nest_sum_kernel(const uint8_t* __restrict__ bm, const float* __restrict__ prob, const int* __restrict__ L,
                const int* __restrict__ touch, Slot s, int n, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const unsigned long long q = prob_q32(prob[t]);
    const long ibase = (t / per) * per;
    int node;           // global index of the first pixel of the current ancestor
    bool fg = bm[t] != 0;
    if (fg) node = L[t];
    else {
      node = canon_bg(L, touch, t);
      if (node == kFrame) continue;
    }
    for (int depth = 0; depth < 4096; ++depth) {
      atomicAdd(&s.sum[node], q);
      atomicAdd(&s.cnt[node], 1);
      if (fg) {  // parent of a foreground component: the background left of its first pixel
        const int local = int(node - ibase);
        const int b = (local % w) == 0 ? kFrame : canon_bg(L, touch, long(node) - 1);
        if (b == kFrame) break;
        node = b;
        fg = false;
      } else {   // parent of a hole: the foreground component left of its first pixel
        node = L[node - 1];
        fg = true;
      }
    }
  }
}

// ---------------------------------------------------------------- 2. border pixels -> start pixel
__global__ void __launch_bounds__(256)
mark_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ L, const int* __restrict__ touch, Slot s,
            int n, int h, int w, const float* __restrict__ prob) {
  pdl_trigger();
  pdl_wait();
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    if (bm[t] == 0) continue;
    const int r = int(t % per);
    const int y = r / w, x = r - y * w;
    const int C = L[t];
    const int bout = outer_bg(L, touch, C, w);
    int seen[4];
    int ns = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dx = (k == 2) ? -1 : (k == 3 ? 1 : 0);
      const int dy = (k == 0) ? -1 : (k == 1 ? 1 : 0);
      const int qx = x + dx, qy = y + dy;
      int B;
      if (qx < 0 || qy < 0 || qx >= w || qy >= h) B = kFrame;
      else {
        const long q = t + dy * w + dx;
        if (bm[q] != 0) continue;
        B = canon_bg(L, touch, q);
      }
      bool dup = false;
      for (int j = 0; j < ns; ++j) dup |= seen[j] == B;
      if (dup) continue;
      seen[ns++] = B;
      const int slot = (B == bout) ? C : B - 1;  // outer border of C, or border of hole B
      if (slot < 0) continue;  // cannot happen: a component that touches the frame has the frame as outer background
      if (atomicExch(&s.flag[slot], 1) == 0) {  // first border pixel of this contour: register its start pixel
        const int img = int(t / per);
        const int pos = atomicAdd(&s.cand_n[img], 1);
        if (pos < kCandCap) s.cand[img * kCandCap + pos] = slot;
      }
      atomicMin(&s.x0[slot], x); atomicMin(&s.y0[slot], y);
      atomicMax(&s.x1[slot], x); atomicMax(&s.y1[slot], y);
      if (s.sum && B != bout) {  // "slow" score: the hole border's own pixels
        atomicAdd(&s.sum[slot], prob_q32(prob[t]));
        atomicAdd(&s.cnt[slot], 1);
      }
    }
  }
}

// ---------------------------------------------------------------- 3. ordered candidate list
// Fast path: the start pixels mark_kernel appended (unordered) are sorted in shared memory, descending = the reference's
// contour order; the first max_cand stay.  Only when an image has more than kCandCap contours (noise) does the
// scan-based list_kernel below do the work.
__global__ void __launch_bounds__(1024)
sort_candidates_kernel(const int* __restrict__ cand, const int* __restrict__ cand_n, int max_cand, int* __restrict__ counts,
                       int* __restrict__ list) {
  pdl_trigger();
  pdl_wait();
  __shared__ int key[kCandCap];
  const int img = blockIdx.x;
  const int n = cand_n[img];
  if (n > kCandCap) return;  // list_kernel takes this image
  int m = 1;
  while (m < n) m <<= 1;     // bitonic network size
  for (int i = threadIdx.x; i < m; i += blockDim.x) key[i] = i < n ? cand[img * kCandCap + i] : -1;
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool desc = (i & k) == 0;  // descending overall
          const int a = key[i], b = key[p];
          if (desc ? a < b : a > b) { key[i] = b; key[p] = a; }
        }
      }
      __syncthreads();
    }
  const int keep = n < max_cand ? n : max_cand;
  for (int i = threadIdx.x; i < keep; i += blockDim.x) list[long(img) * max_cand + i] = key[i];
  if (threadIdx.x == 0) counts[img] = keep;
}

// One CTA per image; walks the flag array from the last pixel backwards in chunks, block-scans the flags and
// appends start pixels (global indices) until max_candidates are collected.
__global__ void __launch_bounds__(1024)
list_kernel(const int* __restrict__ flag, int h, int w, int max_cand, int* __restrict__ counts, int* __restrict__ list,
            const int* __restrict__ cand_n) {
  pdl_trigger();
  pdl_wait();
  constexpr int kPer = 8;
  __shared__ int warp_sums[32];
  __shared__ int base;
  const int img = blockIdx.x;
  if (cand_n[img] <= kCandCap) return;  // sort_candidates_kernel already listed this image
  const long per = long(h) * w;
  const int* f = flag + img * per;
  int* out = list + long(img) * max_cand;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long hi = per; hi > 0; hi -= long(blockDim.x) * kPer) {
    // thread i covers reversed positions [i*kPer, (i+1)*kPer) counted from hi-1 downwards
    int bits = 0, cnt = 0;
    const long first = hi - 1 - long(threadIdx.x) * kPer;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const long p = first - k;
      if (p >= 0 && f[p]) { bits |= 1 << k; ++cnt; }
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int v = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      warp_sums[lane] = v;
    }
    __syncthreads();
    const int b0 = base;
    int pos = b0 + inc - cnt + (warp ? warp_sums[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (bits & (1 << k)) {
        if (pos < max_cand) out[pos] = int(img * per + (first - k));
        ++pos;
      }
    __syncthreads();
    if (threadIdx.x == 0) base = b0 + warp_sums[31];
    __syncthreads();
    if (base >= max_cand) break;
  }
  if (threadIdx.x == 0) counts[img] = base < max_cand ? base : max_cand;
}

// ---------------------------------------------------------------- 4. one CTA per candidate
// 128 threads: the candidates of a batch (~10 per card) then all run at the same time, five CTAs per SM.  Measured with 512
// threads (one CTA per SM at 112 registers): 134 us instead of 81 us for 32 cards -- the kernel's duration is the latency
// of its slowest candidate (thread 0's hull / calipers / Clipper offset in double precision), not the pixel scans.
constexpr int kBoxThreads = 128;
constexpr int kHullCap = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0;
  for (int i = 0; i < kBoxThreads / 32; ++i) r += sh[i];
  return r;
}

// one candidate (image `img`, position `ord` in the ordered list) by one CTA
__device__ void boxes_one(const DbPostParams& P, const float* __restrict__ prob, const uint8_t* __restrict__ bm,
                          const int* __restrict__ L, const int* __restrict__ touch, const Slot& s,
                          const int* __restrict__ list, const DbImageInfo* __restrict__ info, DbBox* __restrict__ boxes,
                          int img, int ord, int* dyn) {
  __shared__ geom::P2 hull[kHullCap];
  __shared__ float sc_a[kHullCap], sc_b[kHullCap], sc_c[kHullCap];
  __shared__ double red[kBoxThreads / 32];
  __shared__ int sh_i[16];
  __shared__ float sh_f[16];
  const int h = P.h, w = P.w;
  const long per = long(h) * w;
  const long ibase = img * per;
  const int slot = list[long(img) * P.max_candidates + ord];
  DbBox* ob = boxes + long(img) * P.max_candidates + ord;
  const int x0 = s.x0[slot], y0 = s.y0[slot], x1 = s.x1[slot], y1 = s.y1[slot];
  const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
  int* rowmin = dyn;
  int* rowmax = dyn + h;
  for (int i = threadIdx.x; i < bh; i += blockDim.x) { rowmin[i] = 0x7fffffff; rowmax[i] = -1; }
  if (threadIdx.x == 0) { ob->valid = 0; ob->start = int(slot - ibase); ob->score = 0.f; sh_i[0] = 0; }
  __syncthreads();
  // which border is this?  slot is the first pixel of a foreground component -> its outer border;
  // otherwise the border of the hole whose first pixel is slot + 1.
  const bool outer = L[slot] == slot;
  const int C = outer ? slot : L[slot];
  const int Bkey = outer ? outer_bg(L, touch, C, w) : slot + 1;
  int nb = 0;
  // The scan is latency-bound (a chain of dependent loads per pixel): four pixels per thread are in flight at a time,
  // and a pixel's four neighbour bytes are requested together before any of them is looked at.
  const int total_px = bw * bh;
  for (int base = threadIdx.x; base < total_px; base += 4 * blockDim.x) {
    long tt[4];
    int xs[4], ys[4];
    bool comp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * blockDim.x;
      const bool ok = i < total_px;
      const int yy = ok ? i / bw : 0, xx = ok ? i - yy * bw : 0;
      xs[u] = x0 + xx; ys[u] = y0 + yy;
      tt[u] = ibase + long(ys[u]) * w + xs[u];
      comp[u] = ok && bm[tt[u]] != 0 && L[tt[u]] == C;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!comp[u]) continue;
      const int x = xs[u], y = ys[u];
      const long t = tt[u];
      const uint8_t nl = x > 0 ? bm[t - 1] : uint8_t(1), nr = x + 1 < w ? bm[t + 1] : uint8_t(1);
      const uint8_t nu = y > 0 ? bm[t - w] : uint8_t(1), nd = y + 1 < h ? bm[t + w] : uint8_t(1);
      bool on = false;
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) on = Bkey == kFrame;
      if (!on && nl == 0) on = canon_bg(L, touch, t - 1) == Bkey;
      if (!on && nr == 0) on = canon_bg(L, touch, t + 1) == Bkey;
      if (!on && nu == 0) on = canon_bg(L, touch, t - w) == Bkey;
      if (!on && nd == 0) on = canon_bg(L, touch, t + w) == Bkey;
      if (!on) continue;
      ++nb;
      atomicMin(&rowmin[y - y0], x);
      atomicMax(&rowmax[y - y0], x);
    }
  }
  atomicAdd(&sh_i[0], nb);
  __syncthreads();
  // ---- thread 0: contour size test, hull, min-area rect, mini box
  if (threadIdx.x == 0) {
    int ok = 1;
    const int nborder = sh_i[0];
    // cv::findContours(CHAIN_APPROX_SIMPLE) yields <= 2 points exactly for a lone pixel and for straight
    // one-pixel-wide runs (horizontal, vertical, diagonal); the reference skips those (postprocess_op.cpp:277)
    if (bw == 1 || bh == 1 || (bw == bh && nborder == bw)) ok = 0;
    int nh = 0;
    if (ok) {
      auto pt = [&](int i) {
        const int row = i >> 1;
        geom::P2 p;
        p.x = float((i & 1) ? rowmax[row] : rowmin[row]);
        p.y = float(y0 + row);
        return p;
      };
      nh = geom::convex_hull_sorted_yx(pt, 2 * bh, hull, kHullCap);
      if (nh <= 0) ok = 0;
      else {
        geom::P2 st;  // the contour's start pixel: the slot itself (outer: first pixel; hole: the pixel left of the hole)
        st.x = float(int((slot - ibase) % w));
        st.y = float(int((slot - ibase) / w));
        geom::hull_order_like_cv(hull, nh, outer, st);
      }
    }
    if (ok) {
      const geom::RotRect rr = geom::min_area_rect_hull(hull, nh, sc_a, sc_b, sc_c);
      geom::P2 mb[4];
      const float ssid = geom::mini_box(rr, mb);
      if (ssid < 3.f) ok = 0;
      for (int k = 0; k < 4; ++k) { sh_f[2 * k] = mb[k].x; sh_f[2 * k + 1] = mb[k].y; }
    }
    sh_i[1] = ok;
  }
  __syncthreads();
  if (!sh_i[1]) return;
  geom::P2 mb[4];
  for (int k = 0; k < 4; ++k) { mb[k].x = sh_f[2 * k]; mb[k].y = sh_f[2 * k + 1]; }
  if (P.score_slow) {
    // ---- PolygonScoreAcc: mean over the filled contour polygon = nested-region sums (see nest_sum_kernel)
    if (threadIdx.x == 0) {
      unsigned long long qs;
      long long qc;
      if (outer) { qs = s.sum[C]; qc = s.cnt[C]; }
      else { qs = s.sum[slot + 1] + s.sum[slot]; qc = (long long)s.cnt[slot + 1] + s.cnt[slot]; }
      const float score = qc > 0 ? float(double(qs) / 4294967296.0 / double(qc)) : 0.f;
      sh_f[8] = score;
      sh_i[1] = !(score < P.box_thresh);
    }
  } else {
    // ---- BoxScoreFast: masked mean of the probability map over the filled (integer-truncated) quad
    float mnx = mb[0].x, mxx = mb[0].x, mny = mb[0].y, mxy = mb[0].y;
    for (int k = 1; k < 4; ++k) {
      mnx = fminf(mnx, mb[k].x); mxx = fmaxf(mxx, mb[k].x);
      mny = fminf(mny, mb[k].y); mxy = fmaxf(mxy, mb[k].y);
    }
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const int xmin = clampi(int(floorf(mnx)), 0, w - 1), xmax = clampi(int(ceilf(mxx)), 0, w - 1);
    const int ymin = clampi(int(floorf(mny)), 0, h - 1), ymax = clampi(int(ceilf(mxy)), 0, h - 1);
    int qx[4], qy[4];
    for (int k = 0; k < 4; ++k) { qx[k] = int(mb[k].x) - xmin; qy[k] = int(mb[k].y) - ymin; }
    const int mw = xmax - xmin + 1, mh = ymax - ymin + 1;
    geom::QuadMask qm;
    qm.init(qx, qy, mw, mh);
    double sum = 0;
    int cnt = 0;
    for (int i = threadIdx.x; i < mw * mh; i += blockDim.x) {
      const int yy = i / mw, xx = i - yy * mw;
      if (qm.inside(xx, yy)) {
        sum += double(prob[ibase + long(ymin + yy) * w + (xmin + xx)]);
        ++cnt;
      }
    }
    const double tsum = block_sum(sum, red);
    const double tcnt = block_sum(double(cnt), red);
    if (threadIdx.x == 0) {
      const float score = tcnt > 0 ? float(tsum / tcnt) : 0.f;
      sh_f[8] = score;
      sh_i[1] = !(score < P.box_thresh);
    }
  }
  __syncthreads();
  if (!sh_i[1] || threadIdx.x != 0) return;
  // ---- thread 0: unclip -> min-area rect -> mini box -> final integer box
  const float dist = geom::unclip_distance(mb, P.unclip_ratio);
  long long qx[4], qy[4];
  for (int k = 0; k < 4; ++k) { qx[k] = (long long)(int(mb[k].x)); qy[k] = (long long)(int(mb[k].y)); }
  float* ox = sc_a;
  float* oy = sc_b;
  const int m = geom::offset_round(qx, qy, double(dist), ox, oy, geom::kMaxOffsetPts);
  geom::RotRect ur;
  if (m <= 0) { ur.cx = 0.f; ur.cy = 0.f; ur.w = 1.f; ur.h = 1.f; ur.angle = 0.f; }
  else {
    geom::sort_yx(ox, oy, m);
    auto pt = [&](int i) { geom::P2 p; p.x = ox[i]; p.y = oy[i]; return p; };
    const int nh = geom::convex_hull_sorted_yx(pt, m, hull, kHullCap);
    // scratch for the calipers must not alias the hull input: reuse sc_c and the tail of sc_a / sc_b
    ur = geom::min_area_rect_hull(hull, nh, sc_c, sc_a + geom::kMaxOffsetPts, sc_b + geom::kMaxOffsetPts);
  }
  if (ur.h < 1.001f && ur.w < 1.001f) return;
  geom::P2 cb[4];
  const float ssid2 = geom::mini_box(ur, cb);
  if (ssid2 < 5.f) return;
  const DbImageInfo inf = info[img];
  int pts[8];
  if (!geom::finish_box(cb, w, h, inf.ratio_w, inf.ratio_h, inf.src_w, inf.src_h, pts)) return;
  for (int k = 0; k < 8; ++k) ob->pts[k] = pts[k];
  ob->score = sh_f[8];
  ob->valid = 1;
}

// The grid is a fixed number of CTAs per image that walk over the image's candidates: a text image has 10-50 of them,
// a grid of max_candidates (1000) CTAs per image would be 97 % empty CTAs whose launch costs more than the real work.
__global__ void __launch_bounds__(kBoxThreads)
boxes_kernel(DbPostParams P, const float* __restrict__ prob, const uint8_t* __restrict__ bm,
             const int* __restrict__ L, const int* __restrict__ touch, Slot s, const int* __restrict__ counts,
             const int* __restrict__ list, const DbImageInfo* __restrict__ info, DbBox* __restrict__ boxes) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ int dyn[];  // rowmin[h], rowmax[h]
  const int img = blockIdx.y;
  const int cnt = counts[img];
  for (int ord = blockIdx.x; ord < cnt; ord += gridDim.x) {
    boxes_one(P, prob, bm, L, touch, s, list, info, boxes, img, ord, dyn);
    __syncthreads();  // the shared scratch is reused by the next candidate
  }
}

__global__ void __launch_bounds__(256)
dilate2x2_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h, int w) {
  pdl_trigger();
  pdl_wait();
  // cv::dilate with a 2x2 rectangle, anchor (1,1): out(y,x) = max over rows y-1..y, cols x-1..x
  const long per = long(h) * w, total = per * n;
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < total; t += long(gridDim.x) * blockDim.x) {
    const int r = int(t % per);
    const int y = r / w, x = r - y * w;
    uint8_t v = in[t];
    if (x > 0) v = max(v, in[t - 1]);
    if (y > 0) { v = max(v, in[t - w]); if (x > 0) v = max(v, in[t - w - 1]); }
    out[t] = v;
  }
}

__global__ void __launch_bounds__(256)
threshold_kernel(const float* __restrict__ prob, long n, int thresh_u8, uint8_t* __restrict__ bm) {
  pdl_trigger();
  pdl_wait();
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x)
    bm[t] = (int((unsigned char)(prob[t] * 255.f)) > thresh_u8) ? 255 : 0;
}

inline int grid_for(long total, int threads = 256) {
  long b = (total + threads - 1) / threads;
  const long cap = 148L * 8;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}
inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

// workspace layout: L | touch | flag | x0 | y0 | x1 | y1 (int32 per pixel each) | list [n*max_candidates]
//                   | cand [n*kCandCap] | cand_n [n]
//                   [| sum (u64 per pixel) | cnt (int32 per pixel)   in "slow" score mode]
size_t dbpost_workspace_bytes(const DbPostParams& p) {
  const size_t px = size_t(p.n) * p.h * p.w;
  return 7 * al(px * 4) + al(size_t(p.n) * p.max_candidates * 4) + al(size_t(p.n) * kCandCap * 4) + al(size_t(p.n) * 4) +
         (p.score_slow ? al(px * 8) + al(px * 4) : 0);
}

void launch_dbpost(const DbPostParams& p, const float* prob, const uint8_t* bitmap, const DbImageInfo* info_dev,
                   void* workspace, int* counts_dev, DbBox* boxes_dev, cudaStream_t st) {
  const size_t px = size_t(p.n) * p.h * p.w;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* L = reinterpret_cast<int*>(ws);
  int* touch = reinterpret_cast<int*>(ws + al(px * 4));
  Slot s;
  s.flag = reinterpret_cast<int*>(ws + 2 * al(px * 4));
  s.x0 = reinterpret_cast<int*>(ws + 3 * al(px * 4));
  s.y0 = reinterpret_cast<int*>(ws + 4 * al(px * 4));
  s.x1 = reinterpret_cast<int*>(ws + 5 * al(px * 4));
  s.y1 = reinterpret_cast<int*>(ws + 6 * al(px * 4));
  int* list = reinterpret_cast<int*>(ws + 7 * al(px * 4));
  s.cand = reinterpret_cast<int*>(ws + 7 * al(px * 4) + al(size_t(p.n) * p.max_candidates * 4));
  s.cand_n = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s.cand) + al(size_t(p.n) * kCandCap * 4));
  cudaMemsetAsync(s.cand_n, 0, size_t(p.n) * 4, st);
  s.sum = nullptr; s.cnt = nullptr;
  if (p.score_slow) {
    uint8_t* extra = reinterpret_cast<uint8_t*>(s.cand_n) + al(size_t(p.n) * 4);
    s.sum = reinterpret_cast<unsigned long long*>(extra);
    s.cnt = reinterpret_cast<int*>(extra + al(px * 8));
  }
  const int g = grid_for(long(px));
  if (p.w % 32 == 0) launch_k(ccl_runs_init_kernel, dim3(g), dim3(256), 0, st, bitmap, L, touch, long(px));
  else launch_k(ccl_init_kernel, dim3(g), dim3(256), 0, st, L, touch, long(px));
  launch_k(slot_init_kernel, dim3(g), dim3(256), 0, st, s, long(px));
  if (p.w % 32 == 0) launch_k(ccl_runs_merge_kernel, dim3(g), dim3(256), 0, st, bitmap, L, p.n, p.h, p.w);
  else launch_k(ccl_merge_kernel, dim3(g), dim3(256), 0, st, bitmap, L, p.n, p.h, p.w);
  launch_k(ccl_flatten_kernel, dim3(g), dim3(256), 0, st, bitmap, L, touch, p.n, p.h, p.w);
  launch_k(mark_kernel, dim3(g), dim3(256), 0, st, bitmap, L, touch, s, p.n, p.h, p.w, prob);
  if (p.score_slow) launch_k(nest_sum_kernel, dim3(g), dim3(256), 0, st, bitmap, prob, L, touch, s, p.n, p.h, p.w);
  launch_k(sort_candidates_kernel, dim3(p.n), dim3(1024), 0, st, s.cand, s.cand_n, p.max_candidates, counts_dev, list);
  launch_k(list_kernel, dim3(p.n), dim3(1024), 0, st, s.flag, p.h, p.w, p.max_candidates, counts_dev, list, s.cand_n);
  const size_t smem = size_t(2) * p.h * sizeof(int);
  const int per_image = std::min(p.max_candidates, std::max(16, (148 * 8 + p.n - 1) / p.n));
  launch_k(boxes_kernel, dim3(dim3(per_image, p.n)), dim3(kBoxThreads), smem, st, p, prob, bitmap, L, touch, s, counts_dev, list, info_dev,
                                                                 boxes_dev);
}

void launch_threshold(const float* prob, long n, int thresh_u8, uint8_t* bitmap, cudaStream_t s) {
  launch_k(threshold_kernel, dim3(grid_for(n)), dim3(256), 0, s, prob, n, thresh_u8, bitmap);
}

void launch_dilate2x2(const uint8_t* in, uint8_t* out, int n, int h, int w, cudaStream_t s) {
  launch_k(dilate2x2_kernel, dim3(grid_for(long(n) * h * w)), dim3(256), 0, s, in, out, n, h, w);
}

}  // namespace b200ocr
