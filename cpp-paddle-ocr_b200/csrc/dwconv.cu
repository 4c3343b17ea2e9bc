// Depthwise convolution (3x3 / 5x5, strides 1 or 2 per axis) for NHWC fp16 activations on sm_100a.
//
// Replaces Paddle's depthwise_conv2d (+ folded BN / scalar affine / activation) in the det / cls / rec backbones.
//
// A 5x5 depthwise layer does 25 multiply-adds per element it moves: on B200 that puts it at the balance point
// between the fp32 FMA rate (128 FMA/clk/SM -> one warp instruction per scheduler per clock) and HBM, so every
// instruction that is not an FMA costs throughput.  Layout of the work:
//   * one CTA (128 threads) = a tile of RB*R x TW output pixels of one channel chunk; its input halo tile is copied
//     ONCE from global to shared memory with 16-byte cp.async (zero fill = conv padding and tensor edge); several
//     CTAs per SM overlap one tile's copy with another tile's arithmetic,
//   * one lane = one channel PAIR (a half2 word): the lanes of a warp read consecutive words of one pixel
//     (bank-conflict free), every word feeds up to K*K*2 FMAs, the K*K weights of the pair stay in registers
//     (FHFMA: fp16 x fp16 + fp32, operands taken straight from the packed words),
//   * one thread = S=4 consecutive output columns x R output rows: the input window slides down row by row, every
//     shared-memory word is read once per thread and reused across rows and columns in registers; rows below the
//     tensor are skipped warp-uniformly.
// Chunks narrower than 64 channels put several column strips into one warp; one pad pixel per strip width in the
// shared-memory row keeps those strips on different banks.
#include "kernels.h"
#include "pdl.h"

#include <algorithm>
#include <cstdlib>

namespace b200ocr {

namespace {

constexpr int kDwThreads = 128;
constexpr int kS = 4;  // output columns per thread

// fp16 x fp16 + fp32 -> fp32 in one instruction (FHFMA): the product of two halves is exact in fp32, so this
// equals converting both operands and issuing an FFMA -- without the conversions.
__device__ __forceinline__ void fhfma2(uint32_t x, uint32_t w, float& a0, float& a1) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a0) : "h"((unsigned short)(x & 0xffffu)), "h"((unsigned short)(w & 0xffffu)));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a1) : "h"((unsigned short)(x >> 16)), "h"((unsigned short)(w >> 16)));
}

struct DwArgs {
  TV in, out;
  const float* wb;     // fp32 [taps][cp] weights, then [cp] bias
  const __half* wh;    // fp16 [taps][cp] weights (WF32 = false)
  int cp;              // weight row stride (channels, padded)
  int ph, pw;
  int wc_log2;         // log2(words per pixel in the chunk): chunk = 2^(wc_log2+1) channels
  int chunks;          // channel chunks per pixel
  int rbn;             // row blocks per tile (1 or 2)
  int tw;              // output columns per tile
  int tiles_x, tiles_y;
  int iht;             // halo tile height
  int iw, phys_iw;     // halo tile width (pixels) and its padded shared-memory pitch (pixels)
  int act;             // 0 none, 1 relu, 2 hard-swish
  float s2, t2;        // y = s2 * act(acc + bias) + t2   (hard-swish: s2 already holds the 1/6)
  const int* vw;
};

// RB_T / WCL_T: row blocks per tile and log2(words per pixel of the chunk) fixed at compile time (every shared-memory
// offset becomes an immediate and the tile exactly covers R rows: no per-row tests), or 0 / -1 = taken from DwArgs.
template <int K, int SH, int SW, int R, bool WF32, int RB_T, int WCL_T>
__global__ void __launch_bounds__(kDwThreads, 4) dwconv_tile_kernel(const DwArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dw_smem[];
  constexpr bool STATIC = RB_T > 0 && WCL_T >= 0;
  constexpr int WIN = (kS - 1) * SW + K;       // input columns one thread needs per row
  constexpr int IHR = (R - 1) * SH + K;        // input rows one thread needs
  constexpr int SPAN = kS * SW;                // input columns between neighbouring strips
  constexpr int TW_T = STATIC ? kS * (kDwThreads >> (WCL_T < 0 ? 0 : WCL_T)) / (RB_T > 0 ? RB_T : 1) : 0;
  constexpr int IW_T = (TW_T - 1) * SW + K;
  const int wcl = STATIC ? WCL_T : a.wc_log2;
  const int rbn = STATIC ? RB_T : a.rbn;
  const int tw = STATIC ? TW_T : a.tw;
  const int iw = STATIC ? IW_T : a.iw;
  const int phys_iw = STATIC ? IW_T + IW_T / SPAN + 1 : a.phys_iw;
  const int iht = STATIC ? (RB_T * R - 1) * SH + K : a.iht;

  int b = blockIdx.x;
  const int chunk = b % a.chunks; b /= a.chunks;
  const int tx = b % a.tiles_x; b /= a.tiles_x;
  const int ty = b % a.tiles_y;
  const int n = b / a.tiles_y;
  const int c0 = chunk << (wcl + 1);     // first channel of the chunk
  const int c8 = (a.in.c + 7) & ~7;
  const int ox0 = tx * tw, oy0 = ty * rbn * R;
  const int ix0 = ox0 * SW - a.pw, iy0 = oy0 * SH - a.ph;

  // ---- halo tile -> shared memory: [iht][phys_iw][wc] words in 16-byte pieces, zero fill outside the tensor.
  // A thread keeps its (piece, column) and walks down the rows; only the row test changes per copy.
  {
    const int pc_log2 = wcl - 2;         // 16-byte pieces per pixel
    const int piece = threadIdx.x & ((1 << pc_log2) - 1);
    const int cols = kDwThreads >> pc_log2;
    const bool c_ok = c0 + piece * 8 < c8;
    const uint32_t sbase = uint32_t(__cvta_generic_to_shared(dw_smem)) + piece * 16;
    const long row_stride = long(a.in.w) * a.in.pitch;
    const __half* img = a.in.p + long(n) * a.in.h * row_stride + c0 + piece * 8;
    const uint32_t srow = uint32_t(phys_iw) << (wcl + 2);
    for (int x = threadIdx.x >> pc_log2; x < iw; x += cols) {
      const int gx = ix0 + x;
      const bool x_ok = c_ok && gx >= 0 && gx < a.in.w;
      const __half* src = img + long(iy0) * row_stride + long(gx) * a.in.pitch;
      uint32_t dst = sbase + (uint32_t(x + x / SPAN) << (wcl + 2));
      for (int y = 0; y < iht; ++y, src += row_stride, dst += srow) {
        if (unsigned(iy0 + y) >= unsigned(a.in.h)) continue;  // rows above / below the tensor are never read (see below)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(x_ok ? src : a.in.p), "r"(x_ok ? 16 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- this thread's outputs: channel pair `lane_c`, strip `sx` (kS columns), row block `rb` (R rows)
  const int wc = 1 << wcl;
  const int lane_c = threadIdx.x & (wc - 1);
  const int strip = threadIdx.x >> wcl;
  const int sxn = (kDwThreads >> wcl) / rbn;   // strips per row block
  const int rb = strip / sxn, sx = strip - rb * sxn;
  const int ch = c0 + 2 * lane_c;
  const bool ch_ok = ch < c8;
  const int chc = ch_ok ? ch : 0;
  const int row0 = oy0 + rb * R;
  const int nrows = min(R, a.out.h - row0);            // warp-uniform (a warp never spans two row blocks)
  const int col0 = ox0 + sx * kS;

  // weights of this channel pair stay in registers for the whole tile
  uint32_t wq[WF32 ? 1 : K * K];
  float2 wf[WF32 ? K * K : 1];
  if (WF32) {
#pragma unroll
    for (int t = 0; t < K * K; ++t) wf[t] = __ldg(reinterpret_cast<const float2*>(a.wb + long(t) * a.cp + chc));
  } else {
#pragma unroll
    for (int t = 0; t < K * K; ++t) wq[t] = __ldg(reinterpret_cast<const uint32_t*>(a.wh + long(t) * a.cp + chc));
  }
  const float2 bias = __ldg(reinterpret_cast<const float2*>(a.wb + long(K * K) * a.cp + chc));

  float acc[R][kS][2];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int s = 0; s < kS; ++s) { acc[r][s][0] = bias.x; acc[r][s][1] = bias.y; }

  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (nrows <= 0 || col0 >= a.out.w || !ch_ok) return;

  const uint32_t* tile = reinterpret_cast<const uint32_t*>(dw_smem);
  const int row_words = phys_iw << wcl;
  const uint32_t* base = tile + (rb * R * SH) * row_words + ((sx * (SPAN + 1)) << wcl) + lane_c;
  // input rows above / below the tensor are all zero (the conv's vertical padding): a whole feature map per tile
  // (the recognizer: 7 rows, 11 with the 5x5 halo) spends a sixth of its multiply-adds on them unless they are skipped
  // Static stride-1 variants cover the whole height (in.h == RB * R, padding K / 2: checked on the host), so which of a
  // row block's rows are padding is known at compile time for RB == 1 and costs nothing; otherwise a warp-uniform test.
  constexpr bool kStaticRows = STATIC && SH == 1 && RB_T == 1;
  const int vy0 = -(iy0 + rb * R * SH), vy1 = a.in.h - (iy0 + rb * R * SH);  // valid rows of this thread's window
#pragma unroll
  for (int iy = 0; iy < IHR; ++iy) {
    if (!STATIC && iy > (nrows - 1) * SH + K - 1) break;  // below the last row that matters
    if (kStaticRows) {
      if (iy < K / 2 || iy >= K / 2 + R) continue;
    } else if (iy < vy0 || iy >= vy1) {
      continue;                                            // warp-uniform
    }
    uint32_t win[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) win[j] = base[iy * row_words + ((j + j / SPAN) << wcl)];
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      if ((iy - ky) % SH != 0) continue;
      const int r = (iy - ky) / SH;
      if (r < 0 || r >= R) continue;
      if (STATIC || r < nrows) {
#pragma unroll
        for (int s = 0; s < kS; ++s)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const uint32_t x = win[s * SW + kx];
            if (WF32) {
              const float2 xf = __half22float2(*reinterpret_cast<const __half2*>(&x));
              acc[r][s][0] = fmaf(xf.x, wf[ky * K + kx].x, acc[r][s][0]);
              acc[r][s][1] = fmaf(xf.y, wf[ky * K + kx].y, acc[r][s][1]);
            } else {
              fhfma2(x, wq[ky * K + kx], acc[r][s][0], acc[r][s][1]);
            }
          }
      }
    }
  }

  // ---- epilogue: y = s2 * act(acc) + t2, zero beyond the logical channels / the row's valid width
  const int vwn = a.vw ? min(a.vw[n], a.out.w) : a.out.w;
  const bool c_lo = ch < a.out.c, c_hi = ch + 1 < a.out.c;
  const uint32_t keep = (c_lo ? 0x0000ffffu : 0u) | (c_hi ? 0xffff0000u : 0u);
  uint32_t cmask[kS];   // per column: keep mask, or 0 beyond the row's valid width
  bool cstore[kS];
#pragma unroll
  for (int s = 0; s < kS; ++s) { cmask[s] = col0 + s < vwn ? keep : 0u; cstore[s] = col0 + s < a.out.w; }
  // 32-bit element offsets from the (warp-uniform) tensor base: the host only takes this kernel for tensors < 2^31 elements
  uint32_t off = uint32_t(((n * a.out.h + row0) * a.out.w + col0) * a.out.pitch + ch);
  const uint32_t off_row = uint32_t(a.out.w * a.out.pitch);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r >= nrows) break;
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      float v0 = acc[r][s][0], v1 = acc[r][s][1];
      if (a.act == 2) {  // x * relu6(x + 3) / 6 == x * saturate(x / 6 + 0.5): one FFMA.SAT + one FMUL
        v0 *= __saturatef(fmaf(v0, 1.f / 6.f, 0.5f));
        v1 *= __saturatef(fmaf(v1, 1.f / 6.f, 0.5f));
      } else if (a.act == 1) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      const __half2 h = __floats2half2_rn(fmaf(a.s2, v0, a.t2), fmaf(a.s2, v1, a.t2));
      if (cstore[s])
        *reinterpret_cast<uint32_t*>(a.out.p + (off + uint32_t(s * a.out.pitch))) = *reinterpret_cast<const uint32_t*>(&h) & cmask[s];
    }
    off += off_row;
  }
}

template <int K, int SH, int SW, int R, bool WF32, int RB_T, int WCL_T>
void dw_launch(DwArgs& a, cudaStream_t s) {
  if (RB_T > 0) { a.rbn = RB_T; a.wc_log2 = WCL_T; }
  const int wc = 1 << a.wc_log2;
  a.chunks = (((a.out.c + 7) & ~7) + 2 * wc - 1) / (2 * wc);
  a.tw = kS * (kDwThreads / wc) / a.rbn;
  a.iht = (a.rbn * R - 1) * SH + K;
  a.iw = (a.tw - 1) * SW + K;
  a.phys_iw = a.iw + a.iw / (kS * SW) + 1;
  a.tiles_x = (a.out.w + a.tw - 1) / a.tw;
  a.tiles_y = (a.out.h + a.rbn * R - 1) / (a.rbn * R);
  const size_t smem = size_t(a.iht) * a.phys_iw * wc * 4;
  auto kern = dwconv_tile_kernel<K, SH, SW, R, WF32, RB_T, WCL_T>;
  if (smem > 48 * 1024) {
    static int done_dev[64] = {0};  // per instantiation; the attribute lives in the device context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done_dev[dev]) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      done_dev[dev] = 1;
    }
  }
  const long blocks = long(a.out.n) * a.tiles_y * a.tiles_x * a.chunks;
  launch_k(kern, dim3(unsigned(blocks)), dim3(kDwThreads), smem, s, a);
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant of the static full-height kernels (recognizer: fp16 taps, 64-channel chunks, one row block, column
// stride 1).  A CTA stays on ONE channel chunk (blockIdx.y) and walks over (image, column tile) pairs: the K*K weights,
// the bias and all index arithmetic are set up once, and the halo tile is double-buffered so that tile i+1 streams in
// (cp.async) while tile i is multiplied -- the load -> barrier -> compute bubble of the one-tile-per-CTA kernel is gone.
template <int K, int SH, int R>
__global__ void __launch_bounds__(kDwThreads, 3) dwconv_persist_kernel(const DwArgs a, const int n_tiles) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dw_smem[];
  constexpr int SW = 1, TW = 16, WIN = (kS - 1) * SW + K, IHR = (R - 1) * SH + K, IW = (TW - 1) * SW + K;
  constexpr uint32_t kTileBytes = IHR * IW * 128;
  const int c0 = blockIdx.y * 64;
  const int c8 = (a.in.c + 7) & ~7;
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(dw_smem));
  const long row_stride = long(a.in.w) * a.in.pitch;
  const int vy0 = a.ph, vy1 = a.ph + a.in.h;   // tile rows that hold real input rows (the rest is the conv's zero padding)

  auto load_tile = [&](int tile, int buf) {
    const int tx = tile % a.tiles_x, n = tile / a.tiles_x;
    const int piece = threadIdx.x & 7;
    const bool c_ok = c0 + piece * 8 < c8;
    const __half* img = a.in.p + long(n) * a.in.h * row_stride + c0 + piece * 8;
    const int ix0 = tx * TW * SW - a.pw;
#pragma unroll
    for (int pass = 0; pass < (IW + 15) / 16; ++pass) {
      const int x = (threadIdx.x >> 3) + 16 * pass;
      if (x < IW) {
        const int gx = ix0 + x;
        const bool x_ok = c_ok && gx >= 0 && gx < a.in.w;
        const __half* src = img + long(gx) * a.in.pitch;      // input row 0
        uint32_t dst = sbase + buf * kTileBytes + (uint32_t(vy0 * IW + x) << 7) + piece * 16;
        for (int gy = 0; gy < a.in.h && vy0 + gy < IHR; ++gy, src += row_stride, dst += IW * 128)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(x_ok ? src : a.in.p), "r"(x_ok ? 16 : 0)
                       : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int tile = blockIdx.x;
  if (tile < n_tiles) load_tile(tile, 0);

  const int lane_c = threadIdx.x & 31, sx = threadIdx.x >> 5;
  const int ch = c0 + 2 * lane_c;
  const bool ch_ok = ch < c8;
  const int chc = ch_ok ? ch : 0;
  uint32_t wq[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wq[t] = __ldg(reinterpret_cast<const uint32_t*>(a.wh + long(t) * a.cp + chc));
  const float2 bias = __ldg(reinterpret_cast<const float2*>(a.wb + long(K * K) * a.cp + chc));
  const bool c_lo = ch < a.out.c, c_hi = ch + 1 < a.out.c;
  const uint32_t keep = (c_lo ? 0x0000ffffu : 0u) | (c_hi ? 0xffff0000u : 0u);
  const uint32_t off_row = uint32_t(a.out.w * a.out.pitch);
  constexpr bool kStaticRows = SH == 1;  // in.h == R, padding K / 2 (host-checked): padding rows known at compile time

  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int next = tile + gridDim.x;
    if (next < n_tiles) {
      load_tile(next, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int tx = tile % a.tiles_x, n = tile / a.tiles_x;
    const int col0 = tx * TW + sx * kS;
    if (ch_ok && col0 < a.out.w) {
      float acc[R][kS][2];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int s_ = 0; s_ < kS; ++s_) { acc[r][s_][0] = bias.x; acc[r][s_][1] = bias.y; }
      const uint32_t* base = reinterpret_cast<const uint32_t*>(dw_smem + (it & 1) * kTileBytes) + sx * kS * SW * 32 + lane_c;
#pragma unroll
      for (int iy = 0; iy < IHR; ++iy) {
        if (kStaticRows) {
          if (iy < K / 2 || iy >= K / 2 + R) continue;
        } else if (iy < vy0 || iy >= vy1) {
          continue;
        }
        uint32_t win[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) win[j] = base[(iy * IW + j) * 32];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          if ((iy - ky) % SH != 0) continue;
          const int r = (iy - ky) / SH;
          if (r < 0 || r >= R) continue;
#pragma unroll
          for (int s_ = 0; s_ < kS; ++s_)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) fhfma2(win[s_ * SW + kx], wq[ky * K + kx], acc[r][s_][0], acc[r][s_][1]);
        }
      }
      const int vwn = a.vw ? min(a.vw[n], a.out.w) : a.out.w;
      uint32_t cmask[kS];
      bool cstore[kS];
#pragma unroll
      for (int s_ = 0; s_ < kS; ++s_) { cmask[s_] = col0 + s_ < vwn ? keep : 0u; cstore[s_] = col0 + s_ < a.out.w; }
      uint32_t off = uint32_t((n * a.out.h * a.out.w + col0) * a.out.pitch + ch);
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int s_ = 0; s_ < kS; ++s_) {
          float v0 = acc[r][s_][0], v1 = acc[r][s_][1];
          if (a.act == 2) {
            v0 *= __saturatef(fmaf(v0, 1.f / 6.f, 0.5f));
            v1 *= __saturatef(fmaf(v1, 1.f / 6.f, 0.5f));
          } else if (a.act == 1) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          const __half2 h = __floats2half2_rn(fmaf(a.s2, v0, a.t2), fmaf(a.s2, v1, a.t2));
          if (cstore[s_])
            *reinterpret_cast<uint32_t*>(a.out.p + (off + uint32_t(s_ * a.out.pitch))) = *reinterpret_cast<const uint32_t*>(&h) & cmask[s_];
        }
        off += off_row;
      }
    }
    __syncthreads();  // everyone is done with this buffer before the load of tile it+2 overwrites it
  }
}

template <int K, int SH, int R>
void dw_persist_launch(DwArgs& a, cudaStream_t s) {
  constexpr int IHR = (R - 1) * SH + K, IW = 15 + K;
  a.chunks = (((a.out.c + 7) & ~7) + 63) / 64;
  a.tiles_x = (a.out.w + 15) / 16;
  const size_t smem = 2 * size_t(IHR) * IW * 128;
  auto kern = dwconv_persist_kernel<K, SH, R>;
  if (smem > 48 * 1024) {
    static int done_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done_dev[dev]) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      done_dev[dev] = 1;
    }
  }
  const int n_tiles = a.out.n * a.tiles_x;
  const int per_chunk = std::max(1, std::min(n_tiles, (148 * 4 + a.chunks - 1) / a.chunks));
  launch_k(kern, dim3(dim3(unsigned(per_chunk), unsigned(a.chunks))), dim3(kDwThreads), smem, s, a, n_tiles);
}

// fraction of the tile's lanes that hold real channels x real columns, for a chunk of 2^(l2+1) channels
double lane_fill(int c8, int w, int l2, int rbn) {
  const int cc = 2 << l2, wc = 1 << l2;
  const int tw = kS * (kDwThreads / wc) / rbn;
  return double(c8) / double((c8 + cc - 1) / cc * cc) * double(w) / double((w + tw - 1) / tw * tw);
}

}  // namespace

bool launch_dwconv_tile(const TV& in, const TV& out, const float* wb, const __half* wh, const ConvGeom& g, const Epi& e,
                        cudaStream_t s, const int* vw) {
  static const bool disabled = getenv("B200OCR_OLD_DWCONV") != nullptr;
  static const bool no_static = getenv("B200OCR_DWCONV_GENERIC") != nullptr;
  if (disabled) return false;
  if (e.res || g.kh != g.kw || (g.kh != 3 && g.kh != 5)) return false;
  if (g.sh < 1 || g.sh > 2 || g.sw < 1 || g.sw > 2) return false;
  if (in.pitch % 8 || out.pitch % 8 || (reinterpret_cast<uintptr_t>(in.p) & 15) || (reinterpret_cast<uintptr_t>(out.p) & 3)) return false;
  if (e.act < 0 || e.act > 2) return false;
  if (g.cout_pad % 2) return false;
  if (double(out.n) * out.h * out.w * out.pitch >= 2147483648.0) return false;
  const bool f32 = wh == nullptr;
  const int K = g.kh;
  DwArgs a{};
  a.in = in; a.out = out; a.wb = wb; a.wh = wh; a.cp = g.cout_pad; a.ph = g.ph; a.pw = g.pw;
  a.act = e.act; a.s2 = e.s2; a.t2 = e.t2; a.vw = vw;
  const int c8 = (out.c + 7) & ~7;
#define DW_CASE(K_, SH_, SW_, R_, F_, RB_, WCL_) \
  { dw_launch<K_, SH_, SW_, R_, F_, RB_, WCL_>(a, s); return true; }
  // ---- fully static variants for the recognizer's feature-map heights (14 / 7 / 4 / 2 at rec_img_h = 28):
  // R rows per thread = the whole height, 64-channel chunks (16 for the two narrow 3x3 layers)
  static const bool no_persist = getenv("B200OCR_DWCONV_NO_PERSIST") != nullptr;
  if (!f32 && !no_static && !no_persist && g.ph == K / 2 && g.pw == K / 2 && g.sw == 1 && (g.sh != 1 || in.h == out.h) &&
      lane_fill(c8, out.w, 5, 1) >= 0.8) {
    if (K == 5 && g.sh == 1 && out.h == 7) { dw_persist_launch<5, 1, 7>(a, s); return true; }
    if (K == 5 && g.sh == 1 && out.h == 4) { dw_persist_launch<5, 1, 4>(a, s); return true; }
    if (K == 5 && g.sh == 1 && out.h == 2) { dw_persist_launch<5, 1, 2>(a, s); return true; }
    if (K == 5 && g.sh == 2 && out.h == 4) { dw_persist_launch<5, 2, 4>(a, s); return true; }
    if (K == 5 && g.sh == 2 && out.h == 2) { dw_persist_launch<5, 2, 2>(a, s); return true; }
    if (K == 3 && g.sh == 1 && out.h == 7) { dw_persist_launch<3, 1, 7>(a, s); return true; }
    if (K == 3 && g.sh == 2 && out.h == 7) { dw_persist_launch<3, 2, 7>(a, s); return true; }
  }
  if (!f32 && !no_static && g.ph == K / 2 && g.pw == K / 2 && (g.sh != 1 || in.h == out.h) && lane_fill(c8, out.w, 5, 1) >= 0.8) {
    if (K == 5 && g.sh == 1 && g.sw == 1 && out.h == 7) DW_CASE(5, 1, 1, 7, false, 1, 5)
    if (K == 5 && g.sh == 1 && g.sw == 1 && out.h == 4) DW_CASE(5, 1, 1, 4, false, 1, 5)
    if (K == 5 && g.sh == 1 && g.sw == 1 && out.h == 2) DW_CASE(5, 1, 1, 2, false, 1, 5)
    if (K == 5 && g.sh == 2 && g.sw == 1 && out.h == 4) DW_CASE(5, 2, 1, 4, false, 1, 5)
    if (K == 5 && g.sh == 2 && g.sw == 1 && out.h == 2) DW_CASE(5, 2, 1, 2, false, 1, 5)
    if (K == 3 && g.sh == 1 && g.sw == 1 && out.h == 14) DW_CASE(3, 1, 1, 7, false, 2, 5)
    if (K == 3 && g.sh == 1 && g.sw == 1 && out.h == 7) DW_CASE(3, 1, 1, 7, false, 1, 5)
    if (K == 3 && g.sh == 2 && g.sw == 1 && out.h == 7) DW_CASE(3, 2, 1, 7, false, 1, 5)
    if (K == 3 && g.sh == 1 && g.sw == 2 && out.h == 7) DW_CASE(3, 1, 2, 7, false, 1, 5)
  }
  if (!f32 && !no_static && g.ph == 1 && g.pw == 1 && in.h == out.h && K == 3 && g.sh == 1 && g.sw == 1 && out.h == 14 &&
      lane_fill(c8, out.w, 3, 2) >= 0.8)
    DW_CASE(3, 1, 1, 7, false, 2, 3)
  // ---- generic variants: rows per thread 8 (4 with fp32 weights or a vertical stride: register budget), chunk =
  // the one that wastes the fewest lanes (channels x columns), the larger one when it is within 3 %
  const int R = f32 || g.sh == 2 ? 4 : 8;
  a.rbn = out.h > R ? 2 : 1;
  double best = -1;
  for (int l2 = 2; l2 <= 5; ++l2) {
    const double fill = lane_fill(c8, out.w, l2, a.rbn);
    if (fill >= best - 0.03) { best = std::max(best, fill); a.wc_log2 = l2; }
  }
#define DW_GEN(K_, SH_, SW_, R_, F_) \
  if (K == K_ && g.sh == SH_ && g.sw == SW_ && R == R_ && f32 == F_) DW_CASE(K_, SH_, SW_, R_, F_, 0, -1)
  // detector / classifier: fp32 weights, strides (1,1), (2,2), (2,1)
  DW_GEN(3, 1, 1, 4, true) DW_GEN(3, 2, 2, 4, true) DW_GEN(5, 1, 1, 4, true) DW_GEN(5, 2, 2, 4, true)
  DW_GEN(3, 2, 1, 4, true) DW_GEN(5, 2, 1, 4, true)
  // recognizer: fp16 weights, strides (1,1), (2,1), (1,2)
  DW_GEN(3, 1, 1, 8, false) DW_GEN(3, 2, 1, 4, false) DW_GEN(3, 1, 2, 8, false)
  DW_GEN(5, 1, 1, 8, false) DW_GEN(5, 2, 1, 4, false)
#undef DW_GEN
#undef DW_CASE
  return false;
}

}  // namespace b200ocr
