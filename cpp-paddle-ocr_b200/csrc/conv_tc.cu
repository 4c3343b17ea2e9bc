// tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x BN channels] (TMEM, fp32) += A[128 x 64] (smem, fp16) * B[BN x 64]^T (smem, fp16)
//
// A tiles are fetched by TMA straight from the NHWC activation view as a 4-D box
// {64 channels, TW, TH, TN}; a filter tap (ky,kx) is just a shifted box origin and the conv's zero
// padding is TMA out-of-bounds fill, so no im2col buffer ever exists.  B tiles are rows of the
// folded fp16 filter [cout_pad][taps][cin_pad].  Both land in 128B-swizzled K-major smem; one
// elected thread issues tcgen05.mma (M=128, N=BN, K=16); accumulators live in TMEM and four
// epilogue warps read them back with tcgen05.ld, apply bias/activation/affine/residual and store
// NHWC fp16.  Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issue, 2..5 = epilogue.
#include "kernels.h"
#include "pdl.h"

#include <cuda.h>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <stdexcept>

namespace b200ocr {

namespace {

constexpr int kStagesMax = 4;
constexpr int kStagesMaxP = 8;  // persistent kernel: the A ring is as deep as shared memory allows (bytes in flight per SM)
// The epilogue is a chain of dependent instructions per 16-column group; with one warp per scheduler every latency
// is exposed.  16 epilogue warps (4 per TMEM lane quadrant, each taking every 4th column group) keep all four
// schedulers busy.  Block = 2 control warps + kEpiWarps.
constexpr int kEpiWarps = 16;
constexpr int kThreadsTc = 64 + 32 * kEpiWarps;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 fp16

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// K-major, 128B swizzle: 8-row atoms of 1024 B (SBO = 1024), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMA with the descriptors split into their constant high word and the low word that carries the start
// address: advancing 32 bytes along K is `lo + 2`, so the single issuing thread spends one add per operand and MMA
// instead of rebuilding both 64-bit descriptors (the issue loop, not the tensor pipe, bounds the narrow-N layers:
// ~170 clk per tcgen05.mma measured on the 96 -> 24 3x3 layers).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024, descriptor version 1, 128B swizzle
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- 2-CTA cluster helpers (filter multicast, conv_tc_kernel<ACT, 1>)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one L2 read, delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <int ACT>
__device__ __forceinline__ float act_fn(float v, float a, float b) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));  // x * relu6(x + 3) / 6: one FFMA.SAT + one FMUL
  if (ACT == 3) return v / (1.f + __expf(-v));
  if (ACT == 4) return __saturatef(fmaf(v, a, b));
  if (ACT == 5) return 1.f / (1.f + __expf(-v));
  return v;
}

struct ConvTcArgs {
  // tile geometry over the output view (for 1x1 convs the view is flattened to n=1,h=1,w=P)
  int on, oh, ow;          // output dims as tiled
  int tw, th, tn;          // box: tw*th*tn <= 128 rows
  int tiles_x, tiles_y;    // tiles_n = pixel tiles / (tiles_x*tiles_y)
  int n_tiles;             // N tiles (non-persistent kernel: interleaved in blockIdx.x)
  // halo mode (persistent kernel, kh*kw > 1): ONE TMA box of (tw+kw-1) x (th+kh-1) pixels per K chunk instead of one
  // shifted box per filter tap.  The MMA's 128 rows walk the PADDED row pitch row_w = tw+kw-1 (outputs at x >= tw are
  // discarded), so filter tap (ky,kx) is the same tile read (ky*row_w + kx) rows further down: the descriptor's start
  // address moves by that many 128-byte rows.  Measured on B200 (tools/diag_halo.py): the 128B-swizzle XOR is taken from
  // the absolute shared-memory address bits, so a start address that is not 1024-byte aligned needs NO base-offset
  // field (setting it to (addr >> 7) & 7, as the field's description suggests, gives wrong data).
  int halo;                // 0 / 1
  int row_w;               // pixels per tile row as the MMA sees them: tw, or tw+kw-1 in halo mode
  int a_stage;             // bytes per A ring stage
  int kh, kw, ph, pw;
  int cin, cin_pad;        // logical input channels; weight row stride per tap
  int bn;                  // N tile (multiple of 16, <= 256)
  int tmem_cols;           // power of two >= max(32, bn)
  int stages;
  int cout;                // logical output channels
  __half* out;
  int out_pitch;
  const float* bias;
  Epi epi;
  // ragged batches: valid output width per image; pixel -> (image, x) through the un-flattened dims
  const int* vw;
  int mask_w, mask_hw;
};

// Drain one accumulator tile: the calling warp owns TMEM lanes 32*(warp%4) .. +31 (row r of the tile = lane of D).
// `stage` != nullptr: the fp16 results go to shared memory instead (rows of 128 B per 64-channel chunk, 16-byte pieces
// XOR-swizzled with the row like CU_TENSOR_MAP_SWIZZLE_128B expects) and leave with one TMA store per chunk.
template <int ACT>
__device__ __forceinline__ void epilogue_tile(const ConvTcArgs& a, const float* __restrict__ bias_blk, uint32_t tmem_base,
                                              int x0, int y0, int n0, int nblk, int warp, int lane,
                                              uint8_t* stage = nullptr) {
  const int q = warp & 3;               // TMEM lane quadrant this warp may read (hardware rule: warp id % 4)
  const int cgrp = (warp - 2) >> 2;     // which share of the 16-column groups
  const int r = q * 32 + lane;
  const int rows = a.row_w * a.th * a.tn;
  const int tw_ = r % a.row_w, th_ = (r / a.row_w) % a.th, tn_ = r / (a.row_w * a.th);
  const int x = x0 + tw_, y = y0 + th_, n = n0 + tn_;
  const bool valid = r < rows && tw_ < a.tw && x < a.ow && y < a.oh && n < a.on;
  const long pix = (long(n) * a.oh + y) * a.ow + x;
  const bool masked = valid && a.vw && int(pix % a.mask_w) >= a.vw[pix / a.mask_hw];
  const int c8lim = (a.cout + 7) & ~7;
  const bool st32 = (a.out_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 31) == 0;  // 32-byte aligned groups
  const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
  for (int col = cgrp * 16; col < a.bn; col += 16 * (kEpiWarps / 4)) {
    uint32_t v[16];
    tmem_ld16(trow + col, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int c0 = nblk * a.bn + col;
    if ((!valid && !stage) || c0 >= c8lim) continue;  // staged tiles: rows outside the tensor are clipped by the store
    float f[16];
    {
      const float4* bp = reinterpret_cast<const float4*>(bias_blk + col);  // bias is padded to a multiple of 16
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 b = bp[g];
        f[4 * g] = __uint_as_float(v[4 * g]) + b.x;
        f[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + b.y;
        f[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + b.z;
        f[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = a.epi.s2 * act_fn<ACT>(f[i], a.epi.a, a.epi.b) + a.epi.t2;
    const bool second = c0 + 8 < c8lim;
    if (a.epi.res && valid) {
      const __half* rp = a.epi.res + pix * a.epi.res_pitch + c0;
      uint4 r0 = *reinterpret_cast<const uint4*>(rp);
      const __half2* h = reinterpret_cast<const __half2*>(&r0);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 p = __half22float2(h[i]); f[2 * i] += p.x; f[2 * i + 1] += p.y; }
      if (second) {
        uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
        const __half2* g = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 p = __half22float2(g[i]); f[8 + 2 * i] += p.x; f[8 + 2 * i + 1] += p.y; }
      }
    }
    if (masked) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = 0.f;
    } else if (c0 + 16 > a.cout) {  // only the last channel group of a layer can be partial
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i >= a.cout) f[i] = 0.f;
    }
    __half* op = a.out + pix * a.out_pitch + c0;
    uint4 o0, o1;
    __half2* h0 = reinterpret_cast<__half2*>(&o0);
    __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h0[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      h1[i] = __floats2half2_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
    }
    if (stage) {
      uint8_t* row = stage + size_t(col >> 6) * kATileBytes + r * 128;
      const int j = (col & 63) >> 3;
      *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = o0;
      *reinterpret_cast<uint4*>(row + (((j + 1) ^ (r & 7)) << 4)) = o1;
    } else if (second && st32) {  // one full 32-byte sector per lane instead of two half-sector writes
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(o0.x), "r"(o0.y), "r"(o0.z),
                   "r"(o0.w), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w)
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(op) = o0;
      if (second) *reinterpret_cast<uint4*>(op + 8) = o1;
    }
  }
}

// MC = 1: launched as clusters of two CTAs that work on two neighbouring pixel tiles of the SAME N tile.  Each CTA fetches
// half of every filter tile and multicasts it into both CTAs' shared memory, so a filter byte crosses the L2 -> SM fabric
// once per 256 pixels instead of once per 128.  For the layers whose filter block does not fit next to the A ring (the
// recognizer's 480 -> 480 convolutions re-stream 230 KB of filter per pixel tile and N tile and are bound by exactly that
// traffic, profiles/r02_notes.md section 10) this removes a third of the bytes a CTA pulls per K step (46 -> 31 KB).
// A stage is free again when BOTH CTAs' MMAs have read it (the peer writes into it): the empty barriers count two
// arrivals and every tcgen05.commit is multicast to the pair.
template <int ACT, int MC>
__global__ void __launch_bounds__(kThreadsTc)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvTcArgs a) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  // [barriers | tmem ptr] then 1024-aligned tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + kStagesMax;
  uint64_t* tmem_full = empty_bar + kStagesMax;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const uint32_t tiles_base = (smem_u32(smem_raw) + 128 + 1023) & ~1023u;
  uint8_t* tiles = smem_raw + (tiles_base - smem_u32(smem_raw));
  const int stage_bytes = kATileBytes + a.bn * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates.  The N tiles of one pixel tile are neighbours in launch order (blockIdx.x = tile * n_tiles +
  // nblk): they run at the same time and the second one finds the activation tile in L2.
  int t, nblk;
  uint32_t rank = 0;
  if (MC) {  // pair p = blockIdx.x / 2 -> (pixel tile pair, N tile); the pair's CTAs take pixel tiles 2j and 2j + 1
    rank = cluster_ctarank();
    const int p = int(blockIdx.x >> 1);
    t = (p / a.n_tiles) * 2 + int(rank);
    nblk = p % a.n_tiles;
  } else {
    t = blockIdx.x / a.n_tiles;
    nblk = blockIdx.x - t * a.n_tiles;  // N tile index
  }
  const int tx = t % a.tiles_x; t /= a.tiles_x;
  const int ty = t % a.tiles_y; t /= a.tiles_y;
  const int x0 = tx * a.tw, y0 = ty * a.th, n0 = t * a.tn;

  const int kchunks = (a.cin + 63) >> 6;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC ? 2 : 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(uint32_t(a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (MC) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above (barriers, TMEM, tensor maps) overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = uint32_t(a.tw * a.th * a.tn * 128 + a.bn * 128);
      const int half_rows = a.bn >> 1;
      int it = 0;
      for (int ky = 0; ky < a.kh; ++ky)
        for (int kx = 0; kx < a.kw; ++kx)
          for (int kc = 0; kc < kchunks; ++kc, ++it) {
            const int s = it % a.stages;
            const uint32_t ph = uint32_t(it / a.stages) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* sa = tiles + size_t(s) * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            mbar_expect_tx(&full_bar[s], tx_bytes);
            tma_load_4d(&tmA, &full_bar[s], sa, kc * 64, x0 + kx - a.pw, y0 + ky - a.ph, n0);
            if (MC)   // tmB's box is half an N tile here: this CTA's half goes to both CTAs
              tma_load_2d_mc(&tmB, &full_bar[s], sb + size_t(rank) * half_rows * 128, (ky * a.kw + kx) * a.cin_pad + kc * 64,
                             nblk * a.bn + int(rank) * half_rows, uint16_t(3));
            else
              tma_load_2d(&tmB, &full_bar[s], sb, (ky * a.kw + kx) * a.cin_pad + kc * 64, nblk * a.bn);
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (uint32_t(a.bn >> 3) << 17) | (uint32_t(128 >> 4) << 24);
      int it = 0;
      for (int tap = 0; tap < a.kh * a.kw; ++tap)
        for (int kc = 0; kc < kchunks; ++kc, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = uint32_t(it / a.stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = tiles_base + uint32_t(s) * stage_bytes;
          const uint32_t sb = sa + kATileBytes;
          const int rem = a.cin - kc * 64;
          const int k16 = rem >= 64 ? 4 : (rem + 15) >> 4;
          const uint32_t alo = umma_desc_lo(sa), blo = umma_desc_lo(sb);
          for (int k = 0; k < k16; ++k) umma_f16_lo(tmem_base, alo + 2 * k, blo + 2 * k, idesc, (it | k) != 0);
          if (MC) umma_commit_mc(&empty_bar[s], uint16_t(3));  // ... in both CTAs: the peer's producer writes this stage too
          else umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
      umma_commit(tmem_full);
    }
  } else {
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue_tile<ACT>(a, a.bias + nblk * a.bn, tmem_base, x0, y0, n0, nblk, warp, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's last commits arrive on OUR barriers: stay resident until it is done too
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(uint32_t(a.tmem_cols))
                 : "memory");
  }
}

// ---------------------------------------------------------------- persistent variant
// For layers whose whole filter block (all taps x all K chunks of one N tile) fits in shared memory next to the A
// ring: one CTA per SM slot keeps the filter resident, walks over its share of the pixel tiles and double-buffers
// the accumulator in TMEM, so the TMA loads of tile i+1, the MMAs of tile i+1 and the epilogue of tile i overlap
// and the per-CTA set-up (barriers, TMEM allocation, descriptor fetch, filter load) is paid once.
template <int ACT>
__global__ void __launch_bounds__(kThreadsTc)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const ConvTcArgs a, const int n_mtiles,
                       const int kt /* taps * K chunks */, const int tma_store) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_full = b_full + 1;
  uint64_t* a_empty = a_full + kStagesMaxP;
  uint64_t* t_full = a_empty + kStagesMaxP;  // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem_raw + 256);  // this N tile's folded bias (<= 256 floats), read by every tile
  const uint32_t tiles_base = (smem_u32(smem_raw) + 256 + 1024 + 1023) & ~1023u;
  uint8_t* tiles = smem_raw + (tiles_base - smem_u32(smem_raw));
  const int b_chunk = a.bn * 128;
  const uint32_t a_ring = uint32_t(kt) * b_chunk;  // offset of the A ring behind the resident filter
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = blockIdx.y;
  const int kchunks = (a.cin + 63) >> 6;
  for (int i = threadIdx.x; i < a.bn; i += blockDim.x)
    sbias[i] = nblk * a.bn + i < ((a.cout + 15) & ~15) ? a.bias[nblk * a.bn + i] : 0.f;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(b_full, 1);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(uint32_t(2 * a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_full, uint32_t(kt) * b_chunk);
      for (int tap = 0, it = 0; tap < a.kh * a.kw; ++tap)
        for (int kc = 0; kc < kchunks; ++kc, ++it)
          tma_load_2d(&tmB, b_full, tiles + size_t(it) * b_chunk, tap * a.cin_pad + kc * 64, nblk * a.bn);
      pdl_wait();  // the filter block is on its way; activations only after the previous kernel has finished
      const uint32_t a_bytes = a.halo ? uint32_t(a.row_w * (a.th + a.kh - 1) * 128) : uint32_t(a.tw * a.th * a.tn * 128);
      int g = 0;
      for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % a.tiles_x; t /= a.tiles_x;
        const int ty = t % a.tiles_y; t /= a.tiles_y;
        const int x0 = tx * a.tw, y0 = ty * a.th, n0 = t * a.tn;
        if (a.halo) {
          for (int kc = 0; kc < kchunks; ++kc, ++g) {
            const int s = g % a.stages;
            mbar_wait(&a_empty[s], (uint32_t(g / a.stages) & 1u) ^ 1u);
            mbar_expect_tx(&a_full[s], a_bytes);
            tma_load_4d(&tmA, &a_full[s], tiles + a_ring + size_t(s) * a.a_stage, kc * 64, x0 - a.pw, y0 - a.ph, n0);
          }
          continue;
        }
        for (int ky = 0; ky < a.kh; ++ky)
          for (int kx = 0; kx < a.kw; ++kx)
            for (int kc = 0; kc < kchunks; ++kc, ++g) {
              const int s = g % a.stages;
              mbar_wait(&a_empty[s], (uint32_t(g / a.stages) & 1u) ^ 1u);
              mbar_expect_tx(&a_full[s], a_bytes);
              tma_load_4d(&tmA, &a_full[s], tiles + a_ring + size_t(s) * a.a_stage, kc * 64, x0 + kx - a.pw,
                          y0 + ky - a.ph, n0);
            }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (uint32_t(a.bn >> 3) << 17) | (uint32_t(128 >> 4) << 24);
      mbar_wait(b_full, 0);
      int g = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&t_empty[buf], (uint32_t(lt >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dst = tmem_base + uint32_t(buf * a.tmem_cols);
        if (a.halo) {
          for (int kc = 0; kc < kchunks; ++kc, ++g) {
            const int s = g % a.stages;
            mbar_wait(&a_full[s], uint32_t(g / a.stages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = tiles_base + a_ring + uint32_t(s) * a.a_stage;
            const int rem = a.cin - kc * 64;
            const int k16 = rem >= 64 ? 4 : (rem + 15) >> 4;
            for (int ky = 0; ky < a.kh; ++ky)
              for (int kx = 0; kx < a.kw; ++kx) {
                const uint32_t sb = tiles_base + uint32_t((ky * a.kw + kx) * kchunks + kc) * b_chunk;
                const uint32_t sat = sa + uint32_t(ky * a.row_w + kx) * 128;
                const uint32_t alo = umma_desc_lo(sat), blo = umma_desc_lo(sb);
                for (int k = 0; k < k16; ++k) umma_f16_lo(dst, alo + 2 * k, blo + 2 * k, idesc, (kc | ky | kx | k) != 0);
              }
            umma_commit(&a_empty[s]);
          }
          umma_commit(&t_full[buf]);
          continue;
        }
        for (int tap = 0, it = 0; tap < a.kh * a.kw; ++tap)
          for (int kc = 0; kc < kchunks; ++kc, ++it, ++g) {
            const int s = g % a.stages;
            mbar_wait(&a_full[s], uint32_t(g / a.stages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = tiles_base + a_ring + uint32_t(s) * a.a_stage;
            const uint32_t sb = tiles_base + uint32_t(it) * b_chunk;
            const int rem = a.cin - kc * 64;
            const int k16 = rem >= 64 ? 4 : (rem + 15) >> 4;
            const uint32_t alo = umma_desc_lo(sa), blo = umma_desc_lo(sb);
            for (int k = 0; k < k16; ++k) umma_f16_lo(dst, alo + 2 * k, blo + 2 * k, idesc, (it | k) != 0);
            umma_commit(&a_empty[s]);
          }
        umma_commit(&t_full[buf]);
      }
    }
  } else {
    pdl_wait();  // residual reads / output writes: only after the previous kernel has finished
    // output staging tile behind the A ring (only with tma_store): ceil(bn / 64) chunks of 128 rows x 128 B
    uint8_t* stage = tma_store ? tiles + a_ring + size_t(a.stages) * a.a_stage : nullptr;
    const bool leader = warp == 2 && lane == 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++lt) {
      int t = tile;
      const int tx = t % a.tiles_x; t /= a.tiles_x;
      const int ty = t % a.tiles_y; t /= a.tiles_y;
      const int buf = lt & 1;
      if (stage) {  // the previous tile's stores must have finished READING the staging tile
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
      }
      mbar_wait(&t_full[buf], uint32_t(lt >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue_tile<ACT>(a, sbias, tmem_base + uint32_t(buf * a.tmem_cols), tx * a.tw, ty * a.th, t * a.tn, nblk, warp, lane,
                         stage);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&t_empty[buf]);
      if (stage) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
        if (leader) {
          const int c8lim = (a.cout + 7) & ~7;
          for (int c = 0; c * 64 < a.bn && nblk * a.bn + c * 64 < c8lim; ++c)
            tma_store_4d(&tmC, stage + size_t(c) * kATileBytes, nblk * a.bn + c * 64, tx * a.tw, ty * a.th, t * a.tn);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (stage && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(uint32_t(2 * a.tmem_cols))
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !p) {
      return;  // reported below: the library never exits the host process (capi_util.h)
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
  return fn;
}

void encode(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, cuuint32_t(rank), base, dims, strides_bytes,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed with %d (rank %d dims %llu %llu box %u %u)", int(r),
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    throw std::runtime_error(msg);
  }
}

}  // namespace

struct ConvTcPlanImpl {
  CUtensorMap tmA, tmB, tmC;
  bool tma_store = false;  // persistent kernel: results leave through a staged TMA store (no residual input)
  ConvTcArgs args;
  dim3 grid;
  size_t smem;
  bool persistent = false;
  bool multicast = false;  // non-persistent kernel launched as 2-CTA clusters sharing the filter stream
  int n_mtiles = 0, kt = 0;
};

bool conv_tc_eligible(const TV& in, const TV& out, const ConvGeom& g) {
  if (g.sh != 1 || g.sw != 1) return false;
  if (in.c < 8) return false;  // stems (C_in = 3) stay on the CUDA-core kernel
  if (in.pitch % 8 || out.pitch % 8) return false;
  if ((reinterpret_cast<uintptr_t>(in.p) & 15) || (reinterpret_cast<uintptr_t>(out.p) & 15)) return false;
  if (out.h != in.h + 2 * g.ph - g.kh + 1 || out.w != in.w + 2 * g.pw - g.kw + 1) return false;
  return true;
}

ConvTcPlan make_conv_tc_plan(const TV& in_, const TV& out_, const __half* w, const ConvGeom& g) {
  TV in = in_, out = out_;
  const bool pointwise = g.kh == 1 && g.kw == 1 && g.ph == 0 && g.pw == 0;
  if (pointwise) {  // plain GEMM over all pixels
    const long P = long(in.n) * in.h * in.w;
    in.n = out.n = 1;
    in.h = out.h = 1;
    in.w = out.w = int(P);
  }
  ConvTcArgs a{};
  a.on = out.n; a.oh = out.h; a.ow = out.w;
  // pick the box that wastes the fewest of the 128 MMA rows
  double best = -1;
  for (int tw = 1; tw <= 128 && tw <= out.w; ++tw) {
    int th = 128 / tw;
    if (th > out.h) th = out.h;
    int tn = 1;
    if (tw == out.w && th == out.h) { tn = 128 / (tw * th); if (tn > out.n) tn = out.n; }
    const long tiles = long((out.w + tw - 1) / tw) * ((out.h + th - 1) / th) * ((out.n + tn - 1) / tn);
    const double eff = double(long(out.n) * out.h * out.w) / double(tiles * 128);
    if (eff > best + 1e-9) { best = eff; a.tw = tw; a.th = th; a.tn = tn; }
  }
  a.tiles_x = (out.w + a.tw - 1) / a.tw;
  a.tiles_y = (out.h + a.th - 1) / a.th;
  const int tiles_n = (out.n + a.tn - 1) / a.tn;
  a.kh = g.kh; a.kw = g.kw; a.ph = g.ph; a.pw = g.pw;
  a.cin = in.c; a.cin_pad = g.cin_pad;
  const int n_tiles = (g.cout_pad + 255) / 256;
  a.bn = ((g.cout_pad + n_tiles - 1) / n_tiles + 15) & ~15;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.bn) a.tmem_cols <<= 1;
  const int k_iters = g.kh * g.kw * ((in.c + 63) / 64);
  a.stages = k_iters < kStagesMax ? k_iters : kStagesMax;
  // Two CTAs per SM (one loads / multiplies while the other drains its accumulator) beat one CTA with a deep
  // ring: keep the ring at what fits twice into shared memory (and twice 256 TMEM columns always fit).
  while (a.stages > 2 && size_t(a.stages) * (kATileBytes + a.bn * 128) + 1024 + 128 > 110 * 1024) --a.stages;
  a.cout = out.c;
  a.out = out.p;
  a.out_pitch = out.pitch;
  a.vw = nullptr;
  a.mask_w = out_.w;
  a.mask_hw = out_.h * out_.w;

  auto* impl = new ConvTcPlanImpl();
  const int c8 = (in.c + 7) & ~7;
  cuuint64_t dA[4] = {cuuint64_t(c8), cuuint64_t(in.w), cuuint64_t(in.h), cuuint64_t(in.n)};
  cuuint64_t sA[3] = {cuuint64_t(in.pitch) * 2, cuuint64_t(in.pitch) * 2 * in.w,
                      cuuint64_t(in.pitch) * 2 * in.w * in.h};
  cuuint32_t bA[4] = {64, cuuint32_t(a.tw), cuuint32_t(a.th), cuuint32_t(a.tn)};
  encode(&impl->tmA, in.p, 4, dA, sA, bA);
  const int taps = g.kh * g.kw;
  cuuint64_t dB[2] = {cuuint64_t(taps) * g.cin_pad, cuuint64_t(g.cout_pad)};
  cuuint64_t sB[1] = {cuuint64_t(taps) * g.cin_pad * 2};
  cuuint32_t bB[2] = {64, cuuint32_t(a.bn)};
  encode(&impl->tmB, const_cast<__half*>(w), 2, dB, sB, bB);
  {  // output view with the same tiling as the input box (rows of a tile = tw x th x tn pixels, 64 channels per store)
    const int oc8 = (out.c + 7) & ~7;
    cuuint64_t dC[4] = {cuuint64_t(oc8), cuuint64_t(out.w), cuuint64_t(out.h), cuuint64_t(out.n)};
    cuuint64_t sC[3] = {cuuint64_t(out.pitch) * 2, cuuint64_t(out.pitch) * 2 * out.w, cuuint64_t(out.pitch) * 2 * out.w * out.h};
    encode(&impl->tmC, out.p, 4, dC, sC, bA);
  }
  a.n_tiles = n_tiles;
  a.halo = 0; a.row_w = a.tw; a.a_stage = kATileBytes;
  impl->grid = dim3(unsigned(a.tiles_x * a.tiles_y * tiles_n * n_tiles), 1u);
  impl->smem = size_t(a.stages) * (kATileBytes + a.bn * 128) + 1024 + 128;
  // persistent variant: filter block resident + A ring + double-buffered accumulator
  {
    const int m_tiles = a.tiles_x * a.tiles_y * tiles_n;
    const size_t b_bytes = size_t(k_iters) * a.bn * 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // CTAs per SM: bounded by TMEM (two accumulators each) and by shared memory (resident filter + >= 2 A stages);
    // the A ring then takes what is left: an HBM-bound layer needs as many bytes in flight per SM as it can get
    const size_t stage_bytes = size_t((a.bn + 63) / 64) * kATileBytes;  // output staging tile of the TMA-store epilogue
    size_t fixed = b_bytes + 1024 + 1024 + 256;
    int per_sm = std::max(1, std::min(512 / (2 * a.tmem_cols), 4));
    auto budget = [](int ctas) { return (size_t(227) * 1024 - size_t(ctas) * 1024) / size_t(ctas); };  // 1 KB/CTA is reserved
    while (per_sm > 1 && fixed + 2 * kATileBytes > budget(per_sm)) --per_sm;
    // measured on B200 (profiles/r01_notes.md): with a single staging tile the two extra CTA-wide barriers per tile cost
    // more than the full-line stores save (39 us vs 31 us on the 240 -> 240 layers), so the staged path is opt-in
    static const bool want_tma_store = getenv("B200OCR_TMA_STORE") != nullptr;
    const bool staged = want_tma_store && fixed + stage_bytes + 2 * kATileBytes <= budget(per_sm);
    if (staged) fixed += stage_bytes;
    int st = 0;
    if (fixed + 2 * kATileBytes <= budget(per_sm))
      st = int(std::min<size_t>(kStagesMaxP, (budget(per_sm) - fixed) / kATileBytes));
    static const int st_cap = getenv("B200OCR_CONV_STAGES") ? atoi(getenv("B200OCR_CONV_STAGES")) : kStagesMaxP;
    st = std::min(st, std::max(2, st_cap));
    const size_t need = fixed + size_t(st) * kATileBytes;
    static const int min_tiles_x3 = getenv("B200OCR_CONV_PERSIST_MIN") ? atoi(getenv("B200OCR_CONV_PERSIST_MIN")) : 6;  // persistent from 2 tiles per SM on (measured: fewer tiles run faster one tile per CTA)
    if (2 * a.tmem_cols <= 512 && st >= 2 && 3 * m_tiles >= min_tiles_x3 * sms && !getenv("B200OCR_NO_PERSISTENT_CONV")) {
      const int ctas = std::min(m_tiles, sms * per_sm);
      impl->persistent = true;
      impl->tma_store = staged;
      impl->n_mtiles = m_tiles;
      impl->kt = k_iters;
      a.stages = st;
      impl->grid = dim3(unsigned(ctas), unsigned(n_tiles));
      impl->smem = need;
    }
    // ---- halo mode for kxk filters: one input box per K chunk instead of one per filter tap (see ConvTcArgs::halo).
    static const bool want_halo = getenv("B200OCR_CONV_HALO") != nullptr;
    if (impl->persistent && !staged && want_halo && taps > 1 && g.kw <= 5 && g.kh <= 5) {
      // output tile = th rows x tw columns with (tw + kw - 1) * th <= 128 MMA rows; the best cover of the output wins
      int btw = 0, bth = 0;
      double bcov = 0;
      for (int tw = 1; tw + g.kw - 1 <= 128 && tw <= out.w; ++tw) {
        const int pwid = tw + g.kw - 1;
        int th = std::min(128 / pwid, out.h);
        if (th < 1 || pwid > 256 || th + g.kh - 1 > 256) continue;
        const long tiles = long((out.w + tw - 1) / tw) * ((out.h + th - 1) / th) * out.n;
        const double cov = double(long(out.n) * out.h * out.w) / double(tiles * 128);
        if (cov > bcov + 1e-9) { bcov = cov; btw = tw; bth = th; }
      }
      const int pwid = btw + g.kw - 1;
      const size_t stage = ((size_t((g.kh - 1) * pwid + g.kw - 1 + 128) * 128) + 1023) & ~size_t(1023);
      const size_t fixed_h = b_bytes + 1024 + 1024 + 256;
      int per_sm_h = std::max(1, std::min(512 / (2 * a.tmem_cols), 4));
      while (per_sm_h > 1 && fixed_h + 2 * stage > budget(per_sm_h)) --per_sm_h;
      const int kch = (in.c + 63) / 64;
      if (btw > 0 && bcov >= 0.5 && fixed_h + 2 * stage <= budget(per_sm_h)) {
        const int st_h = int(std::min<size_t>(kStagesMaxP, (budget(per_sm_h) - fixed_h) / stage));  // as deep as fits (kch stages = one tile)
        a.tw = btw; a.th = bth; a.tn = 1;
        a.tiles_x = (out.w + a.tw - 1) / a.tw;
        a.tiles_y = (out.h + a.th - 1) / a.th;
        a.halo = 1; a.row_w = pwid; a.a_stage = int(stage); a.stages = st_h;
        cuuint32_t bH[4] = {64, cuuint32_t(pwid), cuuint32_t(a.th + g.kh - 1), 1};
        encode(&impl->tmA, in.p, 4, dA, sA, bH);
        impl->n_mtiles = a.tiles_x * a.tiles_y * out.n;
        impl->grid = dim3(unsigned(std::min(impl->n_mtiles, sms * per_sm_h)), unsigned(n_tiles));
        impl->smem = fixed_h + size_t(st_h) * stage;
        impl->tma_store = false;
      }
    }
  }
  // ---- filter multicast for the layers that stay on the non-persistent kernel (see conv_tc_kernel<ACT, 1>)
  {
    static const int mc_min = getenv("B200OCR_CONV_MULTICAST_MIN") ? atoi(getenv("B200OCR_CONV_MULTICAST_MIN")) : 296;
    static const bool mc_off = !(getenv("B200OCR_CONV_MULTICAST") && atoi(getenv("B200OCR_CONV_MULTICAST")) != 0);  // opt-in until measured
    const int m_tiles = a.tiles_x * a.tiles_y * tiles_n;
    if (!impl->persistent && !mc_off && pointwise && m_tiles >= mc_min && (a.bn & 15) == 0 && a.stages >= 2) {
      impl->multicast = true;
      cuuint32_t bH[2] = {64, cuuint32_t(a.bn / 2)};
      encode(&impl->tmB, const_cast<__half*>(w), 2, dB, sB, bH);
      impl->grid = dim3(unsigned(2 * ((m_tiles + 1) / 2) * n_tiles), 1u);
    }
  }
  impl->args = a;
  // function attributes live in the device context: set once per device
  {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 64 || !done[dev]) {
    cudaFuncSetAttribute(conv_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (dev < 64) done[dev] = true;
    }
  }
  ConvTcPlan p;
  p.impl = impl;
  return p;
}

void free_conv_tc_plan(ConvTcPlan* p) {
  delete p->impl;
  p->impl = nullptr;
}

void launch_conv_tc(const ConvTcPlan& p, const float* bias, const Epi& e, cudaStream_t s, const int* vw) {
  ConvTcArgs a = p.impl->args;
  a.bias = bias;
  a.epi = e;
  a.vw = vw;
  if (p.impl->persistent) {
    const int nm = p.impl->n_mtiles, kt = p.impl->kt;
    const int ts = p.impl->tma_store && e.res == nullptr ? 1 : 0;
    switch (e.act) {
      case 1: launch_k(conv_tc_persist_kernel<1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
      case 2: launch_k(conv_tc_persist_kernel<2>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
      case 3: launch_k(conv_tc_persist_kernel<3>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
      case 4: launch_k(conv_tc_persist_kernel<4>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
      case 5: launch_k(conv_tc_persist_kernel<5>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
      default: launch_k(conv_tc_persist_kernel<0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, p.impl->tmC, a, nm, kt, ts); break;
    }
    return;
  }
  if (p.impl->multicast) {
    switch (e.act) {
      case 1: launch_k_cluster(conv_tc_kernel<1, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
      case 2: launch_k_cluster(conv_tc_kernel<2, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
      case 3: launch_k_cluster(conv_tc_kernel<3, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
      case 4: launch_k_cluster(conv_tc_kernel<4, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
      case 5: launch_k_cluster(conv_tc_kernel<5, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
      default: launch_k_cluster(conv_tc_kernel<0, 1>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, 2, p.impl->tmA, p.impl->tmB, a); break;
    }
    return;
  }
  switch (e.act) {
    case 1: launch_k(conv_tc_kernel<1, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
    case 2: launch_k(conv_tc_kernel<2, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
    case 3: launch_k(conv_tc_kernel<3, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
    case 4: launch_k(conv_tc_kernel<4, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
    case 5: launch_k(conv_tc_kernel<5, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
    default: launch_k(conv_tc_kernel<0, 0>, dim3(p.impl->grid), dim3(kThreadsTc), p.impl->smem, s, p.impl->tmA, p.impl->tmB, a); break;
  }
}

}  // namespace b200ocr
