// tcgen05 CTC head: logits = feat[M,120] . W[6625,120]^T + b fused with the softmax-max / arg-max of the greedy
// decode (reference src/ocr_rec.cpp:97-121 runs Utility::argmax and std::max_element over the materialised
// [N,T,6625] softmax).  The [M,6625] logits only ever exist in TMEM:
//
//   CTA = 128 tokens.  A (128 x 128 fp16, K padded 120 -> 128 by TMA zero fill) is loaded once; the class
//   dimension streams through in tiles of 256 classes: TMA -> 2-stage smem ring -> one tcgen05.mma chain
//   (M=128, N=256, K=16 x 8) per tile into one of two 256-column TMEM accumulators, while the four epilogue
//   warps drain the other one (tcgen05.ld) into a per-token running (max, first arg-max, sum exp).
//   Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issue, 2..17 = epilogue (a token = one TMEM lane; the four
//   warps of a lane quadrant split the column groups and merge their (max, arg-max, sum) through shared memory).
#include "kernels.h"
#include "pdl.h"

#include <cuda.h>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace b200ocr {

namespace {

constexpr int kTileN = 256;
constexpr int kStages = 2;
// The epilogue (one exponential per token and class) is a dependent chain per 16-column group: 16 epilogue warps
// (4 per TMEM lane quadrant, each taking every 4th column group) keep all four schedulers and the MUFU busy.
constexpr int kCtcEpiWarps = 16;
constexpr int kCtcThreads = 64 + 32 * kCtcEpiWarps;
constexpr int kABytes = 128 * 128 * 2;      // 128 tokens x 128 channels fp16 (two 64-channel swizzle atoms)
constexpr int kBBytes = kTileN * 128 * 2;   // 256 classes x 128 channels

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // K-major, 128B swizzle, SBO 1024
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CtcArgs {
  long rows;       // tokens
  int T;           // tokens per sequence (ragged mask)
  int ncls_pad;    // classes incl. padding (bias of padded classes = -30000)
  int ntiles;
  const float* bias;   // log2 domain: bias * log2(e)
  const int* vw;
  int* idx;
  float* prob;
};

__global__ void __launch_bounds__(kCtcThreads)
ctc_head_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const CtcArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* b_full = a_full + 1;           // [kStages]
  uint64_t* b_empty = b_full + kStages;    // [kStages]
  uint64_t* t_full = b_empty + kStages;    // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  const uint32_t tiles_base = (smem_u32(smem_raw) + 128 + 1023) & ~1023u;
  uint8_t* tiles = smem_raw + (tiles_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row0 = long(blockIdx.x) * 128;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(a_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 32 * kCtcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, kABytes);
      tma_load_2d(&tmA, a_full, tiles, 0, int(row0));
      tma_load_2d(&tmA, a_full, tiles + kABytes / 2, 64, int(row0));
      for (int t = 0; t < a.ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&b_empty[s], ((t / kStages) & 1) ^ 1);
        uint8_t* sb = tiles + kABytes + size_t(s) * kBBytes;
        mbar_expect_tx(&b_full[s], kBBytes);
        tma_load_2d(&tmB, &b_full[s], sb, 0, t * kTileN);
        tma_load_2d(&tmB, &b_full[s], sb + kBBytes / 2, 64, t * kTileN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (uint32_t(kTileN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
      mbar_wait(a_full, 0);
      for (int t = 0; t < a.ntiles; ++t) {
        const int s = t % kStages, buf = t & 1;
        mbar_wait(&t_empty[buf], ((t >> 1) & 1) ^ 1);
        mbar_wait(&b_full[s], (t / kStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tiles_base, sb = tiles_base + kABytes + uint32_t(s) * kBBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // K chunk k (16 channels): atoms of 64 channels are kABytes/2 (A) or kBBytes/2 (B) apart
          const uint32_t ao = (k >> 2) * (kABytes / 2) + (k & 3) * 32;
          const uint32_t bo = (k >> 2) * (kBBytes / 2) + (k & 3) * 32;
          umma_f16(tmem_base + buf * kTileN, umma_desc(sa + ao), umma_desc(sb + bo), idesc, k != 0);
        }
        umma_commit(&b_empty[s]);
        umma_commit(&t_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;             // TMEM lane quadrant (hardware rule: warp id % 4)
    const int cg = (warp - 2) >> 2;     // this warp's share of the 16-column groups
    float mx = -FLT_MAX, sum = 0.f;
    int am = 0x7fffffff;
    for (int t = 0; t < a.ntiles; ++t) {
      const int buf = t & 1;
      mbar_wait(&t_full[buf], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + buf * kTileN;
      const int cbase = t * kTileN;
      const int ncol = min(kTileN, a.ncls_pad - cbase);
      for (int col = cg * 16; col < ncol; col += 16 * (kCtcEpiWarps / 4)) {
        uint32_t v[16];
        tmem_ld16(trow + col, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // log2 domain: f = logit * log2(e) (bias pre-scaled on the host) in one FFMA per element; the order of the
        // logits is unchanged, and exp(logit - max) = ex2(f - fmax) is a bare MUFU.EX2 (ex2.approx.ftz: none of the
        // scale / denormal guards __expf wraps around it -- they were 5 of the ~10 instructions per element)
        const float4* bp = reinterpret_cast<const float4*>(a.bias + cbase + col);
        float f[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 b = __ldg(bp + g);
          f[4 * g] = fmaf(__uint_as_float(v[4 * g]), kLog2e, b.x);
          f[4 * g + 1] = fmaf(__uint_as_float(v[4 * g + 1]), kLog2e, b.y);
          f[4 * g + 2] = fmaf(__uint_as_float(v[4 * g + 2]), kLog2e, b.z);
          f[4 * g + 3] = fmaf(__uint_as_float(v[4 * g + 3]), kLog2e, b.w);
        }
        // group maximum first (no transcendental), one rescale when the running maximum moves (rare), then 16
        // independent exponentials
        float m16 = f[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) m16 = fmaxf(m16, f[i]);
        if (m16 > mx) {
          sum *= ex2_ftz(mx - m16);
          mx = m16;
          int first = 15;
#pragma unroll
          for (int i = 14; i >= 0; --i) first = (f[i] == m16) ? i : first;  // first maximum wins (Utility::argmax)
          am = cbase + col + first;
        }
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          e0 += ex2_ftz(f[i] - mx); e1 += ex2_ftz(f[i + 1] - mx); e2 += ex2_ftz(f[i + 2] - mx); e3 += ex2_ftz(f[i + 3] - mx);
        }
        sum += (e0 + e1) + (e2 + e3);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&t_empty[buf]);
    }
    // combine the column shares of a token: the pipeline's shared memory is idle by now (all tiles consumed)
    float* pm = reinterpret_cast<float*>(tiles);                 // [4][128] running maxima
    float* ps = pm + 4 * 128;                                     // [4][128] sums
    int* pa = reinterpret_cast<int*>(ps + 4 * 128);               // [4][128] arg-max
    const int rloc = q * 32 + lane;
    pm[cg * 128 + rloc] = mx; ps[cg * 128 + rloc] = sum; pa[cg * 128 + rloc] = am;
    asm volatile("bar.sync 1, %0;" ::"r"(32 * kCtcEpiWarps) : "memory");
    if (cg == 0) {
      const long row = row0 + rloc;
      float M = pm[rloc];
#pragma unroll
      for (int p = 1; p < 4; ++p) M = fmaxf(M, pm[p * 128 + rloc]);
      float total = 0.f;
      int best = 0x7fffffff;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float mp = pm[p * 128 + rloc];
        total += ps[p * 128 + rloc] * ex2_ftz(mp - M);
        if (mp == M) best = min(best, pa[p * 128 + rloc]);       // first maximum wins across the shares too
      }
      if (row < a.rows) {
        const bool pad = a.vw && int(row % a.T) >= a.vw[row / a.T];
        a.idx[row] = pad ? 0 : best;
        a.prob[row] = pad ? 0.f : 1.f / total;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
      return;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
  return fn;
}
void encode2d(CUtensorMap* tm, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t stride1_bytes, cuuint32_t b0,
              cuuint32_t b1) {
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with " + std::to_string(int(r)));
}

}  // namespace

bool ctc_tc_eligible(const TV& feat, int cin_pad) {
  return cin_pad == 128 && feat.c <= 128 && feat.c > 64 && (feat.pitch % 8) == 0 &&
         (reinterpret_cast<uintptr_t>(feat.p) & 15) == 0;
}

void launch_ctc_head_tc(const TV& feat, const __half* w, const float* bias, int cin_pad, int ncls, int ncls_pad, int* idx,
                        float* prob, cudaStream_t s, const int* vw) {
  (void)ncls;
  const long rows = long(feat.n) * feat.h * feat.w;
  CUtensorMap tmA, tmB;
  encode2d(&tmA, feat.p, cuuint64_t((feat.c + 7) & ~7), cuuint64_t(rows), cuuint64_t(feat.pitch) * 2, 64, 128);
  encode2d(&tmB, w, cuuint64_t(cin_pad), cuuint64_t(ncls_pad), cuuint64_t(cin_pad) * 2, 64, kTileN);
  CtcArgs a;
  a.rows = rows; a.T = feat.h * feat.w; a.ncls_pad = ncls_pad;
  a.ntiles = (ncls_pad + kTileN - 1) / kTileN;
  a.bias = bias; a.vw = vw; a.idx = idx; a.prob = prob;
  const size_t smem = kABytes + size_t(kStages) * kBBytes + 1024 + 128;
  cudaFuncSetAttribute(ctc_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  launch_k(ctc_head_tc_kernel, dim3(unsigned((rows + 127) / 128)), dim3(kCtcThreads), smem, s, tmA, tmB, a);
}

}  // namespace b200ocr
