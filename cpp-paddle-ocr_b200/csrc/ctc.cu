// CTC greedy-decode collapse (reference src/ocr_rec.cpp:97-128), one warp per text line.
// The arg-max / max-probability part of the decode is fused into the CTC head kernels (the [N,T,6625]
// softmax is never materialised); this kernel applies the blank / repeat rule with warp ballots and
// accumulates the score in time order in fp32, exactly like the reference's sequential loop.
#include "kernels.h"
#include "pdl.h"

namespace b200ocr {

namespace {

__global__ void __launch_bounds__(128)
ctc_collapse_kernel(const int* __restrict__ idx, const float* __restrict__ prob, int n, int T,
                    int* __restrict__ out_idx, int* __restrict__ out_len, float* __restrict__ out_score) {
  pdl_trigger();
  pdl_wait();
  const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (line >= n) return;
  const int* a = idx + long(line) * T;
  const float* p = prob + long(line) * T;
  int* o = out_idx + long(line) * T;
  int count = 0;
  float score = 0.f;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const int cur = t < T ? a[t] : 0;
    const int prev = (t > 0 && t < T) ? a[t - 1] : 0;
    // emit when idx > 0 && !(t > 0 && idx == last_index)
    const bool keep = t < T && cur > 0 && !(t > 0 && cur == prev);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) o[count + __popc(m & ((1u << lane) - 1))] = cur;
    // score += max_value in time order (fp32, sequential like the reference)
    if (lane == 0) {
      unsigned mm = m;
      while (mm) {
        const int b = __ffs(mm) - 1;
        score += p[t0 + b];
        mm &= mm - 1;
      }
    }
    count += __popc(m);
  }
  if (lane == 0) {
    out_len[line] = count;
    out_score[line] = count ? score / float(count) : 0.f;  // count == 0: NaN in the reference -> slot left at 0
  }
}

}  // namespace

void launch_ctc_collapse(const int* idx, const float* prob, int n, int T, int* out_idx, int* out_len,
                         float* out_score, cudaStream_t s) {
  const int warps_per_block = 4;
  launch_k(ctc_collapse_kernel, dim3((n + warps_per_block - 1) / warps_per_block), dim3(128), 0, s, idx, prob, n, T, out_idx, out_len,
                                                                                out_score);
}

}  // namespace b200ocr
