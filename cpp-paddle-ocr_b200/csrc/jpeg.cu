// Device half of the JPEG decoder (see jpeg.h): Huffman decoding, ISLOW inverse DCT, fancy upsampling + YCbCr -> BGR.
// Integer arithmetic throughout, restating libjpeg's jdhuff.c / jidctint.c / jdsample.c / jdcolor.c the way
// oracle/jpeg_decode.py does; results are bit-identical to cv2.imdecode.
#include "jpeg.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "stages.h"

namespace b200ocr {

namespace {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------ entropy decoding
// MSB-first bit reader over an entropy-coded segment: 0xFF00 is un-stuffed, any other marker feeds zero bits without
// advancing (libjpeg's behaviour at the end of a segment; oracle/jpeg_decode.py::_Bits).
constexpr int kRingBytes = 16384;   // shared-memory staging ring of the single-interval path: 4 chunks of 4 KB
constexpr int kChunkBytes = 4096;
constexpr int kRingChunks = kRingBytes / kChunkBytes;

// RING = false: bytes come straight from global memory (many restart intervals per image, one thread each).
// RING = true : the image is ONE interval, decoded by one thread; the CTA's second warp streams the bytes through a
//               shared-memory ring ahead of it, so the decoder's dependent chain sees ~30-cycle shared-memory loads
//               instead of L2 round trips.
template <bool RING>
struct BitReader {
  const uint8_t* d;       // RING: the ring (shared memory); else the image's bytes (global, 4-byte aligned, padded)
  int p, end;
  uint64_t acc;
  int n;
  int pad;                // zero bits appended after the end of the data (they sit at the tail of acc)
  // ring bookkeeping
  volatile int* prod;     // chunks staged so far (written by the staging warp)
  volatile int* cons;     // chunk the decoder is reading (written here)
  int ready;              // bytes known to be staged
  int chunk;
  __device__ __forceinline__ void need(int upto) {  // make sure bytes [.., upto) are in the ring
    if (RING) {
      if (upto > ready) {
        int c;
        while ((c = *prod) * kChunkBytes < upto) {}
        __threadfence_block();
        ready = c * kChunkBytes;
      }
      const int ch = p / kChunkBytes;
      if (ch != chunk) { chunk = ch; *cons = ch; }
    }
  }
  __device__ __forceinline__ uint32_t word(int idx) const {
    if (RING) return reinterpret_cast<const uint32_t*>(d)[idx & (kRingBytes / 4 - 1)];
    return __ldg(reinterpret_cast<const uint32_t*>(d) + idx);
  }
  __device__ __forceinline__ uint32_t byte(int pos) const { return RING ? d[pos & (kRingBytes - 1)] : d[pos]; }
  __device__ __forceinline__ void fill() {
    if (n > 32) return;
    need(p + 8);
    if (p + 4 <= end) {
      // four bytes at once when none of them is 0xFF (no stuffing, no marker): two aligned words, funnel-shifted
      const uint32_t le = __funnelshift_r(word(p >> 2), word((p >> 2) + 1), (p & 3) * 8);
      const uint32_t inv = ~le;
      if (((inv - 0x01010101u) & ~inv & 0x80808080u) == 0) {
        acc = (acc << 32) | __byte_perm(le, 0, 0x0123);
        n += 32;
        p += 4;
        return;
      }
    }
    while (n <= 56) {
      uint32_t b = 0;
      if (p < end) {
        need(p + 2);
        b = byte(p);
        if (b == 0xFF) {
          const uint32_t nx = p + 1 < end ? byte(p + 1) : 0xD9u;
          if (nx == 0) p += 2; else { b = 0; pad += 8; }
        } else {
          ++p;
        }
      } else {
        pad += 8;
      }
      acc = (acc << 8) | b;
      n += 8;
    }
  }
  __device__ __forceinline__ uint32_t peek(int k) const { return uint32_t(acc >> (n - k)) & ((1u << k) - 1u); }
  __device__ __forceinline__ void skip(int k) { n -= k; }
  __device__ __forceinline__ int receive_extend(int s) {   // F.2.2.1 / F.2.2.4
    if (s == 0) return 0;
    const int v = int(peek(s));
    skip(s);
    return v >= (1 << (s - 1)) ? v : v - (1 << s) + 1;
  }
};

template <class BR>
__device__ __forceinline__ int decode_symbol(BR& br, const JpegHuffLut& t) {
  const uint32_t look = br.peek(9);
  const uint32_t e = t.fast[look];
  if (e) {
    br.skip(int(e >> 8));
    return int(e & 255u);
  }
  // codes longer than 9 bits (F.2.2.3)
  for (int len = 10; len <= 16; ++len) {
    const int code = int(br.peek(len));
    if (code <= t.maxcode[len]) {
      br.skip(len);
      return t.vals[(code + t.valoff[len]) & 255];
    }
  }
  br.skip(16);  // invalid code: libjpeg warns and returns 0
  return 0;
}

constexpr int kHuffThreads = 64;

// Decodes one restart interval (JPEG F.2.2): DC difference + AC run/size pairs per block, blocks in MCU order.
template <bool RING>
__device__ __forceinline__ void decode_interval(BitReader<RING>& br, const JpegSeg sg, const JpegHuffLut* lut,
                                                const long long* s_off, const int* s_bw, const int* s_hs, const int* s_vs,
                                                int ncomp, int mcux, int16_t* __restrict__ coef) {
  br.p = sg.begin; br.end = sg.end; br.acc = 0; br.n = 0; br.pad = 0;
  int pred[3] = {0, 0, 0};
  int my = sg.mcu0 / mcux, mx = sg.mcu0 - my * mcux;
  for (int m = 0; m < sg.nmcu; ++m) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c >= ncomp) break;
      const JpegHuffLut& tdc = lut[2 * c];
      const JpegHuffLut& tac = lut[2 * c + 1];
      const int hs = s_hs[c], vs = s_vs[c], bw = s_bw[c];
      for (int by = 0; by < vs; ++by)
        for (int bx = 0; bx < hs; ++bx) {
          int16_t* blk = coef + (s_off[c] + (long long)(my * vs + by) * bw + (mx * hs + bx)) * 64;
          br.fill();
          const int t = decode_symbol(br, tdc) & 15;
          br.fill();
          pred[c] += br.receive_extend(t);
          blk[0] = int16_t(pred[c]);
          int k = 1;
          while (k < 64) {
            br.fill();
            const int rs = decode_symbol(br, tac);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
              if (r != 15) break;
              k += 16;
              continue;
            }
            k += r;
            const int v = br.receive_extend(s);
            if (k < 64) blk[c_zigzag[k]] = int16_t(v);
            ++k;
          }
        }
    }
    if (++mx == mcux) { mx = 0; ++my; }
    // Out of data (jdhuff.c `insufficient_data`): the MCU in which the decoder ran past the end of the segment is
    // completed with zero bits, the MCUs after it are not decoded at all (their coefficients stay zero: mid grey).
    if (br.pad > br.n) break;
  }
}

// One CTA per image: its six look-up tables are staged in shared memory, then every thread decodes restart intervals
// (thread t takes intervals t, t + 64, ...).  Files without restart markers have one interval: one thread decodes
// the whole image from a shared-memory ring the CTA's second warp keeps filled -- latency-bound per image and meant
// to run beside other work (a batch of images = that many decoding threads on as many SMs).
__global__ void __launch_bounds__(kHuffThreads)
jpeg_huffman_kernel(const JpegImage* __restrict__ imgs, const JpegSeg* __restrict__ segs, const uint8_t* __restrict__ bytes,
                    const int* __restrict__ flags, int16_t* __restrict__ coef) {
  __shared__ JpegHuffLut lut[6];
  const JpegImage& im = imgs[blockIdx.x];
  if (im.seg_count == 1) {
    // decoded by jpeg_parallel_huffman_kernel unless that kernel asked for the sequential decoder (truncated / corrupt
    // stream: libjpeg's out-of-data behaviour is only restated here); its partial output is cleared first
    if (!flags[im.index]) return;
    uint4* z = reinterpret_cast<uint4*>(coef + im.block_begin * 64);
    for (long long i = threadIdx.x; i < im.nblocks * 8; i += kHuffThreads) z[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(im.lut);
    uint32_t* dst = reinterpret_cast<uint32_t*>(lut);
    const int words = int(sizeof(JpegHuffLut) * 2 * im.ncomp / 4);
    for (int i = threadIdx.x; i < words; i += kHuffThreads) dst[i] = src[i];
  }
  __shared__ long long s_off[3];
  __shared__ int s_bw[3], s_hs[3], s_vs[3];
  if (threadIdx.x < im.ncomp) {
    s_off[threadIdx.x] = im.comp[threadIdx.x].coef_off;
    s_bw[threadIdx.x] = im.comp[threadIdx.x].bw;
    s_hs[threadIdx.x] = im.comp[threadIdx.x].hs;
    s_vs[threadIdx.x] = im.comp[threadIdx.x].vs;
  }
  __syncthreads();
  const int ncomp = im.ncomp, mcux = im.mcux, nseg = im.seg_count, seg0 = im.seg_begin;
  const uint8_t* data = bytes + im.data_off;
  if (nseg == 1) {
    // one interval: thread 0 decodes from the ring, warp 1 stages the bytes, the rest of warp 0 has nothing to do
    __shared__ __align__(16) uint8_t ring[kRingBytes];
    __shared__ volatile int prod, cons;
    if (threadIdx.x == 0) { prod = 0; cons = 0; }
    __syncthreads();
    const JpegSeg sg = segs[seg0];
    const int nchunks = (sg.end + 8 + kChunkBytes - 1) / kChunkBytes;
    if (threadIdx.x >= 32) {
      const int lane = threadIdx.x - 32;
      for (int c = 0; c < nchunks; ++c) {
        while (c - cons >= kRingChunks) {}   // the slot's previous chunk (c - kRingChunks) must be behind the decoder
        const uint4* src = reinterpret_cast<const uint4*>(data + size_t(c) * kChunkBytes);
        uint4* dst = reinterpret_cast<uint4*>(ring + (c % kRingChunks) * kChunkBytes);
#pragma unroll
        for (int i = 0; i < kChunkBytes / 16 / 32; ++i) dst[lane + 32 * i] = __ldg(src + lane + 32 * i);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) prod = c + 1;
      }
    } else if (threadIdx.x == 0) {
      BitReader<true> br;
      br.d = ring; br.prod = &prod; br.cons = &cons; br.ready = 0; br.chunk = 0;
      decode_interval<true>(br, sg, lut, s_off, s_bw, s_hs, s_vs, ncomp, mcux, coef);
      cons = 1 << 28;  // release the staging warp if it is still waiting for ring space
    }
    return;
  }
  for (int si = threadIdx.x; si < nseg; si += kHuffThreads) {
    BitReader<false> br;
    br.d = data; br.prod = nullptr; br.cons = nullptr; br.ready = 0; br.chunk = 0;
    decode_interval<false>(br, segs[seg0 + si], lut, s_off, s_bw, s_hs, s_vs, ncomp, mcux, coef);
  }
}

// ------------------------------------------------------------------------------------------------ parallel entropy decode
// Files without restart markers are one long Huffman stream.  JPEG's codes self-synchronise: a decoder that starts at an
// arbitrary bit with the wrong state falls into step with the true symbol sequence after a few hundred bits.  Following
// Weissenberger & Schmidt (ICPP 2018) the stream is cut into subsequences of 1024 bits:
//   phase 0  the CTA removes the 0xFF00 byte stuffing (block-wide stream compaction), so that bit positions are plain;
//   phase 1  every subsequence is decoded from its first bit with the state "start of an MCU";
//   phase 2  repeat: subsequence i is decoded again from the end position / state its LEFT neighbour reached in the
//            previous round, until no end state changes (subsequence 0 starts from the true state, so the fixed point
//            is the true decode; a subsequence whose input did not change is not decoded again);
//   phase 3  a prefix sum over the blocks completed per subsequence gives every subsequence its first block number; one
//            more pass writes the coefficients.  DC values are written as DIFFERENCES; jpeg_dc_scan_kernel turns them
//            into predictions with a per-component prefix sum.
// One CTA (256 threads) per image.  If the blocks found do not add up to the frame (truncated / corrupt file) the image
// is handed to the sequential decoder, which restates libjpeg's out-of-data behaviour.
constexpr int kParThreads = 256;
constexpr int kSubBits = kJpegSubBytes * 8;

struct ParReader {   // MSB-first reader over the un-stuffed copy (zero padded); position = bits consumed so far
  const uint32_t* w;
  int p4;
  uint64_t acc;
  int n;
  __device__ __forceinline__ void init(const uint8_t* clean, int pos) {
    w = reinterpret_cast<const uint32_t*>(clean);
    p4 = pos >> 5;
    acc = __byte_perm(w[p4], 0, 0x0123);
    ++p4;
    n = 32 - (pos & 31);
  }
  __device__ __forceinline__ void fill() {
    if (n <= 32) { acc = (acc << 32) | __byte_perm(w[p4], 0, 0x0123); ++p4; n += 32; }
  }
  __device__ __forceinline__ int pos() const { return p4 * 32 - n; }
  __device__ __forceinline__ uint32_t peek(int k) const { return uint32_t(acc >> (n - k)) & ((1u << k) - 1u); }
  __device__ __forceinline__ void skip(int k) { n -= k; }
  __device__ __forceinline__ int receive_extend(int s) {
    if (s == 0) return 0;
    const int v = int(peek(s));
    skip(s);
    return v >= (1 << (s - 1)) ? v : v - (1 << s) + 1;
  }
};

struct ParGeom {   // per image, in shared memory
  long long off[3];
  int bw[3], hs[3], vs[3];
  int comp[12], by[12], bx[12];
  int bpm, mcux;
  long long total_blocks;
};

// Decodes from the reader's position up to bit `end` (symbols that START before `end`).  State: block-in-MCU `blk`,
// next coefficient index `k` (0 = a DC symbol is due).  WRITE: coefficients go to the block with sequence number q.
template <bool WRITE>
__device__ __forceinline__ void run_sub(ParReader& br, int end, int& blk, int& k, int& nblk, long long q, const JpegHuffLut* lut,
                                        const ParGeom& g, int16_t* __restrict__ coef) {
  int16_t* bp = nullptr;
  auto locate = [&](long long qq) {
    const long long mcu = qq / g.bpm;
    const int b = int(qq - mcu * g.bpm), c = g.comp[b];
    const int my = int(mcu / g.mcux), mx = int(mcu - (long long)my * g.mcux);
    return coef + (g.off[c] + (long long)(my * g.vs[c] + g.by[b]) * g.bw[c] + (mx * g.hs[c] + g.bx[b])) * 64;
  };
  if (WRITE) { if (q >= g.total_blocks) return; bp = locate(q); }
  while (br.pos() < end) {
    const int c = g.comp[blk];
    br.fill();
    if (k == 0) {
      const int t = decode_symbol(br, lut[2 * c]) & 15;
      br.fill();
      const int v = br.receive_extend(t);
      if (WRITE) bp[0] = int16_t(v);   // the DC DIFFERENCE
      k = 1;
    } else {
      const int rs = decode_symbol(br, lut[2 * c + 1]);
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        k = r == 15 ? k + 16 : 64;
      } else {
        k += r;
        const int v = br.receive_extend(s);
        if (WRITE && k < 64) bp[c_zigzag[k]] = int16_t(v);
        ++k;
      }
    }
    if (k >= 64) {
      k = 0;
      ++nblk;
      blk = blk + 1 == g.bpm ? 0 : blk + 1;
      if (WRITE) {
        if (++q >= g.total_blocks) return;
        bp = locate(q);
      }
    }
  }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kParThreads / 32; ++i) {
    const int s = warp_sums[i];
    if (i < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kParThreads)
jpeg_parallel_huffman_kernel(const JpegImage* __restrict__ imgs, const uint8_t* __restrict__ bytes, uint8_t* __restrict__ clean_all,
                             uint4* __restrict__ sub_all, long long n_sub_total, int* __restrict__ flags,
                             int16_t* __restrict__ coef) {
  const JpegImage& im = imgs[blockIdx.x];
  if (im.seg_count != 1) {
    if (threadIdx.x == 0) flags[im.index] = 0;
    return;
  }
  __shared__ JpegHuffLut lut[6];
  __shared__ ParGeom g;
  __shared__ int warp_sums[kParThreads / 32];
  __shared__ int s_carry;
  const int tid = threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(im.lut);
    uint32_t* dst = reinterpret_cast<uint32_t*>(lut);
    const int words = int(sizeof(JpegHuffLut) * 2 * im.ncomp / 4);
    for (int i = tid; i < words; i += kParThreads) dst[i] = src[i];
  }
  if (tid == 0) {
    int b = 0;
    for (int c = 0; c < im.ncomp; ++c) {
      g.off[c] = im.comp[c].coef_off; g.bw[c] = im.comp[c].bw; g.hs[c] = im.comp[c].hs; g.vs[c] = im.comp[c].vs;
      for (int y = 0; y < im.comp[c].vs; ++y)
        for (int x = 0; x < im.comp[c].hs; ++x, ++b) { g.comp[b] = c; g.by[b] = y; g.bx[b] = x; }
    }
    g.bpm = b;
    g.mcux = im.mcux;
    g.total_blocks = (long long)im.mcux * im.mcuy * b;
    s_carry = 0;
  }
  __syncthreads();
  // ---- phase 0: remove byte stuffing (0x00 after 0xFF) and 0xFF fill bytes
  const uint8_t* raw = bytes + im.data_off;
  uint8_t* clean = clean_all + im.data_off;
  const int len = im.data_len;
  for (int base = 0; base < len; base += kParThreads * 16) {
    const int i0 = base + tid * 16;
    uint8_t b[16];
    int cnt = 0;
    unsigned keep = 0;
    if (i0 < len) {
      const uint4 v = *reinterpret_cast<const uint4*>(raw + i0);   // (the batch buffer is padded: reading past len is safe)
      memcpy(b, &v, 16);
      uint8_t prev = i0 > 0 ? raw[i0 - 1] : 0;
      const int m = min(16, len - i0);
      for (int j = 0; j < m; ++j) {
        const bool drop = (prev == 0xFF && b[j] == 0x00) || (b[j] == 0xFF && i0 + j + 1 < len && raw[i0 + j + 1] == 0xFF);
        if (!drop) { keep |= 1u << j; ++cnt; }
        prev = b[j];
      }
    }
    int total;
    const int excl = block_exclusive_scan(cnt, warp_sums, &total);
    uint8_t* o = clean + s_carry + excl;
    for (int j = 0; j < 16; ++j)
      if (keep >> j & 1) *o++ = b[j];
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  const int clen = s_carry;
  for (int i = tid; i < 64; i += kParThreads) clean[clen + i] = 0;   // zero bits after the end, like the sequential reader
  const int total_bits = clen * 8;
  const int nsub = min((total_bits + kSubBits - 1) / kSubBits, im.sub_max);
  uint4* info_a = sub_all + im.sub_begin;
  uint4* info_b = sub_all + n_sub_total + im.sub_begin;
  int* first_blk = reinterpret_cast<int*>(sub_all + 2 * n_sub_total) + im.sub_begin;
  __syncthreads();
  // ---- phase 1
  for (int i = tid; i < nsub; i += kParThreads) {
    ParReader br;
    br.init(clean, i * kSubBits);
    int blk = 0, k = 0, nb = 0;
    run_sub<false>(br, min((i + 1) * kSubBits, total_bits), blk, k, nb, 0, lut, g, coef);
    info_a[i] = make_uint4(unsigned(br.pos()), unsigned(blk | (k << 8)), unsigned(nb), 1u);
  }
  __syncthreads();
  // ---- phase 2
  uint4* cur = info_a;
  uint4* nxt = info_b;
  for (int iter = 0; iter <= nsub; ++iter) {
    int changed = 0;
    for (int i = tid; i < nsub; i += kParThreads) {
      const uint4 mine = cur[i];
      if (i == 0) { nxt[0] = make_uint4(mine.x, mine.y, mine.z, 0u); continue; }
      const uint4 prev = cur[i - 1];
      if (prev.w == 0u) { nxt[i] = make_uint4(mine.x, mine.y, mine.z, 0u); continue; }  // same input as last round
      ParReader br;
      br.init(clean, int(prev.x));
      int blk = int(prev.y & 255u), k = int(prev.y >> 8), nb = 0;
      run_sub<false>(br, min((i + 1) * kSubBits, total_bits), blk, k, nb, 0, lut, g, coef);
      const uint4 now = make_uint4(unsigned(br.pos()), unsigned(blk | (k << 8)), unsigned(nb), 0u);
      const unsigned ch = (now.x != mine.x || now.y != mine.y || now.z != mine.z) ? 1u : 0u;
      nxt[i] = make_uint4(now.x, now.y, now.z, ch);
      changed |= int(ch);
    }
    uint4* t = cur; cur = nxt; nxt = t;
    if (!__syncthreads_or(changed)) break;
  }
  // ---- phase 3: first block number of every subsequence, then the writing pass
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nsub; base += kParThreads) {
    const int i = base + tid;
    const int v = i < nsub ? int(cur[i].z) : 0;
    int total;
    const int excl = block_exclusive_scan(v, warp_sums, &total);
    if (i < nsub) first_blk[i] = s_carry + excl;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  const bool complete = (long long)s_carry >= g.total_blocks;
  if (tid == 0) flags[im.index] = complete ? 0 : 1;
  if (!complete) return;
  for (int i = tid; i < nsub; i += kParThreads) {
    ParReader br;
    int blk = 0, k = 0, nb = 0, start = 0;
    if (i > 0) { const uint4 prev = cur[i - 1]; start = int(prev.x); blk = int(prev.y & 255u); k = int(prev.y >> 8); }
    br.init(clean, start);
    run_sub<true>(br, min((i + 1) * kSubBits, total_bits), blk, k, nb, (long long)first_blk[i], lut, g, coef);
  }
}

// DC prediction of the images the parallel kernel decoded: per component, an inclusive prefix sum over the DC
// differences in scan order (F.2.2.1: DIFF = DC - PRED).  One CTA per (component, image).
__global__ void __launch_bounds__(kParThreads)
jpeg_dc_scan_kernel(const JpegImage* __restrict__ imgs, const int* __restrict__ flags, int16_t* __restrict__ coef) {
  const JpegImage& im = imgs[blockIdx.y];
  const int c = blockIdx.x;
  if (im.seg_count != 1 || flags[im.index] || c >= im.ncomp) return;
  __shared__ int warp_sums[kParThreads / 32];
  __shared__ int s_carry;
  const JpegComp& cp = im.comp[c];
  const int per = cp.hs * cp.vs, mcux = im.mcux;
  const long long n = (long long)im.mcux * im.mcuy * per;
  auto at = [&](long long j) {
    const long long mcu = j / per;
    const int b = int(j - mcu * per), by = b / cp.hs, bx = b - by * cp.hs;
    const int my = int(mcu / mcux), mx = int(mcu - (long long)my * mcux);
    return coef + (cp.coef_off + (long long)(my * cp.vs + by) * cp.bw + (mx * cp.hs + bx)) * 64;
  };
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  constexpr int kPer = 8;
  for (long long base = 0; base < n; base += kParThreads * kPer) {
    const long long j0 = base + (long long)threadIdx.x * kPer;
    int v[kPer], sum = 0;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      v[e] = j0 + e < n ? int(at(j0 + e)[0]) : 0;
      sum += v[e];
      v[e] = sum;
    }
    int total;
    const int excl = block_exclusive_scan(sum, warp_sums, &total) + s_carry;
#pragma unroll
    for (int e = 0; e < kPer; ++e)
      if (j0 + e < n) at(j0 + e)[0] = int16_t(excl + v[e]);
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ inverse DCT
// jidctint.c (JDCT_ISLOW): CONST_BITS 13, PASS1_BITS 2
__device__ __forceinline__ void idct8(const int x[8], int out[8], int shift) {
  constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
                F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
  int z2 = x[2], z3 = x[6];
  int z1 = (z2 + z3) * F0541;
  const int tmp2 = z1 - z3 * F1847;
  const int tmp3 = z1 + z2 * F0765;
  const int tmp0 = (x[0] + x[4]) << 13;
  const int tmp1 = (x[0] - x[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = x[7], t1 = x[5], t2 = x[3], t3 = x[1];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F1175;
  t0 *= F0298; t1 *= F2053; t2 *= F3072; t3 *= F1501;
  z1 *= -F0899; z2 *= -F2562;
  z3 = z3 * -F1961 + z5;
  z4 = z4 * -F0390 + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  out[0] = (tmp10 + t3 + rnd) >> shift; out[7] = (tmp10 - t3 + rnd) >> shift;
  out[1] = (tmp11 + t2 + rnd) >> shift; out[6] = (tmp11 - t2 + rnd) >> shift;
  out[2] = (tmp12 + t1 + rnd) >> shift; out[5] = (tmp12 - t1 + rnd) >> shift;
  out[3] = (tmp13 + t0 + rnd) >> shift; out[4] = (tmp13 - t0 + rnd) >> shift;
}

constexpr int kIdctThreads = 256;  // 32 blocks of 8x8 per CTA, 8 threads each

// image of a batch-wide block number: binary search over the images' first blocks
__device__ __forceinline__ int find_image(const JpegImage* imgs, int n, long long b) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (imgs[mid].block_begin <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kIdctThreads)
jpeg_idct_kernel(const JpegImage* __restrict__ imgs, int n_images, long long total_blocks, const int16_t* __restrict__ coef,
                 uint8_t* __restrict__ planes) {
  __shared__ int ws[kIdctThreads / 8][8][9];
  const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
  const long long b = (long long)blockIdx.x * (kIdctThreads / 8) + g;
  const bool live = b < total_blocks;
  int comp = 0;
  long long local = 0;
  const JpegImage* im = nullptr;
  if (live) {
    im = &imgs[find_image(imgs, n_images, b)];
    local = b - im->block_begin;  // == index into the coefficient buffer relative to comp[0].coef_off
    for (int c = im->ncomp - 1; c > 0; --c)
      if (local >= im->comp[c].coef_off - im->comp[0].coef_off) { comp = c; break; }
  }
  int row[8], v[8];
  if (live) {
    // thread j: row j of the block, dequantised
    const JpegComp& cp = im->comp[comp];
    const int4 raw = *reinterpret_cast<const int4*>(coef + (im->comp[0].coef_off + local) * 64 + j * 8);
    const int16_t* r16 = reinterpret_cast<const int16_t*>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) ws[g][j][k] = int(r16[k]) * int(cp.q[j * 8 + k]);
  }
  __syncwarp();
  if (live) {
    // pass 1: column j
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k] = ws[g][k][j];
    idct8(row, v, 13 - 2);
  }
  __syncwarp();
  if (live) {
#pragma unroll
    for (int k = 0; k < 8; ++k) ws[g][k][j] = v[k];
  }
  __syncwarp();
  if (live) {
    // pass 2: row j
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k] = ws[g][j][k];
    idct8(row, v, 13 + 2 + 3);
    const JpegComp& cp = im->comp[comp];
    const long long cb = local - (cp.coef_off - im->comp[0].coef_off);
    const int by = int(cb / cp.bw), bx = int(cb - (long long)by * cp.bw);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= uint32_t(min(max(v[k] + 128, 0), 255)) << (8 * k);
      hi |= uint32_t(min(max(v[k + 4] + 128, 0), 255)) << (8 * k);
    }
    uint8_t* dst = planes + cp.plane_off + (long long)(by * 8 + j) * (cp.bw * 8) + bx * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------------ upsampling + colour
// jdsample.c fancy upsampling of one chroma sample at full-resolution position (y, x); the component's last real
// row / column is replicated as context.
__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ p, int pitch, int cw, int ch, int fh, int fv, int y, int x) {
  if (fh == 1) return p[(long long)y * pitch + x];
  const int cx = x >> 1;
  const int xo = (x & 1) ? min(cx + 1, cw - 1) : max(cx - 1, 0);
  if (fv == 1) {
    const uint8_t* r = p + (long long)y * pitch;
    return (x & 1) ? (3 * r[cx] + r[xo] + 2) >> 2 : (3 * r[cx] + r[xo] + 1) >> 2;
  }
  const int cy = y >> 1;
  const int yo = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
  const uint8_t* r0 = p + (long long)cy * pitch;
  const uint8_t* r1 = p + (long long)yo * pitch;
  const int cs = 3 * r0[cx] + r1[cx], co = 3 * r0[xo] + r1[xo];
  return (x & 1) ? (3 * cs + co + 7) >> 4 : (3 * cs + co + 8) >> 4;
}

__global__ void __launch_bounds__(256)
jpeg_color_kernel(const JpegImage* __restrict__ imgs, const uint8_t* __restrict__ planes, uint8_t* __restrict__ out) {
  const JpegImage& im = imgs[blockIdx.z];
  const int x = blockIdx.x * 64 + (threadIdx.x & 63);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= im.width || y >= im.height) return;
  const JpegComp& c0 = im.comp[0];
  const int Y = planes[c0.plane_off + (long long)y * (c0.bw * 8) + x];
  int b = Y, g = Y, r = Y;
  if (im.ncomp == 3) {
    const JpegComp& c1 = im.comp[1];
    const JpegComp& c2 = im.comp[2];
    const int fh = im.hmax / c1.hs, fv = im.vmax / c1.vs;
    const int cb = chroma_at(planes + c1.plane_off, c1.bw * 8, c1.cw, c1.ch, fh, fv, y, x) - 128;
    const int cr = chroma_at(planes + c2.plane_off, c2.bw * 8, c2.cw, c2.ch, fh, fv, y, x) - 128;
    // jdcolor.c, SCALEBITS 16: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554
    r = Y + ((91881 * cr + 32768) >> 16);
    b = Y + ((116130 * cb + 32768) >> 16);
    g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    r = min(max(r, 0), 255); g = min(max(g, 0), 255); b = min(max(b, 0), 255);
  }
  uint8_t* o = out + im.out_off + (long long)y * im.out_stride + 3 * x;
  o[0] = uint8_t(b); o[1] = uint8_t(g); o[2] = uint8_t(r);
}

}  // namespace

void launch_jpeg_decode(const JpegImage* imgs_dev, const JpegImage* imgs_host, int n_images, const JpegSeg* segs_dev,
                        int n_segs, const uint8_t* bytes_dev, uint8_t* clean_dev, void* sub_dev, long long n_sub,
                        int16_t* coef_dev, size_t coef_bytes, uint8_t* planes_dev, uint8_t* out_dev, cudaStream_t s) {
  (void)n_segs;
  if (n_images < 1) return;
  // only the non-zero coefficients are written by the entropy decoders
  if (cudaMemsetAsync(coef_dev, 0, coef_bytes, s) != cudaSuccess) throw std::runtime_error("jpeg: cudaMemsetAsync failed");
  uint4* sub = static_cast<uint4*>(sub_dev);
  int* flags = reinterpret_cast<int*>(sub + 3 * n_sub);
  static const bool sequential_only = getenv("B200OCR_JPEG_SEQUENTIAL") != nullptr;  // A/B: one thread per interval only
  bool any_single = false;
  for (int i = 0; i < n_images; ++i) any_single |= imgs_host[i].seg_count == 1;
  if (any_single && !sequential_only) {
    jpeg_parallel_huffman_kernel<<<n_images, kParThreads, 0, s>>>(imgs_dev, bytes_dev, clean_dev, sub, n_sub, flags, coef_dev);
  } else if (cudaMemsetAsync(flags, 0xff, sizeof(int) * size_t(n_images), s) != cudaSuccess) {   // "use the sequential decoder"
    throw std::runtime_error("jpeg: cudaMemsetAsync failed");
  }
  jpeg_huffman_kernel<<<n_images, kHuffThreads, 0, s>>>(imgs_dev, segs_dev, bytes_dev, flags, coef_dev);
  if (any_single && !sequential_only) jpeg_dc_scan_kernel<<<dim3(3, n_images), kParThreads, 0, s>>>(imgs_dev, flags, coef_dev);
  const JpegImage& last = imgs_host[n_images - 1];
  const long long total_blocks = last.block_begin + last.nblocks;
  const int per = kIdctThreads / 8;
  jpeg_idct_kernel<<<unsigned((total_blocks + per - 1) / per), kIdctThreads, 0, s>>>(imgs_dev, n_images, total_blocks, coef_dev,
                                                                                    planes_dev);
  int wmax = 0, hmax = 0;
  for (int i = 0; i < n_images; ++i) { wmax = std::max(wmax, imgs_host[i].width); hmax = std::max(hmax, imgs_host[i].height); }
  dim3 grid((wmax + 63) / 64, (hmax + 3) / 4, n_images);
  jpeg_color_kernel<<<grid, 256, 0, s>>>(imgs_dev, planes_dev, out_dev);
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("jpeg decode launch: ") + cudaGetErrorString(cudaGetLastError()));
}

// ------------------------------------------------------------------------------------------------ batch front end
struct JpegBatch::Impl {
  DevBuf meta, coef, planes, out, clean, sub;
  DevBuf stage{true};
  cudaEvent_t copied = nullptr;
  bool pending = false;
};

JpegBatch::JpegBatch() : impl_(new Impl) {}
JpegBatch::~JpegBatch() {
  if (impl_->copied) cudaEventDestroy(impl_->copied);
  delete impl_;
}

int JpegBatch::decode(const uint8_t* const* data, const size_t* sizes, int n, cudaStream_t s, std::vector<DevImg>* out,
                      std::vector<std::string>* why) {
  out->assign(size_t(n), DevImg());
  why->assign(size_t(n), std::string());
  h2d_bytes_ = 0;
  std::vector<JpegImage> imgs;
  std::vector<JpegSeg> segs;
  std::vector<int> index;
  std::vector<std::pair<size_t, size_t>> ecs;
  imgs.reserve(size_t(n));
  auto a256 = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t bytes_total = 0, coef_blocks = 0, plane_bytes = 0, out_bytes = 0, n_sub = 0;
  for (int i = 0; i < n; ++i) {
    JpegImage im;
    std::vector<JpegSeg> sg;
    size_t b = 0, e = 0;
    if (!data[i] || sizes[i] == 0) { (*why)[i] = "Empty image data provided"; continue; }
    if (!jpeg_parse(data[i], sizes[i], &im, &b, &e, &sg, &(*why)[i])) continue;
    im.data_off = (long long)bytes_total;
    im.data_len = int(e - b);
    // 16-byte aligned start; 96 bytes of slack: the readers look up to 16 bytes past the end and the parallel decoder
    // zero-pads 64 bytes behind its un-stuffed copy, which lives at the same offsets in a second buffer
    bytes_total += (size_t(im.data_len) + 96 + 15) & ~size_t(15);
    im.seg_begin = int(segs.size());
    im.seg_count = int(sg.size());
    for (auto& x : sg) { x.image = int(imgs.size()); segs.push_back(x); }
    im.block_begin = (long long)coef_blocks;
    for (int c = 0; c < im.ncomp; ++c) {
      im.comp[c].coef_off = (long long)coef_blocks;
      coef_blocks += size_t(im.comp[c].bw) * im.comp[c].bh;
      im.comp[c].plane_off = (long long)plane_bytes;
      plane_bytes += a256(size_t(im.comp[c].bw) * 8 * im.comp[c].bh * 8);
    }
    im.nblocks = (long long)coef_blocks - im.block_begin;
    im.index = int(imgs.size());
    im.sub_begin = (long long)n_sub;
    im.sub_max = sg.size() == 1 ? (im.data_len + kJpegSubBytes - 1) / kJpegSubBytes + 1 : 0;
    n_sub += size_t(im.sub_max);
    im.out_off = (long long)out_bytes;
    im.out_stride = (long long)im.width * 3;
    out_bytes += a256(size_t(im.width) * im.height * 3);
    imgs.push_back(im);
    index.push_back(i);
    ecs.emplace_back(b, e);
  }
  const int m = int(imgs.size());
  if (m == 0) return 0;
  Impl& I = *impl_;
  if (!I.copied) cuda_check(cudaEventCreateWithFlags(&I.copied, cudaEventDisableTiming), "cudaEventCreate");
  if (I.pending) { cuda_check(cudaEventSynchronize(I.copied), "jpeg staging reuse"); I.pending = false; }
  const size_t off_segs = a256(sizeof(JpegImage) * size_t(m));
  const size_t off_bytes = off_segs + a256(sizeof(JpegSeg) * segs.size());
  const size_t total = off_bytes + bytes_total + 2 * kChunkBytes;  // the staging warp reads whole 4 KB chunks
  I.stage.ensure(total);
  I.meta.ensure(total);
  I.coef.ensure(coef_blocks * 128);
  I.planes.ensure(plane_bytes);
  I.out.ensure(out_bytes);
  I.clean.ensure(bytes_total + 2 * kChunkBytes);
  I.sub.ensure(n_sub * 48 + size_t(m) * 16 + 256);
  uint8_t* h = I.stage.as<uint8_t>();
  memcpy(h, imgs.data(), sizeof(JpegImage) * size_t(m));
  memcpy(h + off_segs, segs.data(), sizeof(JpegSeg) * segs.size());
  for (int k = 0; k < m; ++k) {
    uint8_t* dst = h + off_bytes + imgs[k].data_off;
    memcpy(dst, data[index[k]] + ecs[k].first, size_t(imgs[k].data_len));
    memset(dst + imgs[k].data_len, 0, 16);
  }
  cuda_check(cudaMemcpyAsync(I.meta.p, h, total, cudaMemcpyHostToDevice, s), "jpeg upload");
  cuda_check(cudaEventRecord(I.copied, s), "cudaEventRecord");
  I.pending = true;
  h2d_bytes_ = total;
  const uint8_t* d = I.meta.as<uint8_t>();
  launch_jpeg_decode(reinterpret_cast<const JpegImage*>(d), imgs.data(), m, reinterpret_cast<const JpegSeg*>(d + off_segs),
                     int(segs.size()), d + off_bytes, I.clean.as<uint8_t>(), I.sub.p, (long long)n_sub, I.coef.as<int16_t>(),
                     coef_blocks * 128, I.planes.as<uint8_t>(), I.out.as<uint8_t>(), s);
  launches += 5;
  for (int k = 0; k < m; ++k) {
    DevImg& o = (*out)[size_t(index[k])];
    o.p = I.out.as<uint8_t>() + imgs[k].out_off;
    o.rows = imgs[k].height; o.cols = imgs[k].width; o.stride = long(imgs[k].out_stride);
  }
  return m;
}

}  // namespace b200ocr
