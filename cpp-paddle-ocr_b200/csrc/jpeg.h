// Device-side baseline JPEG decoding: the ingest step in front of the OCR path.
//
// Replaces cv::imread / cv::imdecode in the reference's request handler (src/ocr_ipc_service.cpp:336-344), whose
// arithmetic is libjpeg(-turbo)'s with OpenCV's defaults: Huffman decoding (JPEG Annex F), JDCT_ISLOW integer inverse
// DCT (jidctint.c), "fancy" triangle-filter chroma upsampling (jdsample.c h2v1 / h2v2) and the fixed-point
// YCbCr -> BGR conversion (jdcolor.c).  Output is bit-identical to cv2.imdecode (checked through oracle/jpeg_decode.py).
//
// Split of the work:
//   host  : marker parsing only (tables, frame, scan header, restart-marker positions) -- a few hundred bytes per file;
//           the entropy-coded bytes are uploaded as they are (about 10x fewer PCIe bytes than the BGR pixels).
//   device: entropy decoding -> int16 coefficients: files with restart markers by jpeg_huffman_kernel (one thread per
//           restart interval); files without by jpeg_parallel_huffman_kernel (self-synchronising decode: 256 threads
//           per image work on 1024-bit subsequences, see jpeg.cu) + jpeg_dc_scan_kernel (DC prediction as a prefix sum);
//           jpeg_idct_kernel (8 threads per 8x8 block) -> component planes;
//           jpeg_color_kernel (upsampling + colour conversion) -> BGR u8 in the layout the det / cls / rec
//           pre-processing kernels read.
// Scope: what the restated algorithm covers -- baseline / extended sequential, Huffman, 8-bit, one interleaved scan,
// grey or YCbCr with 4:4:4 / 4:2:2 / 4:2:0 sampling, EXIF orientation absent or 1.  Anything else (progressive,
// arithmetic, CMYK, PNG, ...) is reported as unsupported so that the caller can decode it the reference's way.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200ocr {

struct JpegHuffLut {        // one Huffman table, canonical form (JPEG Annex C / F.2.2.3)
  uint16_t fast[512];       // 9-bit look-ahead: (length << 8) | symbol; 0 = code longer than 9 bits
  int32_t maxcode[18];      // largest code of each length (-1: none); [17] = sentinel
  int32_t valoff[17];       // index of the first symbol of a length minus its first code
  uint8_t vals[256];
};

struct JpegComp {
  int hs, vs;               // sampling factors
  int bw, bh;               // blocks per row / column over the whole MCU grid (mcux * hs, mcuy * vs)
  int cw, ch;               // real extent of the component in samples (ceil(width * hs / hmax), ...)
  long long coef_off;       // first block of this component in the coefficient buffer (units of 64 int16)
  long long plane_off;      // byte offset of its plane [bh * 8][bw * 8] in the plane buffer
  uint16_t q[64];           // quantisation table, natural (row-major) order
};

struct JpegImage {          // POD shared by host and device
  int width, height, ncomp, hmax, vmax, mcux, mcuy, restart_interval;
  JpegComp comp[3];
  JpegHuffLut lut[6];       // component c: DC table lut[2c], AC table lut[2c + 1]
  long long data_off;       // offset of the entropy-coded segment in the batch's byte buffer
  int data_len;
  int seg_begin, seg_count; // its restart intervals in the batch's segment list
  long long block_begin;    // first 8x8 block of this image in the batch-wide block numbering
  long long nblocks;
  long long out_off;        // byte offset of the BGR image in the output buffer
  long long out_stride;
  long long sub_begin;      // single-interval images: first entry of the image's subsequence table (parallel decode)
  int sub_max;              // its capacity: ceil(data_len / kJpegSubBytes)
  int index;                // position of the image in the batch (flags)
};

constexpr int kJpegSubBytes = 128;  // the parallel entropy decoder cuts an interval into subsequences of 1024 bits

struct JpegSeg { int image; int begin, end; int mcu0, nmcu; };  // [begin, end) relative to the image's data_off

// Parses the markers of one file.  On success fills `img` (everything except the batch-relative offsets), the byte
// range of the entropy-coded segment inside `data` and the restart-interval list (offsets relative to that range).
// Returns false with a reason when the file is not in the supported subset.
bool jpeg_parse(const uint8_t* data, size_t size, JpegImage* img, size_t* ecs_begin, size_t* ecs_end,
                std::vector<JpegSeg>* segs, std::string* why);

// Device side.  All pointers are device memory; `imgs` / `segs` were uploaded by the caller.
// `clean_dev`: as large as the byte buffer (un-stuffed copies); `sub_dev`: 3 x 16 bytes per subsequence + 4 per image
// (see jpeg.cu); `n_sub`: total subsequences of the batch.
void launch_jpeg_decode(const JpegImage* imgs_dev, const JpegImage* imgs_host, int n_images, const JpegSeg* segs_dev,
                        int n_segs, const uint8_t* bytes_dev, uint8_t* clean_dev, void* sub_dev, long long n_sub,
                        int16_t* coef_dev, size_t coef_bytes, uint8_t* planes_dev, uint8_t* out_dev, cudaStream_t s);

struct DevBuf;
struct DevImg;

// A batch of encoded files -> device-resident BGR images (the layout ImageBatch::upload produces).  One instance per
// worker; buffers are reused from call to call.  Not thread-safe (like the worker that owns it).
class JpegBatch {
 public:
  JpegBatch();
  ~JpegBatch();
  // out[i].p == nullptr and why[i] set when file i is outside the supported subset; the rest are decoded on `s`
  // (asynchronously: the images are valid for work queued on `s` afterwards).  Returns the number decoded.
  int decode(const uint8_t* const* data, const size_t* sizes, int n, cudaStream_t s, std::vector<DevImg>* out,
             std::vector<std::string>* why);
  size_t h2d_bytes() const { return h2d_bytes_; }   // bytes of the last call's host -> device copy
  long launches = 0;
 private:
  struct Impl;
  Impl* impl_;
  size_t h2d_bytes_ = 0;
};

}  // namespace b200ocr
