/* b200ocr — C ABI of the B200-native det -> cls -> rec OCR path.
 *
 * Drop-in boundary for the three stage classes and the worker of sssxyd/cpp-paddle-ocr
 * (file:line citations are relative to that repository):
 *
 *   b200ocr_det_*     replaces  PaddleOCR::DBDetector      include/paddle_ocr/ocr_det.h:60-97,  src/ocr_det.cpp:93-176
 *   b200ocr_cls_*     replaces  PaddleOCR::Classifier      include/paddle_ocr/ocr_cls.h:57-82,  src/ocr_cls.cpp:23-106
 *   b200ocr_rec_*     replaces  PaddleOCR::CRNNRecognizer  include/paddle_ocr/ocr_rec.h:61-95,  src/ocr_rec.cpp:24-135
 *   b200ocr_worker_*  replaces  PaddleOCR::OCRWorker::processRequest + result JSON  src/ocr_worker.cpp:133-311
 *   b200ocr_pool_*    replaces  PaddleOCR::GPUWorkerPool   include/paddle_ocr/gpu_worker_pool.h:14-31, src/gpu_worker_pool.cpp:8-59
 *
 * Conventions
 *   - Plain C types only.  Images are 8-bit, 3-channel, BGR, row-major host memory (what
 *     cv::Mat::data / rows / cols / step describe).
 *   - Every function returns 0 on success or a B200OCR_ERR_* code; b200ocr_last_error() returns
 *     a thread-local message.  Nothing in the library calls exit() (the reference does on a missing
 *     model, src/ocr_det.cpp:41-45).
 *   - Handles are single-threaded like the reference stage objects (one set per worker);
 *     only b200ocr_pool_* is thread-safe.
 *   - All arithmetic runs on the GPU (sm_100a).  There is no CPU fallback: creating a handle without
 *     a usable CUDA device fails with B200OCR_ERR_RUNTIME.
 */
#ifndef B200OCR_H_
#define B200OCR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200OCR_OK 0
#define B200OCR_ERR_INVALID 1 /* bad argument */
#define B200OCR_ERR_RUNTIME 2 /* CUDA / model / shape error; see b200ocr_last_error() */
#define B200OCR_ERR_NOMEM 3

const char* b200ocr_last_error(void);
const char* b200ocr_version(void);
/* Frees strings / buffers the library returned through an out-pointer. */
void b200ocr_free(void* p);

/* Lists the persistable parameters of a .pdmodel in `.pdiparams` order (ascending name) as JSON
 * [{"name":..., "dims":[...]}, ...].  Host-only (no GPU needed); free *json with b200ocr_free. */
int b200ocr_model_params_json(const char* pdmodel_path, char** json);
/* Builds the fused-layer plan of <model_dir> on the host (no GPU needed) and returns its text dump
 * (one line per fused kernel).  Free *text with b200ocr_free. */
int b200ocr_model_plan_text(const char* model_dir, char** text);

/* ------------------------------------------------------------------ images
 * What cv::Mat::data / rows / cols / step describe: 8-bit, 3 channels, BGR, row-major HOST memory
 * (pinned memory from b200ocr_host_alloc makes the upload asynchronous; any host pointer works). */
typedef struct b200ocr_image {
  const uint8_t* data;
  int rows, cols;
  size_t step; /* bytes per row */
} b200ocr_image;

void* b200ocr_host_alloc(size_t bytes); /* page-locked host memory; NULL on failure */
void b200ocr_host_free(void* p);
int b200ocr_device_count(void);

/* ------------------------------------------------------------------ DBDetector
 * Constructor arguments of PaddleOCR::DBDetector (include/paddle_ocr/ocr_det.h:60-69), same order and meaning.
 * use_gpu, gpu_mem, cpu_math_library_num_threads, use_mkldnn, use_tensorrt and precision are accepted for
 * signature compatibility and ignored: this implementation always runs on GPU `gpu_id` in fp16 with fp32
 * accumulation.  det_db_score_mode: "fast" (the worker's setting; BoxScoreFast) or "slow" (PolygonScoreAcc); anything else
 * returns B200OCR_ERR_INVALID.  A C++ caller can use include/paddle_ocr/b200ocr_shim.h instead: the same three classes
 * with the reference's constructor and Run() signatures over these entry points. */
typedef struct b200ocr_det_config {
  const char* model_dir;
  int use_gpu, gpu_id, gpu_mem, cpu_math_library_num_threads, use_mkldnn;
  const char* limit_type; /* "max" | "min" */
  int limit_side_len;
  double det_db_thresh, det_db_box_thresh, det_db_unclip_ratio;
  const char* det_db_score_mode;
  int use_dilation, use_tensorrt;
  const char* precision;
} b200ocr_det_config;
typedef struct b200ocr_det* b200ocr_det_t;
int b200ocr_det_create(const b200ocr_det_config* cfg, b200ocr_det_t* out);
void b200ocr_det_destroy(b200ocr_det_t det);
/* DBDetector::Run (ocr_det.h:95-97, src/ocr_det.cpp:93-176).  boxes: [cap][4][2] int32, points ordered
 * tl,tr,br,bl in source pixels, in the reference's order; *n_boxes = number found (may exceed cap; only cap are
 * written).  times: 3 values appended by the reference (ms: pre, infer, post); may be NULL. */
int b200ocr_det_run(b200ocr_det_t det, const b200ocr_image* img, int32_t* boxes, int cap, int* n_boxes, double times[3]);
/* Batched form: n images in one pass; boxes [n][cap][4][2], counts [n]. */
int b200ocr_det_run_batch(b200ocr_det_t det, const b200ocr_image* imgs, int n, int32_t* boxes, int cap, int* counts,
                          double times[3]);
/* Post-processing only (BoxesFromBitmap + FilterTagDetRes, src/postprocess_op.cpp:255-362) on a caller-supplied
 * probability map pred [h][w] fp32 (host): bitmap = (uchar)(p*255) > det_db_thresh*255, as src/ocr_det.cpp:143-154.
 * src_h/src_w: size of the image the boxes are mapped back to. */
int b200ocr_det_postprocess(b200ocr_det_t det, const float* pred, int h, int w, int src_h, int src_w, int32_t* boxes,
                            int cap, int* n_boxes, uint8_t* bitmap_out /* [h][w] or NULL */);
/* Pre-processing only (ResizeImgType0 + Normalize + Permute): writes the network input as fp32 NCHW [3][rh][rw]
 * (host; pass NULL to query the size) and the resize ratios. */
int b200ocr_det_preprocess(b200ocr_det_t det, const b200ocr_image* img, float* nchw, int* rh, int* rw, float* ratio_h,
                           float* ratio_w);

/* ------------------------------------------------------------------ Classifier
 * PaddleOCR::Classifier (include/paddle_ocr/ocr_cls.h:57-62). */
typedef struct b200ocr_cls_config {
  const char* model_dir;
  int use_gpu, gpu_id, gpu_mem, cpu_math_library_num_threads, use_mkldnn;
  double cls_thresh;
  int use_tensorrt;
  const char* precision;
  int cls_batch_num;
} b200ocr_cls_config;
typedef struct b200ocr_cls* b200ocr_cls_t;
int b200ocr_cls_create(const b200ocr_cls_config* cfg, b200ocr_cls_t* out);
void b200ocr_cls_destroy(b200ocr_cls_t cls);
/* Classifier::Run (ocr_cls.h:81-82, src/ocr_cls.cpp:23-106): cls_labels / cls_scores are caller-sized [n]. */
int b200ocr_cls_run(b200ocr_cls_t cls, const b200ocr_image* imgs, int n, int* cls_labels, float* cls_scores,
                    double times[3]);

/* ------------------------------------------------------------------ CRNNRecognizer
 * PaddleOCR::CRNNRecognizer (include/paddle_ocr/ocr_rec.h:61-68). */
typedef struct b200ocr_rec_config {
  const char* model_dir;
  int use_gpu, gpu_id, gpu_mem, cpu_math_library_num_threads, use_mkldnn;
  const char* label_path;
  int use_tensorrt;
  const char* precision;
  int rec_batch_num, rec_img_h, rec_img_w;
} b200ocr_rec_config;
typedef struct b200ocr_rec* b200ocr_rec_t;
int b200ocr_rec_create(const b200ocr_rec_config* cfg, b200ocr_rec_t* out);
void b200ocr_rec_destroy(b200ocr_rec_t rec);
/* CRNNRecognizer::Run (ocr_rec.h:92-95, src/ocr_rec.cpp:24-135).  rec_texts: [n] UTF-8 strings, each allocated by the
 * library (free every entry with b200ocr_free); rec_text_scores: caller-sized [n].  Lines that decode to nothing
 * keep "" and 0, like the reference's untouched pre-sized vectors. */
int b200ocr_rec_run(b200ocr_rec_t rec, const b200ocr_image* imgs, int n, char** rec_texts, float* rec_text_scores,
                    double times[3]);

/* ------------------------------------------------------------------ OCRWorker
 * PaddleOCR::OCRWorker(worker_id, model_dir, use_gpu, gpu_id, enable_cls) (include/paddle_ocr/ocr_worker.h:57) with the
 * hyper-parameters of src/ocr_worker.cpp:21-63.  process() = processRequest + the result JSON of
 * src/ocr_worker.cpp:155-190 (one compact line, jsoncpp key order).  Free *json with b200ocr_free. */
typedef struct b200ocr_worker* b200ocr_worker_t;
int b200ocr_worker_create(int worker_id, const char* model_dir, int use_gpu, int gpu_id, int enable_cls,
                          b200ocr_worker_t* out);
/* The same worker with other stage hyper-parameters than the ones src/ocr_worker.cpp:21-63 hard-codes: fields that are
 * 0 / NULL keep the reference's value (limit "max" 512, thresh 0.2, box_thresh 0.4, unclip 1.8, "fast", no dilation,
 * cls batch 8, rec batch 16 at 28x192, 64 images per device batch).  Dense 2048x2048 pages (BASELINE config 5) use
 * limit_side_len 960: at 512 their text lines shrink to 6 px. */
typedef struct b200ocr_worker_params {
  const char* limit_type;
  int limit_side_len;
  double det_db_thresh, det_db_box_thresh, det_db_unclip_ratio;
  const char* det_db_score_mode;
  int use_dilation, cls_batch_num, rec_batch_num, rec_img_h, rec_img_w, max_batch;
} b200ocr_worker_params;
int b200ocr_worker_create_ex(int worker_id, const char* model_dir, int gpu_id, int enable_cls,
                             const b200ocr_worker_params* params, b200ocr_worker_t* out);
void b200ocr_worker_destroy(b200ocr_worker_t w);
int b200ocr_worker_process(b200ocr_worker_t w, int request_id, const b200ocr_image* img, char** json);
/* Throughput form: n requests through the GPU together (results are per image, identical to n process() calls). */
int b200ocr_worker_process_batch(b200ocr_worker_t w, const int* request_ids, const b200ocr_image* imgs, int n,
                                 char** jsons);
/* Device-resident inputs: upload once, process many times (what a caller that already has its frames in HBM uses;
 * bench.py's device-resident throughput figure).  The batch is not modified by processing. */
typedef struct b200ocr_batch* b200ocr_batch_t;
int b200ocr_batch_upload(int device, const b200ocr_image* imgs, int n, b200ocr_batch_t* out);
void b200ocr_batch_destroy(b200ocr_batch_t batch);
int b200ocr_worker_process_resident(b200ocr_worker_t w, b200ocr_batch_t batch, const int* request_ids, char** jsons);
/* ENCODED inputs -- the bytes the reference hands to cv::imread / cv::imdecode (src/ocr_ipc_service.cpp:336-344): only
 * the file's bytes cross PCIe; baseline JPEG (sequential Huffman, 8-bit, grey / 4:4:4 / 4:2:2 / 4:2:0, EXIF orientation
 * absent or 1) is decoded on the device, bit-identical to cv::imdecode.  A file outside that subset gets a result line
 * with success=false and error "Unsupported image encoding: <reason>": decode it the reference's way (cv::imdecode) and
 * call b200ocr_worker_process. */
typedef struct b200ocr_blob { const uint8_t* data; size_t size; } b200ocr_blob;
int b200ocr_worker_process_encoded(b200ocr_worker_t w, const int* request_ids, const b200ocr_blob* blobs, int n,
                                   char** jsons);
/* bytes of the last process_encoded call's host -> device copy (tables + entropy-coded data) */
long long b200ocr_worker_last_encoded_h2d_bytes(b200ocr_worker_t w);
/* The decoder on its own: *rows / *cols always; bgr [rows][cols][3] (host) when not NULL.  Unsupported files return
 * B200OCR_ERR_INVALID with the reason in b200ocr_last_error(). */
int b200ocr_jpeg_decode(int device, const uint8_t* data, size_t size, int* rows, int* cols, uint8_t* bgr);
/* The worker's CUDA stream (a cudaStream_t), so that a caller can bracket its work with CUDA events. */
void* b200ocr_worker_stream(b200ocr_worker_t w);
/* kernels launched by this worker so far */
long long b200ocr_worker_launches(b200ocr_worker_t w);

/* Stage calls on device-resident inputs (b200ocr_batch_upload; every image of the batch is one input of the call) --
 * what bench.py's device-resident figures for the detection-only / recognition-only configurations time.  Outputs as
 * in b200ocr_det_run_batch / b200ocr_rec_run.  The *_stream accessors return the stage's cudaStream_t. */
int b200ocr_det_run_resident(b200ocr_det_t det, b200ocr_batch_t batch, int32_t* boxes, int cap, int* counts, double times[3]);
int b200ocr_rec_run_resident(b200ocr_rec_t rec, b200ocr_batch_t batch, char** rec_texts, float* rec_text_scores,
                             double times[3]);
void* b200ocr_det_stream(b200ocr_det_t det);
void* b200ocr_rec_stream(b200ocr_rec_t rec);
long long b200ocr_det_launches(b200ocr_det_t det);
long long b200ocr_rec_launches(b200ocr_rec_t rec);
/* Per-fused-layer device time of the stage's network at the largest forward pass of its last run, in the format of
 * b200ocr_worker_profile: {"det":{"shape":[n,h,w],"layers":[...]}} / {"rec":{...}}; free with b200ocr_free. */
int b200ocr_det_profile(b200ocr_det_t det, int warmup, int reps, char** json);
int b200ocr_rec_profile(b200ocr_rec_t rec, int warmup, int reps, char** json);

/* ------------------------------------------------------------------ per-device worker pool
 * Replaces PaddleOCR::GPUWorkerPool (include/paddle_ocr/gpu_worker_pool.h:14-31, src/gpu_worker_pool.cpp:8-59), which
 * pins every worker to GPU 0: here workers are spread over `n_devices` GPUs (devices[i], or 0..n-1 when NULL),
 * `workers_per_device` each; submit() is thread-safe, clones the image (like OCRRequest's clone) and returns a
 * ticket; a worker drains up to `max_batch` queued requests at a time.  Dispatch: the device with the shortest queue
 * (the reference's idle-first / round-robin, generalised).
 * The clone goes straight into the memory of the device the request is dispatched to, through that device's uploader
 * thread: when `img->data` is page-locked (b200ocr_host_alloc / cudaHostAlloc / cudaHostRegister) it is one DMA and
 * no CPU core touches the pixels; ordinary pageable memory is first copied into a page-locked bounce buffer by the
 * calling thread.  submit() returns when the pixels have landed: the caller's buffer is free again. */
typedef struct b200ocr_pool* b200ocr_pool_t;
int b200ocr_pool_create(const char* model_dir, int n_devices, const int* devices, int workers_per_device,
                        int enable_cls, int max_batch, b200ocr_pool_t* out);
void b200ocr_pool_destroy(b200ocr_pool_t pool);
int b200ocr_pool_submit(b200ocr_pool_t pool, int request_id, const b200ocr_image* img, long long* ticket);
/* The same for an encoded file (see b200ocr_worker_process_encoded); the bytes are copied. */
int b200ocr_pool_submit_encoded(b200ocr_pool_t pool, int request_id, const uint8_t* data, size_t size, long long* ticket);
/* Blocks until the request is done; *json is the worker's result line (free with b200ocr_free).  A ticket that was
 * never issued or was already consumed returns B200OCR_ERR_INVALID; b200ocr_pool_destroy wakes pending waiters, which
 * return B200OCR_ERR_RUNTIME. */
int b200ocr_pool_wait(b200ocr_pool_t pool, long long ticket, char** json);
/* Same with a time limit: returns B200OCR_OK with *json == NULL when `timeout_ms` (>= 0) elapsed first; the ticket stays
 * valid.  timeout_ms < 0 waits forever. */
int b200ocr_pool_wait_for(b200ocr_pool_t pool, long long ticket, int timeout_ms, char** json);
int b200ocr_pool_worker_count(b200ocr_pool_t pool);
/* Service counters as one compact JSON object (free with b200ocr_free); the reference's getStatusInfo
 * (src/ocr_ipc_service.cpp:438-448) plus what it declares but never updates:
 * {"average_processing_time_ms","batches","failed_requests","idle_workers","queued_requests","running",
 *  "stage_ms_per_image":{"cls","det","rec"},"successful_requests","total_requests","uptime_s","workers"}; the average
 * is submission -> completion; stage_ms_per_image is the workers' host wall time per image by stage (the `times` the
 * reference's stages report and its worker drops; det includes the wait for its boxes, cls is enqueue only, rec
 * includes the wait for cls + rec). */
int b200ocr_pool_status(b200ocr_pool_t pool, char** json);
int b200ocr_pool_idle_count(b200ocr_pool_t pool);

/* ------------------------------------------------------------------ stand-alone image ops (test / utility surface)
 * cv::resize(INTER_LINEAR) of an 8-bit BGR image on the GPU (the kernel inside det/cls/rec pre-processing). */
int b200ocr_resize_u8(int device, const b200ocr_image* src, int dst_rows, int dst_cols, uint8_t* dst);
/* Utility::GetRotateCropImage (src/utility.cpp:137-190; defined by the reference but never called by its worker, which
 * crops the bounding rectangle instead): perspective crop of `box` (4 points tl,tr,br,bl), transposed + flipped when
 * height >= 1.5 x width.  Call with dst == NULL to get the output size. */
int b200ocr_rotate_crop(int device, const b200ocr_image* src, const int32_t box[8], int* dst_rows, int* dst_cols, uint8_t* dst);
/* Rec / cls pre-processing of one batch of crops to fp32 NCHW [n][3][img_h][img_w] (host), for parity tests:
 * kind 0 = rec (CrnnResizeImg, pad -1 after normalisation of u8 zeros), 1 = cls (ClsResizeImg, pad 0). */
int b200ocr_crop_preprocess(int device, const b200ocr_image* crops, int n, int kind, int img_h, int img_w, float* nchw);

/* ------------------------------------------------------------------ network-level entry points
 * One loaded .pdmodel/.pdiparams pair executing on one GPU; what `paddle_infer::Predictor`
 * is to the reference stages (predictor_->Run(), src/ocr_det.cpp:120).  Used by the stages below
 * and by the parity tests, which compare every fused layer against the CPU oracle. */
typedef struct b200ocr_net* b200ocr_net_t;

#define B200OCR_NET_KEEP_ALL 1   /* keep every intermediate tensor fetchable (no buffer reuse) */
#define B200OCR_NET_FORCE_SIMT 2 /* run dense convolutions on the CUDA-core kernel (A/B check of the tcgen05 path) */
#define B200OCR_NET_NO_GRAPH 4   /* launch kernels one by one instead of replaying a CUDA graph */

int b200ocr_net_create(const char* model_dir, int device, int flags, b200ocr_net_t* out);
void b200ocr_net_destroy(b200ocr_net_t net);
/* "det" | "cls" | "rec" */
int b200ocr_net_kind(b200ocr_net_t net, char* buf, int cap);
int b200ocr_net_plan_dump(b200ocr_net_t net, char* buf, int cap, int* needed);
/* Forward pass on a host fp32 NCHW tensor [n,3,height,width] (already normalised).
 * det: thresh_u8 >= 0 additionally produces the bitmap `(uchar)(p*255) > thresh_u8` (src/ocr_det.cpp:143-154). */
int b200ocr_net_forward(b200ocr_net_t net, const float* nchw, int n, int height, int width, int thresh_u8);
/* Ragged batch (sequence graphs): row i is widths[i] <= width columns wide and must be zero beyond that; every row's
 * result is bit-identical to running it in a dense batch of its own width (tests/test_net_parity_gpu.py). */
int b200ocr_net_forward_ragged(b200ocr_net_t net, const float* nchw, int n, int height, int width, const int* widths);
/* det: (n, H, W)   cls: (n, 1, 1)   rec: (n, 1, T) */
int b200ocr_net_out_shape(b200ocr_net_t net, int shape[3]);
/* det: out_f32 = probability map [n,H,W], out_bitmap = [n,H,W] (0/255)
 * cls: out_f32 = softmax [n,2]
 * rec: out_f32 = softmax probability of the arg-max class per time step [n,T], out_idx = that class [n,T]
 * Any pointer may be NULL. */
int b200ocr_net_output(b200ocr_net_t net, float* out_f32, uint8_t* out_bitmap, int32_t* out_idx);
/* Copies an intermediate tensor (Paddle variable name) of the last forward to host as fp32 NCHW.
 * Needs B200OCR_NET_KEEP_ALL.  out may be NULL to query dims. */
int b200ocr_net_fetch(b200ocr_net_t net, const char* var, float* out, size_t cap_elems, int dims[4]);
/* Per-fused-layer device time (CUDA events around every launch, averaged over `reps` after `warmup` passes) of the last
 * forward's shape, as JSON [{"name","kind","ms","flops","bytes","tensor_core"}...]; free with b200ocr_free. */
int b200ocr_net_profile(b200ocr_net_t net, int warmup, int reps, char** json);
/* Same for the three networks of a worker at the shapes of its last process call:
 * {"det":[...],"cls":[...],"rec":[...]} . */
int b200ocr_worker_profile(b200ocr_worker_t w, int warmup, int reps, char** json);
/* Number of kernels one forward pass launches for the last shape. */
int b200ocr_net_launches(b200ocr_net_t net);

/* ------------------------------------------------------------------ kernel-level entry points (parity tests)
 * One hot kernel on host tensors (fp32 in / out; fp16 NHWC activations and fp32 accumulation on the device, exactly
 * what the networks run), so that tests can sweep shapes the shipped graphs do not contain.
 * Depthwise convolution = Paddle depthwise_conv2d (+ folded bias, activation 0 none / 1 relu / 2 hard-swish, scalar
 * affine): x [n,c,h,w], filt [c][k*k], bias [c], k in {3,5}, padding k/2, strides 1|2.  fp16_weights = 0 keeps fp32
 * filter taps (det / cls), 1 rounds them to fp16 (rec).  out_widths (may be NULL): ragged rows -- out[i] is zero at
 * x >= out_widths[i].  out [n,c,out_h,out_w]. */
int b200ocr_kernel_dwconv(int device, const float* x, int n, int c, int h, int w, const float* filt, const float* bias,
                          int k, int sh, int sw, int act, float post_scale, float post_shift, int fp16_weights,
                          const int* out_widths, float* out, int* out_h, int* out_w);
/* Dense convolution = Paddle conv2d (stride 1, "same" padding, odd kh x kw; + folded bias, activation 0 none / 1 relu /
 * 2 hard-swish / 3 swish, scalar affine, optional residual [n,cout,h,w] added after the affine): x [n,cin,h,w], filt
 * [cout][cin][kh][kw].  force_simt = 0: the engine's choice (the mma.sync pixel stream for 1x1 filters with <= 48 input
 * and <= 64 output channels, else the tcgen05 implicit-GEMM kernel when the shape is eligible (cin >= 8), else the CUDA-core kernel);
 * 1: the CUDA-core kernel; 2: as 0 without the narrow-1x1 kernel.  out_widths as above.  out [n,cout,h,w]. */
int b200ocr_kernel_conv(int device, const float* x, int n, int cin, int h, int w, const float* filt, const float* bias,
                        int cout, int kh, int kw, int act, float post_scale, float post_shift, const float* residual,
                        const int* out_widths, int force_simt, float* out);
/* SVTR self-attention on packed qkv rows [n][t][3][heads][head_dim] -> out [n][t][heads*head_dim];
 * valid (may be NULL): tokens per sequence, keys beyond it are ignored and queries beyond it give zeros. */
int b200ocr_kernel_attention(int device, const float* qkv, int n, int t, int heads, int head_dim, float scale,
                             const int* valid, float* out);
/* Recognizer head + greedy decode (reference src/ocr_rec.cpp:97-128) on its own: feat [n][t][cin] (values are rounded to
 * fp16 like the activations the networks hold), w [cin][ncls] (Paddle linear layout, rounded to fp16), bias [ncls] ->
 * idx [n][t] arg-max class (first maximum wins), prob [n][t] soft-max probability of that class, then the blank / repeat
 * collapse: collapsed [n][t] label ids (first lens[i] entries of a row valid), scores [n] mean probability of the kept
 * steps (0 when nothing is kept).  force_simt = 0: the tcgen05 kernel (error when the shape is not eligible), 1: the
 * CUDA-core kernel.  collapsed / lens / scores may be NULL. */
int b200ocr_kernel_ctc_head(int device, const float* feat, int n, int t, int cin, const float* w, const float* bias,
                            int ncls, int force_simt, int32_t* idx, float* prob, int32_t* collapsed, int32_t* lens,
                            float* scores);

#ifdef __cplusplus
}
#endif
#endif /* B200OCR_H_ */
