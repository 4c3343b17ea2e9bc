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
/* Number of kernels one forward pass launches for the last shape. */
int b200ocr_net_launches(b200ocr_net_t net);

#ifdef __cplusplus
}
#endif
#endif /* B200OCR_H_ */
