// The three pipeline stages and the worker, B200-native.
//
//   DetStage  <->  PaddleOCR::DBDetector::Run      reference src/ocr_det.cpp:93-176
//   ClsStage  <->  PaddleOCR::Classifier::Run      reference src/ocr_cls.cpp:23-106
//   RecStage  <->  PaddleOCR::CRNNRecognizer::Run  reference src/ocr_rec.cpp:24-135
//   Worker    <->  PaddleOCR::OCRWorker::processRequest + result JSON  reference src/ocr_worker.cpp:133-311
//
// Unlike the reference (one image, one stage call at a time, every tensor through host memory) a stage
// takes a whole batch of device-resident images / ROIs and runs it as a handful of kernels; only boxes,
// labels and decoded label ids ever return to the host.  Results are defined to be what the reference
// computes for each image on its own: batching never changes them (rec rows are only merged into one
// launch when they share the padded width the reference would have used).
#pragma once
#include <array>
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

namespace b200ocr {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  explicit DevBuf(bool pinned_host = false) : pinned(pinned_host) {}
  ~DevBuf();
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  // Grows the buffer (contents are NOT kept).  The old allocation may still be read by work queued on the owner's
  // stream, and a device-wide synchronisation is not allowed while another worker's stream is being captured into a
  // CUDA graph, so it is only retired here and freed with the object.
  void ensure(size_t bytes);
  template <class T> T* as() const { return static_cast<T*>(p); }
 private:
  std::vector<void*> retired_;
};

struct HostImage {  // what cv::Mat::data / rows / cols / step describe (8-bit BGR)
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
};
struct DevImg {
  uint8_t* p = nullptr;
  int rows = 0, cols = 0;
  long stride = 0;
};
struct Roi {
  int img = 0;  // index into the DevImg array
  int x = 0, y = 0, w = 0, h = 0;
};
using Box = std::array<int, 8>;  // 4 points (x,y): tl, tr, br, bl in source-image pixels

struct DetParams {  // DBDetector ctor arguments that affect results (reference include/paddle_ocr/ocr_det.h:60-69)
  std::string limit_type = "max";
  int limit_side_len = 960;
  double det_db_thresh = 0.3, det_db_box_thresh = 0.5, det_db_unclip_ratio = 2.0;
  std::string det_db_score_mode = "slow";
  bool use_dilation = false;
};

// Uploads host images into one device buffer (rows tightly packed, stride = cols*3).
class ImageBatch {
 public:
  void upload(const HostImage* imgs, int n, cudaStream_t s);
  const std::vector<DevImg>& images() const { return imgs_; }
  size_t bytes() const { return bytes_; }
 private:
  DevBuf dev_;
  std::vector<DevImg> imgs_;
  size_t bytes_ = 0;
};

class DetStage {
 public:
  DetStage(const std::string& model_dir, int device, const DetParams& p);
  // boxes[i] = the reference's `boxes` for image i.  times (ms): pre, infer, post (appended, 3 values).
  void run(const std::vector<DevImg>& imgs, std::vector<std::vector<Box>>* boxes, cudaStream_t s,
           std::vector<double>* times = nullptr);
  static void resized_dims(int rows, int cols, const std::string& limit_type, int limit_side_len, int* rh, int* rw,
                           float* ratio_h, float* ratio_w);
  Net& net() { return net_; }
  int prof_n = 0, prof_h = 0, prof_w = 0;  // largest forward pass of the last run() (what b200ocr_worker_profile times)
  int max_batch = 32;  // images per forward pass
  long launches = 0;   // kernels launched so far (bench accounting)
 private:
  void run_group(const std::vector<DevImg>& imgs, const std::vector<int>& idx, int rh, int rw,
                 std::vector<std::vector<Box>>* boxes, cudaStream_t s, double* t_ms);
  Net net_;
  DetParams p_;
  int thresh_u8_;
  DevBuf items_, info_, ws_, counts_, boxes_, dil_;
  DevBuf h_items_{true}, h_info_{true}, h_counts_{true}, h_boxes_{true};
};

class ClsStage {
 public:
  ClsStage(const std::string& model_dir, int device, int cls_batch_num, float cls_thresh);
  // labels/scores for every ROI; the device label array stays valid until the next run (for rotate_rois).
  void run(const std::vector<DevImg>& imgs, const std::vector<Roi>& rois, std::vector<int>* labels,
           std::vector<float>* scores, cudaStream_t s, bool fetch_to_host = true, std::vector<double>* times = nullptr);
  // cv::rotate(ROTATE_180) in place on every ROI whose label is 1, sequentially in ROI order per image
  // (reference src/ocr_worker.cpp:277-281; ROIs alias the image, order matters where they overlap).
  void rotate_rois(const std::vector<DevImg>& imgs, const std::vector<Roi>& rois, cudaStream_t s);
  Net& net() { return net_; }
  int prof_n = 0, prof_h = 0, prof_w = 0;  // largest forward pass of the last run() (what b200ocr_worker_profile times)
  int max_batch = 512;
  long launches = 0;
 private:
  Net net_;
  int batch_num_;
  float thresh_;
  DevBuf items_, labels_, probs_, rot_;
  DevBuf h_items_{true}, h_out_{true}, h_rot_{true};
};

class RecStage {
 public:
  RecStage(const std::string& model_dir, int device, const std::string& label_path, int rec_batch_num, int rec_img_h,
           int rec_img_w);
  // One entry of `calls` = one CRNNRecognizer::Run call of the reference (the ROIs of one image): the
  // aspect-ratio sort and the batches of rec_batch_num are formed inside a call, exactly like the reference.
  // texts/scores are written per ROI in caller order ("" / 0 when nothing was decoded).
  void run(const std::vector<DevImg>& imgs, const std::vector<std::vector<Roi>>& calls,
           std::vector<std::vector<std::string>>* texts, std::vector<std::vector<float>>* scores, cudaStream_t s,
           std::vector<double>* times = nullptr);
  const std::vector<std::string>& labels() const { return label_list_; }
  Net& net() { return net_; }
  int prof_n = 0, prof_h = 0, prof_w = 0;  // largest forward pass of the last run() (what b200ocr_worker_profile times)
  int max_rows = 1024;       // rows per forward pass
  long max_cols = 400000;    // rows x padded width per forward pass (bounds the activation arena)
  int last_chunks = 0; long last_cols = 0, last_real_cols = 0;  // trace: ragged chunking of the last run()
  double min_fill = 0.92;    // a ragged chunk is cut where its real columns / (rows x widest row) would drop below this
  long launches = 0;
 private:
  Net net_;
  int batch_num_, img_h_, img_w_;
  std::vector<std::string> label_list_;
  DevBuf items_, cidx_, clen_, cscore_;
  DevBuf h_items_{true}, h_cidx_{true}, h_clen_{true}, h_cscore_{true};
};

class JpegBatch;
struct WordOut { std::string text; float confidence; Box box; };

struct WorkerOptions {
  bool enable_cls = false;
  int max_batch = 64;  // images processed together by process_batch
  // Stage hyper-parameters; the defaults are what the reference OCRWorker hard-codes (src/ocr_worker.cpp:21-63).
  // b200ocr_worker_create_ex overrides them (a dense 2048x2048 page needs limit_side_len 960 to keep its lines legible).
  std::string det_limit_type = "max";
  int det_limit_side_len = 512;
  double det_db_thresh = 0.2, det_db_box_thresh = 0.4, det_db_unclip_ratio = 1.8;
  std::string det_db_score_mode = "fast";
  bool use_dilation = false;
  int cls_batch_num = 8;
  double cls_thresh = 0.98;
  int rec_batch_num = 16, rec_img_h = 28, rec_img_w = 192;
};

// Same hyper-parameters as the reference OCRWorker constructor (src/ocr_worker.cpp:21-63).
class Worker {
 public:
  Worker(int worker_id, const std::string& model_dir, int device, const WorkerOptions& opt);
  ~Worker();
  // One result JSON per image (reference schema, src/ocr_worker.cpp:155-190).
  void process_batch(const int* request_ids, const HostImage* imgs, int n, std::vector<std::string>* json);
  // Same, for images that already live in device memory (`resident` is not modified: when the classifier is
  // enabled its in-place ROI rotations happen on a device-side copy, like the reference's cloned request image).
  void process_resident(const int* request_ids, const std::vector<DevImg>& resident, std::vector<std::string>* json);
  // Same for ENCODED images (the bytes cv::imread / cv::imdecode would be given, reference
  // src/ocr_ipc_service.cpp:336-344): baseline JPEG is decoded on the device (jpeg.h); a file outside that subset gets
  // success=false with error "Unsupported image encoding: <reason>" and is the caller's to decode the reference's way.
  void process_encoded(const int* request_ids, const uint8_t* const* data, const size_t* sizes, int n,
                       std::vector<std::string>* json);
  size_t last_encoded_h2d_bytes() const;
  cudaStream_t stream() const { return stream_; }
  int worker_id() const { return worker_id_; }
  int device() const { return device_; }
  long launches() const;
  // cumulative host wall time (us) this worker spent in det / cls / rec and the images it processed: the per-stage
  // times the reference measures (`times`, ocr_worker.cpp:233-289) and then drops, kept for the pool's status
  void stage_totals(long long out[4]) const {
    for (int i = 0; i < 3; ++i) out[i] = stage_us_[i].load();
    out[3] = images_.load();
  }
  DetStage& det() { return *det_; }
  RecStage& rec() { return *rec_; }
  ClsStage* cls() { return cls_.get(); }
 private:
  int worker_id_, device_;
  WorkerOptions opt_;
  cudaStream_t stream_ = nullptr;
  std::unique_ptr<DetStage> det_;
  std::unique_ptr<ClsStage> cls_;
  std::unique_ptr<RecStage> rec_;
  void run_device(const std::vector<DevImg>& dimgs, std::vector<std::vector<WordOut>>* words);
  void recover_after_failure();
  ImageBatch batch_;
  std::unique_ptr<JpegBatch> jpeg_;
  DevBuf copy_;
  std::atomic<long long> stage_us_[3] = {{0}, {0}, {0}}, images_{0};
};

// jsoncpp-compatible compact writer pieces (StreamWriterBuilder, indentation "", emitUTF8 true)
std::string json_quote(const std::string& s);
std::string json_double(double v);

std::string result_json(int request_id, int worker_id, bool success, int width, int height, double ms,
                        const std::vector<WordOut>& words, const std::string& error);

std::vector<std::string> read_dict(const std::string& path);  // Utility::ReadDict, reference src/utility.cpp:32-48

}  // namespace b200ocr
