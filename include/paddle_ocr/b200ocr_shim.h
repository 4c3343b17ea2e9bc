// Header-only replacement for the reference's three stage classes over the C ABI of b200ocr.h.
//
// Same namespace, class names, constructor argument lists and Run() signatures as
//   include/paddle_ocr/ocr_det.h:60-69,95-97   (DBDetector)
//   include/paddle_ocr/ocr_cls.h:57-62,81-82   (Classifier)
//   include/paddle_ocr/ocr_rec.h:61-68,92-95   (CRNNRecognizer)
// of sssxyd/cpp-paddle-ocr, so that src/ocr_worker.cpp and everything above it compile unchanged against this header
// instead of the three originals (link with -lb200ocr instead of paddle_inference).  Needs <opencv2/core.hpp> only for
// cv::Mat (data, rows, cols, step of an 8-bit BGR image).
//
// Behaviour kept from the reference: Run() is noexcept; `boxes` is replaced; cls_labels / cls_scores / rec_texts /
// rec_text_scores are written by index into vectors the CALLER has sized to img_list.size(); `times` gets exactly
// three values appended (pre-process, inference, post-process, in ms).  Differences: a missing model throws
// std::runtime_error from the constructor instead of exit(1); use_gpu / gpu_mem / cpu_math_library_num_threads /
// use_mkldnn / use_tensorrt / precision are accepted and ignored (there is no CPU path).
#pragma once
#include <opencv2/core.hpp>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200ocr.h"

namespace PaddleOCR {

namespace b200ocr_detail {
inline b200ocr_image as_image(const cv::Mat& m) {
  b200ocr_image im;
  im.data = m.data;
  im.rows = m.rows;
  im.cols = m.cols;
  im.step = size_t(m.step);
  return im;
}
inline std::vector<b200ocr_image> as_images(const std::vector<cv::Mat>& v) {
  std::vector<b200ocr_image> out;
  out.reserve(v.size());
  for (const cv::Mat& m : v) out.push_back(as_image(m));
  return out;
}
[[noreturn]] inline void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + b200ocr_last_error()); }
}  // namespace b200ocr_detail

class DBDetector {
 public:
  explicit DBDetector(const std::string& model_dir, const bool& use_gpu, const int& gpu_id, const int& gpu_mem,
                      const int& cpu_math_library_num_threads, const bool& use_mkldnn, const std::string& limit_type,
                      const int& limit_side_len, const double& det_db_thresh, const double& det_db_box_thresh,
                      const double& det_db_unclip_ratio, const std::string& det_db_score_mode, const bool& use_dilation,
                      const bool& use_tensorrt, const std::string& precision) {
    b200ocr_det_config c;
    c.model_dir = model_dir.c_str();
    c.use_gpu = use_gpu; c.gpu_id = gpu_id; c.gpu_mem = gpu_mem;
    c.cpu_math_library_num_threads = cpu_math_library_num_threads; c.use_mkldnn = use_mkldnn;
    c.limit_type = limit_type.c_str(); c.limit_side_len = limit_side_len;
    c.det_db_thresh = det_db_thresh; c.det_db_box_thresh = det_db_box_thresh; c.det_db_unclip_ratio = det_db_unclip_ratio;
    c.det_db_score_mode = det_db_score_mode.c_str();
    c.use_dilation = use_dilation; c.use_tensorrt = use_tensorrt; c.precision = precision.c_str();
    if (b200ocr_det_create(&c, &h_) != B200OCR_OK) b200ocr_detail::fail("DBDetector");
  }
  ~DBDetector() { b200ocr_det_destroy(h_); }
  DBDetector(const DBDetector&) = delete;
  DBDetector& operator=(const DBDetector&) = delete;

  void Run(const cv::Mat& img, std::vector<std::vector<std::vector<int>>>& boxes, std::vector<double>& times) noexcept {
    int32_t buf[1000 * 8];
    int n = 0;
    double t[3] = {0, 0, 0};
    const b200ocr_image im = b200ocr_detail::as_image(img);
    if (b200ocr_det_run(h_, &im, buf, 1000, &n, t) != B200OCR_OK) n = 0;
    std::vector<std::vector<std::vector<int>>> out(size_t(n), std::vector<std::vector<int>>(4, std::vector<int>(2)));
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 4; ++k) { out[i][k][0] = buf[i * 8 + 2 * k]; out[i][k][1] = buf[i * 8 + 2 * k + 1]; }
    boxes = std::move(out);                // the reference replaces `boxes` (src/ocr_det.cpp:161)
    times.insert(times.end(), t, t + 3);   // and appends three times (src/ocr_det.cpp:168-175)
  }

 private:
  b200ocr_det_t h_ = nullptr;
};

class Classifier {
 public:
  explicit Classifier(const std::string& model_dir, const bool& use_gpu, const int& gpu_id, const int& gpu_mem,
                      const int& cpu_math_library_num_threads, const bool& use_mkldnn, const double& cls_thresh,
                      const bool& use_tensorrt, const std::string& precision, const int& cls_batch_num) {
    b200ocr_cls_config c;
    c.model_dir = model_dir.c_str();
    c.use_gpu = use_gpu; c.gpu_id = gpu_id; c.gpu_mem = gpu_mem;
    c.cpu_math_library_num_threads = cpu_math_library_num_threads; c.use_mkldnn = use_mkldnn;
    c.cls_thresh = cls_thresh; c.use_tensorrt = use_tensorrt; c.precision = precision.c_str();
    c.cls_batch_num = cls_batch_num;
    this->cls_thresh = cls_thresh;
    if (b200ocr_cls_create(&c, &h_) != B200OCR_OK) b200ocr_detail::fail("Classifier");
  }
  ~Classifier() { b200ocr_cls_destroy(h_); }
  Classifier(const Classifier&) = delete;
  Classifier& operator=(const Classifier&) = delete;

  // cls_labels / cls_scores must already hold img_list.size() elements (src/ocr_cls.cpp:97-98 writes by index)
  void Run(const std::vector<cv::Mat>& img_list, std::vector<int>& cls_labels, std::vector<float>& cls_scores,
           std::vector<double>& times) noexcept {
    double t[3] = {0, 0, 0};
    const std::vector<b200ocr_image> v = b200ocr_detail::as_images(img_list);
    if (!v.empty() && cls_labels.size() >= v.size() && cls_scores.size() >= v.size())
      b200ocr_cls_run(h_, v.data(), int(v.size()), cls_labels.data(), cls_scores.data(), t);
    times.insert(times.end(), t, t + 3);
  }

  double cls_thresh = 0.9;  // public like the reference's member (ocr_cls.h:76); stored and, like there, never consulted

 private:
  b200ocr_cls_t h_ = nullptr;
};

class CRNNRecognizer {
 public:
  explicit CRNNRecognizer(const std::string& model_dir, const bool& use_gpu, const int& gpu_id, const int& gpu_mem,
                          const int& cpu_math_library_num_threads, const bool& use_mkldnn, const std::string& label_path,
                          const bool& use_tensorrt, const std::string& precision, const int& rec_batch_num,
                          const int& rec_img_h, const int& rec_img_w) {
    b200ocr_rec_config c;
    c.model_dir = model_dir.c_str();
    c.use_gpu = use_gpu; c.gpu_id = gpu_id; c.gpu_mem = gpu_mem;
    c.cpu_math_library_num_threads = cpu_math_library_num_threads; c.use_mkldnn = use_mkldnn;
    c.label_path = label_path.c_str(); c.use_tensorrt = use_tensorrt; c.precision = precision.c_str();
    c.rec_batch_num = rec_batch_num; c.rec_img_h = rec_img_h; c.rec_img_w = rec_img_w;
    if (b200ocr_rec_create(&c, &h_) != B200OCR_OK) b200ocr_detail::fail("CRNNRecognizer");
  }
  ~CRNNRecognizer() { b200ocr_rec_destroy(h_); }
  CRNNRecognizer(const CRNNRecognizer&) = delete;
  CRNNRecognizer& operator=(const CRNNRecognizer&) = delete;

  // rec_texts / rec_text_scores must already hold img_list.size() elements (src/ocr_rec.cpp:126-127 writes by index;
  // lines that decode to nothing keep the caller's "" / 0)
  void Run(const std::vector<cv::Mat>& img_list, std::vector<std::string>& rec_texts, std::vector<float>& rec_text_scores,
           std::vector<double>& times) noexcept {
    double t[3] = {0, 0, 0};
    const std::vector<b200ocr_image> v = b200ocr_detail::as_images(img_list);
    if (!v.empty() && rec_texts.size() >= v.size() && rec_text_scores.size() >= v.size()) {
      std::vector<char*> s(v.size(), nullptr);
      std::vector<float> sc(v.size(), 0.f);
      if (b200ocr_rec_run(h_, v.data(), int(v.size()), s.data(), sc.data(), t) == B200OCR_OK) {
        for (size_t i = 0; i < v.size(); ++i) {
          if (s[i] && s[i][0]) { rec_texts[i] = s[i]; rec_text_scores[i] = sc[i]; }
          b200ocr_free(s[i]);
        }
      }
    }
    times.insert(times.end(), t, t + 3);
  }

 private:
  b200ocr_rec_t h_ = nullptr;
};

}  // namespace PaddleOCR
