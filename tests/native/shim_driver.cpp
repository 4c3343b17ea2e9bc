// Drives the three reference-signature classes of include/paddle_ocr/b200ocr_shim.h the way OCRWorker::processRequest
// does (reference src/ocr_worker.cpp:228-300): det -> boundingRect crops (ROI views) -> cls -> rec, on a raw BGR file.
//   shim_driver <model_dir> <raw_bgr_file> <rows> <cols>
// Prints one JSON line {"boxes":[...],"labels":[...],"texts":[...],"scores":[...],"times":N}.  Without a CUDA device
// the constructors throw (no CPU fallback) and the program prints {"error": "..."} and exits 3.
#include <paddle_ocr/b200ocr_shim.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: shim_driver model_dir raw_bgr rows cols\n"); return 2; }
  const std::string models = argv[1];
  const int rows = atoi(argv[3]), cols = atoi(argv[4]);
  std::vector<uint8_t> pix(size_t(rows) * cols * 3);
  std::ifstream f(argv[2], std::ios::binary);
  f.read(reinterpret_cast<char*>(pix.data()), std::streamsize(pix.size()));
  if (!f) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
  cv::Mat img(rows, cols, pix.data(), size_t(cols) * 3);
  try {
    // the worker's settings (reference src/ocr_worker.cpp:21-63)
    PaddleOCR::DBDetector det(models + "/det", true, 0, 4000, 2, false, "max", 512, 0.2, 0.4, 1.8, "fast", false, false, "fp32");
    PaddleOCR::Classifier cls(models + "/cls", true, 0, 4000, 1, false, 0.98, false, "fp32", 8);
    PaddleOCR::CRNNRecognizer rec(models + "/rec", true, 0, 4000, 2, false, models + "/rec/ppocr_keys_v1.txt", false, "fp32",
                                  16, 28, 192);
    std::vector<std::vector<std::vector<int>>> boxes;
    std::vector<double> times;
    det.Run(img, boxes, times);
    std::vector<cv::Mat> crops;
    for (auto& b : boxes) {
      int x0 = b[0][0], x1 = b[0][0], y0 = b[0][1], y1 = b[0][1];
      for (auto& p : b) { x0 = std::min(x0, p[0]); x1 = std::max(x1, p[0]); y0 = std::min(y0, p[1]); y1 = std::max(y1, p[1]); }
      x0 = std::max(x0, 0); y0 = std::max(y0, 0);
      x1 = std::min(x1 + 1, cols); y1 = std::min(y1 + 1, rows);
      crops.push_back(img.roi(x0, y0, x1 - x0, y1 - y0));
    }
    std::vector<int> labels(crops.size(), 0);
    std::vector<float> cscores(crops.size(), 0.f);
    cls.Run(crops, labels, cscores, times);
    std::vector<std::string> texts(crops.size(), "");
    std::vector<float> scores(crops.size(), 0.f);
    rec.Run(crops, texts, scores, times);
    std::cout << "{\"boxes\":[";
    for (size_t i = 0; i < boxes.size(); ++i) {
      std::cout << (i ? "," : "") << "[";
      for (int k = 0; k < 4; ++k) std::cout << (k ? "," : "") << "[" << boxes[i][k][0] << "," << boxes[i][k][1] << "]";
      std::cout << "]";
    }
    std::cout << "],\"labels\":[";
    for (size_t i = 0; i < labels.size(); ++i) std::cout << (i ? "," : "") << labels[i];
    std::cout << "],\"texts\":[";
    for (size_t i = 0; i < texts.size(); ++i) {
      std::cout << (i ? "," : "") << "\"";
      for (char ch : texts[i]) { if (ch == '"' || ch == '\\') std::cout << '\\'; std::cout << ch; }
      std::cout << "\"";
    }
    std::cout << "],\"scores\":[";
    for (size_t i = 0; i < scores.size(); ++i) std::cout << (i ? "," : "") << scores[i];
    std::cout << "],\"times\":" << times.size() << "}" << std::endl;
  } catch (const std::exception& e) {
    std::string m = e.what();
    for (char& ch : m) if (ch == '"') ch = '\'';
    std::cout << "{\"error\":\"" << m << "\"}" << std::endl;
    return 3;
  }
  return 0;
}
