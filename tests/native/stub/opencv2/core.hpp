// Minimal stand-in for <opencv2/core.hpp> (OpenCV's C++ headers are not in this image): just the four cv::Mat fields
// the shim reads, for an 8-bit 3-channel image that borrows its pixels.  Test infrastructure only.
#pragma once
#include <cstddef>
#include <cstdint>
namespace cv {
struct Mat {
  uint8_t* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
  Mat() = default;
  Mat(int r, int c, uint8_t* p, size_t s) : data(p), rows(r), cols(c), step(s) {}
  bool empty() const { return !data || rows <= 0 || cols <= 0; }
  // ROI view, like cv::Mat::operator()(cv::Rect)
  Mat roi(int x, int y, int w, int h) const { return Mat(h, w, data + size_t(y) * step + size_t(x) * 3, step); }
};
}  // namespace cv
