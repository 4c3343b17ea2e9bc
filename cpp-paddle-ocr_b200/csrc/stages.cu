#include "stages.h"
#include "jpeg.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <stdexcept>

namespace b200ocr {

namespace {
using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }
}  // namespace

// ------------------------------------------------------------------------------------------------ buffers
DevBuf::~DevBuf() {
  if (p) retired_.push_back(p);
  for (void* q : retired_) {
    if (pinned) cudaFreeHost(q); else cudaFree(q);
  }
}
void DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return;
  if (p) {
    retired_.push_back(p);
    p = nullptr;
  }
  size_t want = std::max(bytes * 2, size_t(4096));  // doubling keeps the retired allocations below the live one
  if (pinned) cuda_check(cudaMallocHost(&p, want), "cudaMallocHost");
  else cuda_check(cudaMalloc(&p, want), "cudaMalloc");
  cap = want;
}

void ImageBatch::upload(const HostImage* imgs, int n, cudaStream_t s) {
  size_t total = 0;
  std::vector<size_t> off(n);
  for (int i = 0; i < n; ++i) {
    off[i] = total;
    total += (size_t(imgs[i].rows) * imgs[i].cols * 3 + 255) & ~size_t(255);
  }
  dev_.ensure(total);
  imgs_.resize(n);
  bytes_ = 0;
  // images that follow each other in host memory the way they do in the device buffer (a caller's frame array) travel
  // as ONE copy: the run [run0, i) is flushed when image i does not continue it
  int run0 = 0;
  size_t run_bytes = 0;
  auto flush = [&](int upto) {
    if (run_bytes) cuda_check(cudaMemcpyAsync(imgs_[run0].p, imgs[run0].data, run_bytes, cudaMemcpyHostToDevice, s), "image upload");
    run0 = upto; run_bytes = 0;
  };
  for (int i = 0; i < n; ++i) {
    DevImg d;
    d.p = dev_.as<uint8_t>() + off[i];
    d.rows = imgs[i].rows; d.cols = imgs[i].cols; d.stride = long(imgs[i].cols) * 3;
    const size_t row = size_t(d.cols) * 3, size = row * d.rows;
    imgs_[i] = d;
    bytes_ += size;
    if (imgs[i].step != row) {
      flush(i + 1);
      cuda_check(cudaMemcpy2DAsync(d.p, row, imgs[i].data, imgs[i].step, row, d.rows, cudaMemcpyHostToDevice, s), "image upload");
      continue;
    }
    if (run_bytes && (imgs[i].data != imgs[run0].data + run_bytes || off[i] != off[run0] + run_bytes)) flush(i);
    run_bytes += size;
    if ((size & 255) != 0) flush(i + 1);   // the device slot is padded to 256 bytes: the next image cannot continue this run
  }
  flush(n);
}

// ------------------------------------------------------------------------------------------------ det
void DetStage::resized_dims(int rows, int cols, const std::string& limit_type, int limit_side_len, int* rh, int* rw,
                            float* ratio_h, float* ratio_w) {
  // ResizeImgType0::Run, reference src/preprocess_op.cpp:57-93 (float32 arithmetic, C round())
  const int w = cols, h = rows;
  float ratio = 1.f;
  if (limit_type == "min") {
    if (std::min(h, w) < limit_side_len) ratio = h < w ? float(limit_side_len) / float(h) : float(limit_side_len) / float(w);
  } else {
    if (std::max(h, w) > limit_side_len) ratio = h > w ? float(limit_side_len) / float(h) : float(limit_side_len) / float(w);
  }
  int resize_h = int(float(h) * ratio);
  int resize_w = int(float(w) * ratio);
  resize_h = std::max(int(std::round(float(resize_h) / 32) * 32), 32);
  resize_w = std::max(int(std::round(float(resize_w) / 32) * 32), 32);
  *rh = resize_h; *rw = resize_w;
  *ratio_h = float(resize_h) / float(h);
  *ratio_w = float(resize_w) / float(w);
}

DetStage::DetStage(const std::string& model_dir, int device, const DetParams& p)
    : net_(model_dir, device, NetOptions()), p_(p) {
  if (net_.kind() != "det") throw std::runtime_error("model in " + model_dir + " is not a DB detector graph");
  if (p.det_db_score_mode != "fast" && p.det_db_score_mode != "slow")
    throw std::invalid_argument("det_db_score_mode \"" + p.det_db_score_mode + "\" is neither \"fast\" nor \"slow\"");
  // cv::threshold on 8-bit data floors the threshold: bit = cbuf > floor(thresh * 255)  (src/ocr_det.cpp:151-154)
  thresh_u8_ = int(std::floor(double(float(p.det_db_thresh)) * 255));
}

void DetStage::run(const std::vector<DevImg>& imgs, std::vector<std::vector<Box>>* boxes, cudaStream_t s,
                   std::vector<double>* times) {
  boxes->assign(imgs.size(), {});
  prof_n = prof_h = prof_w = 0;
  std::map<std::pair<int, int>, std::vector<int>> groups;  // resized (h, w) -> image indices
  for (size_t i = 0; i < imgs.size(); ++i) {
    int rh, rw;
    float a, b;
    resized_dims(imgs[i].rows, imgs[i].cols, p_.limit_type, p_.limit_side_len, &rh, &rw, &a, &b);
    groups[{rh, rw}].push_back(int(i));
  }
  double t[3] = {0, 0, 0};
  for (auto& g : groups)
    for (size_t b0 = 0; b0 < g.second.size(); b0 += size_t(max_batch)) {
      std::vector<int> idx(g.second.begin() + b0, g.second.begin() + std::min(g.second.size(), b0 + size_t(max_batch)));
      run_group(imgs, idx, g.first.first, g.first.second, boxes, s, t);
    }
  if (times) times->insert(times->end(), t, t + 3);
}

void DetStage::run_group(const std::vector<DevImg>& imgs, const std::vector<int>& idx, int rh, int rw,
                         std::vector<std::vector<Box>>* boxes, cudaStream_t s, double* t_ms) {
  const int n = int(idx.size());
  auto t0 = Clock::now();
  h_items_.ensure(sizeof(DetPreItem) * n);
  h_info_.ensure(sizeof(DbImageInfo) * n);
  items_.ensure(sizeof(DetPreItem) * n);
  info_.ensure(sizeof(DbImageInfo) * n);
  for (int k = 0; k < n; ++k) {
    const DevImg& im = imgs[idx[k]];
    h_items_.as<DetPreItem>()[k] = DetPreItem{im.p, im.cols, im.rows, im.stride};
    int a, b;
    DbImageInfo inf;
    resized_dims(im.rows, im.cols, p_.limit_type, p_.limit_side_len, &a, &b, &inf.ratio_h, &inf.ratio_w);
    inf.src_h = im.rows; inf.src_w = im.cols;
    h_info_.as<DbImageInfo>()[k] = inf;
  }
  if (long(n) * rh * rw > long(prof_n) * prof_h * prof_w) { prof_n = n; prof_h = rh; prof_w = rw; }
  __half* in = net_.prepare(n, rh, rw, nullptr, s);
  cuda_check(cudaMemcpyAsync(items_.p, h_items_.p, sizeof(DetPreItem) * n, cudaMemcpyHostToDevice, s), "det items");
  cuda_check(cudaMemcpyAsync(info_.p, h_info_.p, sizeof(DbImageInfo) * n, cudaMemcpyHostToDevice, s), "det info");
  static const float mean[3] = {0.485f, 0.456f, 0.406f};                     // reference ocr_det.h:121
  static const float scale[3] = {1 / 0.229f, 1 / 0.224f, 1 / 0.225f};        // reference ocr_det.h:122
  // resize + normalise run inside the first convolution when the graph's stem allows it (kernels_simt.cu: fused_stem_kernel)
  const bool fuse = net_.stem_fusable();
  StemSource src;
  src.kind = 1; src.items = items_.p; src.np = make_norm(mean, scale);
  if (!fuse) launch_det_preprocess(items_.as<DetPreItem>(), n, rh, rw, src.np, in, s);
  t_ms[0] += ms_since(t0);
  t0 = Clock::now();
  net_.run(s, thresh_u8_, fuse ? &src : nullptr);
  launches += net_.launches_per_run() + (fuse ? 0 : 1);
  t_ms[1] += ms_since(t0);
  t0 = Clock::now();
  DbPostParams pp;
  pp.n = n; pp.h = rh; pp.w = rw;
  pp.box_thresh = float(p_.det_db_box_thresh);
  pp.unclip_ratio = float(p_.det_db_unclip_ratio);
  pp.max_candidates = 1000;
  pp.score_slow = p_.det_db_score_mode == "slow";  // PolygonScoreAcc instead of BoxScoreFast (postprocess_op.cpp:285-288)
  ws_.ensure(dbpost_workspace_bytes(pp));
  counts_.ensure(sizeof(int) * n);
  boxes_.ensure(sizeof(DbBox) * size_t(n) * pp.max_candidates);
  h_counts_.ensure(sizeof(int) * n);
  h_boxes_.ensure(sizeof(DbBox) * size_t(n) * pp.max_candidates);
  const uint8_t* bitmap = net_.out_bitmap();
  if (p_.use_dilation) {
    dil_.ensure(size_t(n) * rh * rw);
    launch_dilate2x2(bitmap, dil_.as<uint8_t>(), n, rh, rw, s);
    bitmap = dil_.as<uint8_t>();
    ++launches;
  }
  launch_dbpost(pp, net_.out_f32(), bitmap, info_.as<DbImageInfo>(), ws_.p, counts_.as<int>(), boxes_.as<DbBox>(), s);
  launches += 8 + pp.score_slow;
  cuda_check(cudaMemcpyAsync(h_counts_.p, counts_.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s), "det counts");
  host_wait(s, "det post-process");
  // second copy sized by what was found (boxes of image k live at [k * max_candidates, +count))
  int maxc = 0;
  for (int k = 0; k < n; ++k) maxc = std::max(maxc, h_counts_.as<int>()[k]);
  if (maxc > 0) {
    cuda_check(cudaMemcpy2DAsync(h_boxes_.p, sizeof(DbBox) * maxc, boxes_.p, sizeof(DbBox) * pp.max_candidates,
                                 sizeof(DbBox) * maxc, n, cudaMemcpyDeviceToHost, s), "det boxes");
    host_wait(s, "det boxes");
  }
  for (int k = 0; k < n; ++k) {
    std::vector<Box>& out = (*boxes)[idx[k]];
    const DbBox* b = h_boxes_.as<DbBox>() + size_t(k) * maxc;
    for (int c = 0; c < h_counts_.as<int>()[k]; ++c) {
      if (!b[c].valid) continue;
      Box bx;
      for (int j = 0; j < 8; ++j) bx[j] = b[c].pts[j];
      out.push_back(bx);
    }
  }
  t_ms[2] += ms_since(t0);
}

// ------------------------------------------------------------------------------------------------ cls
namespace {
// ClsResizeImg / CrnnResizeImg width rule (reference src/preprocess_op.cpp:104-112, :128-134)
int resize_width(int img_h, int img_w, int crop_w, int crop_h) {
  const float ratio = float(crop_w) / float(crop_h);
  if (std::ceil(float(img_h) * ratio) > float(img_w)) return img_w;
  return int(std::ceil(float(img_h) * ratio));
}
const float kMean05[3] = {0.5f, 0.5f, 0.5f};             // reference ocr_cls.h:93, ocr_rec.h:108
const float kScale2[3] = {1 / 0.5f, 1 / 0.5f, 1 / 0.5f};  // reference ocr_cls.h:94, ocr_rec.h:109

__global__ void cls_argmax_kernel(const float* __restrict__ prob, int n, int ncls, int* __restrict__ label,
                                  float* __restrict__ score) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // Utility::argmax = first maximum (std::max_element), reference src/ocr_cls.cpp:90-96
  int best = 0;
  float mx = prob[long(i) * ncls];
  for (int c = 1; c < ncls; ++c) {
    const float v = prob[long(i) * ncls + c];
    if (v > mx) { mx = v; best = c; }
  }
  label[i] = best;
  score[i] = mx;
}

// One CTA per image: ROIs of that image are rotated one after the other (they may overlap).
struct RotItem { uint8_t* img; long stride; int x, y, w, h; int image; };
__global__ void __launch_bounds__(1024)
rotate_seq_kernel(const RotItem* __restrict__ items, const int* __restrict__ first, const int* __restrict__ labels) {
  const int img = blockIdx.x;
  for (int k = first[img]; k < first[img + 1]; ++k) {
    if (labels[k] == 1) {
      const RotItem it = items[k];
      const long total = long(it.w) * it.h, half = total / 2;
      for (long t = threadIdx.x; t < half; t += blockDim.x) {
        const long u = total - 1 - t;
        const int ay = int(t / it.w), ax = int(t - long(ay) * it.w);
        const int by = int(u / it.w), bx = int(u - long(by) * it.w);
        uint8_t* a = it.img + long(it.y + ay) * it.stride + long(it.x + ax) * 3;
        uint8_t* b = it.img + long(it.y + by) * it.stride + long(it.x + bx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const uint8_t v = a[c]; a[c] = b[c]; b[c] = v; }
      }
    }
    __syncthreads();
  }
}
}  // namespace

ClsStage::ClsStage(const std::string& model_dir, int device, int cls_batch_num, float cls_thresh)
    : net_(model_dir, device, NetOptions()), batch_num_(cls_batch_num), thresh_(cls_thresh) {
  if (net_.kind() != "cls") throw std::runtime_error("model in " + model_dir + " is not an angle classifier graph");
  (void)batch_num_; (void)thresh_;  // cls_thresh is stored but never consulted by the reference either
}

void ClsStage::run(const std::vector<DevImg>& imgs, const std::vector<Roi>& rois, std::vector<int>* labels,
                   std::vector<float>* scores, cudaStream_t s, bool fetch_to_host, std::vector<double>* times) {
  const int n = int(rois.size());
  double t[3] = {0, 0, 0};
  labels_.ensure(sizeof(int) * std::max(n, 1));
  probs_.ensure(sizeof(float) * std::max(n, 1));
  // Every row is resized/padded to the fixed 48x192 input on its own, so the reference's batches of
  // cls_batch_num (src/ocr_cls.cpp:35-62) can be merged into larger launches without changing any value.
  // all items are staged once: the pinned buffer must not be rewritten while an earlier async copy may still read it
  h_items_.ensure(sizeof(CropItem) * std::max(n, 1));
  items_.ensure(sizeof(CropItem) * std::max(n, 1));
  for (int k = 0; k < n; ++k) {
    const Roi& r = rois[k];
    const DevImg& im = imgs[r.img];
    h_items_.as<CropItem>()[k] = CropItem{im.p, im.stride, r.x, r.y, r.w, r.h, resize_width(48, 192, r.w, r.h), 192};
  }
  if (n > 0) cuda_check(cudaMemcpyAsync(items_.p, h_items_.p, sizeof(CropItem) * n, cudaMemcpyHostToDevice, s), "cls items");
  for (int b0 = 0; b0 < n; b0 += max_batch) {
    const int nb = std::min(max_batch, n - b0);
    auto t0 = Clock::now();
    if (b0 == 0) { prof_n = nb; prof_h = 48; prof_w = 192; }
    __half* in = net_.prepare(nb, 48, 192, nullptr, s);
    // pad value 0.0: the classifier pads AFTER normalisation (src/ocr_cls.cpp:52-56)
    const bool fuse = net_.stem_fusable();
    StemSource src;
    src.kind = 2; src.items = items_.as<CropItem>() + b0; src.np = make_norm(kMean05, kScale2); src.pad_value = 0.f;
    if (!fuse) launch_crop_preprocess(items_.as<CropItem>() + b0, nb, 48, 192, src.np, 0.f, in, s);
    t[0] += ms_since(t0);
    t0 = Clock::now();
    net_.run(s, -1, fuse ? &src : nullptr);
    t[1] += ms_since(t0);
    t0 = Clock::now();
    const int ncls = net_.plan().layers.back().cout;
    cls_argmax_kernel<<<(nb + 127) / 128, 128, 0, s>>>(net_.out_f32(), nb, ncls, labels_.as<int>() + b0,
                                                       probs_.as<float>() + b0);
    launches += net_.launches_per_run() + 2;
    t[2] += ms_since(t0);
  }
  if (fetch_to_host && n > 0) {
    auto t0 = Clock::now();
    h_out_.ensure(size_t(n) * 8);
    cuda_check(cudaMemcpyAsync(h_out_.p, labels_.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s), "cls labels");
    cuda_check(cudaMemcpyAsync(h_out_.as<uint8_t>() + size_t(n) * 4, probs_.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s),
               "cls scores");
    host_wait(s, "cls");
    if (labels) labels->assign(h_out_.as<int>(), h_out_.as<int>() + n);
    if (scores) scores->assign(reinterpret_cast<float*>(h_out_.as<uint8_t>() + size_t(n) * 4),
                               reinterpret_cast<float*>(h_out_.as<uint8_t>() + size_t(n) * 4) + n);
    t[2] += ms_since(t0);
  }
  if (times) times->insert(times->end(), t, t + 3);
}

void ClsStage::rotate_rois(const std::vector<DevImg>& imgs, const std::vector<Roi>& rois, cudaStream_t s) {
  const int n = int(rois.size());
  if (!n) return;
  // rois arrive grouped by image in ROI order; first[i] = index of image i's first ROI
  const int nimg = int(imgs.size());
  const size_t bytes = sizeof(RotItem) * n + sizeof(int) * (nimg + 1);
  h_rot_.ensure(bytes);
  rot_.ensure(bytes);
  RotItem* it = h_rot_.as<RotItem>();
  int* first = reinterpret_cast<int*>(it + n);
  int k = 0;
  for (int i = 0; i < nimg; ++i) {
    first[i] = k;
    while (k < n && rois[k].img == i) {
      it[k] = RotItem{imgs[i].p, imgs[i].stride, rois[k].x, rois[k].y, rois[k].w, rois[k].h, i};
      ++k;
    }
  }
  first[nimg] = k;
  if (k != n) throw std::runtime_error("rotate_rois: ROIs must be grouped by ascending image index");
  cuda_check(cudaMemcpyAsync(rot_.p, h_rot_.p, bytes, cudaMemcpyHostToDevice, s), "rotate items");
  rotate_seq_kernel<<<nimg, 1024, 0, s>>>(rot_.as<RotItem>(), reinterpret_cast<const int*>(rot_.as<RotItem>() + n),
                                          labels_.as<int>());
  ++launches;
}

// ------------------------------------------------------------------------------------------------ rec
std::vector<std::string> read_dict(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("no such label file: " + path);
  std::vector<std::string> v;
  std::string line;
  while (std::getline(in, line)) v.emplace_back(line);
  return v;
}

RecStage::RecStage(const std::string& model_dir, int device, const std::string& label_path, int rec_batch_num,
                   int rec_img_h, int rec_img_w)
    : net_(model_dir, device, NetOptions()), batch_num_(rec_batch_num), img_h_(rec_img_h), img_w_(rec_img_w) {
  if (net_.kind() != "rec") throw std::runtime_error("model in " + model_dir + " is not a CRNN/SVTR recognizer graph");
  // label list = ["#"] + dictionary lines + [" "]  (reference include/paddle_ocr/ocr_rec.h:82-84)
  label_list_ = read_dict(label_path);
  label_list_.insert(label_list_.begin(), "#");
  label_list_.emplace_back(" ");
  const int ncls = net_.plan().layers.back().cout;
  if (int(label_list_.size()) != ncls)
    throw std::runtime_error("label list has " + std::to_string(label_list_.size()) + " entries but the CTC head has " +
                             std::to_string(ncls) + " classes");
}

void RecStage::run(const std::vector<DevImg>& imgs, const std::vector<std::vector<Roi>>& calls,
                   std::vector<std::vector<std::string>>* texts, std::vector<std::vector<float>>* scores, cudaStream_t s,
                   std::vector<double>* times) {
  double t[3] = {0, 0, 0};
  auto t0 = Clock::now();
  texts->assign(calls.size(), {});
  scores->assign(calls.size(), {});
  struct Row { int call, roi, width; CropItem item; };
  std::vector<Row> rows;
  for (size_t c = 0; c < calls.size(); ++c) {
    const std::vector<Roi>& rois = calls[c];
    const size_t m = rois.size();
    (*texts)[c].assign(m, std::string());
    (*scores)[c].assign(m, 0.f);
    // aspect-ratio sort (reference src/ocr_rec.cpp:35-40; Utility::argsort, src/utility.cpp:192-203)
    std::vector<float> ratio(m);
    for (size_t i = 0; i < m; ++i) ratio[i] = float(rois[i].w) / float(rois[i].h);
    std::vector<size_t> order(m);
    std::iota(order.begin(), order.end(), size_t(0));
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ratio[a] < ratio[b]; });
    for (size_t beg = 0; beg < m; beg += size_t(batch_num_)) {
      const size_t end = std::min(m, beg + size_t(batch_num_));
      float max_wh_ratio = float(img_w_ * 1.0 / img_h_);
      for (size_t k = beg; k < end; ++k) {
        const Roi& r = rois[order[k]];
        max_wh_ratio = std::max(max_wh_ratio, float(r.w * 1.0 / r.h));
      }
      const int img_w = int(float(img_h_) * max_wh_ratio);  // CrnnResizeImg: imgW = int(imgH * wh_ratio)
      const int batch_width = std::max(img_w_, img_w);
      // a batch narrower than rec_img_w would be padded by PermuteBatch's zero-filled tensor with 0.0 instead of
      // -1.0; int(imgH * (imgW/imgH)) equals imgW for the shipped settings, so that case is rejected, not emulated
      if (img_w < batch_width) throw std::runtime_error("rec_img_w/rec_img_h combination pads with two different values");
      for (size_t k = beg; k < end; ++k) {
        const Roi& r = rois[order[k]];
        const DevImg& im = imgs[r.img];
        Row row;
        row.call = int(c); row.roi = int(order[k]); row.width = batch_width;
        row.item = CropItem{im.p, im.stride, r.x, r.y, r.w, r.h, resize_width(img_h_, img_w, r.w, r.h), batch_width};
        rows.push_back(row);
      }
    }
  }
  // Rows of every reference batch keep the padded width of THEIR batch; rows of different widths share one launch
  // as a ragged batch (Net::prepare with per-row widths), which leaves every row's values unchanged.  Sorting by
  // width and cutting chunks where padding would exceed ~25% bounds the wasted columns.
  std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.width < b.width; });
  struct Chunk { size_t b0; int nb, wmax, T; size_t off; };
  std::vector<Chunk> chunks;
  size_t total_ids = 0;
  for (size_t b0 = 0; b0 < rows.size();) {
    size_t e = b0;
    long sum_w = 0;
    while (e < rows.size() && int(e - b0) < max_rows) {
      const long w_new = rows[e].width;
      if (e > b0 && double(sum_w + w_new) < min_fill * double(long(e - b0 + 1) * w_new)) break;  // too much padding
      if (e > b0 && long(e - b0 + 1) * w_new > max_cols) break;
      sum_w += w_new;
      ++e;
    }
    Chunk ch;
    ch.b0 = b0; ch.nb = int(e - b0); ch.wmax = rows[e - 1].width; ch.T = 0; ch.off = 0;
    chunks.push_back(ch);
    b0 = e;
  }
  last_chunks = int(chunks.size()); last_cols = 0; last_real_cols = 0;
  for (const Chunk& ch : chunks) last_cols += long(ch.nb) * ch.wmax;
  for (const Row& r : rows) last_real_cols += r.width;
  t[0] += ms_since(t0);
  t0 = Clock::now();
  h_items_.ensure(sizeof(CropItem) * std::max<size_t>(rows.size(), 1));
  items_.ensure(sizeof(CropItem) * std::max<size_t>(rows.size(), 1));
  for (size_t k = 0; k < rows.size(); ++k) h_items_.as<CropItem>()[k] = rows[k].item;
  if (!rows.empty())
    cuda_check(cudaMemcpyAsync(items_.p, h_items_.p, sizeof(CropItem) * rows.size(), cudaMemcpyHostToDevice, s), "rec items");
  // worst-case output size: T <= wmax / 8 + 1
  for (Chunk& ch : chunks) { ch.off = total_ids; total_ids += size_t(ch.nb) * (ch.wmax / 8 + 2); }
  cidx_.ensure(sizeof(int) * std::max<size_t>(total_ids, 1));
  clen_.ensure(sizeof(int) * std::max<size_t>(rows.size(), 1));
  cscore_.ensure(sizeof(float) * std::max<size_t>(rows.size(), 1));
  h_cidx_.ensure(sizeof(int) * std::max<size_t>(total_ids, 1));
  h_clen_.ensure(sizeof(int) * std::max<size_t>(rows.size(), 1));
  h_cscore_.ensure(sizeof(float) * std::max<size_t>(rows.size(), 1));
  std::vector<int> widths;
  prof_n = prof_h = prof_w = 0;
  for (Chunk& ch : chunks) {
    if (long(ch.nb) * ch.wmax > long(prof_n) * prof_w) { prof_n = ch.nb; prof_h = img_h_; prof_w = ch.wmax; }
    widths.resize(ch.nb);
    for (int k = 0; k < ch.nb; ++k) widths[k] = rows[ch.b0 + k].width;
    __half* in = net_.prepare(ch.nb, img_h_, ch.wmax, widths.data(), s);
    // pad value -1.0: CrnnResizeImg pads with u8 zeros BEFORE normalisation (src/preprocess_op.cpp:115-117)
    const bool fuse = net_.stem_fusable();
    StemSource src;
    src.kind = 2; src.items = items_.as<CropItem>() + ch.b0; src.np = make_norm(kMean05, kScale2); src.pad_value = -1.f;
    if (!fuse) launch_crop_preprocess(items_.as<CropItem>() + ch.b0, ch.nb, img_h_, ch.wmax, src.np, -1.f, in, s);
    net_.run(s, -1, fuse ? &src : nullptr);
    ch.T = net_.out_shape().w;
    if (size_t(ch.nb) * ch.T > size_t(ch.nb) * (ch.wmax / 8 + 2)) throw std::runtime_error("rec: unexpected sequence length");
    launch_ctc_collapse(net_.out_idx(), net_.out_f32(), ch.nb, ch.T, cidx_.as<int>() + ch.off, clen_.as<int>() + ch.b0,
                        cscore_.as<float>() + ch.b0, s);
    launches += net_.launches_per_run() + 2;
  }
  if (!rows.empty()) {
    cuda_check(cudaMemcpyAsync(h_cidx_.p, cidx_.p, sizeof(int) * total_ids, cudaMemcpyDeviceToHost, s), "rec ids");
    cuda_check(cudaMemcpyAsync(h_clen_.p, clen_.p, sizeof(int) * rows.size(), cudaMemcpyDeviceToHost, s), "rec len");
    cuda_check(cudaMemcpyAsync(h_cscore_.p, cscore_.p, sizeof(float) * rows.size(), cudaMemcpyDeviceToHost, s), "rec score");
    host_wait(s, "rec");
  }
  t[1] += ms_since(t0);
  t0 = Clock::now();
  for (const Chunk& ch : chunks)
    for (int k = 0; k < ch.nb; ++k) {
      const int len = h_clen_.as<int>()[ch.b0 + k];
      if (len == 0) continue;  // score is NaN in the reference -> the caller's "" / 0 stay (src/ocr_rec.cpp:122-125)
      std::string str;
      const int* ids = h_cidx_.as<int>() + ch.off + size_t(k) * ch.T;
      for (int j = 0; j < len; ++j) str += label_list_[ids[j]];
      const Row& r = rows[ch.b0 + k];
      (*texts)[r.call][r.roi] = std::move(str);
      (*scores)[r.call][r.roi] = h_cscore_.as<float>()[ch.b0 + k];
    }
  t[2] += ms_since(t0);
  if (times) times->insert(times->end(), t, t + 3);
}

// ------------------------------------------------------------------------------------------------ JSON
std::string json_quote(const std::string& s) {
  // jsoncpp valueToQuotedStringN with emitUTF8 = true
  std::string o = "\"";
  static const char* hex = "0123456789abcdef";
  for (unsigned char c : s) {
    switch (c) {
      case '\"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\b': o += "\\b"; break;
      case '\f': o += "\\f"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { o += "\\u00"; o += hex[c >> 4]; o += hex[c & 15]; }
        else o += char(c);
    }
  }
  return o + "\"";
}

std::string json_double(double v) {
  // jsoncpp valueToString(double, useSpecialFloats=false, precision=17, significantDigits)
  if (!std::isfinite(v)) return std::isnan(v) ? "null" : (v < 0 ? "-1e+9999" : "1e+9999");
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string s = buf;
  if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s += ".0";
  return s;
}

std::string result_json(int request_id, int worker_id, bool success, int width, int height, double ms,
                        const std::vector<WordOut>& words, const std::string& error) {
  // jsoncpp keeps object members in a std::map -> keys come out in alphabetical order
  std::string o = "{";
  if (!success) o += "\"error\":" + json_quote(error) + ",";
  o += "\"height\":" + std::to_string(height) + ",";
  o += "\"processing_time_ms\":" + json_double(ms) + ",";
  o += "\"request_id\":" + std::to_string(request_id) + ",";
  o += std::string("\"success\":") + (success ? "true" : "false") + ",";
  o += "\"width\":" + std::to_string(width) + ",";
  if (success) {
    o += "\"words\":[";
    for (size_t i = 0; i < words.size(); ++i) {
      const WordOut& w = words[i];
      if (i) o += ",";
      o += "{\"box\":[";
      for (int k = 0; k < 4; ++k) {
        if (k) o += ",";
        o += "[" + std::to_string(w.box[2 * k]) + "," + std::to_string(w.box[2 * k + 1]) + "]";
      }
      o += "],\"confidence\":" + json_double(double(w.confidence)) + ",\"text\":" + json_quote(w.text) + "}";
    }
    o += "],";
  }
  o += "\"worker_id\":" + std::to_string(worker_id) + "}";
  return o;
}

// ------------------------------------------------------------------------------------------------ worker
Worker::Worker(int worker_id, const std::string& model_dir, int device, const WorkerOptions& opt)
    : worker_id_(worker_id), device_(device), opt_(opt) {
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  cuda_check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  // reference src/ocr_worker.cpp:21-63
  DetParams dp;
  dp.limit_type = opt.det_limit_type;
  dp.limit_side_len = opt.det_limit_side_len;
  dp.det_db_thresh = opt.det_db_thresh;
  dp.det_db_box_thresh = opt.det_db_box_thresh;
  dp.det_db_unclip_ratio = opt.det_db_unclip_ratio;
  dp.det_db_score_mode = opt.det_db_score_mode;
  dp.use_dilation = opt.use_dilation;
  det_ = std::make_unique<DetStage>(model_dir + "/det", device, dp);
  if (opt.enable_cls) cls_ = std::make_unique<ClsStage>(model_dir + "/cls", device, opt.cls_batch_num, float(opt.cls_thresh));
  rec_ = std::make_unique<RecStage>(model_dir + "/rec", device, model_dir + "/rec/ppocr_keys_v1.txt", opt.rec_batch_num,
                                    opt.rec_img_h, opt.rec_img_w);
  // tuning knobs (launch granularity only; results do not depend on them)
  if (const char* v = getenv("B200OCR_DET_MAX_BATCH")) det_->max_batch = std::max(1, atoi(v));
  if (const char* v = getenv("B200OCR_CLS_MAX_BATCH")) { if (cls_) cls_->max_batch = std::max(1, atoi(v)); }
  if (const char* v = getenv("B200OCR_REC_MAX_COLS")) rec_->max_cols = std::max(1000L, atol(v));
  if (const char* v = getenv("B200OCR_REC_MAX_ROWS")) rec_->max_rows = std::max(1, atoi(v));
  if (const char* v = getenv("B200OCR_REC_FILL")) rec_->min_fill = std::min(1.0, std::max(0.1, atof(v)));
}

Worker::~Worker() {
  cudaSetDevice(device_);
  det_.reset(); cls_.reset(); rec_.reset();
  if (stream_) cudaStreamDestroy(stream_);
}

long Worker::launches() const {
  return det_->launches + rec_->launches + (cls_ ? cls_->launches : 0) + (jpeg_ ? jpeg_->launches : 0);
}

// det -> ROI -> (cls -> rotate) -> rec for one batch of device images (reference src/ocr_worker.cpp:228-300)
void Worker::run_device(const std::vector<DevImg>& dimgs, std::vector<std::vector<WordOut>>* words) {
  const int nb = int(dimgs.size());
  words->assign(nb, {});
  std::vector<std::vector<Box>> boxes;
  static const bool trace = getenv("B200OCR_TRACE") != nullptr;  // host wall time per phase, to stderr
  const auto t_begin = Clock::now();
  std::vector<double> t_det, t_cls, t_rec;
  det_->run(dimgs, &boxes, stream_, trace ? &t_det : nullptr);
  const double ms_det = ms_since(t_begin);
  // ROI = cv::boundingRect(points) & image (src/ocr_worker.cpp:244-259); boundingRect of integer-valued
  // float points is (minx, miny, maxx - minx + 1, maxy - miny + 1)
  std::vector<std::vector<Roi>> calls(nb);
  std::vector<Roi> all;
  for (int i = 0; i < nb; ++i)
    for (const Box& b : boxes[i]) {
      int minx = b[0], maxx = b[0], miny = b[1], maxy = b[1];
      for (int k = 1; k < 4; ++k) {
        minx = std::min(minx, b[2 * k]); maxx = std::max(maxx, b[2 * k]);
        miny = std::min(miny, b[2 * k + 1]); maxy = std::max(maxy, b[2 * k + 1]);
      }
      const int x0 = std::max(minx, 0), y0 = std::max(miny, 0);
      const int x1 = std::min(maxx + 1, dimgs[i].cols), y1 = std::min(maxy + 1, dimgs[i].rows);
      Roi r;
      r.img = i; r.x = x0; r.y = y0; r.w = x1 - x0; r.h = y1 - y0;
      if (r.w > 0 && r.h > 0) { calls[i].push_back(r); all.push_back(r); }
    }
  const double ms_roi = ms_since(t_begin);
  if (cls_ && !all.empty()) {
    cls_->run(dimgs, all, nullptr, nullptr, stream_, /*fetch_to_host=*/false, trace ? &t_cls : nullptr);
    cls_->rotate_rois(dimgs, all, stream_);
  }
  const double ms_cls = ms_since(t_begin);
  std::vector<std::vector<std::string>> texts;
  std::vector<std::vector<float>> scores;
  rec_->run(dimgs, calls, &texts, &scores, stream_, trace ? &t_rec : nullptr);
  {
    // det waits for its boxes, cls is only enqueued (its GPU time is waited for inside rec): host wall time per stage
    const double ms_end = ms_since(t_begin);
    stage_us_[0] += (long long)(ms_det * 1e3);
    stage_us_[1] += (long long)((ms_cls - ms_roi) * 1e3);
    stage_us_[2] += (long long)((ms_end - ms_cls) * 1e3);
    images_ += nb;
  }
  if (trace) {
    auto v = [](const std::vector<double>& t, int i) { return int(t.size()) > i ? t[i] : 0.0; };
    fprintf(stderr, "[b200ocr trace] worker %d: %d images, %zu rois | det %.2f (pre %.2f net %.2f post+sync %.2f) roi %.2f "
            "cls-enqueue %.2f (pre %.2f net %.2f post %.2f) rec %.2f (plan %.2f run+sync %.2f decode %.2f) ms; rec chunks %d, cols %ld (real %ld)\n",
            worker_id_, nb, all.size(), ms_det, v(t_det, 0), v(t_det, 1), v(t_det, 2), ms_roi - ms_det, ms_cls - ms_roi,
            v(t_cls, 0), v(t_cls, 1), v(t_cls, 2), ms_since(t_begin) - ms_cls, v(t_rec, 0), v(t_rec, 1), v(t_rec, 2),
            rec_->last_chunks, rec_->last_cols, rec_->last_real_cols);
  }
  for (int i = 0; i < nb; ++i) {
    std::vector<WordOut>& w = (*words)[i];
    // words[i] = (rec_texts[i], rec_scores[i], det_boxes[i])  (src/ocr_worker.cpp:293-300)
    for (size_t k = 0; k < texts[i].size(); ++k) w.push_back(WordOut{texts[i][k], scores[i][k], boxes[i][k]});
  }
}

void Worker::process_resident(const int* request_ids, const std::vector<DevImg>& resident, std::vector<std::string>* json) {
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  const int n = int(resident.size());
  json->assign(n, std::string());
  const auto t_start = Clock::now();
  std::vector<std::vector<WordOut>> words(n);
  std::vector<std::string> errors(n);
  auto run_range = [&](int b0, int nb) {
    std::vector<DevImg> dimgs(resident.begin() + b0, resident.begin() + b0 + nb);
    if (cls_) {  // rotations are in place: work on a copy
      size_t total = 0;
      for (auto& d : dimgs) total += (size_t(d.rows) * d.stride + 255) & ~size_t(255);
      copy_.ensure(total);
      size_t off = 0;
      for (auto& d : dimgs) {
        uint8_t* dst = copy_.as<uint8_t>() + off;
        cuda_check(cudaMemcpyAsync(dst, d.p, size_t(d.rows) * d.stride, cudaMemcpyDeviceToDevice, stream_), "batch copy");
        off += (size_t(d.rows) * d.stride + 255) & ~size_t(255);
        d.p = dst;
      }
    }
    std::vector<std::vector<WordOut>> w;
    run_device(dimgs, &w);
    for (int i = 0; i < nb; ++i) words[b0 + i] = std::move(w[i]);
  };
  for (int b0 = 0; b0 < n; b0 += opt_.max_batch) {
    const int nb = std::min(opt_.max_batch, n - b0);
    try {
      run_range(b0, nb);
    } catch (const std::exception& e) {
      // The reference handles one request at a time, so a failure there is isolated (src/ocr_worker.cpp:192-206).
      // Keep that: the images of the failed sub-batch are retried one by one and only the offender reports the error.
      recover_after_failure();
      if (nb == 1) { errors[b0] = e.what(); continue; }
      for (int i = b0; i < b0 + nb; ++i) {
        try {
          run_range(i, 1);
        } catch (const std::exception& e1) {
          recover_after_failure();
          words[i].clear();
          errors[i] = e1.what();
        }
      }
    }
  }
  const double ms = ms_since(t_start);
  for (int i = 0; i < n; ++i)
    (*json)[i] = result_json(request_ids[i], worker_id_, errors[i].empty(), resident[i].cols, resident[i].rows, ms, words[i],
                             errors[i]);
}

// After an exception inside a batch: drain the stream and drop the (non-sticky) CUDA error so that the retries start clean.
void Worker::recover_after_failure() {
  cudaStreamSynchronize(stream_);
  cudaGetLastError();
}

size_t Worker::last_encoded_h2d_bytes() const { return jpeg_ ? jpeg_->h2d_bytes() : 0; }

void Worker::process_encoded(const int* request_ids, const uint8_t* const* data, const size_t* sizes, int n,
                             std::vector<std::string>* json) {
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  json->assign(n, std::string());
  const auto t_start = Clock::now();
  if (!jpeg_) jpeg_ = std::make_unique<JpegBatch>();
  std::vector<DevImg> dec;
  std::vector<std::string> why;
  std::vector<int> live;
  std::vector<std::string> errors(n);
  std::vector<std::vector<WordOut>> words(n);
  try {
    static const bool trace = getenv("B200OCR_TRACE") != nullptr;
    const auto t_dec = Clock::now();
    jpeg_->decode(data, sizes, n, stream_, &dec, &why);
    if (trace) {  // (the synchronisation is part of the trace only)
      const double enq = ms_since(t_dec);
      cudaStreamSynchronize(stream_);
      fprintf(stderr, "[b200ocr trace] worker %d: jpeg decode of %d files: enqueue %.2f ms, done after %.2f ms, %zu bytes uploaded\n",
              worker_id_, n, enq, ms_since(t_dec), jpeg_->h2d_bytes());
    }
  } catch (const std::exception& e) {
    recover_after_failure();
    dec.assign(n, DevImg());
    why.assign(n, e.what());
  }
  for (int i = 0; i < n; ++i) {
    if (dec[i].p) live.push_back(i);
    else errors[i] = (!data[i] || sizes[i] == 0) ? "Empty image data provided" : "Unsupported image encoding: " + why[i];
  }
  auto run_range = [&](size_t b0, int nb) {
    std::vector<DevImg> dimgs(nb);
    for (int k = 0; k < nb; ++k) dimgs[k] = dec[live[b0 + k]];
    std::vector<std::vector<WordOut>> w;
    run_device(dimgs, &w);   // (the classifier's in-place rotations land in the decoder's own output buffer)
    for (int k = 0; k < nb; ++k) words[live[b0 + k]] = std::move(w[k]);
  };
  for (size_t b0 = 0; b0 < live.size(); b0 += size_t(opt_.max_batch)) {
    const int nb = int(std::min(live.size() - b0, size_t(opt_.max_batch)));
    try {
      run_range(b0, nb);
    } catch (const std::exception& e) {
      recover_after_failure();
      if (nb == 1) { errors[live[b0]] = e.what(); continue; }
      // NOTE: a retry sees ROIs the failed pass may already have rotated; decode again for a clean retry
      for (size_t i = b0; i < b0 + size_t(nb); ++i) {
        try {
          std::vector<DevImg> one;
          std::vector<std::string> w1;
          const int src = live[i];
          jpeg_->decode(data + src, sizes + src, 1, stream_, &one, &w1);
          if (!one[0].p) throw std::runtime_error(w1[0]);
          std::vector<std::vector<WordOut>> w;
          run_device(one, &w);
          words[src] = std::move(w[0]);
        } catch (const std::exception& e1) {
          recover_after_failure();
          words[live[i]].clear();
          errors[live[i]] = e1.what();
        }
      }
      // the single-image decodes reused the batch buffers: the images of later sub-batches are gone -> decode them again
      if (b0 + size_t(nb) < live.size()) {
        std::vector<DevImg> again;
        std::vector<std::string> w2;
        jpeg_->decode(data, sizes, n, stream_, &again, &w2);
        for (size_t i = b0 + size_t(nb); i < live.size(); ++i) dec[live[i]] = again[live[i]];
      }
    }
  }
  const double ms = ms_since(t_start);
  for (int i = 0; i < n; ++i) {
    const bool ok = errors[i].empty();
    (*json)[i] = result_json(request_ids[i], worker_id_, ok, ok || dec[i].p ? dec[i].cols : 0, ok || dec[i].p ? dec[i].rows : 0, ms,
                             words[i], errors[i]);
  }
}

void Worker::process_batch(const int* request_ids, const HostImage* imgs, int n, std::vector<std::string>* json) {
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  json->assign(n, std::string());
  const auto t_start = Clock::now();
  // empty images fail on their own (src/ocr_worker.cpp:223-226); the rest go through the GPU together
  std::vector<int> live;
  std::vector<HostImage> live_imgs;
  for (int i = 0; i < n; ++i) {
    if (!imgs[i].data || imgs[i].rows <= 0 || imgs[i].cols <= 0)
      (*json)[i] = result_json(request_ids[i], worker_id_, false, 0, 0, 0.0, {}, "Empty image data provided");
    else { live.push_back(i); live_imgs.push_back(imgs[i]); }
  }
  if (live.empty()) return;
  std::vector<std::string> errors(live.size());
  std::vector<std::vector<WordOut>> words(live.size());
  auto run_range = [&](size_t b0, int nb) {
    batch_.upload(live_imgs.data() + b0, nb, stream_);
    std::vector<std::vector<WordOut>> w;
    run_device(batch_.images(), &w);
    for (int i = 0; i < nb; ++i) words[b0 + i] = std::move(w[i]);
  };
  for (size_t b0 = 0; b0 < live.size(); b0 += size_t(opt_.max_batch)) {
    const int nb = int(std::min(live.size() - b0, size_t(opt_.max_batch)));
    try {
      run_range(b0, nb);
    } catch (const std::exception& e) {
      // failures stay isolated per request, as in the reference (see process_resident)
      recover_after_failure();
      if (nb == 1) { errors[b0] = e.what(); continue; }
      for (size_t i = b0; i < b0 + size_t(nb); ++i) {
        try {
          run_range(i, 1);
        } catch (const std::exception& e1) {
          recover_after_failure();
          words[i].clear();
          errors[i] = e1.what();
        }
      }
    }
  }
  const double ms = ms_since(t_start);
  for (size_t k = 0; k < live.size(); ++k) {
    const int i = live[k];
    if (!errors[k].empty())
      (*json)[i] = result_json(request_ids[i], worker_id_, false, imgs[i].cols, imgs[i].rows, ms, {}, errors[k]);
    else
      (*json)[i] = result_json(request_ids[i], worker_id_, true, imgs[i].cols, imgs[i].rows, ms, words[k], "");
  }
}

}  // namespace b200ocr
