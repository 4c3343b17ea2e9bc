// C ABI of the stages, the worker and the per-device pool (declared in include/b200ocr.h).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <set>
#include <mutex>
#include <thread>

#include "../../include/b200ocr.h"
#include "capi_util.h"
#include "stages.h"
#include "jpeg.h"

using namespace b200ocr;

namespace {

char* dup_string(const std::string& s) {
  char* p = static_cast<char*>(malloc(s.size() + 1));
  if (!p) throw std::bad_alloc();
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

HostImage to_host(const b200ocr_image& im) {
  HostImage h;
  h.data = im.data; h.rows = im.rows; h.cols = im.cols;
  h.step = im.step ? im.step : size_t(im.cols) * 3;
  return h;
}

void need_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
    throw std::runtime_error("no CUDA device is available: b200ocr has no CPU fallback");
  if (device < 0 || device >= n) throw std::invalid_argument("gpu_id " + std::to_string(device) + " out of range");
}

struct StageBase {
  int device = 0;
  cudaStream_t stream = nullptr;
  ImageBatch batch;
  void init(int dev) {
    need_device(dev);
    device = dev;
    cuda_check(cudaSetDevice(dev), "cudaSetDevice");
    cuda_check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~StageBase() { if (stream) { cudaSetDevice(device); cudaStreamDestroy(stream); } }
  const std::vector<DevImg>& upload(const b200ocr_image* imgs, int n) {
    std::vector<HostImage> h(n);
    for (int i = 0; i < n; ++i) {
      if (!imgs[i].data || imgs[i].rows <= 0 || imgs[i].cols <= 0) throw std::invalid_argument("empty image");
      h[i] = to_host(imgs[i]);
    }
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    batch.upload(h.data(), n, stream);
    return batch.images();
  }
};

std::vector<Roi> full_rois(const std::vector<DevImg>& d) {
  std::vector<Roi> r(d.size());
  for (size_t i = 0; i < d.size(); ++i) { r[i].img = int(i); r[i].x = 0; r[i].y = 0; r[i].w = d[i].cols; r[i].h = d[i].rows; }
  return r;
}

}  // namespace

struct b200ocr_det : StageBase { std::unique_ptr<DetStage> st; DetParams p; };
struct b200ocr_cls : StageBase { std::unique_ptr<ClsStage> st; };
struct b200ocr_rec : StageBase { std::unique_ptr<RecStage> st; };
struct b200ocr_worker { std::unique_ptr<Worker> w; };
struct b200ocr_batch { int device = 0; ImageBatch b; };
std::string profile_json(Net& net, cudaStream_t stream, int warmup, int reps, int thresh);

// ------------------------------------------------------------------------------------------------ pool
// The pool deep-copies every submitted image (like OCRRequest's clone, reference include/paddle_ocr/ocr_worker.h:28-29).
// Pixels are cloned STRAIGHT INTO THE MEMORY OF THE DEVICE the request is dispatched to (b200ocr_pool_submit): a host-side
// clone costs a 2 MB memcpy per card, and the host's copy bandwidth (measured ~20 GB/s over all cores of the box,
// profiles/r02_notes.md section 8) capped a two-GPU pool at 6.8 k images/s; a page-locked caller buffer now moves as one
// DMA and no CPU core touches the pixels.  Encoded files (tens of KB, parsed on the host) are cloned into page-locked
// host memory.  Both kinds of buffer are recycled through free lists (cudaHostAlloc / cudaMalloc cost milliseconds
// and serialise the whole process), sized in 256 KB steps.
class BufferPool {
 public:
  BufferPool(bool on_device, int device) : on_device_(on_device), device_(device) {}
  ~BufferPool() {
    if (on_device_) cudaSetDevice(device_);
    for (auto& kv : free_) for (void* p : kv.second) { if (on_device_) cudaFree(p); else cudaFreeHost(p); }
  }
  // device pools: the caller has made `device` current
  uint8_t* take(size_t bytes, size_t* cap) {
    const size_t c = std::max<size_t>(1, (bytes + (size_t(256) << 10) - 1) >> 18) << 18;
    *cap = c;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = free_.find(c);
      if (it != free_.end() && !it->second.empty()) { void* p = it->second.back(); it->second.pop_back(); return static_cast<uint8_t*>(p); }
    }
    void* p = nullptr;
    const cudaError_t e = on_device_ ? cudaMalloc(&p, c) : cudaHostAlloc(&p, c, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
    return static_cast<uint8_t*>(p);
  }
  void give(uint8_t* p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu_);
    free_[cap].push_back(p);
  }
 private:
  const bool on_device_;
  const int device_;
  std::mutex mu_;
  std::map<size_t, std::vector<void*>> free_;
};

struct b200ocr_pool {
  BufferPool pinned{false, 0};
  // What b200ocr_pool_wait sleeps on: one per request, so a finished batch wakes exactly the threads that wait for ITS
  // requests (a single condition variable for the pool woke every waiter on every batch: with 64 client threads and
  // 10^4 batches/s the clients did nothing but wake up and go back to sleep, profiles/r02_notes.md section 8).
  struct Pending {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    std::string json;
  };
  struct Request {
    std::shared_ptr<Pending> pending;
    long long ticket;
    int request_id;
    BufferPool* owner = nullptr;
    uint8_t* data = nullptr;      // deep copy: the pixels in DEVICE memory, or the encoded file in page-locked host memory
    size_t size = 0, cap = 0;
    bool encoded = false;         // data holds the encoded file (b200ocr_pool_submit_encoded)
    int rows, cols;
    std::chrono::steady_clock::time_point t_submit;
    ~Request() { if (owner) owner->give(data, cap); }
  };
  // One uploader thread per device owns the H2D clones: submitting threads hand it a job and sleep until the pixels
  // have landed.  (Issuing the copy from the submitting threads themselves -- 8 to 16 extra threads calling into the
  // CUDA runtime next to the workers -- serialised on the driver at ~115 us per image and slowed the workers' own
  // launches three-fold: profiles/r02_notes.md section 8.)  See upload_loop.
  struct CopyJob {
    std::shared_ptr<Request> r;
    const uint8_t* src = nullptr;
    size_t step = 0, row = 0;
    bool done = false;
    std::string error;
    std::condition_variable cv;   // waited on with the device's copy_mu: only this job's submitter wakes
  };
  struct Device {
    explicit Device(int dev) : device(dev), pixels(true, dev) {}
    ~Device() { queue.clear(); }   // requests give their buffers back before `pixels` goes
    int device;
    BufferPool pixels;              // declared before the queues: destroyed after the requests that point into it
    std::mutex copy_mu;
    std::condition_variable copy_cv;
    std::deque<CopyJob*> copy_q;
    std::atomic<int> copying{0};    // jobs handed to the uploader and not yet in `queue` (counted as load by the dispatch)
    std::thread uploader;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::shared_ptr<Request>> queue;
    std::vector<std::thread> threads;
    std::vector<std::unique_ptr<Worker>> workers;
    std::atomic<int> busy{0};
  };
  std::vector<std::unique_ptr<Device>> devs;
  std::atomic<bool> running{true};
  std::atomic<long long> next_ticket{1};
  std::atomic<size_t> rr{0};
  std::mutex res_mu;
  std::condition_variable res_cv;    // destructor <- last waiter
  std::map<long long, std::shared_ptr<Pending>> outstanding;   // issued and not yet consumed by b200ocr_pool_wait (guarded by res_mu)
  int waiters = 0;                   // threads inside b200ocr_pool_wait (guarded by res_mu)
  int max_batch = 64;
  // status counters (reference OCRIPCService::getStatusInfo, src/ocr_ipc_service.cpp:438-448 -- whose success / time
  // counters are declared but never updated; these are)
  std::atomic<long long> total_requests{0}, successful_requests{0}, failed_requests{0}, batches{0};
  std::atomic<long long> total_time_us{0};   // sum over requests of (completion - submission)
  std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();

  // Copies are queued on one stream as their jobs arrive, each followed by its own event; the uploader sleeps on the
  // OLDEST copy in flight, hands that request to the workers and wakes its submitter, while the DMA engine already works
  // on the copies behind it.  (Completing whole groups of copies at once made a submit wait for ~16 copies: 1.4 ms per
  // request on 8 GPUs, which capped 64 submitting threads at 42 k images/s; profiles/r02_notes.md section 8.)
  void upload_loop(Device* d) {
    cudaStream_t stream = nullptr;
    bool ok = cudaSetDevice(d->device) == cudaSuccess && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) cudaGetLastError();   // every job fails with a message below
    std::vector<cudaEvent_t> spare;   // blocking-sync events, recycled: the uploader sleeps while its DMAs run
    struct InFlight { CopyJob* job; cudaEvent_t landed; };
    std::deque<InFlight> inflight;
    std::vector<CopyJob*> fresh;
    auto complete = [&](CopyJob* j) {
      if (j->error.empty()) {  // the clone is on the device: hand the request to the workers ...
        { std::lock_guard<std::mutex> lk(d->mu); d->queue.push_back(j->r); }
        d->cv.notify_one();
      }
      d->copying -= 1;
      // ... then release the submitting thread (the job lives on its stack: notify under the lock)
      std::lock_guard<std::mutex> lk(d->copy_mu);
      j->done = true;
      j->cv.notify_one();
    };
    while (true) {
      fresh.clear();
      {
        std::unique_lock<std::mutex> lk(d->copy_mu);
        // with copies in flight do not sleep here: take what is waiting (possibly nothing) and go on completing
        if (inflight.empty()) d->copy_cv.wait(lk, [&] { return !d->copy_q.empty() || !running; });
        if (d->copy_q.empty() && inflight.empty()) break;   // shutting down and drained
        while (!d->copy_q.empty() && fresh.size() < 32) { fresh.push_back(d->copy_q.front()); d->copy_q.pop_front(); }
      }
      for (CopyJob* j : fresh) {
        cudaEvent_t ev = nullptr;
        try {
          if (!ok) throw std::runtime_error("the pool's copy stream could not be created");
          Request& r = *j->r;
          r.owner = &d->pixels;
          r.data = d->pixels.take(r.size, &r.cap);
          if (j->step == j->row) cuda_check(cudaMemcpyAsync(r.data, j->src, r.size, cudaMemcpyHostToDevice, stream), "image clone");
          else cuda_check(cudaMemcpy2DAsync(r.data, j->row, j->src, j->step, j->row, r.rows, cudaMemcpyHostToDevice, stream), "image clone");
          if (spare.empty()) cuda_check(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreate");
          else { ev = spare.back(); spare.pop_back(); }
          cuda_check(cudaEventRecord(ev, stream), "image clone");
          inflight.push_back(InFlight{j, ev});
        } catch (const std::exception& e) {
          if (ev) spare.push_back(ev);
          j->error = e.what()[0] ? e.what() : "image clone failed";
          cudaStreamSynchronize(stream);   // nothing of this job may still be reading the caller's buffer
          cudaGetLastError();
          complete(j);
        }
      }
      // complete the oldest copy (sleeping until it has landed), then everything that has landed since; go back for
      // new jobs as soon as some are waiting, so that the DMA queue never runs dry
      bool first = true;
      while (!inflight.empty()) {
        InFlight f = inflight.front();
        if (first) {
          if (cudaEventSynchronize(f.landed) != cudaSuccess) f.job->error = std::string("image clone: ") + cudaGetErrorString(cudaGetLastError());
        } else {
          const cudaError_t q = cudaEventQuery(f.landed);
          if (q == cudaErrorNotReady) break;
          if (q != cudaSuccess) f.job->error = std::string("image clone: ") + cudaGetErrorString(cudaGetLastError());
        }
        first = false;
        inflight.pop_front();
        spare.push_back(f.landed);
        complete(f.job);
      }
    }
    for (cudaEvent_t ev : spare) cudaEventDestroy(ev);
    if (stream) cudaStreamDestroy(stream);
  }

  void loop(Device* d, Worker* w) {
    while (true) {
      std::vector<std::shared_ptr<Request>> take;
      {
        std::unique_lock<std::mutex> lk(d->mu);
        d->cv.wait(lk, [&] { return !d->queue.empty() || !running; });
        if (d->queue.empty()) { if (!running) return; continue; }
        while (!d->queue.empty() && int(take.size()) < max_batch) { take.push_back(d->queue.front()); d->queue.pop_front(); }
        d->busy += 1;
      }
      std::vector<int> ids(take.size());
      for (size_t i = 0; i < take.size(); ++i) ids[i] = take[i]->request_id;
      std::vector<std::string> out(take.size());
      try {
        // raw (already on this device), empty and encoded requests of one batch go through the worker as three groups
        std::vector<size_t> raw, enc, empty;
        for (size_t i = 0; i < take.size(); ++i) (take[i]->encoded ? enc : take[i]->size ? raw : empty).push_back(i);
        if (!raw.empty()) {
          std::vector<int> rid(raw.size());
          std::vector<DevImg> imgs(raw.size());
          for (size_t k = 0; k < raw.size(); ++k) {
            const Request& r = *take[raw[k]];
            rid[k] = r.request_id;
            imgs[k].p = r.data;
            imgs[k].rows = r.rows; imgs[k].cols = r.cols; imgs[k].stride = long(r.cols) * 3;
          }
          std::vector<std::string> o;
          w->process_resident(rid.data(), imgs, &o);
          for (size_t k = 0; k < raw.size(); ++k) out[raw[k]] = std::move(o[k]);
        }
        if (!empty.empty()) {   // "Empty image data provided" (src/ocr_worker.cpp:223-226), no GPU work
          std::vector<int> rid(empty.size());
          std::vector<HostImage> imgs(empty.size());
          for (size_t k = 0; k < empty.size(); ++k) {
            const Request& r = *take[empty[k]];
            rid[k] = r.request_id;
            imgs[k].data = nullptr; imgs[k].rows = r.rows; imgs[k].cols = r.cols; imgs[k].step = 0;
          }
          std::vector<std::string> o;
          w->process_batch(rid.data(), imgs.data(), int(empty.size()), &o);
          for (size_t k = 0; k < empty.size(); ++k) out[empty[k]] = std::move(o[k]);
        }
        if (!enc.empty()) {
          std::vector<int> rid(enc.size());
          std::vector<const uint8_t*> data(enc.size());
          std::vector<size_t> sizes(enc.size());
          for (size_t k = 0; k < enc.size(); ++k) {
            const Request& r = *take[enc[k]];
            rid[k] = r.request_id; data[k] = r.data; sizes[k] = r.size;
          }
          std::vector<std::string> o;
          w->process_encoded(rid.data(), data.data(), sizes.data(), int(enc.size()), &o);
          for (size_t k = 0; k < enc.size(); ++k) out[enc[k]] = std::move(o[k]);
        }
      } catch (const std::exception& e) {
        // reference src/ocr_worker.cpp:192-206: request_id, success=false, error, worker_id
        out.assign(take.size(), std::string());
        for (size_t i = 0; i < take.size(); ++i)
          out[i] = "{\"error\":" + json_quote(e.what()) + ",\"request_id\":" + std::to_string(ids[i]) +
                   ",\"success\":false,\"worker_id\":" + std::to_string(w->worker_id()) + "}";
      }
      {
        const auto now = std::chrono::steady_clock::now();
        for (size_t i = 0; i < take.size(); ++i) {
          (out[i].find("\"success\":true") != std::string::npos ? successful_requests : failed_requests) += 1;
          total_time_us += std::chrono::duration_cast<std::chrono::microseconds>(now - take[i]->t_submit).count();
        }
        batches += 1;
        for (size_t i = 0; i < take.size(); ++i) {
          Pending& p = *take[i]->pending;
          { std::lock_guard<std::mutex> lk(p.mu); p.json = std::move(out[i]); p.done = true; }
          p.cv.notify_all();
        }
      }
      d->busy -= 1;
    }
  }

  ~b200ocr_pool() {
    running = false;
    {  // waiters return B200OCR_ERR_RUNTIME instead of touching a destroyed condition variable
      std::unique_lock<std::mutex> lk(res_mu);
      for (auto& kv : outstanding) {
        { std::lock_guard<std::mutex> pl(kv.second->mu); }
        kv.second->cv.notify_all();
      }
      res_cv.wait(lk, [&] { return waiters == 0; });
    }
    for (auto& d : devs) {
      { std::lock_guard<std::mutex> lk(d->copy_mu); }
      d->copy_cv.notify_all();
      if (d->uploader.joinable()) d->uploader.join();
      { std::lock_guard<std::mutex> lk(d->mu); }
      d->cv.notify_all();
      for (auto& t : d->threads) if (t.joinable()) t.join();
    }
  }
};

extern "C" {

int b200ocr_pool_worker_count(b200ocr_pool_t pool);
int b200ocr_pool_idle_count(b200ocr_pool_t pool);

void* b200ocr_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void b200ocr_host_free(void* p) { if (p) cudaFreeHost(p); }
int b200ocr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---- det
int b200ocr_det_create(const b200ocr_det_config* c, b200ocr_det_t* out) {
  return capi_guard([&] {
    if (!c || !out || !c->model_dir) throw std::invalid_argument("null argument");
    auto h = std::make_unique<b200ocr_det>();
    h->init(c->gpu_id);
    h->p.limit_type = c->limit_type ? c->limit_type : "max";
    h->p.limit_side_len = c->limit_side_len;
    h->p.det_db_thresh = c->det_db_thresh;
    h->p.det_db_box_thresh = c->det_db_box_thresh;
    h->p.det_db_unclip_ratio = c->det_db_unclip_ratio;
    h->p.det_db_score_mode = c->det_db_score_mode ? c->det_db_score_mode : "slow";
    h->p.use_dilation = c->use_dilation != 0;
    h->st = std::make_unique<DetStage>(c->model_dir, c->gpu_id, h->p);
    *out = h.release();
  });
}
void b200ocr_det_destroy(b200ocr_det_t d) { if (d) { cudaSetDevice(d->device); delete d; } }

int b200ocr_det_run_batch(b200ocr_det_t d, const b200ocr_image* imgs, int n, int32_t* boxes, int cap, int* counts,
                          double times[3]) {
  return capi_guard([&] {
    if (!d || !imgs || n < 0 || !counts || (cap > 0 && !boxes)) throw std::invalid_argument("null argument");
    const auto& dev = d->upload(imgs, n);
    std::vector<std::vector<Box>> res;
    std::vector<double> t;
    d->st->run(dev, &res, d->stream, &t);
    for (int i = 0; i < n; ++i) {
      counts[i] = int(res[i].size());
      for (int k = 0; k < counts[i] && k < cap; ++k) memcpy(boxes + (size_t(i) * cap + k) * 8, res[i][k].data(), 32);
    }
    if (times) for (int k = 0; k < 3; ++k) times[k] = t[k];
  });
}
int b200ocr_det_run(b200ocr_det_t d, const b200ocr_image* img, int32_t* boxes, int cap, int* n_boxes, double times[3]) {
  return b200ocr_det_run_batch(d, img, 1, boxes, cap, n_boxes, times);
}

int b200ocr_det_preprocess(b200ocr_det_t d, const b200ocr_image* img, float* nchw, int* rh, int* rw, float* ratio_h,
                           float* ratio_w) {
  return capi_guard([&] {
    if (!d || !img || !rh || !rw) throw std::invalid_argument("null argument");
    float a, b;
    DetStage::resized_dims(img->rows, img->cols, d->p.limit_type, d->p.limit_side_len, rh, rw, &a, &b);
    if (ratio_h) *ratio_h = a;
    if (ratio_w) *ratio_w = b;
    if (!nchw) return;
    const auto& dev = d->upload(img, 1);
    DevBuf items, in, out;
    items.ensure(sizeof(DetPreItem));
    in.ensure(size_t(*rh) * *rw * 8 * sizeof(__half));
    out.ensure(size_t(*rh) * *rw * 3 * sizeof(float));
    DetPreItem it{dev[0].p, dev[0].cols, dev[0].rows, dev[0].stride};
    cuda_check(cudaMemcpyAsync(items.p, &it, sizeof it, cudaMemcpyHostToDevice, d->stream), "items");
    static const float mean[3] = {0.485f, 0.456f, 0.406f};
    static const float scale[3] = {1 / 0.229f, 1 / 0.224f, 1 / 0.225f};
    launch_det_preprocess(items.as<DetPreItem>(), 1, *rh, *rw, make_norm(mean, scale), in.as<__half>(), d->stream);
    TV v;
    v.p = in.as<__half>(); v.n = 1; v.h = *rh; v.w = *rw; v.c = 3; v.pitch = 8;
    launch_nhwc_to_nchw_f32(v, out.as<float>(), d->stream);
    cuda_check(cudaMemcpyAsync(nchw, out.p, size_t(*rh) * *rw * 3 * sizeof(float), cudaMemcpyDeviceToHost, d->stream), "copy");
    cuda_check(cudaStreamSynchronize(d->stream), "det preprocess");
  });
}

int b200ocr_det_postprocess(b200ocr_det_t d, const float* pred, int h, int w, int src_h, int src_w, int32_t* boxes,
                            int cap, int* n_boxes, uint8_t* bitmap_out) {
  return capi_guard([&] {
    if (!d || !pred || h < 1 || w < 1 || !n_boxes) throw std::invalid_argument("bad argument");
    cuda_check(cudaSetDevice(d->device), "cudaSetDevice");
    cudaStream_t s = d->stream;
    const size_t px = size_t(h) * w;
    DevBuf dp, db, dd, info, ws, counts, bx;
    dp.ensure(px * 4); db.ensure(px); dd.ensure(px);
    cuda_check(cudaMemcpyAsync(dp.p, pred, px * 4, cudaMemcpyHostToDevice, s), "pred upload");
    const int thr = int(std::floor(double(float(d->p.det_db_thresh)) * 255));
    launch_threshold(dp.as<float>(), long(px), thr, db.as<uint8_t>(), s);
    const uint8_t* bm = db.as<uint8_t>();
    if (d->p.use_dilation) { launch_dilate2x2(bm, dd.as<uint8_t>(), 1, h, w, s); bm = dd.as<uint8_t>(); }
    DbPostParams pp;
    pp.n = 1; pp.h = h; pp.w = w;
    pp.box_thresh = float(d->p.det_db_box_thresh);
    pp.unclip_ratio = float(d->p.det_db_unclip_ratio);
    pp.max_candidates = 1000;
    pp.score_slow = d->p.det_db_score_mode == "slow";
    DbImageInfo inf;
    // the detector feeds a map of the resized size; ratios as DBDetector::Run computes them (src/preprocess_op.cpp:91-92)
    inf.ratio_h = float(h) / float(src_h);
    inf.ratio_w = float(w) / float(src_w);
    inf.src_h = src_h; inf.src_w = src_w;
    info.ensure(sizeof inf);
    cuda_check(cudaMemcpyAsync(info.p, &inf, sizeof inf, cudaMemcpyHostToDevice, s), "info");
    ws.ensure(dbpost_workspace_bytes(pp));
    counts.ensure(4);
    bx.ensure(sizeof(DbBox) * pp.max_candidates);
    launch_dbpost(pp, dp.as<float>(), bm, info.as<DbImageInfo>(), ws.p, counts.as<int>(), bx.as<DbBox>(), s);
    int cnt = 0;
    std::vector<DbBox> hb(pp.max_candidates);
    cuda_check(cudaMemcpyAsync(&cnt, counts.p, 4, cudaMemcpyDeviceToHost, s), "counts");
    cuda_check(cudaMemcpyAsync(hb.data(), bx.p, sizeof(DbBox) * pp.max_candidates, cudaMemcpyDeviceToHost, s), "boxes");
    if (bitmap_out) cuda_check(cudaMemcpyAsync(bitmap_out, bm, px, cudaMemcpyDeviceToHost, s), "bitmap");
    cuda_check(cudaStreamSynchronize(s), "det postprocess");
    int m = 0;
    for (int c = 0; c < cnt; ++c) {
      if (!hb[c].valid) continue;
      if (m < cap && boxes) memcpy(boxes + size_t(m) * 8, hb[c].pts, 32);
      ++m;
    }
    *n_boxes = m;
  });
}

// ---- cls
int b200ocr_cls_create(const b200ocr_cls_config* c, b200ocr_cls_t* out) {
  return capi_guard([&] {
    if (!c || !out || !c->model_dir) throw std::invalid_argument("null argument");
    auto h = std::make_unique<b200ocr_cls>();
    h->init(c->gpu_id);
    h->st = std::make_unique<ClsStage>(c->model_dir, c->gpu_id, c->cls_batch_num > 0 ? c->cls_batch_num : 1,
                                       float(c->cls_thresh));
    *out = h.release();
  });
}
void b200ocr_cls_destroy(b200ocr_cls_t c) { if (c) { cudaSetDevice(c->device); delete c; } }
int b200ocr_cls_run(b200ocr_cls_t c, const b200ocr_image* imgs, int n, int* cls_labels, float* cls_scores,
                    double times[3]) {
  return capi_guard([&] {
    if (!c || (n > 0 && (!imgs || !cls_labels || !cls_scores)) || n < 0) throw std::invalid_argument("null argument");
    std::vector<double> t;
    if (n > 0) {
      const auto& dev = c->upload(imgs, n);
      std::vector<int> labels;
      std::vector<float> scores;
      c->st->run(dev, full_rois(dev), &labels, &scores, c->stream, true, &t);
      memcpy(cls_labels, labels.data(), sizeof(int) * n);
      memcpy(cls_scores, scores.data(), sizeof(float) * n);
    } else t.assign(3, 0.0);
    if (times) for (int k = 0; k < 3; ++k) times[k] = t[k];
  });
}

// ---- rec
int b200ocr_rec_create(const b200ocr_rec_config* c, b200ocr_rec_t* out) {
  return capi_guard([&] {
    if (!c || !out || !c->model_dir || !c->label_path) throw std::invalid_argument("null argument");
    auto h = std::make_unique<b200ocr_rec>();
    h->init(c->gpu_id);
    h->st = std::make_unique<RecStage>(c->model_dir, c->gpu_id, c->label_path, c->rec_batch_num > 0 ? c->rec_batch_num : 1,
                                       c->rec_img_h, c->rec_img_w);
    *out = h.release();
  });
}
void b200ocr_rec_destroy(b200ocr_rec_t r) { if (r) { cudaSetDevice(r->device); delete r; } }
int b200ocr_rec_run(b200ocr_rec_t r, const b200ocr_image* imgs, int n, char** rec_texts, float* rec_text_scores,
                    double times[3]) {
  return capi_guard([&] {
    if (!r || (n > 0 && (!imgs || !rec_texts || !rec_text_scores)) || n < 0) throw std::invalid_argument("null argument");
    std::vector<double> t(3, 0.0);
    if (n > 0) {
      const auto& dev = r->upload(imgs, n);
      std::vector<std::vector<Roi>> calls(1, full_rois(dev));
      std::vector<std::vector<std::string>> texts;
      std::vector<std::vector<float>> scores;
      t.clear();
      r->st->run(dev, calls, &texts, &scores, r->stream, &t);
      for (int i = 0; i < n; ++i) { rec_texts[i] = dup_string(texts[0][i]); rec_text_scores[i] = scores[0][i]; }
    }
    if (times) for (int k = 0; k < 3; ++k) times[k] = t[k];
  });
}

// ---- device-resident forms of the stage calls (inputs uploaded once with b200ocr_batch_upload)
int b200ocr_det_run_resident(b200ocr_det_t d, b200ocr_batch_t batch, int32_t* boxes, int cap, int* counts, double times[3]) {
  return capi_guard([&] {
    if (!d || !batch || !counts || (cap > 0 && !boxes)) throw std::invalid_argument("null argument");
    if (batch->device != d->device) throw std::invalid_argument("batch and detector live on different devices");
    cuda_check(cudaSetDevice(d->device), "cudaSetDevice");
    const auto& dev = batch->b.images();
    std::vector<std::vector<Box>> res;
    std::vector<double> t;
    d->st->run(dev, &res, d->stream, &t);
    for (size_t i = 0; i < dev.size(); ++i) {
      counts[i] = int(res[i].size());
      for (int k = 0; k < counts[i] && k < cap; ++k) memcpy(boxes + (i * size_t(cap) + k) * 8, res[i][k].data(), 32);
    }
    if (times) for (int k = 0; k < 3; ++k) times[k] = t[k];
  });
}
int b200ocr_rec_run_resident(b200ocr_rec_t r, b200ocr_batch_t batch, char** rec_texts, float* rec_text_scores,
                             double times[3]) {
  return capi_guard([&] {
    if (!r || !batch || !rec_texts || !rec_text_scores) throw std::invalid_argument("null argument");
    if (batch->device != r->device) throw std::invalid_argument("batch and recognizer live on different devices");
    cuda_check(cudaSetDevice(r->device), "cudaSetDevice");
    const auto& dev = batch->b.images();
    std::vector<std::vector<Roi>> calls(1, full_rois(dev));
    std::vector<std::vector<std::string>> texts;
    std::vector<std::vector<float>> scores;
    std::vector<double> t;
    r->st->run(dev, calls, &texts, &scores, r->stream, &t);
    for (size_t i = 0; i < dev.size(); ++i) { rec_texts[i] = dup_string(texts[0][i]); rec_text_scores[i] = scores[0][i]; }
    if (times) for (int k = 0; k < 3; ++k) times[k] = t[k];
  });
}
void* b200ocr_det_stream(b200ocr_det_t d) { return d ? d->stream : nullptr; }
void* b200ocr_rec_stream(b200ocr_rec_t r) { return r ? r->stream : nullptr; }
long long b200ocr_det_launches(b200ocr_det_t d) { return d ? d->st->launches : 0; }
long long b200ocr_rec_launches(b200ocr_rec_t r) { return r ? r->st->launches : 0; }

// ---- worker
int b200ocr_worker_create(int worker_id, const char* model_dir, int use_gpu, int gpu_id, int enable_cls,
                          b200ocr_worker_t* out) {
  return capi_guard([&] {
    (void)use_gpu;
    if (!model_dir || !out) throw std::invalid_argument("null argument");
    need_device(gpu_id);
    auto h = std::make_unique<b200ocr_worker>();
    WorkerOptions o;
    o.enable_cls = enable_cls != 0;
    h->w = std::make_unique<Worker>(worker_id, model_dir, gpu_id, o);
    *out = h.release();
  });
}
int b200ocr_worker_create_ex(int worker_id, const char* model_dir, int gpu_id, int enable_cls,
                             const b200ocr_worker_params* p, b200ocr_worker_t* out) {
  return capi_guard([&] {
    if (!model_dir || !out) throw std::invalid_argument("null argument");
    need_device(gpu_id);
    auto h = std::make_unique<b200ocr_worker>();
    WorkerOptions o;
    o.enable_cls = enable_cls != 0;
    if (p) {
      if (p->limit_type) o.det_limit_type = p->limit_type;
      if (o.det_limit_type != "max" && o.det_limit_type != "min") throw std::invalid_argument("limit_type must be max or min");
      if (p->limit_side_len > 0) o.det_limit_side_len = p->limit_side_len;
      if (p->det_db_thresh > 0) o.det_db_thresh = p->det_db_thresh;
      if (p->det_db_box_thresh > 0) o.det_db_box_thresh = p->det_db_box_thresh;
      if (p->det_db_unclip_ratio > 0) o.det_db_unclip_ratio = p->det_db_unclip_ratio;
      if (p->det_db_score_mode) o.det_db_score_mode = p->det_db_score_mode;
      if (o.det_db_score_mode != "fast" && o.det_db_score_mode != "slow") throw std::invalid_argument("det_db_score_mode must be fast or slow");
      o.use_dilation = p->use_dilation != 0;
      if (p->cls_batch_num > 0) o.cls_batch_num = p->cls_batch_num;
      if (p->rec_batch_num > 0) o.rec_batch_num = p->rec_batch_num;
      if (p->rec_img_h > 0) o.rec_img_h = p->rec_img_h;
      if (p->rec_img_w > 0) o.rec_img_w = p->rec_img_w;
      if (p->max_batch > 0) o.max_batch = p->max_batch;
    }
    h->w = std::make_unique<Worker>(worker_id, model_dir, gpu_id, o);
    *out = h.release();
  });
}
void b200ocr_worker_destroy(b200ocr_worker_t w) { delete w; }
int b200ocr_worker_process_batch(b200ocr_worker_t w, const int* request_ids, const b200ocr_image* imgs, int n,
                                 char** jsons) {
  return capi_guard([&] {
    if (!w || n < 0 || (n > 0 && (!request_ids || !imgs || !jsons))) throw std::invalid_argument("null argument");
    std::vector<HostImage> h(n);
    for (int i = 0; i < n; ++i) h[i] = to_host(imgs[i]);
    std::vector<std::string> out;
    w->w->process_batch(request_ids, h.data(), n, &out);
    for (int i = 0; i < n; ++i) jsons[i] = dup_string(out[i]);
  });
}
int b200ocr_worker_process(b200ocr_worker_t w, int request_id, const b200ocr_image* img, char** json) {
  return b200ocr_worker_process_batch(w, &request_id, img, 1, json);
}
long long b200ocr_worker_launches(b200ocr_worker_t w) { return w ? w->w->launches() : 0; }
void* b200ocr_worker_stream(b200ocr_worker_t w) { return w ? static_cast<void*>(w->w->stream()) : nullptr; }

int b200ocr_batch_upload(int device, const b200ocr_image* imgs, int n, b200ocr_batch_t* out) {
  return capi_guard([&] {
    if (!imgs || n < 1 || !out) throw std::invalid_argument("bad argument");
    need_device(device);
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    auto h = std::make_unique<b200ocr_batch>();
    h->device = device;
    std::vector<HostImage> hi(n);
    for (int i = 0; i < n; ++i) {
      if (!imgs[i].data || imgs[i].rows <= 0 || imgs[i].cols <= 0) throw std::invalid_argument("empty image");
      hi[i] = to_host(imgs[i]);
    }
    // per-thread stream: never a device-wide synchronisation (another worker may be capturing a CUDA graph)
    h->b.upload(hi.data(), n, cudaStreamPerThread);
    cuda_check(cudaStreamSynchronize(cudaStreamPerThread), "batch upload");
    *out = h.release();
  });
}
void b200ocr_batch_destroy(b200ocr_batch_t b) { if (b) { cudaSetDevice(b->device); delete b; } }

int b200ocr_worker_process_resident(b200ocr_worker_t w, b200ocr_batch_t batch, const int* request_ids, char** jsons) {
  return capi_guard([&] {
    if (!w || !batch || !request_ids || !jsons) throw std::invalid_argument("null argument");
    if (batch->device != w->w->device()) throw std::invalid_argument("batch and worker live on different devices");
    std::vector<std::string> out;
    w->w->process_resident(request_ids, batch->b.images(), &out);
    for (size_t i = 0; i < out.size(); ++i) jsons[i] = dup_string(out[i]);
  });
}

int b200ocr_worker_process_encoded(b200ocr_worker_t w, const int* request_ids, const b200ocr_blob* blobs, int n,
                                   char** jsons) {
  return capi_guard([&] {
    if (!w || !request_ids || !blobs || !jsons || n < 0) throw std::invalid_argument("null argument");
    std::vector<const uint8_t*> data(n);
    std::vector<size_t> sizes(n);
    for (int i = 0; i < n; ++i) { data[i] = blobs[i].data; sizes[i] = blobs[i].size; }
    std::vector<std::string> out;
    w->w->process_encoded(request_ids, data.data(), sizes.data(), n, &out);
    for (int i = 0; i < n; ++i) jsons[i] = dup_string(out[i]);
  });
}
long long b200ocr_worker_last_encoded_h2d_bytes(b200ocr_worker_t w) { return w ? (long long)w->w->last_encoded_h2d_bytes() : 0; }

int b200ocr_jpeg_decode(int device, const uint8_t* data, size_t size, int* rows, int* cols, uint8_t* bgr) {
  return capi_guard([&] {
    if (!data || !rows || !cols) throw std::invalid_argument("null argument");
    JpegImage im;
    std::vector<JpegSeg> segs;
    size_t b = 0, e = 0;
    std::string why;
    if (!jpeg_parse(data, size, &im, &b, &e, &segs, &why)) throw std::invalid_argument("Unsupported image encoding: " + why);
    *rows = im.height; *cols = im.width;
    if (!bgr) return;
    StageBase sb;
    sb.init(device);
    JpegBatch jb;
    std::vector<DevImg> out;
    std::vector<std::string> w2;
    jb.decode(&data, &size, 1, sb.stream, &out, &w2);
    if (!out[0].p) throw std::runtime_error(w2[0]);
    cuda_check(cudaMemcpyAsync(bgr, out[0].p, size_t(im.height) * im.width * 3, cudaMemcpyDeviceToHost, sb.stream), "copy");
    cuda_check(cudaStreamSynchronize(sb.stream), "jpeg decode");
  });
}

int b200ocr_worker_profile(b200ocr_worker_t w, int warmup, int reps, char** json) {
  return capi_guard([&] {
    if (!w || !json || reps < 1) throw std::invalid_argument("bad argument");
    Worker& k = *w->w;
    cuda_check(cudaSetDevice(k.device()), "cudaSetDevice");
    // every network is timed at the largest forward pass of the worker's last call (dense: a ragged rec chunk is timed
    // as the dense batch of its widest row)
    auto timed = [&](Net& net, int n, int h, int wd, int thresh) {
      if (n > 0) net.prepare(n, h, wd, nullptr, k.stream());
      char shape[96];
      snprintf(shape, sizeof shape, "{\"shape\":[%d,%d,%d],\"layers\":", n, h, wd);
      return std::string(shape) + profile_json(net, k.stream(), warmup, reps, thresh) + "}";
    };
    std::string o = "{\"det\":" + timed(k.det().net(), k.det().prof_n, k.det().prof_h, k.det().prof_w, 51);
    if (k.cls()) o += ",\"cls\":" + timed(k.cls()->net(), k.cls()->prof_n, k.cls()->prof_h, k.cls()->prof_w, -1);
    o += ",\"rec\":" + timed(k.rec().net(), k.rec().prof_n, k.rec().prof_h, k.rec().prof_w, -1) + "}";
    *json = dup_string(o);
  });
}

// per-layer profile of a stage's network at the largest forward pass of its last run
static std::string stage_profile(Net& net, int n, int h, int wd, int thresh, cudaStream_t s, int warmup, int reps) {
  if (n <= 0) throw std::runtime_error("profile before the first run");
  net.prepare(n, h, wd, nullptr, s);
  char shape[96];
  snprintf(shape, sizeof shape, "{\"shape\":[%d,%d,%d],\"layers\":", n, h, wd);
  return std::string(shape) + profile_json(net, s, warmup, reps, thresh) + "}";
}
int b200ocr_det_profile(b200ocr_det_t d, int warmup, int reps, char** json) {
  return capi_guard([&] {
    if (!d || !json || reps < 1) throw std::invalid_argument("bad argument");
    cuda_check(cudaSetDevice(d->device), "cudaSetDevice");
    *json = dup_string("{\"det\":" + stage_profile(d->st->net(), d->st->prof_n, d->st->prof_h, d->st->prof_w, 51, d->stream,
                                                    warmup, reps) + "}");
  });
}
int b200ocr_rec_profile(b200ocr_rec_t r, int warmup, int reps, char** json) {
  return capi_guard([&] {
    if (!r || !json || reps < 1) throw std::invalid_argument("bad argument");
    cuda_check(cudaSetDevice(r->device), "cudaSetDevice");
    *json = dup_string("{\"rec\":" + stage_profile(r->st->net(), r->st->prof_n, r->st->prof_h, r->st->prof_w, -1, r->stream,
                                                    warmup, reps) + "}");
  });
}

// ---- pool
int b200ocr_pool_create(const char* model_dir, int n_devices, const int* devices, int workers_per_device,
                        int enable_cls, int max_batch, b200ocr_pool_t* out) {
  return capi_guard([&] {
    if (!model_dir || !out || n_devices < 1 || workers_per_device < 1) throw std::invalid_argument("bad argument");
    auto p = std::make_unique<b200ocr_pool>();
    p->max_batch = max_batch > 0 ? max_batch : 64;
    int wid = 0;
    for (int i = 0; i < n_devices; ++i) {
      const int dev = devices ? devices[i] : i;
      need_device(dev);
      auto d = std::make_unique<b200ocr_pool::Device>(dev);
      WorkerOptions o;
      o.enable_cls = enable_cls != 0;
      o.max_batch = p->max_batch;
      for (int k = 0; k < workers_per_device; ++k) d->workers.push_back(std::make_unique<Worker>(wid++, model_dir, dev, o));
      p->devs.push_back(std::move(d));
    }
    for (auto& d : p->devs) {
      d->uploader = std::thread(&b200ocr_pool::upload_loop, p.get(), d.get());
      for (auto& w : d->workers) d->threads.emplace_back(&b200ocr_pool::loop, p.get(), d.get(), w.get());
    }
    *out = p.release();
  });
}
void b200ocr_pool_destroy(b200ocr_pool_t pool) { delete pool; }

// shortest queue first; ties broken round-robin (reference src/gpu_worker_pool.cpp:46-59: idle worker, else round-robin)
static b200ocr_pool::Device& pool_pick_device(b200ocr_pool_t pool) {
  const size_t nd = pool->devs.size(), start = pool->rr++ % nd;
  size_t best = start, best_load = ~size_t(0);
  for (size_t k = 0; k < nd; ++k) {
    auto& d = *pool->devs[(start + k) % nd];
    std::lock_guard<std::mutex> lk(d.mu);
    const size_t load = d.queue.size() + size_t(d.copying.load()) + size_t(d.busy.load()) * size_t(pool->max_batch);
    if (load < best_load) { best_load = load; best = (start + k) % nd; }
  }
  return *pool->devs[best];
}
static void pool_enqueue(b200ocr_pool_t pool, b200ocr_pool::Device& d, const std::shared_ptr<b200ocr_pool::Request>& r) {
  r->pending = std::make_shared<b200ocr_pool::Pending>();
  { std::lock_guard<std::mutex> lk(pool->res_mu); pool->outstanding.emplace(r->ticket, r->pending); }
  { std::lock_guard<std::mutex> lk(d.mu); d.queue.push_back(r); }
  d.cv.notify_one();
}

int b200ocr_pool_submit(b200ocr_pool_t pool, int request_id, const b200ocr_image* img, long long* ticket) {
  return capi_guard([&] {
    if (!pool || !img || !ticket) throw std::invalid_argument("null argument");
    auto r = std::make_shared<b200ocr_pool::Request>();
    r->ticket = pool->next_ticket++;
    r->request_id = request_id;
    r->t_submit = std::chrono::steady_clock::now();
    pool->total_requests += 1;
    r->rows = img->rows; r->cols = img->cols;
    auto& d = pool_pick_device(pool);
    if (img->data && img->rows > 0 && img->cols > 0) {
      // the clone: host pixels -> the memory of the device the request goes to; the call returns when the copy has
      // landed, so the caller's buffer is free again (OCRRequest's img.clone())
      const size_t row = size_t(img->cols) * 3, step = img->step ? img->step : row;
      r->size = row * img->rows;
      b200ocr_pool::CopyJob job;
      job.r = r;
      job.src = img->data; job.step = step; job.row = row;
      // Pageable caller memory: the driver would stage it through its own bounce buffer inside the (single) uploader
      // thread; cloning it into page-locked memory HERE keeps that CPU copy on the submitting threads, in parallel.
      cudaPointerAttributes at;
      const bool locked = cudaPointerGetAttributes(&at, img->data) == cudaSuccess && at.type == cudaMemoryTypeHost;
      if (!locked) cudaGetLastError();
      uint8_t* bounce = nullptr;
      size_t bounce_cap = 0;
      if (!locked) {
        bounce = pool->pinned.take(r->size, &bounce_cap);
        if (step == row) memcpy(bounce, img->data, r->size);
        else for (int y = 0; y < img->rows; ++y) memcpy(bounce + y * row, img->data + y * step, row);
        job.src = bounce; job.step = row;
      }
      r->pending = std::make_shared<b200ocr_pool::Pending>();
      { std::lock_guard<std::mutex> lk(pool->res_mu); pool->outstanding.emplace(r->ticket, r->pending); }
      {
        std::unique_lock<std::mutex> lk(d.copy_mu);
        d.copying += 1;
        d.copy_q.push_back(&job);
        d.copy_cv.notify_one();
        job.cv.wait(lk, [&] { return job.done; });
      }
      if (bounce) pool->pinned.give(bounce, bounce_cap);
      if (!job.error.empty()) {
        { std::lock_guard<std::mutex> lk(pool->res_mu); pool->outstanding.erase(r->ticket); }
        throw std::runtime_error(job.error);
      }
      *ticket = r->ticket;
      return;
    }
    pool_enqueue(pool, d, r);
    *ticket = r->ticket;
  });
}

int b200ocr_pool_submit_encoded(b200ocr_pool_t pool, int request_id, const uint8_t* data, size_t size, long long* ticket) {
  return capi_guard([&] {
    if (!pool || !ticket || (!data && size)) throw std::invalid_argument("null argument");
    auto r = std::make_shared<b200ocr_pool::Request>();
    r->ticket = pool->next_ticket++;
    r->request_id = request_id;
    r->t_submit = std::chrono::steady_clock::now();
    pool->total_requests += 1;
    r->rows = r->cols = 0;
    r->encoded = true;
    if (size) {
      r->owner = &pool->pinned;
      r->data = pool->pinned.take(size, &r->cap);
      r->size = size;
      memcpy(r->data, data, size);
    }
    pool_enqueue(pool, pool_pick_device(pool), r);
    *ticket = r->ticket;
  });
}

int b200ocr_pool_wait_for(b200ocr_pool_t pool, long long ticket, int timeout_ms, char** json) {
  return capi_guard([&] {
    if (!pool || !json) throw std::invalid_argument("null argument");
    *json = nullptr;
    std::shared_ptr<b200ocr_pool::Pending> p;
    {
      std::lock_guard<std::mutex> lk(pool->res_mu);
      auto it = pool->outstanding.find(ticket);
      if (it == pool->outstanding.end()) throw std::invalid_argument("unknown or already consumed ticket");
      p = it->second;
      ++pool->waiters;
    }
    struct Count {
      b200ocr_pool* q;
      ~Count() {
        std::lock_guard<std::mutex> lk(q->res_mu);
        if (--q->waiters == 0) q->res_cv.notify_all();
      }
    } count{pool};
    std::string line;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      auto ready = [&] { return p->done || !pool->running; };
      if (timeout_ms < 0) p->cv.wait(lk, ready);
      else if (!p->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready)) return;  // *json stays NULL
      if (!p->done) throw std::runtime_error("the pool is shutting down");
      line = std::move(p->json);
      p->json.clear();
    }
    {
      std::lock_guard<std::mutex> lk(pool->res_mu);
      if (pool->outstanding.erase(ticket) == 0) throw std::invalid_argument("unknown or already consumed ticket");  // a second waiter took it
    }
    *json = dup_string(line);
  });
}
int b200ocr_pool_wait(b200ocr_pool_t pool, long long ticket, char** json) {
  return b200ocr_pool_wait_for(pool, ticket, -1, json);
}
int b200ocr_pool_status(b200ocr_pool_t pool, char** json) {
  return capi_guard([&] {
    if (!pool || !json) throw std::invalid_argument("null argument");
    const long long done = pool->successful_requests.load() + pool->failed_requests.load();
    const double avg_ms = done > 0 ? double(pool->total_time_us.load()) / 1e3 / double(done) : 0.0;
    const double up = std::chrono::duration<double>(std::chrono::steady_clock::now() - pool->t_start).count();
    size_t queued = 0;
    for (auto& d : pool->devs) { std::lock_guard<std::mutex> lk(d->mu); queued += d->queue.size(); }
    long long st[4] = {0, 0, 0, 0};
    for (auto& d : pool->devs)
      for (auto& w : d->workers) {
        long long t[4];
        w->stage_totals(t);
        for (int i = 0; i < 4; ++i) st[i] += t[i];
      }
    auto per_image = [&](int i) { return st[3] > 0 ? double(st[i]) / 1e3 / double(st[3]) : 0.0; };
    // compact, keys in alphabetical order like jsoncpp writes them
    std::string o = "{\"average_processing_time_ms\":" + json_double(avg_ms) +
                    ",\"batches\":" + std::to_string(pool->batches.load()) +
                    ",\"failed_requests\":" + std::to_string(pool->failed_requests.load()) +
                    ",\"idle_workers\":" + std::to_string(b200ocr_pool_idle_count(pool)) +
                    ",\"queued_requests\":" + std::to_string(queued) +
                    ",\"running\":" + (pool->running.load() ? "true" : "false") +
                    ",\"stage_ms_per_image\":{\"cls\":" + json_double(per_image(1)) + ",\"det\":" + json_double(per_image(0)) +
                    ",\"rec\":" + json_double(per_image(2)) + "}" +
                    ",\"successful_requests\":" + std::to_string(pool->successful_requests.load()) +
                    ",\"total_requests\":" + std::to_string(pool->total_requests.load()) +
                    ",\"uptime_s\":" + json_double(up) +
                    ",\"workers\":" + std::to_string(b200ocr_pool_worker_count(pool)) + "}";
    *json = dup_string(o);
  });
}
int b200ocr_pool_worker_count(b200ocr_pool_t pool) {
  int n = 0;
  if (pool) for (auto& d : pool->devs) n += int(d->workers.size());
  return n;
}
int b200ocr_pool_idle_count(b200ocr_pool_t pool) {
  int n = 0;
  if (pool) for (auto& d : pool->devs) n += int(d->workers.size()) - d->busy.load();
  return n;
}

// ---- stand-alone image ops
int b200ocr_resize_u8(int device, const b200ocr_image* src, int dst_rows, int dst_cols, uint8_t* dst) {
  return capi_guard([&] {
    if (!src || !src->data || !dst || dst_rows < 1 || dst_cols < 1) throw std::invalid_argument("bad argument");
    StageBase b;
    b.init(device);
    const auto& dev = b.upload(src, 1);
    DevBuf out;
    out.ensure(size_t(dst_rows) * dst_cols * 3);
    launch_resize_u8(dev[0].p, dev[0].cols, dev[0].rows, dev[0].stride, dst_cols, dst_rows, out.as<uint8_t>(), b.stream);
    cuda_check(cudaMemcpyAsync(dst, out.p, size_t(dst_rows) * dst_cols * 3, cudaMemcpyDeviceToHost, b.stream), "copy");
    cuda_check(cudaStreamSynchronize(b.stream), "resize");
  });
}

int b200ocr_rotate_crop(int device, const b200ocr_image* src, const int32_t box[8], int* dst_rows, int* dst_cols, uint8_t* dst) {
  return capi_guard([&] {
    if (!src || !src->data || !box || !dst_rows || !dst_cols) throw std::invalid_argument("null argument");
    int b[8], cw, ch;
    for (int i = 0; i < 8; ++i) b[i] = box[i];
    rotate_crop_dims(b, dst_rows, dst_cols, &cw, &ch);
    if (!dst) return;
    StageBase sb;
    sb.init(device);
    const auto& dev = sb.upload(src, 1);
    DevBuf out;
    const size_t bytes = size_t(*dst_rows) * *dst_cols * 3;
    out.ensure(bytes);
    launch_rotate_crop(dev[0].p, dev[0].rows, dev[0].cols, dev[0].stride, b, out.as<uint8_t>(), sb.stream);
    cuda_check(cudaMemcpyAsync(dst, out.p, bytes, cudaMemcpyDeviceToHost, sb.stream), "copy");
    cuda_check(cudaStreamSynchronize(sb.stream), "rotate crop");
  });
}

int b200ocr_crop_preprocess(int device, const b200ocr_image* crops, int n, int kind, int img_h, int img_w, float* nchw) {
  return capi_guard([&] {
    if (!crops || n < 1 || !nchw || img_h < 1 || img_w < 1) throw std::invalid_argument("bad argument");
    StageBase b;
    b.init(device);
    const auto& dev = b.upload(crops, n);
    std::vector<CropItem> items(n);
    for (int i = 0; i < n; ++i) {
      const float ratio = float(dev[i].cols) / float(dev[i].rows);
      int rw = std::ceil(float(img_h) * ratio) > float(img_w) ? img_w : int(std::ceil(float(img_h) * ratio));
      items[i] = CropItem{dev[i].p, dev[i].stride, 0, 0, dev[i].cols, dev[i].rows, rw, img_w};
    }
    DevBuf di, in, out;
    di.ensure(sizeof(CropItem) * n);
    const size_t px = size_t(n) * img_h * img_w;
    in.ensure(px * 8 * sizeof(__half));
    out.ensure(px * 3 * sizeof(float));
    cuda_check(cudaMemcpyAsync(di.p, items.data(), sizeof(CropItem) * n, cudaMemcpyHostToDevice, b.stream), "items");
    static const float mean[3] = {0.5f, 0.5f, 0.5f};
    static const float scale[3] = {1 / 0.5f, 1 / 0.5f, 1 / 0.5f};
    launch_crop_preprocess(di.as<CropItem>(), n, img_h, img_w, make_norm(mean, scale), kind == 0 ? -1.f : 0.f,
                           in.as<__half>(), b.stream);
    TV v;
    v.p = in.as<__half>(); v.n = n; v.h = img_h; v.w = img_w; v.c = 3; v.pitch = 8;
    launch_nhwc_to_nchw_f32(v, out.as<float>(), b.stream);
    cuda_check(cudaMemcpyAsync(nchw, out.p, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, b.stream), "copy");
    cuda_check(cudaStreamSynchronize(b.stream), "crop preprocess");
  });
}

}  // extern "C"
