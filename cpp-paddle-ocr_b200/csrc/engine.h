// Net: one loaded Paddle graph executing on one GPU as a sequence of fused sm_100a kernels.
//
// Takes the place of `paddle_infer::Predictor` in the reference stages
// (predictor_->Run(), reference src/ocr_det.cpp:120, src/ocr_cls.cpp:76, src/ocr_rec.cpp:85).
// A Net is single-threaded like the predictor it replaces; one instance per worker.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.h"
#include "plan.h"

namespace b200ocr {

struct NetOptions {
  bool keep_all = false;    // debug: every tensor keeps its own memory so it can be fetched after a run
  bool use_graph = true;    // replay the layer sequence as a CUDA graph per input shape
  bool force_simt = false;  // debug: run dense convs / CTC head on the CUDA-core kernels
};

struct Shape3 {
  int n = 0, h = 0, w = 0;
};

class Net {
 public:
  // Loads <model_dir>/inference.pdmodel + .pdiparams, builds the fused plan, uploads weights.
  Net(const std::string& model_dir, int device, const NetOptions& opt);
  ~Net();
  Net(const Net&) = delete;
  Net& operator=(const Net&) = delete;

  const Plan& plan() const { return plan_; }
  const std::string& kind() const { return plan_.kind; }
  int device() const { return device_; }

  // Prepare buffers (and the CUDA graph) for an input of n x h x w.  The returned pointer is the
  // network input: NHWC fp16 with channel pitch 8 (channels 3..7 must be written as zero).
  // `widths` (optional, host int[n]): ragged batch -- row i is only widths[i] <= w columns wide; every layer then
  // treats the columns beyond a row's (layer-scaled) width as the zero padding it would see at the edge of a
  // tensor of that width, so each row's result is bit-identical to running it in a dense batch of its own
  // width.  The caller must write zeros into the input beyond widths[i].  Sequence graphs (rec) only.
  // `stream`: the stream this net's work is queued on; it is synchronised (never the whole device: another worker may
  // be capturing a CUDA graph) before memory that queued work may still use is released.
  __half* prepare(int n, int h, int w, const int* widths = nullptr, cudaStream_t stream = nullptr);
  // Execute the forward pass for the last prepared shape on `stream`.
  //   det: `thresh_u8` >= 0 also writes the thresholded bitmap.
  // `src` (optional, only when stem_fusable()): the first convolution pre-processes these 8-bit sources on the fly
  // instead of reading the network input; the caller then skips its pre-processing kernel and the input stays unwritten.
  void run(cudaStream_t stream, int thresh_u8 = -1, const StemSource* src = nullptr);
  bool stem_fusable() const;

  // Outputs of the last prepared shape (device pointers, valid until the next prepare()).
  Shape3 out_shape() const;              // det: (n, H, W); cls: (n,1,1); rec: (n, 1, T)
  const float* out_f32() const;          // det: prob [n,H,W]; cls: softmax [n,2]; rec: max prob [n*T]
  const uint8_t* out_bitmap() const;     // det only
  const int* out_idx() const;            // rec only: argmax class per (n,t)

  // debug: copy a tensor (by Paddle var name) of the last run to host as fp32 NCHW.
  bool fetch(const std::string& var, std::vector<float>* out, int dims[4]);

  // Per-layer device time of the last prepared shape: every fused layer is launched on its own between two
  // CUDA events on `stream` (after `warmup` untimed passes, L2 flushed before every timed launch), averaged
  // over `reps`.  Also reports the layer's
  // algorithmic FLOPs and HBM bytes (activations in + out + weights) so that a caller can place it on a roofline.
  struct LayerProfile { std::string name, kind; double ms, flops, bytes; int tensor_core; };
  std::vector<LayerProfile> profile(cudaStream_t stream, int warmup, int reps, int thresh_u8 = -1);

  size_t arena_bytes() const { return arena_bytes_; }
  int launches_per_run() const;

 private:
  struct Inst;
  Inst* instantiate(int n, int h, int w);
  void infer(int n, int h, int w, std::vector<Shape3>* ts, std::vector<int>* splits, std::vector<int>* hw,
             std::vector<int>* gap_src) const;
  void record(Inst& I, cudaStream_t s, int thresh_u8, const std::function<void(int, bool)>* hook = nullptr,
              int only = -1);

  Plan plan_;
  NetOptions opt_;
  int device_ = 0;
  __half* d_wh_ = nullptr;
  float* d_wf_ = nullptr;
  uint8_t* arena_ = nullptr;
  size_t arena_bytes_ = 0;
  std::map<std::tuple<int, int, int>, std::unique_ptr<Inst>> cache_;
  Inst* cur_ = nullptr;
  StemSource stem_;         // source of the last run(): profile() times the same first layer
  // ragged batches: per tensor, per row valid widths (int[nt][n]); host copy is pageable (see prepare)
  bool ragged_ = false;
  int* vw_pin_ = nullptr;
  int* vw_dev_ = nullptr;
  size_t vw_cap_ = 0;
  int* se_cnt_ = nullptr;   // ticket counters of the fused pool + SE-gate kernel (int[n], self-resetting)
  int se_cnt_cap_ = 0;
};

void cuda_check(cudaError_t e, const char* what);
// Host wait for everything queued on `s`.  A worker thread waits twice per batch (detected boxes, decoded ids).
// Default: cudaStreamSynchronize (spins on a host core).  B200OCR_BLOCKING_SYNC=1: record a cudaEventBlockingSync event
// and sleep on it -- frees the core, but measured on one B200 with 16 host cores (profiles/r02_notes.md) the wake-up
// latency at the two sync points costs 34 % of the device-resident rate (5922 -> 3920 images/s) and 5.5 % end to end
// (5703 -> 5386), so it is an opt-in for hosts with fewer cores than workers.
void host_wait(cudaStream_t s, const char* what);

}  // namespace b200ocr
