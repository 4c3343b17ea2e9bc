// Debug / test entry points that expose one loaded graph (a Net) through the C ABI.
// Declared in include/b200ocr.h ("network-level entry points").
#include <cstring>
#include <string>

#include "../../include/b200ocr.h"
#include "capi_util.h"
#include "engine.h"

using namespace b200ocr;

struct b200ocr_net {
  Net* net = nullptr;
  float* d_in = nullptr;
  size_t d_in_bytes = 0;
  cudaStream_t stream = nullptr;
};

std::string profile_json(Net& net, cudaStream_t stream, int warmup, int reps, int thresh) {
  std::string o = "[";
  bool first = true;
  for (const auto& p : net.profile(stream, warmup, reps, thresh)) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s{\"name\":\"%s\",\"kind\":\"%s\",\"ms\":%.6f,\"flops\":%.0f,\"bytes\":%.0f,\"tensor_core\":%d}",
             first ? "" : ",", p.name.c_str(), p.kind.c_str(), p.ms, p.flops, p.bytes, p.tensor_core);
    o += buf;
    first = false;
  }
  return o + "]";
}


extern "C" {

int b200ocr_net_create(const char* model_dir, int device, int flags, b200ocr_net_t* out) {
  return capi_guard([&] {
    if (!model_dir || !out) throw std::invalid_argument("null argument");
    NetOptions o;
    o.keep_all = (flags & B200OCR_NET_KEEP_ALL) != 0;
    o.force_simt = (flags & B200OCR_NET_FORCE_SIMT) != 0;
    o.use_graph = (flags & B200OCR_NET_NO_GRAPH) == 0;
    auto* h = new b200ocr_net();
    try {
      h->net = new Net(model_dir, device, o);
      cuda_check(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    } catch (...) {
      delete h->net;
      delete h;
      throw;
    }
    *out = h;
  });
}

void b200ocr_net_destroy(b200ocr_net_t h) {
  if (!h) return;
  delete h->net;
  cudaFree(h->d_in);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int b200ocr_net_kind(b200ocr_net_t h, char* buf, int cap) {
  return capi_guard([&] {
    if (!h || !buf || cap < 1) throw std::invalid_argument("null argument");
    strncpy(buf, h->net->kind().c_str(), size_t(cap) - 1);
    buf[cap - 1] = 0;
  });
}

int b200ocr_net_plan_dump(b200ocr_net_t h, char* buf, int cap, int* needed) {
  return capi_guard([&] {
    if (!h) throw std::invalid_argument("null handle");
    std::string s = h->net->plan().dump();
    if (needed) *needed = int(s.size()) + 1;
    if (buf && cap > 0) {
      strncpy(buf, s.c_str(), size_t(cap) - 1);
      buf[cap - 1] = 0;
    }
  });
}

static int net_forward_impl(b200ocr_net_t h, const float* nchw, int n, int height, int width, int thresh_u8,
                            const int* widths) {
  return capi_guard([&] {
    if (!h || !nchw) throw std::invalid_argument("null argument");
    Net& net = *h->net;
    __half* in = net.prepare(n, height, width, widths, h->stream);
    const size_t bytes = size_t(n) * 3 * height * width * sizeof(float);
    if (bytes > h->d_in_bytes) {
      cudaFree(h->d_in);
      h->d_in = nullptr;
      cuda_check(cudaMalloc(&h->d_in, bytes), "cudaMalloc input staging");
      h->d_in_bytes = bytes;
    }
    cuda_check(cudaMemcpyAsync(h->d_in, nchw, bytes, cudaMemcpyHostToDevice, h->stream), "input upload");
    launch_nchw3_to_input(h->d_in, n, height, width, in, h->stream);
    net.run(h->stream, thresh_u8);
    cuda_check(cudaStreamSynchronize(h->stream), "forward");
  });
}

int b200ocr_net_forward(b200ocr_net_t h, const float* nchw, int n, int height, int width, int thresh_u8) {
  return net_forward_impl(h, nchw, n, height, width, thresh_u8, nullptr);
}
int b200ocr_net_forward_ragged(b200ocr_net_t h, const float* nchw, int n, int height, int width, const int* widths) {
  if (!widths) { b200ocr::set_last_error("null widths"); return 1; }
  return net_forward_impl(h, nchw, n, height, width, -1, widths);
}

int b200ocr_net_out_shape(b200ocr_net_t h, int shape[3]) {
  return capi_guard([&] {
    if (!h || !shape) throw std::invalid_argument("null argument");
    Shape3 s = h->net->out_shape();
    shape[0] = s.n; shape[1] = s.h; shape[2] = s.w;
  });
}

int b200ocr_net_output(b200ocr_net_t h, float* out_f32, uint8_t* out_bitmap, int32_t* out_idx) {
  return capi_guard([&] {
    if (!h) throw std::invalid_argument("null handle");
    Net& net = *h->net;
    Shape3 s = net.out_shape();
    const std::string& k = net.kind();
    size_t count = k == "det" ? size_t(s.n) * s.h * s.w : k == "cls" ? size_t(s.n) * net.plan().layers.back().cout
                                                                    : size_t(s.n) * s.w;
    if (out_f32) cuda_check(cudaMemcpy(out_f32, net.out_f32(), count * 4, cudaMemcpyDeviceToHost), "output copy");
    if (out_bitmap && net.out_bitmap())
      cuda_check(cudaMemcpy(out_bitmap, net.out_bitmap(), count, cudaMemcpyDeviceToHost), "bitmap copy");
    if (out_idx && net.out_idx())
      cuda_check(cudaMemcpy(out_idx, net.out_idx(), count * 4, cudaMemcpyDeviceToHost), "idx copy");
  });
}

int b200ocr_net_fetch(b200ocr_net_t h, const char* var, float* out, size_t cap_elems, int dims[4]) {
  return capi_guard([&] {
    if (!h || !var || !dims) throw std::invalid_argument("null argument");
    std::vector<float> v;
    if (!h->net->fetch(var, &v, dims)) throw std::runtime_error(std::string("no such tensor: ") + var);
    if (out) {
      if (v.size() > cap_elems) throw std::runtime_error("fetch buffer too small");
      memcpy(out, v.data(), v.size() * 4);
    }
  });
}

int b200ocr_net_profile(b200ocr_net_t h, int warmup, int reps, char** json) {
  return capi_guard([&] {
    if (!h || !json || reps < 1) throw std::invalid_argument("bad argument");
    std::string s = profile_json(*h->net, h->stream, warmup, reps, h->net->kind() == "det" ? 51 : -1);
    *json = static_cast<char*>(malloc(s.size() + 1));
    if (!*json) throw std::bad_alloc();
    memcpy(*json, s.c_str(), s.size() + 1);
  });
}

int b200ocr_net_launches(b200ocr_net_t h) { return h ? h->net->launches_per_run() : 0; }

}  // extern "C"
