#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace b200ocr {

void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

void host_wait(cudaStream_t s, const char* what) {
  static const bool block = [] { const char* v = getenv("B200OCR_BLOCKING_SYNC"); return v && v[0] == '1'; }();
  if (!block) { cuda_check(cudaStreamSynchronize(s), what); return; }
  int dev = 0;
  cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
  thread_local cudaEvent_t ev[64] = {};
  if (dev < 0 || dev >= 64) { cuda_check(cudaStreamSynchronize(s), what); return; }
  if (!ev[dev]) cuda_check(cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventRecord(ev[dev], s), what);
  cuda_check(cudaEventSynchronize(ev[dev]), what);
}

struct Net::Inst {
  int n = 0, h = 0, w = 0;
  std::vector<Shape3> ts;            // per tensor
  std::vector<int> splits, hw;       // per tensor (Gap outputs): partial-sum layout
  std::vector<int> gap_src;          // per tensor (Gap outputs): the pooled tensor
  std::vector<size_t> boff, bbytes;  // per buffer
  std::vector<ConvTcPlan> tc;        // per layer (impl == nullptr -> CUDA-core kernel)
  size_t need = 0;
  cudaGraphExec_t graph = nullptr;
  int graph_thresh = -2;
  StemSource graph_src;              // the 8-bit source baked into the captured stem launch
  int runs = 0;
  int launches = 0;
  ~Inst() {
    for (auto& p : tc)
      if (p.impl) free_conv_tc_plan(&p);
    if (graph) cudaGraphExecDestroy(graph);
  }
};

Net::Net(const std::string& model_dir, int device, const NetOptions& opt) : opt_(opt), device_(device) {
  std::string mfile, pfile;
  if (!find_model_files(model_dir, &mfile, &pfile))
    throw std::runtime_error("No valid model file found in " + model_dir);
  PdProgram prog;
  load_program(mfile, &prog);
  load_params(pfile, &prog);
  build_plan(prog, &plan_);
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  cuda_check(cudaMalloc(&d_wh_, std::max<size_t>(plan_.wh.size(), 8) * sizeof(uint16_t)), "cudaMalloc weights");
  cuda_check(cudaMalloc(&d_wf_, std::max<size_t>(plan_.wf.size(), 4) * sizeof(float)), "cudaMalloc weights");
  cuda_check(cudaMemcpy(d_wh_, plan_.wh.data(), plan_.wh.size() * sizeof(uint16_t), cudaMemcpyHostToDevice),
             "upload weights");
  cuda_check(cudaMemcpy(d_wf_, plan_.wf.data(), plan_.wf.size() * sizeof(float), cudaMemcpyHostToDevice),
             "upload weights");
}

Net::~Net() {
  cache_.clear();
  cudaFree(d_wh_);
  cudaFree(d_wf_);
  cudaFree(arena_);
  cudaFree(vw_dev_);
  cudaFree(se_cnt_);
  free(vw_pin_);
}

namespace {
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

void Net::infer(int n, int h, int w, std::vector<Shape3>* ts_, std::vector<int>* splits, std::vector<int>* hwv,
                std::vector<int>* gap_src) const {
  const int nt = int(plan_.tensors.size());
  std::vector<Shape3>& ts = *ts_;
  ts.assign(nt, Shape3());
  if (splits) splits->assign(nt, 0);
  if (hwv) hwv->assign(nt, 0);
  if (gap_src) gap_src->assign(nt, -1);
  ts[plan_.input] = Shape3{n, h, w};
  auto fail = [&](const Layer& L, const char* why) {
    throw std::runtime_error("shape error at layer " + L.name + ": " + why);
  };
  // ---- shape inference (concat views made by re-homing have no producer: they take a member's shape)
  auto resolve = [&](int t) {
    if (ts[t].n == 0)
      for (int u = 0; u < nt; ++u)
        if (u != t && plan_.tensors[u].buf == plan_.tensors[t].buf && ts[u].n) { ts[t] = ts[u]; break; }
    return ts[t];
  };
  for (const Layer& L : plan_.layers) {
    const Shape3 in = resolve(L.in);
    if (in.n == 0) fail(L, "input shape unknown");
    Shape3 o = in;
    switch (L.kind) {
      case LKind::Conv:
      case LKind::DwConv:
        o.h = (in.h + 2 * L.ph - L.kh) / L.sh + 1;
        o.w = (in.w + 2 * L.pw - L.kw) / L.sw + 1;
        if (o.h < 1 || o.w < 1) fail(L, "input smaller than the filter");
        break;
      case LKind::Pool:
        // Paddle PoolOutputSize with C++ truncating division (28-px-high rec input: (2-3)/3+1 = 1)
        o.h = (in.h - L.kh) / L.sh + 1;
        o.w = (in.w - L.kw) / L.sw + 1;
        if (o.h < 1 || o.w < 1) fail(L, "pool output is empty");
        break;
      case LKind::Gap:
        o = Shape3{in.n, 1, 1};
        if (splits) (*splits)[L.out] = gap_splits(in.n, in.h, in.w, L.cin, plan_.kind == "rec");
        if (hwv) (*hwv)[L.out] = in.h * in.w;
        if (gap_src) (*gap_src)[L.out] = L.in;
        break;
      case LKind::SeFc:
      case LKind::FcSoftmax:
        o = Shape3{in.n, 1, 1};
        break;
      case LKind::UpAdd: {
        const Shape3 b = ts[L.in2];
        if (b.h * 2 != in.h || b.w * 2 != in.w) fail(L, "FPN levels are not 2x apart (input not a multiple of 32?)");
        break;
      }
      case LKind::UpCat: {
        const int sh[4] = {L.kh, L.kw, L.sh, L.sw};
        for (int k = 0; k < 4 && L.ins[k] >= 0; ++k) {
          const Shape3 s = ts[L.ins[k]];
          if ((s.h << sh[k]) != in.h || (s.w << sh[k]) != in.w) fail(L, "concat inputs do not upsample to one size");
        }
        break;
      }
      case LKind::Attn:
        if (in.h != 1) fail(L, "sequence neck needs feature height 1 (rec_img_h must be 28..48-class)");
        break;
      case LKind::CtcHead:
        if (in.h != 1) fail(L, "CTC head needs feature height 1");
        break;
      case LKind::DbHead:
        o = Shape3{in.n, in.h * 4, in.w * 4};
        break;
      default: break;
    }
    ts[L.out] = o;
  }
  for (int t = 0; t < nt; ++t) resolve(t);
}

Net::Inst* Net::instantiate(int n, int h, int w) {
  auto I = std::make_unique<Inst>();
  I->n = n; I->h = h; I->w = w;
  const int nt = int(plan_.tensors.size()), nb = int(plan_.buffers.size()), nl = int(plan_.layers.size());
  infer(n, h, w, &I->ts, &I->splits, &I->hw, &I->gap_src);
  // ---- buffer sizes
  I->boff.assign(nb, 0);
  I->bbytes.assign(nb, 0);
  std::vector<int> first(nb, std::numeric_limits<int>::max()), last(nb, -1);
  for (int t = 0; t < nt; ++t) {
    const TensorDesc& td = plan_.tensors[t];
    if (plan_.buffers[td.buf].c_total == 0) continue;
    const Shape3 s = I->ts[t];
    size_t bytes;
    if (!td.vec) bytes = size_t(s.n) * s.h * s.w * round_up(plan_.buffers[td.buf].c_total, 8) * sizeof(__half);
    else bytes = 0;  // vector buffers are sized by their producing layer below
    I->bbytes[td.buf] = std::max(I->bbytes[td.buf], bytes);
  }
  for (int li = 0; li < nl; ++li) {
    const Layer& L = plan_.layers[li];
    const int ob = plan_.tensors[L.out].buf;
    const Shape3 in = I->ts[L.in];
    switch (L.kind) {
      case LKind::Gap: I->bbytes[ob] = size_t(in.n) * I->splits[L.out] * round_up(L.cin, 8) * 4; break;
      case LKind::SeFc: I->bbytes[ob] = size_t(in.n) * round_up(L.cin, 8) * 4; break;
      case LKind::FcSoftmax: I->bbytes[ob] = size_t(in.n) * L.cout * 4; break;
      case LKind::DbHead: I->bbytes[ob] = align_up(size_t(in.n) * in.h * in.w * 16 * 4, 256) + size_t(in.n) * in.h * in.w * 16; break;
      case LKind::CtcHead: I->bbytes[ob] = align_up(size_t(in.n) * in.w * 4, 256) * 2; break;
      default: break;
    }
    // the pooled partial sums of an SE block stay live until its Scale layer: the fused pool + gate + scale kernel
    // writes the scaled map while other samples' blocks still read / write their partial sums
    if (L.kind == LKind::Gap && li + 2 < nl && plan_.layers[li + 1].kind == LKind::SeFc && plan_.layers[li + 2].kind == LKind::Scale) {
      last[ob] = std::max(last[ob], li + 2);
      const int sb = plan_.tensors[plan_.layers[li + 2].out].buf;   // ... and the scaled map is written from here on
      first[sb] = std::min(first[sb], li);
    }
    auto touch_r = [&](int t) { if (t >= 0) { int b = plan_.tensors[t].buf; last[b] = std::max(last[b], li); } };
    touch_r(L.in); touch_r(L.in2); touch_r(L.residual);
    for (int k = 0; k < 4; ++k) touch_r(L.ins[k]);
    first[ob] = std::min(first[ob], li);
    last[ob] = std::max(last[ob], li);
  }
  first[plan_.tensors[plan_.input].buf] = -1;
  last[plan_.tensors[plan_.output].buf] = nl;
  // ---- arena layout: first-fit over live intervals
  struct Live { size_t off, bytes; int last; };
  std::vector<Live> live;
  std::vector<int> order;
  for (int b = 0; b < nb; ++b)
    if (I->bbytes[b]) order.push_back(b);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return first[a] < first[b]; });
  size_t top = 0;
  for (int b : order) {
    const size_t bytes = align_up(I->bbytes[b], 1024);
    if (!opt_.keep_all)
      live.erase(std::remove_if(live.begin(), live.end(), [&](const Live& l) { return l.last < first[b]; }),
                 live.end());
    std::sort(live.begin(), live.end(), [](const Live& a, const Live& c) { return a.off < c.off; });
    size_t off = 0;
    for (const Live& l : live) {
      if (off + bytes <= l.off) break;
      off = std::max(off, l.off + l.bytes);
    }
    I->boff[b] = off;
    live.push_back(Live{off, bytes, last[b]});
    top = std::max(top, off + bytes);
  }
  I->need = top;
  I->tc.assign(nl, ConvTcPlan());
  auto key = std::make_tuple(n, h, w);
  Inst* raw = I.get();
  cache_[key] = std::move(I);
  return raw;
}

static TV make_tv(uint8_t* arena, const Plan& plan, const std::vector<size_t>& boff,
                  const std::vector<Shape3>& ts, int t) {
  TV v;
  const TensorDesc& td = plan.tensors[t];
  v.pitch = round_up(plan.buffers[td.buf].c_total, 8);
  v.p = reinterpret_cast<__half*>(arena + boff[td.buf]) + td.c_off;
  v.n = ts[t].n; v.h = ts[t].h; v.w = ts[t].w; v.c = td.c;
  return v;
}

__half* Net::prepare(int n, int h, int w, const int* widths, cudaStream_t stream) {
  if (n < 1 || h < 1 || w < 1) throw std::runtime_error("Net::prepare: empty input");
  ragged_ = false;
  if (widths) {
    bool all_full = true;
    for (int i = 0; i < n; ++i) {
      if (widths[i] < 1 || widths[i] > w) throw std::runtime_error("Net::prepare: row width out of range");
      all_full &= widths[i] == w;
    }
    if (!all_full) {
      for (const Layer& L : plan_.layers)
        if (L.kind == LKind::UpAdd || L.kind == LKind::UpCat || L.kind == LKind::DbHead || L.kind == LKind::FcSoftmax)
          throw std::runtime_error("ragged batches are only supported for sequence (rec) graphs");
      const size_t nt = plan_.tensors.size(), need = nt * size_t(n);
      cuda_check(cudaSetDevice(device_), "cudaSetDevice");
      if (need > vw_cap_) {
        cuda_check(cudaStreamSynchronize(stream), "sync before table growth");
        cudaFree(vw_dev_);
        free(vw_pin_);
        vw_cap_ = need + need / 2;
        cuda_check(cudaMalloc(&vw_dev_, vw_cap_ * sizeof(int)), "cudaMalloc width table");
        // PAGEABLE on purpose: cudaMemcpyAsync from pageable memory snapshots the source before it returns, so the
        // next prepare() may overwrite this table while earlier launches are still queued
        vw_pin_ = static_cast<int*>(malloc(vw_cap_ * sizeof(int)));
        if (!vw_pin_) throw std::bad_alloc();
      }
      std::map<int, std::vector<Shape3>> per_width;  // per distinct row width: every tensor's shape
      for (int i = 0; i < n; ++i) {
        auto it = per_width.find(widths[i]);
        if (it == per_width.end()) {
          std::vector<Shape3> ts;
          infer(1, h, widths[i], &ts, nullptr, nullptr, nullptr);
          it = per_width.emplace(widths[i], std::move(ts)).first;
        }
        for (size_t t = 0; t < nt; ++t) vw_pin_[t * n + i] = it->second[t].w;
      }
      ragged_ = true;
    }
  }
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  if (n > se_cnt_cap_) {
    cuda_check(cudaStreamSynchronize(stream), "sync before counter growth");
    cudaFree(se_cnt_);
    se_cnt_cap_ = n + n / 2 + 64;
    cuda_check(cudaMalloc(&se_cnt_, size_t(se_cnt_cap_) * sizeof(int)), "cudaMalloc SE counters");
    cuda_check(cudaMemset(se_cnt_, 0, size_t(se_cnt_cap_) * sizeof(int)), "cudaMemset SE counters");
  }
  auto key = std::make_tuple(n, h, w);
  auto it = cache_.find(key);
  Inst* I = it != cache_.end() ? it->second.get() : instantiate(n, h, w);
  if (I->need > arena_bytes_) {
    // grow the shared arena; every cached instance holds pointers into the old one
    const size_t need = I->need;
    cuda_check(cudaStreamSynchronize(stream), "sync before arena growth");
    cache_.clear();
    cudaFree(arena_);
    arena_ = nullptr;
    arena_bytes_ = align_up(2 * need, size_t(1) << 20);  // generous: growing again means cudaFree + cudaMalloc, which stalls the whole device
    cuda_check(cudaMalloc(&arena_, arena_bytes_), "cudaMalloc activation arena");
    I = instantiate(n, h, w);
  }
  if (cache_.size() > 64) {  // bound the number of cached shapes (variable-width rec batches)
    for (auto c = cache_.begin(); c != cache_.end();) {
      if (c->second.get() != I) c = cache_.erase(c); else ++c;
    }
  }
  cur_ = I;
  return reinterpret_cast<__half*>(arena_ + I->boff[plan_.tensors[plan_.input].buf]);
}

void Net::record(Inst& I, cudaStream_t s, int thresh_u8, const std::function<void(int, bool)>* hook, int only) {
  auto tv = [&](int t) { return make_tv(arena_, plan_, I.boff, I.ts, t); };
  auto vwp = [&](int t) -> const int* { return ragged_ ? vw_dev_ + size_t(t) * I.n : nullptr; };
  auto vecp = [&](int t) { return reinterpret_cast<float*>(arena_ + I.boff[plan_.tensors[t].buf]); };
  int launches = 0;
  static const bool no_se_fuse = getenv("B200OCR_NO_SE_FUSE") != nullptr;
  // Applying the gate inside the pool + gate kernel (one launch less per SE block, one block per sample) measured 2.8 %
  // SLOWER than the separate scale kernel on C4 at 192 cards per step (profiles/r02_notes.md section 13): opt-in.
  static const bool no_se_apply = !(getenv("B200OCR_SE_APPLY_FUSE") && atoi(getenv("B200OCR_SE_APPLY_FUSE")) != 0);
  int scale_fused_at = -1;
  for (size_t li = 0; li < plan_.layers.size(); ++li) {
    if (only >= 0 && int(li) != only) continue;
    const Layer& L = plan_.layers[li];
    if (hook) (*hook)(int(li), true);
    Epi e;
    e.act = int(L.act); e.a = L.act_a; e.b = L.act_b; e.s2 = L.post_scale; e.t2 = L.post_shift;
    if (L.residual >= 0) {
      TV r = tv(L.residual);
      e.res = r.p;
      e.res_pitch = r.pitch;
    }
    ConvGeom g;
    g.kh = L.kh; g.kw = L.kw; g.sh = L.sh; g.sw = L.sw; g.ph = L.ph; g.pw = L.pw;
    g.cin_pad = L.cin_pad; g.cout_pad = L.cout_pad;
    switch (L.kind) {
      case LKind::Conv: {
        TV in = tv(L.in), out = tv(L.out);
        const __half* w = d_wh_ + L.wh_off;
        const float* bias = d_wf_ + L.bias_off;
        if (li == 0 && stem_.kind != 0) {
          // pre-processing fused into the first convolution: the input tensor is never written (kernels_simt.cu)
          if (!fused_stem_eligible(in, out, g, e)) throw std::runtime_error("fused stem: the first layer is not a 3x3 stride-2 stem");
          launch_fused_stem(stem_, in, out, w, bias, g, e, s, vwp(L.out));
        } else if (!opt_.force_simt && launch_pwconv_mma(in, out, w, bias, g, e, s, vwp(L.out))) {
          // narrow 1x1 convolution: streamed on mma.sync
        } else if (!opt_.force_simt && conv_tc_eligible(in, out, g)) {
          if (!I.tc[li].impl) I.tc[li] = make_conv_tc_plan(in, out, w, g);
          launch_conv_tc(I.tc[li], bias, e, s, vwp(L.out));
        } else {
          launch_conv_simt(in, out, w, bias, g, e, s, vwp(L.out));
        }
        break;
      }
      case LKind::DwConv:
        // the detector and the classifier keep fp32 depthwise weights: their outputs (thresholded probability map,
        // soft-max score) have to stay within 1e-2 of the fp32 reference; rec takes the FHFMA kernel (fp16 weights)
        launch_dwconv(tv(L.in), tv(L.out), d_wf_ + L.wf_off, plan_.kind == "rec" ? d_wh_ + L.wh_off : nullptr, g, e, s,
                      vwp(L.out));
        break;
      case LKind::Gap: {
        // pool + SE gate in one launch when the pooled vector only feeds the gate that follows
        const bool fuse = !no_se_fuse && li + 1 < plan_.layers.size() && plan_.layers[li + 1].kind == LKind::SeFc &&
                          plan_.layers[li + 1].in == L.out;
        if (fuse) {
          const Layer& F = plan_.layers[li + 1];
          SeFuse f;
          f.blk = d_wf_ + F.wf_off; f.gate = vecp(F.out); f.counters = se_cnt_;
          f.c = F.cin; f.cmid = F.cmid; f.slope = F.act_a; f.offset = F.act_b;
          f.inv_hw = 1.f / float(I.hw[L.out]);
          f.vw_in = vwp(L.in); f.h = I.ts[L.in].h;
          // ... and the gate is applied there too when the Scale layer follows (x -> pool -> gate -> x * gate) and the
          // batch has enough samples to keep the GPU busy with one block per sample
          if (!no_se_apply && only < 0 && li + 2 < plan_.layers.size() && I.ts[L.in].n >= 16) {
            const Layer& S = plan_.layers[li + 2];
            if (S.kind == LKind::Scale && S.in == L.in && S.in2 == F.out) {
              TV so = tv(S.out);
              f.sc_out = so.p; f.sc_out_pitch = so.pitch; f.sc_add_x = S.scale_residual ? 1 : 0;
              scale_fused_at = int(li + 2);
            }
          }
          launch_gap_partial(tv(L.in), vecp(L.out), I.splits[L.out], plan_.kind == "rec", s, &f);
        } else {
          launch_gap_partial(tv(L.in), vecp(L.out), I.splits[L.out], plan_.kind == "rec", s);
        }
        break;
      }
      case LKind::SeFc:
        if (!no_se_fuse && li > 0 && plan_.layers[li - 1].kind == LKind::Gap && plan_.layers[li - 1].out == L.in) {
          --launches;  // fused into the pooling kernel above: nothing is launched for this layer
          break;
        }
        launch_se_fc(vecp(L.in), I.splits[L.in], I.hw[L.in], I.ts[L.in].n, L.cin, L.cmid, d_wf_ + L.wf_off,
                     L.act_a, L.act_b, vecp(L.out), s, vwp(I.gap_src[L.in]), I.ts[I.gap_src[L.in]].h);
        break;
      case LKind::Scale:
        if (scale_fused_at == int(li)) { --launches; break; }  // done by the pool + gate kernel two layers up
        launch_scale(tv(L.in), vecp(L.in2), L.scale_residual, tv(L.out), s);
        break;
      case LKind::UpAdd: launch_upadd(tv(L.in), tv(L.in2), tv(L.out), s); break;
      case LKind::UpCat: {
        TV ins[4];
        int sh[4] = {L.kh, L.kw, L.sh, L.sw}, nin = 0;
        for (int k = 0; k < 4 && L.ins[k] >= 0; ++k) { ins[k] = tv(L.ins[k]); ++nin; }
        launch_upcat(ins, sh, nin, tv(L.out), s);
        break;
      }
      case LKind::Pool: launch_pool(tv(L.in), tv(L.out), L.kh, L.kw, L.sh, L.sw, L.pool_max, s, vwp(L.out)); break;
      case LKind::Add: launch_add(tv(L.in), tv(L.in2), tv(L.out), s); break;
      case LKind::LayerNorm: launch_layernorm(tv(L.in), tv(L.out), d_wf_ + L.wf_off, L.eps, s, vwp(L.out)); break;
      case LKind::Attn: launch_attention(tv(L.in), tv(L.out), L.heads, L.head_dim, L.attn_scale, s, vwp(L.out)); break;
      case LKind::DbHead: {
        TV in = tv(L.in);
        float* prob = vecp(L.out);
        uint8_t* bm = reinterpret_cast<uint8_t*>(prob) + align_up(size_t(in.n) * in.h * in.w * 16 * 4, 256);
        launch_dbhead(in, d_wf_ + L.wf_off, L.cmid, prob, thresh_u8 >= 0 ? bm : nullptr, thresh_u8, s);
        break;
      }
      case LKind::FcSoftmax:
        launch_fc_softmax(vecp(L.in), I.splits[L.in], I.hw[L.in], I.ts[L.in].n, L.cin, L.cout,
                          d_wf_ + L.wf_off, vecp(L.out), s);
        break;
      case LKind::CtcHead: {
        TV in = tv(L.in);
        float* prob = vecp(L.out);
        int* idx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(prob) + align_up(size_t(in.n) * in.w * 4, 256));
        if (!opt_.force_simt && ctc_tc_eligible(in, L.cin_pad))
          launch_ctc_head_tc(in, d_wh_ + L.wh_off, d_wf_ + L.wf_off, L.cin_pad, L.cout, L.cout_pad, idx, prob, s,
                             vwp(L.in));
        else
          launch_ctc_head_simt(in, d_wh_ + L.wh_off, d_wf_ + L.bias_off, L.cin_pad, L.cout, L.cout_pad, idx, prob, s,
                               vwp(L.in));
        break;
      }
    }
    ++launches;
    if (hook) (*hook)(int(li), false);
  }
  if (only < 0) I.launches = launches;
}

std::vector<Net::LayerProfile> Net::profile(cudaStream_t stream, int warmup, int reps, int thresh_u8) {
  if (!cur_) throw std::runtime_error("Net::profile before prepare");
  Inst& I = *cur_;
  const int nl = int(plan_.layers.size());
  for (int k = 0; k < warmup; ++k) record(I, stream, thresh_u8);
  // Every launch is timed on its own between two CUDA events on `stream`, after a write of a buffer larger than
  // L2 (so a layer never finds its input in cache only because the previous repetition left it there).  All work
  // is enqueued without host synchronisation: the GPU never waits for the host inside a timed window.
  const size_t flush_bytes = size_t(256) << 20;
  void* flush = nullptr;
  cuda_check(cudaMalloc(&flush, flush_bytes), "cudaMalloc L2 flush buffer");
  std::vector<cudaEvent_t> ev(size_t(nl) * reps * 2);
  for (auto& e : ev) cuda_check(cudaEventCreate(&e), "cudaEventCreate");
  // B200OCR_PROFILE_WARM=1 (diagnostic): the layers run in network order, back to back, without the flush -- every
  // layer finds in L2 what its producer left there, as inside the captured graph.  Not used for reported numbers.
  static const bool warm = getenv("B200OCR_PROFILE_WARM") != nullptr;
  if (warm) {
    for (int r = 0; r < reps; ++r)
      for (int li = 0; li < nl; ++li) {
        cudaEventRecord(ev[(size_t(li) * reps + r) * 2], stream);
        record(I, stream, thresh_u8, nullptr, li);
        cudaEventRecord(ev[(size_t(li) * reps + r) * 2 + 1], stream);
      }
  } else {
    for (int li = 0; li < nl; ++li)
      for (int r = 0; r < reps; ++r) {
        cudaMemsetAsync(flush, r, flush_bytes, stream);
        cudaEventRecord(ev[(size_t(li) * reps + r) * 2], stream);
        record(I, stream, thresh_u8, nullptr, li);
        cudaEventRecord(ev[(size_t(li) * reps + r) * 2 + 1], stream);
      }
  }
  cuda_check(cudaStreamSynchronize(stream), "profile pass");
  std::vector<double> ms(nl, 0.0);
  for (int li = 0; li < nl; ++li)
    for (int r = 0; r < reps; ++r) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[(size_t(li) * reps + r) * 2], ev[(size_t(li) * reps + r) * 2 + 1]);
      ms[li] += t;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  cudaFree(flush);
  static const char* kn[] = {"Conv", "DwConv", "Gap", "SeFc", "Scale", "UpAdd", "UpCat", "Pool",
                             "Add", "LayerNorm", "Attn", "DbHead", "FcSoftmax", "CtcHead"};
  std::vector<LayerProfile> out;
  for (int li = 0; li < nl; ++li) {
    const Layer& L = plan_.layers[li];
    const Shape3 in = I.ts[L.in];
    const Shape3 o = I.ts[L.out];
    LayerProfile p;
    p.name = L.name;
    p.kind = kn[int(L.kind)];
    p.ms = ms[li] / reps;
    p.tensor_core = I.tc[li].impl != nullptr;
    const double pin = double(in.n) * in.h * in.w, pout = double(o.n) * o.h * o.w;
    const int taps = L.kh * L.kw;
    p.flops = 0;
    p.bytes = (pin * L.cin + pout * L.cout) * 2;
    switch (L.kind) {
      case LKind::Conv:
        p.flops = 2.0 * pout * L.cout * L.cin * taps;
        p.bytes += double(L.cout) * L.cin * taps * 2 + (L.residual >= 0 ? pout * L.cout * 2 : 0);
        break;
      case LKind::DwConv: p.flops = 2.0 * pout * L.cout * taps; break;
      case LKind::CtcHead:
        p.flops = 2.0 * pin * L.cin * L.cout;
        p.bytes = pin * L.cin * 2 + double(L.cout) * L.cin * 2 + pin * 8;
        break;
      case LKind::DbHead:
        p.flops = 2.0 * pin * 4 * (double(L.cin) * L.cmid + L.cmid * 4);
        p.bytes = pin * L.cin * 2 + pin * 16 * 5;
        break;
      case LKind::Attn:
        p.flops = 4.0 * in.n * double(in.w) * in.w * L.heads * L.head_dim;
        break;
      case LKind::UpAdd: p.bytes = pout * L.cout * 2 * 2 + pout / 4 * L.cout * 2; break;
      case LKind::Add: p.bytes = pout * L.cout * 2 * 3; break;
      case LKind::Gap: p.bytes = pin * L.cin * 2; break;
      case LKind::SeFc: case LKind::FcSoftmax: p.bytes = 0; break;
      default: break;
    }
    out.push_back(p);
  }
  return out;
}

bool Net::stem_fusable() const {
  if (plan_.layers.empty() || plan_.layers[0].kind != LKind::Conv || plan_.layers[0].in != plan_.input) return false;
  const Layer& L = plan_.layers[0];
  TV in, out;
  in.c = L.cin; in.h = 64; in.w = 64;
  out.c = L.cout; out.h = 32; out.w = 32;
  ConvGeom g;
  g.kh = L.kh; g.kw = L.kw; g.sh = L.sh; g.sw = L.sw; g.ph = L.ph; g.pw = L.pw;
  Epi e;
  e.act = int(L.act);
  if (L.residual >= 0) return false;
  // the input tensor must have no other reader
  for (size_t li = 1; li < plan_.layers.size(); ++li) {
    const Layer& M = plan_.layers[li];
    if (M.in == plan_.input || M.in2 == plan_.input || M.residual == plan_.input) return false;
    for (int k = 0; k < 4; ++k) if (M.ins[k] == plan_.input) return false;
  }
  return fused_stem_eligible(in, out, g, e);
}

void Net::run(cudaStream_t stream, int thresh_u8, const StemSource* src) {
  if (!cur_) throw std::runtime_error("Net::run before prepare");
  Inst& I = *cur_;
  ++I.runs;
  stem_ = src ? *src : StemSource();
  if (ragged_)
    cuda_check(cudaMemcpyAsync(vw_dev_, vw_pin_, plan_.tensors.size() * size_t(I.n) * sizeof(int), cudaMemcpyHostToDevice,
                               stream), "width table upload");
  if (!opt_.use_graph || I.runs == 1 || ragged_) {  // ragged shapes rarely repeat: always eager
    // first run of a shape is eager: encodes tensor maps, sets function attributes
    record(I, stream, thresh_u8);
    cuda_check(cudaGetLastError(), "forward launch");
    return;
  }
  if (!I.graph || I.graph_thresh != thresh_u8 || I.graph_src.kind != stem_.kind || I.graph_src.items != stem_.items ||
      I.graph_src.pad_value != stem_.pad_value || memcmp(&I.graph_src.np, &stem_.np, sizeof(NormParams)) != 0) {
    if (I.graph) { cudaGraphExecDestroy(I.graph); I.graph = nullptr; }
    cudaGraph_t g = nullptr;
    cuda_check(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal), "begin capture");
    try {
      record(I, stream, thresh_u8);
    } catch (...) {  // never leave the stream in capture mode
      cudaStreamEndCapture(stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    cuda_check(cudaStreamEndCapture(stream, &g), "end capture");
    cuda_check(cudaGraphInstantiate(&I.graph, g, 0), "graph instantiate");
    cudaGraphDestroy(g);
    I.graph_thresh = thresh_u8;
    I.graph_src = stem_;
  }
  cuda_check(cudaGraphLaunch(I.graph, stream), "graph launch");
}

int Net::launches_per_run() const { return cur_ ? cur_->launches : 0; }

Shape3 Net::out_shape() const {
  if (!cur_) return Shape3();
  const Layer& L = plan_.layers.back();
  const Shape3 in = cur_->ts[L.in];
  if (L.kind == LKind::DbHead) return Shape3{in.n, in.h * 4, in.w * 4};
  if (L.kind == LKind::CtcHead) return Shape3{in.n, 1, in.w};
  return Shape3{in.n, 1, 1};
}

const float* Net::out_f32() const {
  return reinterpret_cast<const float*>(arena_ + cur_->boff[plan_.tensors[plan_.output].buf]);
}

const uint8_t* Net::out_bitmap() const {
  const Layer& L = plan_.layers.back();
  if (L.kind != LKind::DbHead) return nullptr;
  const Shape3 in = cur_->ts[L.in];
  return reinterpret_cast<const uint8_t*>(out_f32()) + align_up(size_t(in.n) * in.h * in.w * 16 * 4, 256);
}

const int* Net::out_idx() const {
  const Layer& L = plan_.layers.back();
  if (L.kind != LKind::CtcHead) return nullptr;
  const Shape3 in = cur_->ts[L.in];
  return reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(out_f32()) +
                                      align_up(size_t(in.n) * in.w * 4, 256));
}

bool Net::fetch(const std::string& var, std::vector<float>* out, int dims[4]) {
  if (!cur_) return false;
  const int t = plan_.find_tensor(var);
  if (t < 0) return false;
  const TensorDesc& td = plan_.tensors[t];
  cuda_check(cudaDeviceSynchronize(), "sync before fetch");
  if (td.vec) {
    // find the producing layer to know the layout
    for (const Layer& L : plan_.layers) {
      if (L.out != t) continue;
      const Shape3 in = cur_->ts[L.in];
      const float* p = reinterpret_cast<const float*>(arena_ + cur_->boff[td.buf]);
      size_t count = 0;
      if (L.kind == LKind::DbHead) { dims[0] = in.n; dims[1] = 1; dims[2] = in.h * 4; dims[3] = in.w * 4; }
      else if (L.kind == LKind::FcSoftmax) { dims[0] = in.n; dims[1] = L.cout; dims[2] = dims[3] = 1; }
      else if (L.kind == LKind::SeFc) { dims[0] = in.n; dims[1] = round_up(L.cin, 8); dims[2] = dims[3] = 1; }
      else if (L.kind == LKind::CtcHead) { dims[0] = in.n; dims[1] = in.w; dims[2] = dims[3] = 1; }
      else return false;
      count = size_t(dims[0]) * dims[1] * dims[2] * dims[3];
      out->resize(count);
      cuda_check(cudaMemcpy(out->data(), p, count * 4, cudaMemcpyDeviceToHost), "fetch");
      return true;
    }
    return false;
  }
  TV v = make_tv(arena_, plan_, cur_->boff, cur_->ts, t);
  const size_t count = size_t(v.n) * v.c * v.h * v.w;
  float* d = nullptr;
  cuda_check(cudaMalloc(&d, count * 4), "fetch scratch");
  launch_nhwc_to_nchw_f32(v, d, 0);
  out->resize(count);
  cuda_check(cudaMemcpy(out->data(), d, count * 4, cudaMemcpyDeviceToHost), "fetch");
  cudaFree(d);
  dims[0] = v.n; dims[1] = v.c; dims[2] = v.h; dims[3] = v.w;
  return true;
}

}  // namespace b200ocr
