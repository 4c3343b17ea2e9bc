// Fused-layer plan built from a Paddle ProgramDesc.
//
// The reference hands the op list to Paddle Inference and lets its IR passes
// fuse (config.SwitchIrOptim(true), reference src/ocr_det.cpp:84).  Here the
// op list is pattern-matched once at load into a short list of fused layers
// (conv + bias/BN/scalar-affine + activation + affine [+ residual], SE block,
// FPN glue, DB head, SVTR attention, CTC head) whose weights are folded and
// repacked for the sm_100a kernels.  Activations are NHWC fp16 views
// (channel pitch a multiple of 8) into reusable buffers.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pd_model.h"

namespace b200ocr {

enum class Act : int { None = 0, Relu = 1, HSwish = 2, Swish = 3, HSigmoid = 4, Sigmoid = 5 };

enum class LKind : int {
  Conv = 0,     // dense conv (1x1 / kxk, any stride) as implicit GEMM, fused epilogue
  DwConv,       // depthwise conv, fused epilogue
  Gap,          // global average pool -> fp32 partial sums
  SeFc,         // SE gate: fc -> relu -> fc -> hard-sigmoid on pooled vector
  Scale,        // x * gate[n,c] (+ x)
  UpAdd,        // a + nearest_up2(b)
  UpCat,        // concat(up8(a), up4(b), up2(c), d)
  Pool,         // avg / max pool, clipped exclusive windows
  Add,          // a + b
  LayerNorm,
  Attn,         // multi-head self attention on packed qkv [N,T,3*heads*d]
  DbHead,       // deconv2x2 + BN + relu + deconv2x2 + sigmoid (+ threshold bitmap)
  FcSoftmax,    // pooled vector -> fc -> softmax (cls head)
  CtcHead,      // fc + softmax-max + argmax (rec head)
};

struct TensorDesc {
  std::string name;   // Paddle variable name (first producer; aliases listed in `aliases`)
  int c = 0;          // logical channels
  int buf = -1;       // buffer id
  int c_off = 0;      // channel offset inside the buffer (concat slices)
  bool vec = false;   // fp32 [N, C] vector (pooled / gate tensors), not NHWC fp16
  // shape relation to the producing layer is resolved at instantiate time
  std::vector<std::string> aliases;
};

struct BufferDesc {
  int c_total = 0;  // logical channels of the whole buffer; pitch = round_up(c_total, 8)
  bool vec = false;
  int like = -1;    // tensor id whose (n,h,w) this buffer takes
};

struct Layer {
  LKind kind = LKind::Conv;
  std::string name;          // name of the head Paddle op's output var
  int in = -1, in2 = -1, out = -1;
  int ins[4] = {-1, -1, -1, -1};
  int residual = -1;
  // conv-like geometry
  int kh = 1, kw = 1, sh = 1, sw = 1, ph = 0, pw = 0;
  int cin = 0, cout = 0, cmid = 0;
  // epilogue: y = post_scale * act(acc + bias[c]) + post_shift (+ residual)
  Act act = Act::None;
  float act_a = 0.f, act_b = 0.f;  // hard-sigmoid slope / offset
  float post_scale = 1.f, post_shift = 0.f;
  // pooling
  bool pool_max = false;
  // attention
  int heads = 0, head_dim = 0;
  float attn_scale = 1.f;
  float eps = 1e-5f;
  bool scale_residual = false;  // Scale: out = x + x*gate
  // weights: offsets (in elements) into Plan::wh (fp16) and Plan::wf (fp32)
  int64_t wh_off = -1;  // Conv / CtcHead: [cout_pad16][kh*kw][cin_pad64] fp16
  int64_t wf_off = -1;  // kind-specific fp32 block (see plan.cpp)
  int64_t bias_off = -1;  // fp32 [cout_pad16] folded bias
  int cin_pad = 0, cout_pad = 0;
};

struct Plan {
  std::vector<TensorDesc> tensors;
  std::vector<BufferDesc> buffers;
  std::vector<Layer> layers;
  std::vector<uint16_t> wh;  // fp16 bit patterns
  std::vector<float> wf;
  int input = -1;   // tensor id of the feed (3 channels)
  int output = -1;  // tensor id of the fetch
  std::string kind; // "det" | "cls" | "rec" (from the output pattern)

  int find_tensor(const std::string& var) const;
  std::string dump() const;  // human-readable, used by tests
};

// Throws std::runtime_error naming the first op that does not fit a pattern.
void build_plan(const PdProgram& prog, Plan* plan);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
uint16_t f32_to_f16_bits(float f);
float f16_bits_to_f32(uint16_t h);

}  // namespace b200ocr
