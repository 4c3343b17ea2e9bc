#include "plan.h"

#include <cmath>
#include <cstring>
#include <set>
#include <sstream>
#include <stdexcept>

namespace b200ocr {

// ---------------------------------------------------------------- fp16 helpers
uint16_t f32_to_f16_bits(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x007FFFFFu;
  int32_t exp = int32_t((x >> 23) & 0xFF) - 127 + 15;
  if (((x >> 23) & 0xFF) == 0xFF) return uint16_t(sign | 0x7C00u | (mant ? 0x200u : 0));
  if (exp >= 31) return uint16_t(sign | 0x7C00u);
  if (exp <= 0) {
    if (exp < -10) return uint16_t(sign);
    mant |= 0x00800000u;
    int shift = 14 - exp;
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return uint16_t(sign | h);
  }
  uint32_t h = (uint32_t(exp) << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;  // RNE; carry into exponent is correct
  return uint16_t(sign | h);
}

float f16_bits_to_f32(uint16_t h) {
  uint32_t sign = uint32_t(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F, mant = h & 0x3FFu, x;
  if (exp == 0) {
    if (mant == 0) x = sign;
    else {
      int e = -1;
      do { mant <<= 1; ++e; } while (!(mant & 0x400u));
      x = sign | uint32_t(127 - 15 - e) << 23 | ((mant & 0x3FFu) << 13);
    }
  } else if (exp == 31) x = sign | 0x7F800000u | (mant << 13);
  else x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}

int Plan::find_tensor(const std::string& var) const {
  for (size_t i = 0; i < tensors.size(); ++i) {
    if (tensors[i].name == var) return int(i);
    for (auto& a : tensors[i].aliases)
      if (a == var) return int(i);
  }
  return -1;
}

namespace {

const char* kind_name(LKind k) {
  static const char* n[] = {"Conv", "DwConv", "Gap", "SeFc", "Scale", "UpAdd", "UpCat", "Pool",
                            "Add", "LayerNorm", "Attn", "DbHead", "FcSoftmax", "CtcHead"};
  return n[int(k)];
}

struct Builder {
  const PdProgram& prog;
  Plan& plan;
  std::map<std::string, std::vector<int>> consumers;  // var -> op indices reading it
  std::map<std::string, int> var2tensor;
  std::vector<char> done;
  struct PendingUp { int src; int shift; };
  std::map<std::string, PendingUp> pending_up;

  Builder(const PdProgram& p, Plan& pl) : prog(p), plan(pl), done(p.ops.size(), 0) {
    for (size_t i = 0; i < p.ops.size(); ++i)
      for (auto& kv : p.ops[i].inputs)
        for (auto& v : kv.second) consumers[v].push_back(int(i));
  }

  [[noreturn]] void fail(int i, const std::string& why) const {
    std::ostringstream os;
    os << "plan: op #" << i << " (" << prog.ops[i].type << "): " << why;
    throw std::runtime_error(os.str());
  }

  bool is_param(const std::string& v) const { return prog.params.count(v) != 0; }
  const std::vector<float>& param(const std::string& v) const {
    auto it = prog.params.find(v);
    if (it == prog.params.end()) throw std::runtime_error("plan: missing parameter " + v);
    return it->second;
  }
  const std::vector<int64_t>& dims(const std::string& v) const { return prog.vars.at(v).dims; }

  int new_tensor(const std::string& name, int c, bool vec = false, int like = -1) {
    BufferDesc b;
    b.c_total = c;
    b.vec = vec;
    plan.buffers.push_back(b);
    TensorDesc t;
    t.name = name;
    t.c = c;
    t.vec = vec;
    t.buf = int(plan.buffers.size()) - 1;
    plan.tensors.push_back(t);
    int id = int(plan.tensors.size()) - 1;
    plan.buffers.back().like = like < 0 ? id : like;
    var2tensor[name] = id;
    return id;
  }
  void alias(const std::string& var, int t) {
    var2tensor[var] = t;
    plan.tensors[t].aliases.push_back(var);
  }
  int tensor_of(int i, const std::string& var) const {
    auto it = var2tensor.find(var);
    if (it == var2tensor.end()) fail(i, "input " + var + " has no planned producer");
    return it->second;
  }
  // the unique not-yet-planned consumer of `var`, or -1
  int sole_consumer(const std::string& var) const {
    auto it = consumers.find(var);
    if (it == consumers.end() || it->second.size() != 1) return -1;
    return done[it->second[0]] ? -1 : it->second[0];
  }
  static bool shape_only(const std::string& t) {
    return t == "shape" || t == "fill_constant";
  }

  // ---- epilogue folding -----------------------------------------------------
  struct Fold {
    std::vector<float> A, B;  // pre-activation per-channel affine of the raw accumulator
    Act act = Act::None;
    float act_a = 0, act_b = 0, s2 = 1, t2 = 0;
    int residual = -1;
    std::string out_var;
  };

  // Absorb the linear chain hanging off `var` (output of op `head`) into `f`.
  void absorb(const std::string& var0, int C, Fold* f) {
    std::string var = var0;
    f->A.assign(C, 1.f);
    f->B.assign(C, 0.f);
    while (true) {
      f->out_var = var;
      int j = sole_consumer(var);
      if (j < 0 || f->residual >= 0) return;
      const PdOp& op = prog.ops[j];
      const std::string& t = op.type;
      bool post = f->act != Act::None;
      if (t == "elementwise_add" || t == "elementwise_mul") {
        const std::string& x = op.in("X");
        const std::string& y = op.in("Y");
        const std::string& other = (x == var) ? y : x;
        if (x != var && y != var) return;
        if (is_param(other)) {
          const std::vector<float>& p = param(other);
          bool scalar = p.size() == 1;
          if (!scalar && (int(p.size()) != C || post)) return;
          if (!scalar) {
            int64_t axis = op.attr_i("axis", -1);
            // per-channel vector: conv bias uses axis=1 on NCHW; linear bias uses the last axis
            if (!(axis == 1 || axis == -1 || axis == int64_t(dims(var).size()) - 1)) return;
          }
          bool add = t == "elementwise_add";
          if (post) {
            if (add) f->t2 += p[0];
            else { f->s2 *= p[0]; f->t2 *= p[0]; }
          } else {
            for (int c = 0; c < C; ++c) {
              float v = scalar ? p[0] : p[c];
              if (add) f->B[c] += v;
              else { f->A[c] *= v; f->B[c] *= v; }
            }
          }
        } else {
          if (t != "elementwise_add") return;
          auto it = var2tensor.find(other);
          if (it == var2tensor.end()) return;  // residual not produced yet -> standalone Add
          const TensorDesc& rt = plan.tensors[it->second];
          if (rt.vec || rt.c != C) return;
          if (dims(other).size() != dims(var).size()) return;
          f->residual = it->second;
        }
      } else if (t == "batch_norm") {
        if (post) return;
        const auto& sc = param(op.in("Scale"));
        const auto& bi = param(op.in("Bias"));
        const auto& me = param(op.in("Mean"));
        const auto& va = param(op.in("Variance"));
        float eps = op.attr_f("epsilon", 1e-5f);
        for (int c = 0; c < C; ++c) {
          float g = sc[c] / std::sqrt(va[c] + eps);
          f->A[c] *= g;
          f->B[c] = (f->B[c] - me[c]) * g + bi[c];
        }
        done[j] = 1;
        var = op.out("Y");
        continue;
      } else if (t == "hard_swish" || t == "relu" || t == "swish" || t == "hard_sigmoid" ||
                 t == "sigmoid") {
        if (post) return;
        if (t == "hard_swish") {
          if (op.attr_f("offset", 3.f) != 3.f || op.attr_f("scale", 6.f) != 6.f ||
              op.attr_f("threshold", 6.f) != 6.f)
            return;
          f->act = Act::HSwish;
        } else if (t == "relu") f->act = Act::Relu;
        else if (t == "swish") {
          if (op.attr_f("beta", 1.f) != 1.f) return;
          f->act = Act::Swish;
        } else if (t == "hard_sigmoid") {
          f->act = Act::HSigmoid;
          f->act_a = op.attr_f("slope", 0.2f);
          f->act_b = op.attr_f("offset", 0.5f);
        } else f->act = Act::Sigmoid;
      } else if (t == "dropout" || t == "assign") {
        // identity at inference (dropout: is_test + upscale_in_train)
      } else {
        return;
      }
      done[j] = 1;
      var = op.out("Out");
    }
  }

  void set_epilogue(Layer* L, const Fold& f) {
    L->act = f.act;
    L->act_a = f.act_a;
    L->act_b = f.act_b;
    L->post_scale = f.s2;
    L->post_shift = f.t2;
    L->residual = f.residual;
  }

  int64_t push_f(const std::vector<float>& v) {
    int64_t off = int64_t(plan.wf.size());
    plan.wf.insert(plan.wf.end(), v.begin(), v.end());
    while (plan.wf.size() % 4) plan.wf.push_back(0.f);  // keep 16 B alignment of every block
    return off;
  }

  // Pack a dense filter as fp16 [cout_pad][taps][cin_pad], scaled by A[co].
  // `get(co, ci, tap)` returns the raw fp32 weight.
  template <class G>
  void pack_conv(Layer* L, const Fold& f, G get) {
    int taps = L->kh * L->kw;
    L->cin_pad = round_up(L->cin, 64);
    L->cout_pad = round_up(L->cout, 16);
    L->wh_off = int64_t(plan.wh.size());
    plan.wh.resize(plan.wh.size() + size_t(L->cout_pad) * taps * L->cin_pad, 0);
    uint16_t* w = plan.wh.data() + L->wh_off;
    for (int co = 0; co < L->cout; ++co)
      for (int t = 0; t < taps; ++t)
        for (int ci = 0; ci < L->cin; ++ci)
          w[(size_t(co) * taps + t) * L->cin_pad + ci] = f32_to_f16_bits(f.A[co] * get(co, ci, t));
    std::vector<float> b(L->cout_pad, 0.f);
    for (int co = 0; co < L->cout; ++co) b[co] = f.B[co];
    L->bias_off = push_f(b);
  }

  // ---- op handlers ----------------------------------------------------------
  void do_conv(int i) {
    const PdOp& op = prog.ops[i];
    bool dw = op.type == "depthwise_conv2d";
    const std::string& wname = op.in("Filter");
    const auto& wd = dims(wname);
    const auto& w = param(wname);
    auto st = op.attr_ints("strides"), pd = op.attr_ints("paddings"), dl = op.attr_ints("dilations");
    if (pd.size() != 2 || st.size() != 2) fail(i, "only 2-element strides/paddings supported");
    if (dl.size() == 2 && (dl[0] != 1 || dl[1] != 1)) fail(i, "dilation unsupported");
    if (op.attr_s("padding_algorithm", "EXPLICIT") != "EXPLICIT") fail(i, "padding_algorithm");
    int in = tensor_of(i, op.in("Input"));
    Layer L;
    L.in = in;
    L.name = op.out("Output");
    L.kh = int(wd[2]); L.kw = int(wd[3]);
    L.sh = int(st[0]); L.sw = int(st[1]);
    L.ph = int(pd[0]); L.pw = int(pd[1]);
    L.cout = int(wd[0]);
    int groups = int(op.attr_i("groups", 1));
    done[i] = 1;
    Fold f;
    absorb(op.out("Output"), L.cout, &f);
    set_epilogue(&L, f);
    int taps = L.kh * L.kw;
    if (dw || (groups > 1 && groups == L.cout && wd[1] == 1)) {
      L.kind = LKind::DwConv;
      L.cin = L.cout;
      if (plan.tensors[in].c != L.cin) fail(i, "depthwise channel mismatch");
      int cp = round_up(L.cout, 8);
      std::vector<float> blk(size_t(taps + 1) * cp, 0.f);
      for (int c = 0; c < L.cout; ++c) {
        for (int t = 0; t < taps; ++t) blk[size_t(t) * cp + c] = f.A[c] * w[size_t(c) * taps + t];
        blk[size_t(taps) * cp + c] = f.B[c];
      }
      L.wf_off = push_f(blk);
      L.cin_pad = L.cout_pad = cp;
      // fp16 copy [tap][cp] for the mixed-precision (fp16 x fp16 + fp32) kernel; the fp32 block keeps the bias
      while (plan.wh.size() % 8) plan.wh.push_back(0);
      L.wh_off = int64_t(plan.wh.size());
      for (int t = 0; t < taps; ++t)
        for (int c = 0; c < cp; ++c) plan.wh.push_back(f32_to_f16_bits(blk[size_t(t) * cp + c]));
    } else {
      if (groups != 1) fail(i, "grouped conv unsupported");
      L.kind = LKind::Conv;
      L.cin = int(wd[1]);
      if (plan.tensors[in].c != L.cin) fail(i, "conv channel mismatch");
      pack_conv(&L, f, [&](int co, int ci, int t) { return w[(size_t(co) * L.cin + ci) * taps + t]; });
    }
    L.out = new_tensor(op.out("Output"), L.cout);
    if (f.out_var != op.out("Output")) alias(f.out_var, L.out);
    plan.layers.push_back(L);
  }

  void do_linear(int i) {  // matmul_v2(X tensor, Y param [K,N])
    const PdOp& op = prog.ops[i];
    if (op.attr_b("trans_x") || op.attr_b("trans_y")) fail(i, "transposed matmul unsupported");
    const std::string& y = op.in("Y");
    if (!is_param(y)) fail(i, "matmul with two activations outside the attention pattern");
    const auto& wd = dims(y);
    const auto& w = param(y);
    int in = tensor_of(i, op.in("X"));
    int K = int(wd[0]), N = int(wd[1]);
    if (plan.tensors[in].c != K) fail(i, "linear channel mismatch");
    done[i] = 1;
    Fold f;
    absorb(op.out("Out"), N, &f);
    if (plan.tensors[in].vec) {
      // pooled vector -> fc -> softmax : the cls head
      int j = sole_consumer(f.out_var);
      if (j < 0 || prog.ops[j].type != "softmax" || f.act != Act::None || f.residual >= 0)
        fail(i, "vector matmul only supported as fc+softmax head");
      done[j] = 1;
      Layer L;
      L.kind = LKind::FcSoftmax;
      L.name = op.out("Out");
      L.in = in;
      L.cin = K; L.cout = N;
      std::vector<float> blk(size_t(K) * N + N);
      for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) blk[size_t(k) * N + n] = w[size_t(k) * N + n] * f.A[n];
      for (int n = 0; n < N; ++n) blk[size_t(K) * N + n] = f.B[n];
      L.wf_off = push_f(blk);
      L.out = new_tensor(prog.ops[j].out("Out"), N, true);
      plan.layers.push_back(L);
      return;
    }
    Layer L;
    L.in = in;
    L.name = op.out("Out");
    L.cin = K; L.cout = N;
    int j = sole_consumer(f.out_var);
    bool ctc = j >= 0 && prog.ops[j].type == "softmax" && f.act == Act::None && f.residual < 0;
    if (ctc) {
      int64_t ax = prog.ops[j].attr_i("axis", -1);
      if (!(ax == -1 || ax == 2)) fail(j, "softmax axis");
      int jj = sole_consumer(prog.ops[j].out("Out"));
      if (jj < 0 || prog.ops[jj].type != "fetch") fail(j, "softmax head must feed fetch");
      done[j] = 1;
      L.kind = LKind::CtcHead;
    } else {
      L.kind = LKind::Conv;
    }
    set_epilogue(&L, f);
    pack_conv(&L, f, [&](int co, int ci, int) { return w[size_t(ci) * N + co]; });
    if (ctc) {
      // padded classes must never win the argmax nor contribute to the softmax sum
      for (int co = L.cout; co < L.cout_pad; ++co) plan.wf[L.bias_off + co] = -30000.f;
      // the tcgen05 head works in the log2 domain (logit * log2(e) in one FFMA, then a bare ex2): bias * log2(e)
      std::vector<float> b2(size_t(L.cout_pad));
      for (int co = 0; co < L.cout_pad; ++co) b2[co] = plan.wf[L.bias_off + co] * 1.4426950408889634f;
      L.wf_off = push_f(b2);
      L.out = new_tensor(prog.ops[j].out("Out"), N, true);  // (prob, idx) per token, not [T,6625]
    } else {
      L.out = new_tensor(op.out("Out"), N);
      if (f.out_var != op.out("Out")) alias(f.out_var, L.out);
    }
    plan.layers.push_back(L);
  }

  void do_pool(int i) {
    const PdOp& op = prog.ops[i];
    int in = tensor_of(i, op.in("X"));
    auto ks = op.attr_ints("ksize"), st = op.attr_ints("strides"), pd = op.attr_ints("paddings");
    std::string ptype = op.attr_s("pooling_type");
    bool global = op.attr_b("global_pooling") ||
                  (op.attr_b("adaptive") && ks.size() == 2 && ks[0] == 1 && ks[1] == 1);
    done[i] = 1;
    int C = plan.tensors[in].c;
    if (global) {
      if (ptype != "avg") fail(i, "global max pool unsupported");
      Layer G;
      G.kind = LKind::Gap;
      G.name = op.out("Out");
      G.in = in;
      G.cin = G.cout = C;
      G.out = new_tensor(op.out("Out"), C, true);
      plan.layers.push_back(G);
      try_se(i, in, G.out);
      return;
    }
    if (op.attr_b("adaptive")) fail(i, "adaptive pool other than 1x1 unsupported");
    if (pd.size() != 2 || pd[0] != 0 || pd[1] != 0) fail(i, "padded pool unsupported");
    if (op.attr_b("ceil_mode")) fail(i, "ceil_mode unsupported");
    Layer L;
    L.kind = LKind::Pool;
    L.name = op.out("Out");
    L.in = in;
    L.cin = L.cout = C;
    L.kh = int(ks[0]); L.kw = int(ks[1]);
    L.sh = int(st[0]); L.sw = int(st[1]);
    L.pool_max = ptype == "max";
    if (!L.pool_max && !op.attr_b("exclusive", true)) fail(i, "inclusive avg pool unsupported");
    L.out = new_tensor(op.out("Out"), C);
    plan.layers.push_back(L);
  }

  // pool -> conv1x1+b -> relu -> conv1x1+b -> hard_sigmoid -> mul(x, gate) [-> add(x, .)]
  void try_se(int i, int x, int pooled) {
    const std::string& pv = prog.ops[i].out("Out");
    int j1 = sole_consumer(pv);
    if (j1 < 0 || prog.ops[j1].type != "conv2d") return;  // plain GAP (cls tail)
    auto conv1x1 = [&](int j, int* co, int* ci) {
      const auto& wd = dims(prog.ops[j].in("Filter"));
      if (wd[2] != 1 || wd[3] != 1) fail(j, "SE conv must be 1x1");
      *co = int(wd[0]); *ci = int(wd[1]);
    };
    int cm, c0, c1, cm1;
    conv1x1(j1, &cm, &c0);
    done[j1] = 1;
    Fold f1;
    absorb(prog.ops[j1].out("Output"), cm, &f1);
    if (f1.act != Act::Relu || f1.s2 != 1.f || f1.t2 != 0.f || f1.residual >= 0) fail(j1, "SE fc1 pattern");
    int j2 = sole_consumer(f1.out_var);
    if (j2 < 0 || prog.ops[j2].type != "conv2d") fail(j1, "SE fc2 missing");
    conv1x1(j2, &c1, &cm1);
    done[j2] = 1;
    Fold f2;
    absorb(prog.ops[j2].out("Output"), c1, &f2);
    if (f2.act != Act::HSigmoid || f2.s2 != 1.f || f2.t2 != 0.f || f2.residual >= 0) fail(j2, "SE fc2 pattern");
    int C = plan.tensors[x].c;
    if (c0 != C || c1 != C || cm1 != cm) fail(j2, "SE channel mismatch");
    const auto& w1 = param(prog.ops[j1].in("Filter"));
    const auto& w2 = param(prog.ops[j2].in("Filter"));
    Layer S;
    S.kind = LKind::SeFc;
    S.name = f2.out_var;
    S.in = pooled;
    S.cin = S.cout = C;
    S.cmid = cm;
    S.act_a = f2.act_a; S.act_b = f2.act_b;
    std::vector<float> blk;
    blk.reserve(size_t(2) * C * cm + C + cm);
    // both matrices are stored with the OUTPUT index fastest (w1t[c][m], w2t[m][c]): in se_fc_kernel one thread owns
    // one output and neighbouring threads read neighbouring words
    for (int c = 0; c < C; ++c)
      for (int m = 0; m < cm; ++m) blk.push_back(w1[size_t(m) * C + c] * f1.A[m]);
    for (int m = 0; m < cm; ++m) blk.push_back(f1.B[m]);
    for (int m = 0; m < cm; ++m)
      for (int c = 0; c < C; ++c) blk.push_back(w2[size_t(c) * cm + m] * f2.A[c]);
    for (int c = 0; c < C; ++c) blk.push_back(f2.B[c]);
    S.wf_off = push_f(blk);
    S.out = new_tensor(f2.out_var, C, true);
    plan.layers.push_back(S);
    // x * gate
    int j3 = sole_consumer(f2.out_var);
    if (j3 < 0 || prog.ops[j3].type != "elementwise_mul") fail(j2, "SE scale missing");
    const PdOp& mul = prog.ops[j3];
    const std::string& xv = (mul.in("X") == f2.out_var) ? mul.in("Y") : mul.in("X");
    if (tensor_of(j3, xv) != x) fail(j3, "SE scale operand is not the pooled tensor");
    done[j3] = 1;
    Layer M;
    M.kind = LKind::Scale;
    M.name = mul.out("Out");
    M.in = x;
    M.in2 = S.out;
    M.cin = M.cout = C;
    std::string outv = mul.out("Out");
    int j4 = sole_consumer(outv);
    if (j4 >= 0 && prog.ops[j4].type == "elementwise_add") {
      const PdOp& add = prog.ops[j4];
      const std::string& o = (add.in("X") == outv) ? add.in("Y") : add.in("X");
      auto it = var2tensor.find(o);
      if (it != var2tensor.end() && it->second == x) {
        M.scale_residual = true;
        done[j4] = 1;
        outv = add.out("Out");
      }
    }
    M.out = new_tensor(mul.out("Out"), C);
    if (outv != mul.out("Out")) alias(outv, M.out);
    plan.layers.push_back(M);
  }

  void do_interp(int i) {
    const PdOp& op = prog.ops[i];
    if (op.attr_s("interp_method") != "nearest" || op.attr_b("align_corners")) fail(i, "interp mode");
    auto sc = op.attr_floats("scale");
    if (sc.size() != 2 || sc[0] != sc[1]) fail(i, "interp scale");
    int shift = sc[0] == 2.f ? 1 : sc[0] == 4.f ? 2 : sc[0] == 8.f ? 3 : -1;
    if (shift < 0) fail(i, "interp scale must be 2/4/8");
    int src = tensor_of(i, op.in("X"));
    done[i] = 1;
    const std::string& ov = op.out("Out");
    int j = sole_consumer(ov);
    if (j >= 0 && prog.ops[j].type == "elementwise_add" && shift == 1) {
      const PdOp& add = prog.ops[j];
      const std::string& o = (add.in("X") == ov) ? add.in("Y") : add.in("X");
      int a = tensor_of(j, o);
      done[j] = 1;
      Layer L;
      L.kind = LKind::UpAdd;
      L.name = add.out("Out");
      L.in = a; L.in2 = src;
      L.cin = L.cout = plan.tensors[a].c;
      L.out = new_tensor(add.out("Out"), L.cout);
      plan.layers.push_back(L);
      return;
    }
    if (j >= 0 && prog.ops[j].type == "concat") {
      pending_up[ov] = PendingUp{src, shift};
      return;
    }
    fail(i, "nearest_interp must feed elementwise_add (x2) or concat");
  }

  void do_concat(int i) {
    const PdOp& op = prog.ops[i];
    if (op.attr_i("axis", 1) != 1) fail(i, "concat axis must be 1 (channels)");
    const auto& xs = op.inputs.at("X");
    done[i] = 1;
    bool any_up = false;
    for (auto& v : xs) any_up |= pending_up.count(v) != 0;
    if (any_up) {
      if (xs.size() > 4) fail(i, "concat of more than 4 inputs");
      Layer L;
      L.kind = LKind::UpCat;
      L.name = op.out("Out");
      int ctot = 0, like = -1;
      L.kh = 0;  // kh..: per-input shifts packed below
      int shifts[4] = {0, 0, 0, 0};
      for (size_t k = 0; k < xs.size(); ++k) {
        auto it = pending_up.find(xs[k]);
        if (it != pending_up.end()) { L.ins[k] = it->second.src; shifts[k] = it->second.shift; }
        else { L.ins[k] = tensor_of(i, xs[k]); like = L.ins[k]; }
        ctot += plan.tensors[L.ins[k]].c;
      }
      if (like < 0) fail(i, "concat needs one full-resolution input");
      L.kh = shifts[0]; L.kw = shifts[1]; L.sh = shifts[2]; L.sw = shifts[3];
      L.in = like;
      L.cin = L.cout = ctot;
      L.out = new_tensor(op.out("Out"), ctot, false);
      plan.buffers[plan.tensors[L.out].buf].like = like;
      plan.layers.push_back(L);
      return;
    }
    // zero-copy concat: re-home the producers into one buffer
    int ctot = 0;
    std::vector<int> ids;
    for (auto& v : xs) {
      int t = tensor_of(i, v);
      const TensorDesc& td = plan.tensors[t];
      if (td.vec || td.c_off != 0 || plan.buffers[td.buf].c_total != td.c || td.c % 8 != 0)
        fail(i, "concat input cannot be re-homed");
      ids.push_back(t);
      ctot += td.c;
    }
    BufferDesc b;
    b.c_total = ctot;
    b.like = ids[0];
    plan.buffers.push_back(b);
    int nb = int(plan.buffers.size()) - 1, off = 0;
    for (int t : ids) {
      plan.buffers[plan.tensors[t].buf].c_total = 0;  // orphaned
      plan.tensors[t].buf = nb;
      plan.tensors[t].c_off = off;
      off += plan.tensors[t].c;
    }
    TensorDesc t;
    t.name = op.out("Out");
    t.c = ctot;
    t.buf = nb;
    plan.tensors.push_back(t);
    var2tensor[t.name] = int(plan.tensors.size()) - 1;
  }

  void do_layernorm(int i) {
    const PdOp& op = prog.ops[i];
    int in = tensor_of(i, op.in("X"));
    int C = plan.tensors[in].c;
    if (op.attr_i("begin_norm_axis", 1) != int64_t(dims(op.in("X")).size()) - 1)
      fail(i, "layer_norm must normalise the last axis");
    const auto& g = param(op.in("Scale"));
    const auto& b = param(op.in("Bias"));
    if (int(g.size()) != C) fail(i, "layer_norm width");
    done[i] = 1;
    Layer L;
    L.kind = LKind::LayerNorm;
    L.name = op.out("Y");
    L.in = in;
    L.cin = L.cout = C;
    L.eps = op.attr_f("epsilon", 1e-5f);
    std::vector<float> blk(g);
    blk.insert(blk.end(), b.begin(), b.end());
    L.wf_off = push_f(blk);
    L.out = new_tensor(op.out("Y"), C);
    plan.layers.push_back(L);
  }

  // reshape2[0,-1,3,H,D] transpose2[2,0,3,1,4] slice x3 scale transpose2[0,1,3,2] matmul softmax
  // dropout matmul transpose2[0,2,1,3] reshape2[0,-1,H*D]   (reference graph: rec ops 236-250)
  bool try_attention(int i) {
    const PdOp& op = prog.ops[i];
    auto shp = op.attr_ints("shape");
    if (shp.size() != 5 || shp[0] != 0 || shp[1] != -1 || shp[2] != 3) return false;
    int heads = int(shp[3]), hd = int(shp[4]);
    int in = tensor_of(i, op.in("X"));
    if (plan.tensors[in].c != 3 * heads * hd) fail(i, "attention qkv width");
    static const char* seq[] = {"transpose2", "slice", "scale", "slice", "slice", "transpose2",
                                "matmul_v2", "softmax", "dropout", "matmul_v2", "transpose2", "reshape2"};
    int j = i;
    float scale = 1.f;
    std::string last_out;
    done[i] = 1;
    for (const char* want : seq) {
      do { ++j; } while (j < int(prog.ops.size()) && (done[j] || shape_only(prog.ops[j].type)));
      if (j >= int(prog.ops.size()) || prog.ops[j].type != want)
        fail(i, std::string("attention pattern broken at expected ") + want);
      const PdOp& o = prog.ops[j];
      if (o.type == "scale") {
        if (o.attr_f("bias", 0.f) != 0.f) fail(j, "attention scale bias");
        scale = o.attr_f("scale", 1.f);
      }
      if (o.type == "transpose2" && j == i + 1) {
        auto ax = o.attr_ints("axis");
        if (ax != std::vector<int64_t>{2, 0, 3, 1, 4}) fail(j, "attention qkv transpose");
      }
      done[j] = 1;
      last_out = o.out("Out");
    }
    Layer L;
    L.kind = LKind::Attn;
    L.name = last_out;
    L.in = in;
    L.cin = 3 * heads * hd;
    L.cout = heads * hd;
    L.heads = heads; L.head_dim = hd;
    L.attn_scale = scale;
    L.out = new_tensor(last_out, L.cout);
    plan.layers.push_back(L);
    return true;
  }

  // conv2d_transpose k2 s2 + b + BN + relu -> conv2d_transpose k2 s2 (->1) + b -> sigmoid
  void do_dbhead(int i) {
    const PdOp& op = prog.ops[i];
    auto chk = [&](int j) {
      const PdOp& o = prog.ops[j];
      auto st = o.attr_ints("strides"), pd = o.attr_ints("paddings");
      const auto& wd = dims(o.in("Filter"));
      if (wd[2] != 2 || wd[3] != 2 || st[0] != 2 || st[1] != 2 || pd[0] != 0 || pd[1] != 0 ||
          o.attr_i("groups", 1) != 1)
        fail(j, "conv2d_transpose must be k2 s2 p0");
    };
    chk(i);
    int in = tensor_of(i, op.in("Input"));
    const auto& w1d = dims(op.in("Filter"));  // [cin, cmid, 2, 2]
    int cin = int(w1d[0]), cmid = int(w1d[1]);
    if (plan.tensors[in].c != cin) fail(i, "deconv channel mismatch");
    done[i] = 1;
    Fold f1;
    absorb(op.out("Output"), cmid, &f1);
    if (f1.act != Act::Relu || f1.s2 != 1.f || f1.t2 != 0.f || f1.residual >= 0) fail(i, "DB head stage 1");
    int j = sole_consumer(f1.out_var);
    if (j < 0 || prog.ops[j].type != "conv2d_transpose") fail(i, "DB head stage 2 missing");
    chk(j);
    const auto& w2d = dims(prog.ops[j].in("Filter"));
    if (w2d[0] != cmid || w2d[1] != 1) fail(j, "DB head stage 2 shape");
    done[j] = 1;
    Fold f2;
    absorb(prog.ops[j].out("Output"), 1, &f2);
    if (f2.act != Act::Sigmoid || f2.s2 != 1.f || f2.t2 != 0.f) fail(j, "DB head sigmoid");
    const auto& w1 = param(op.in("Filter"));
    const auto& w2 = param(prog.ops[j].in("Filter"));
    Layer L;
    L.kind = LKind::DbHead;
    L.name = f2.out_var;
    L.in = in;
    L.cin = cin; L.cmid = cmid; L.cout = 1;
    // block: w1[q][ci][cm] (q = dy*2+dx), b1[cm], w2[cm][4], b2
    std::vector<float> blk;
    for (int q = 0; q < 4; ++q)
      for (int ci = 0; ci < cin; ++ci)
        for (int cm = 0; cm < cmid; ++cm) blk.push_back(w1[(size_t(ci) * cmid + cm) * 4 + q] * f1.A[cm]);
    for (int cm = 0; cm < cmid; ++cm) blk.push_back(f1.B[cm]);
    for (int cm = 0; cm < cmid; ++cm)
      for (int q = 0; q < 4; ++q) blk.push_back(w2[size_t(cm) * 4 + q] * f2.A[0]);
    blk.push_back(f2.B[0]);
    L.wf_off = push_f(blk);
    L.out = new_tensor(f2.out_var, 1, true);  // fp32 [N, 4h, 4w] probability map
    plan.layers.push_back(L);
  }

  void run() {
    const int n = int(prog.ops.size());
    for (int i = 0; i < n; ++i) {
      if (done[i]) continue;
      const PdOp& op = prog.ops[i];
      const std::string& t = op.type;
      if (t == "feed") {
        const std::string& v = op.out("Out");
        const auto& d = dims(v);
        if (d.size() != 4 || d[1] != 3) fail(i, "feed must be [N,3,H,W]");
        plan.input = new_tensor(v, 3);
        done[i] = 1;
      } else if (t == "fetch") {
        plan.output = tensor_of(i, op.in("X"));
        done[i] = 1;
      } else if (t == "conv2d" || t == "depthwise_conv2d") do_conv(i);
      else if (t == "conv2d_transpose") do_dbhead(i);
      else if (t == "pool2d") do_pool(i);
      else if (t == "nearest_interp_v2") do_interp(i);
      else if (t == "concat") do_concat(i);
      else if (t == "layer_norm") do_layernorm(i);
      else if (t == "matmul_v2") do_linear(i);
      else if (shape_only(t)) done[i] = 1;
      else if (t == "slice") {
        // only slices of shape tensors survive to here (attention slices are consumed by the pattern)
        if (var2tensor.count(op.in("Input"))) fail(i, "slice of an activation outside attention");
        done[i] = 1;
      } else if (t == "reshape2") {
        if (try_attention(i)) continue;
        // pure views in NHWC: [N,T,C] <-> [N,1,T,C] and [N,C,1,1] -> [N,C]
        alias(op.out("Out"), tensor_of(i, op.in("X")));
        done[i] = 1;
      } else if (t == "transpose2" || t == "flatten_contiguous_range" || t == "squeeze2" ||
                 t == "assign" || t == "dropout") {
        // Layout no-ops for NHWC storage when H == 1 (checked at instantiate time):
        // NCHW [N,C,1,W] --flatten/squeeze--> [N,C,W] --transpose [0,2,1]--> [N,W,C]
        if (t == "transpose2") {
          auto ax = op.attr_ints("axis");
          bool ok = ax == std::vector<int64_t>{0, 2, 1} || ax == std::vector<int64_t>{0, 3, 1, 2};
          if (!ok) fail(i, "transpose outside a known view pattern");
        }
        alias(op.out("Out"), tensor_of(i, op.in("X")));
        done[i] = 1;
      } else if (t == "elementwise_add") {
        const std::string& x = op.in("X");
        const std::string& y = op.in("Y");
        if (is_param(x) || is_param(y)) fail(i, "stray parameter add");
        Layer L;
        L.kind = LKind::Add;
        L.name = op.out("Out");
        L.in = tensor_of(i, x);
        L.in2 = tensor_of(i, y);
        L.cin = L.cout = plan.tensors[L.in].c;
        if (plan.tensors[L.in2].c != L.cin) fail(i, "add channel mismatch");
        L.out = new_tensor(op.out("Out"), L.cout);
        plan.layers.push_back(L);
        done[i] = 1;
      } else {
        fail(i, "no fused-layer pattern covers this op");
      }
    }
    if (plan.input < 0 || plan.output < 0) throw std::runtime_error("plan: graph has no feed/fetch");
    // shapes the sm_100a kernels do not implement are rejected here, so that *_create fails with a message instead
    // of the first forward pass
    for (const Layer& L : plan.layers) {
      auto bad = [&](const std::string& why) {
        throw std::runtime_error("plan: layer " + L.name + ": " + why + " (not implemented by the sm_100a kernels)");
      };
      if (L.kind == LKind::DwConv && !((L.kh == 3 && L.kw == 3) || (L.kh == 5 && L.kw == 5)))
        bad("depthwise filter " + std::to_string(L.kh) + "x" + std::to_string(L.kw) + ", only 3x3 and 5x5");
      if (L.kind == LKind::DbHead && (L.cin != 24 || L.cmid != 24))
        bad("DB head " + std::to_string(L.cin) + " -> " + std::to_string(L.cmid) + ", only 24 -> 24");
      if (L.kind == LKind::FcSoftmax && L.cout > 8) bad("classifier head with more than 8 classes");
    }
    const Layer& last = plan.layers.back();
    plan.kind = last.kind == LKind::DbHead ? "det" : last.kind == LKind::FcSoftmax ? "cls"
              : last.kind == LKind::CtcHead ? "rec" : "generic";
  }
};

}  // namespace

void build_plan(const PdProgram& prog, Plan* plan) {
  Builder b(prog, *plan);
  b.run();
}

std::string Plan::dump() const {
  std::ostringstream os;
  os << "kind " << kind << " tensors " << tensors.size() << " layers " << layers.size() << " wh "
     << wh.size() << " wf " << wf.size() << "\n";
  for (size_t i = 0; i < layers.size(); ++i) {
    const Layer& L = layers[i];
    // third column: the Paddle variable whose value the output tensor holds (last op of the fused chain)
    const TensorDesc& ot = tensors[L.out];
    os << i << " " << kind_name(L.kind) << " " << (ot.aliases.empty() ? ot.name : ot.aliases.back()) << " in=" << L.in;
    if (L.in2 >= 0) os << " in2=" << L.in2;
    if (L.kind == LKind::UpCat) os << " ins=" << L.ins[0] << "," << L.ins[1] << "," << L.ins[2] << "," << L.ins[3];
    os << " out=" << L.out << " c=" << L.cin << "->" << L.cout;
    if (L.kind == LKind::Conv || L.kind == LKind::DwConv || L.kind == LKind::Pool)
      os << " k=" << L.kh << "x" << L.kw << " s=" << L.sh << "x" << L.sw << " p=" << L.ph << "x" << L.pw;
    os << " act=" << int(L.act) << " post=" << L.post_scale << "," << L.post_shift;
    if (L.residual >= 0) os << " res=" << L.residual;
    if (L.scale_residual) os << " +x";
    os << "\n";
  }
  return os.str();
}

}  // namespace b200ocr
