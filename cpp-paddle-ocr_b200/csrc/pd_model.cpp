#include "pd_model.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <sys/stat.h>

namespace b200ocr {

namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t r = 0;
    int shift = 0;
    while (true) {
      if (p >= end) throw std::runtime_error("pdmodel: truncated varint");
      uint8_t b = *p++;
      r |= uint64_t(b & 0x7F) << shift;
      if (!(b & 0x80)) return r;
      shift += 7;
      if (shift > 63) throw std::runtime_error("pdmodel: varint overflow");
    }
  }
  Cursor sub(uint64_t len) {
    if (uint64_t(end - p) < len) throw std::runtime_error("pdmodel: truncated field");
    Cursor c{p, p + len};
    p += len;
    return c;
  }
};

struct Field {
  int no;
  int wt;
  uint64_t v;   // varint value / raw bits for fixed
  Cursor body;  // for wire type 2
};

bool next_field(Cursor& c, Field* f) {
  if (c.done()) return false;
  uint64_t key = c.varint();
  f->no = int(key >> 3);
  f->wt = int(key & 7);
  f->body = Cursor{nullptr, nullptr};
  switch (f->wt) {
    case 0: f->v = c.varint(); break;
    case 1: {
      Cursor b = c.sub(8);
      memcpy(&f->v, b.p, 8);
      break;
    }
    case 2: {
      uint64_t len = c.varint();
      f->body = c.sub(len);
      break;
    }
    case 5: {
      Cursor b = c.sub(4);
      uint32_t t;
      memcpy(&t, b.p, 4);
      f->v = t;
      break;
    }
    default: throw std::runtime_error("pdmodel: unsupported wire type");
  }
  return true;
}

std::string str_of(const Cursor& c) { return std::string((const char*)c.p, c.end - c.p); }

void packed_ints(const Field& f, std::vector<int64_t>* out) {
  if (f.wt == 0) {
    out->push_back(int64_t(f.v));
    return;
  }
  Cursor c = f.body;
  while (!c.done()) out->push_back(int64_t(c.varint()));
}

float f32_of(uint64_t bits) {
  uint32_t b = uint32_t(bits);
  float f;
  memcpy(&f, &b, 4);
  return f;
}

void parse_attr(Cursor c, std::string* name, PdAttr* a) {
  Field f;
  while (next_field(c, &f)) {
    switch (f.no) {
      case 1: *name = str_of(f.body); break;
      case 2: a->type = int(f.v); break;
      case 3: a->i = int64_t(int32_t(uint32_t(f.v))); break;
      case 4: a->f = f32_of(f.v); break;
      case 5: a->s = str_of(f.body); break;
      case 6: packed_ints(f, &a->ints); break;
      case 7:
        if (f.wt == 2) {
          for (const uint8_t* p = f.body.p; p + 4 <= f.body.end; p += 4) {
            float x;
            memcpy(&x, p, 4);
            a->floats.push_back(x);
          }
        } else {
          a->floats.push_back(f32_of(f.v));
        }
        break;
      case 8: a->strings.push_back(str_of(f.body)); break;
      case 10: a->b = f.v != 0; break;
      case 13: a->i = int64_t(f.v); break;
      case 15: packed_ints(f, &a->ints); break;
      case 19: memcpy(&a->d, &f.v, 8); break;
      default: break;
    }
  }
  if (a->type == 3)  // INTS are int32 on the wire (sign-extended varints)
    for (auto& x : a->ints) x = int64_t(int32_t(uint32_t(uint64_t(x))));
}

void parse_opvar(Cursor c, std::string* slot, std::vector<std::string>* args) {
  Field f;
  while (next_field(c, &f)) {
    if (f.no == 1) *slot = str_of(f.body);
    else if (f.no == 2) args->push_back(str_of(f.body));
  }
}

void parse_op(Cursor c, PdOp* op) {
  Field f;
  while (next_field(c, &f)) {
    if (f.no == 3) {
      op->type = str_of(f.body);
    } else if (f.no == 1 || f.no == 2) {
      std::string slot;
      std::vector<std::string> args;
      parse_opvar(f.body, &slot, &args);
      (f.no == 1 ? op->inputs : op->outputs)[slot] = std::move(args);
    } else if (f.no == 4) {
      std::string name;
      PdAttr a;
      parse_attr(f.body, &name, &a);
      if (name != "op_callstack" && name != "op_namescope") op->attrs[name] = std::move(a);
    }
  }
}

void parse_tensor_desc(Cursor c, int* dtype, std::vector<int64_t>* dims) {
  Field f;
  while (next_field(c, &f)) {
    if (f.no == 1) *dtype = int(f.v);
    else if (f.no == 2) packed_ints(f, dims);
  }
}

void parse_var(Cursor c, PdVar* v) {
  Field f;
  while (next_field(c, &f)) {
    if (f.no == 1) v->name = str_of(f.body);
    else if (f.no == 3) v->persistable = f.v != 0;
    else if (f.no == 2) {  // VarType
      Cursor vt = f.body;
      Field g;
      while (next_field(vt, &g)) {
        if (g.no == 1) v->vtype = int(g.v);
        else if (g.no == 3) {  // LoDTensorDesc
          Cursor lt = g.body;
          Field h;
          while (next_field(lt, &h))
            if (h.no == 1) parse_tensor_desc(h.body, &v->dtype, &v->dims);
        }
      }
    }
  }
}

std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + path);
  in.seekg(0, std::ios::end);
  std::streamoff n = in.tellg();
  in.seekg(0);
  std::vector<uint8_t> buf((size_t)n);
  if (n > 0) in.read((char*)buf.data(), n);
  if (!in) throw std::runtime_error("short read on " + path);
  return buf;
}

const std::string kEmpty;

}  // namespace

const std::string& PdOp::in(const std::string& slot, size_t k) const {
  auto it = inputs.find(slot);
  if (it == inputs.end() || it->second.size() <= k) return kEmpty;
  return it->second[k];
}
const std::string& PdOp::out(const std::string& slot, size_t k) const {
  auto it = outputs.find(slot);
  if (it == outputs.end() || it->second.size() <= k) return kEmpty;
  return it->second[k];
}
bool PdOp::has_in(const std::string& slot) const {
  auto it = inputs.find(slot);
  return it != inputs.end() && !it->second.empty();
}
int64_t PdOp::attr_i(const std::string& n, int64_t dflt) const {
  auto it = attrs.find(n);
  return it == attrs.end() ? dflt : it->second.i;
}
float PdOp::attr_f(const std::string& n, float dflt) const {
  auto it = attrs.find(n);
  if (it == attrs.end()) return dflt;
  return it->second.type == 15 ? float(it->second.d) : it->second.f;
}
bool PdOp::attr_b(const std::string& n, bool dflt) const {
  auto it = attrs.find(n);
  return it == attrs.end() ? dflt : it->second.b;
}
std::string PdOp::attr_s(const std::string& n, const std::string& dflt) const {
  auto it = attrs.find(n);
  return it == attrs.end() ? dflt : it->second.s;
}
std::vector<int64_t> PdOp::attr_ints(const std::string& n) const {
  auto it = attrs.find(n);
  return it == attrs.end() ? std::vector<int64_t>{} : it->second.ints;
}
std::vector<float> PdOp::attr_floats(const std::string& n) const {
  auto it = attrs.find(n);
  return it == attrs.end() ? std::vector<float>{} : it->second.floats;
}

std::vector<std::string> PdProgram::param_names() const {
  std::vector<std::string> names;
  for (auto& kv : vars)
    if (kv.second.persistable && kv.second.vtype == 7 && kv.first != "feed" && kv.first != "fetch")
      names.push_back(kv.first);
  std::sort(names.begin(), names.end());  // std::map is already ordered; keep explicit
  return names;
}

void load_program(const std::string& path, PdProgram* prog) {
  std::vector<uint8_t> buf = read_file(path);
  Cursor c{buf.data(), buf.data() + buf.size()};
  Field f;
  int nblocks = 0;
  try {
    while (next_field(c, &f)) {
      if (f.no != 1) continue;  // BlockDesc
      if (++nblocks > 1) throw std::runtime_error("multi-block programs are not supported");
      Cursor b = f.body;
      Field g;
      while (next_field(b, &g)) {
        if (g.no == 3) {
          PdVar v;
          parse_var(g.body, &v);
          prog->vars[v.name] = std::move(v);
        } else if (g.no == 4) {
          PdOp op;
          parse_op(g.body, &op);
          prog->ops.push_back(std::move(op));
        }
      }
    }
  } catch (const std::exception& e) {
    throw std::runtime_error(path + ": " + e.what());
  }
  if (prog->ops.empty()) throw std::runtime_error(path + ": no ops decoded (not a ProgramDesc?)");
}

void load_params(const std::string& path, PdProgram* prog) {
  std::vector<uint8_t> buf = read_file(path);
  size_t pos = 0;
  auto need = [&](size_t n) {
    if (n > buf.size() - pos) throw std::runtime_error(path + ": truncated parameter stream");  // no wrap-around
  };
  for (const std::string& name : prog->param_names()) {
    // u32 version | u64 lod_level (+ levels) | u32 tensor version | i32 desc_len | TensorDesc | data
    need(4 + 8);
    pos += 4;
    uint64_t lod;
    memcpy(&lod, &buf[pos], 8);
    pos += 8;
    for (uint64_t l = 0; l < lod; ++l) {
      need(8);
      uint64_t sz;
      memcpy(&sz, &buf[pos], 8);
      pos += 8;
      need(sz);
      pos += sz;
    }
    need(4 + 4);
    pos += 4;
    int32_t dl;
    memcpy(&dl, &buf[pos], 4);
    pos += 4;
    if (dl < 0) throw std::runtime_error(path + ": bad TensorDesc length");
    need((size_t)dl);
    int dtype = -1;
    std::vector<int64_t> dims;
    parse_tensor_desc(Cursor{&buf[pos], &buf[pos] + dl}, &dtype, &dims);
    pos += dl;
    const PdVar& v = prog->vars.at(name);
    if (dtype != 5) throw std::runtime_error(path + ": parameter " + name + " is not fp32");
    if (dims != v.dims)
      throw std::runtime_error(path + ": parameter " + name + " dims do not match the graph");
    size_t n = 1;
    for (int64_t d : dims) {
      if (d <= 0) throw std::runtime_error(path + ": parameter " + name + " has a non-positive dimension");
      if ((uint64_t)d > buf.size() / 4 / n) throw std::runtime_error(path + ": truncated parameter stream");
      n *= (size_t)d;
    }
    need(n * 4);
    std::vector<float> data(n);
    memcpy(data.data(), &buf[pos], n * 4);
    pos += n * 4;
    prog->params[name] = std::move(data);
  }
  if (pos != buf.size())
    throw std::runtime_error(path + ": trailing bytes after the last parameter");
}

bool find_model_files(const std::string& dir, std::string* model, std::string* params) {
  static const char* variants[][2] = {{"/inference.pdmodel", "/inference.pdiparams"},
                                      {"/model.pdmodel", "/model.pdiparams"}};
  for (auto& v : variants) {
    struct stat st;
    std::string m = dir + v[0];
    if (stat(m.c_str(), &st) == 0) {
      *model = m;
      *params = dir + v[1];
      return true;
    }
  }
  return false;
}

}  // namespace b200ocr
