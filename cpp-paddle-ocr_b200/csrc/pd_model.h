// Paddle ProgramDesc (.pdmodel) / LoDTensor stream (.pdiparams) reader.
//
// Replaces what the reference delegates to Paddle Inference through
// `config.SetModel(model_file_path, param_file_path)` (reference
// src/ocr_det.cpp:46, src/ocr_cls.cpp:130, src/ocr_rec.cpp:162).  Field numbers
// follow the reference's vendored schema
// include/paddle_inference/internal/framework.pb.h; the reader is schema-less
// (no protobuf runtime) and keeps only what the planner needs.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace b200ocr {

struct PdAttr {
  int type = -1;  // AttrType enum of framework.proto
  int64_t i = 0;
  float f = 0.f;
  double d = 0.0;
  bool b = false;
  std::string s;
  std::vector<int64_t> ints;
  std::vector<float> floats;
  std::vector<std::string> strings;
};

struct PdOp {
  std::string type;
  std::map<std::string, std::vector<std::string>> inputs, outputs;
  std::map<std::string, PdAttr> attrs;

  const std::string& in(const std::string& slot, size_t k = 0) const;
  const std::string& out(const std::string& slot, size_t k = 0) const;
  bool has_in(const std::string& slot) const;
  int64_t attr_i(const std::string& n, int64_t dflt = 0) const;
  float attr_f(const std::string& n, float dflt = 0.f) const;
  bool attr_b(const std::string& n, bool dflt = false) const;
  std::string attr_s(const std::string& n, const std::string& dflt = "") const;
  std::vector<int64_t> attr_ints(const std::string& n) const;
  std::vector<float> attr_floats(const std::string& n) const;
};

struct PdVar {
  std::string name;
  bool persistable = false;
  int vtype = -1;  // VarType::Type; 7 = LOD_TENSOR
  int dtype = -1;  // 5 = FP32
  std::vector<int64_t> dims;
};

struct PdProgram {
  std::map<std::string, PdVar> vars;
  std::vector<PdOp> ops;
  std::map<std::string, std::vector<float>> params;  // filled by load_params

  // Names of persistable LoDTensor vars in `.pdiparams` order (ascending name).
  std::vector<std::string> param_names() const;
};

// Both throw std::runtime_error with a message naming the file on failure.
void load_program(const std::string& pdmodel_path, PdProgram* prog);
void load_params(const std::string& pdiparams_path, PdProgram* prog);

// Locate <dir>/inference.pdmodel etc. the way the reference's LoadModel does
// (reference src/ocr_det.cpp:29-45; the .json variants are not supported here).
bool find_model_files(const std::string& model_dir, std::string* model, std::string* params);

}  // namespace b200ocr
