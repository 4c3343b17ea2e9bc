#include <cstdlib>
#include <string>

#include "../../include/b200ocr.h"
#include "capi_util.h"
#include "pd_model.h"
#include "plan.h"
#include <cstring>
#include <sstream>

namespace b200ocr {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
std::string json_quote(const std::string& s);  // stages.cu
}  // namespace b200ocr

extern "C" {
const char* b200ocr_last_error(void) { return b200ocr::g_last_error.c_str(); }
const char* b200ocr_version(void) { return "b200ocr 0.1 (sm_100a)"; }
void b200ocr_free(void* p) { free(p); }

int b200ocr_model_params_json(const char* pdmodel_path, char** json) {
  return b200ocr::capi_guard([&] {
    if (!pdmodel_path || !json) throw std::invalid_argument("null argument");
    b200ocr::PdProgram prog;
    b200ocr::load_program(pdmodel_path, &prog);
    std::ostringstream os;
    os << "[";
    bool first = true;
    for (const std::string& n : prog.param_names()) {
      if (!first) os << ",";
      first = false;
      os << "{\"name\":" << b200ocr::json_quote(n) << ",\"dims\":[";
      const auto& d = prog.vars.at(n).dims;
      for (size_t i = 0; i < d.size(); ++i) os << (i ? "," : "") << d[i];
      os << "]}";
    }
    os << "]";
    std::string s = os.str();
    *json = static_cast<char*>(malloc(s.size() + 1));
    if (!*json) throw std::bad_alloc();
    memcpy(*json, s.c_str(), s.size() + 1);
  });
}

int b200ocr_model_plan_text(const char* model_dir, char** text) {
  return b200ocr::capi_guard([&] {
    if (!model_dir || !text) throw std::invalid_argument("null argument");
    std::string mfile, pfile;
    if (!b200ocr::find_model_files(model_dir, &mfile, &pfile))
      throw std::runtime_error(std::string("No valid model file found in ") + model_dir);
    b200ocr::PdProgram prog;
    b200ocr::load_program(mfile, &prog);
    b200ocr::load_params(pfile, &prog);
    b200ocr::Plan plan;
    b200ocr::build_plan(prog, &plan);
    std::string s = plan.dump();
    *text = static_cast<char*>(malloc(s.size() + 1));
    if (!*text) throw std::bad_alloc();
    memcpy(*text, s.c_str(), s.size() + 1);
  });
}
}
