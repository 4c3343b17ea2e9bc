// Error plumbing of the C ABI: every entry point returns an int code and never throws or exits
// (the reference stages are noexcept and exit(1) on a missing model, src/ocr_det.cpp:41-45;
// SURVEY.md §8b asks for error codes instead).
#pragma once
#include <exception>
#include <new>
#include <stdexcept>
#include <string>

namespace b200ocr {

void set_last_error(const std::string& msg);

template <class F>
int capi_guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::invalid_argument& e) {
    set_last_error(e.what());
    return 1;  // B200OCR_ERR_INVALID
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return 3;  // B200OCR_ERR_NOMEM
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 2;  // B200OCR_ERR_RUNTIME
  } catch (...) {
    set_last_error("unknown error");
    return 2;
  }
}

}  // namespace b200ocr
