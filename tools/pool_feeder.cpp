// Native client of b200ocr_pool_* (include/b200ocr.h): what a C++ host of the reference's GPUWorkerPool
// (src/gpu_worker_pool.cpp:46-59 submitRequest -> future) does, without an interpreter between the request source and
// the pool.  `bench.py --pool` runs it; it is a measurement tool, not part of the product.
//
//   pool_feeder <model_dir> <raw|rawp|enc> <frames.bin> <index.bin|-> <rows> <cols> <n_items> <n_devices> <workers_per_device>
//               <max_batch> <feeders> <window> <n_warm> <warm_passes>
//
// raw: frames.bin = BGR frames of rows x cols, back to back, held in page-locked memory (b200ocr_host_alloc: the
// pool's clone is then one DMA per frame); rawp: the same in ordinary pageable memory.  enc: frames.bin = the encoded files back to back,
// index.bin = their int64 offsets (one more than files).  Item i is stored frame i mod (frames stored).  Items [0, n_warm) are streamed `warm_passes` times untimed, then items
// [n_warm, n_items) once, timed with the wall clock from the first submit to the last result.  Feeder thread f owns items
// f, f + F, ... and keeps at most `window` of them in flight.  One JSON line on stdout.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "b200ocr.h"

namespace {
using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

std::vector<uint8_t> read_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "pool_feeder: cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b(static_cast<size_t>(n));
  if (n > 0 && fread(b.data(), 1, b.size(), f) != b.size()) { fprintf(stderr, "pool_feeder: short read of %s\n", path); exit(2); }
  fclose(f);
  return b;
}

size_t count_of(const char* s, const char* needle) {
  size_t n = 0;
  const size_t len = strlen(needle);
  for (const char* p = strstr(s, needle); p; p = strstr(p + len, needle)) ++n;
  return n;
}

struct Totals {
  std::atomic<long long> words{0}, fails{0}, submit_us{0}, wait_us{0};
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 15) {
    fprintf(stderr, "usage: pool_feeder model_dir raw|rawp|enc frames.bin index.bin|- rows cols n_items n_devices workers_per_device "
                    "max_batch feeders window n_warm warm_passes\n");
    return 2;
  }
  const char* model_dir = argv[1];
  const bool enc = strcmp(argv[2], "enc") == 0;
  const std::vector<uint8_t> file = read_file(argv[3]);
  const bool pin = strcmp(argv[2], "raw") == 0;
  uint8_t* pinned = pin ? static_cast<uint8_t*>(b200ocr_host_alloc(file.size())) : nullptr;
  if (pin && !pinned) { fprintf(stderr, "pool_feeder: b200ocr_host_alloc(%zu) failed\n", file.size()); return 1; }
  if (pin) memcpy(pinned, file.data(), file.size());
  struct Bytes { const uint8_t* p; size_t n; const uint8_t* data() const { return p; } size_t size() const { return n; } };
  const Bytes frames{pin ? pinned : file.data(), file.size()};
  const int rows = atoi(argv[5]), cols = atoi(argv[6]), n_items = atoi(argv[7]), n_devices = atoi(argv[8]);
  const int wpd = atoi(argv[9]), max_batch = atoi(argv[10]), feeders = atoi(argv[11]), window = atoi(argv[12]);
  const int n_warm = atoi(argv[13]), warm_passes = atoi(argv[14]);
  std::vector<int64_t> index;
  if (enc) {
    const std::vector<uint8_t> ib = read_file(argv[4]);
    index.resize(ib.size() / 8);
    memcpy(index.data(), ib.data(), index.size() * 8);
    if (index.size() < 2) { fprintf(stderr, "pool_feeder: index holds %zu offsets\n", index.size()); return 2; }
  } else if (frames.size() < size_t(rows) * cols * 3 || frames.size() % (size_t(rows) * cols * 3) != 0) {
    fprintf(stderr, "pool_feeder: %zu bytes is not a whole number of %dx%d frames\n", frames.size(), rows, cols);
    return 2;
  }
  // item i is stored frame (file) i mod n_stored: the stream may be longer than the set of distinct inputs
  const size_t n_stored = enc ? index.size() - 1 : frames.size() / (size_t(rows) * cols * 3);
  if (b200ocr_device_count() < n_devices) { fprintf(stderr, "pool_feeder: %d devices visible\n", b200ocr_device_count()); return 2; }

  b200ocr_pool_t pool = nullptr;
  if (b200ocr_pool_create(model_dir, n_devices, nullptr, wpd, 1, max_batch, &pool) != B200OCR_OK) {
    fprintf(stderr, "pool_feeder: %s\n", b200ocr_last_error());
    return 1;
  }

  auto stream = [&](int first, int last, Totals* tot) {
    std::vector<std::thread> ts;
    for (int f = 0; f < feeders; ++f)
      ts.emplace_back([&, f] {
        std::deque<long long> pending;
        long long words = 0, fails = 0, sub_us = 0, wait_us = 0;
        auto collect = [&] {
          char* line = nullptr;
          const auto t0 = Clock::now();
          const int rc = b200ocr_pool_wait(pool, pending.front(), &line);
          wait_us += (long long)(seconds_since(t0) * 1e6);
          pending.pop_front();
          if (rc != B200OCR_OK || !line) { ++fails; return; }
          words += (long long)count_of(line, "\"text\"");
          if (!strstr(line, "\"success\":true")) ++fails;
          b200ocr_free(line);
        };
        for (int i = first + f; i < last; i += feeders) {
          long long ticket = 0;
          const auto t0 = Clock::now();
          int rc;
          const size_t k = size_t(i) % n_stored;
          if (enc) rc = b200ocr_pool_submit_encoded(pool, i, frames.data() + index[k], size_t(index[k + 1] - index[k]), &ticket);
          else {
            b200ocr_image im;
            im.data = frames.data() + k * rows * cols * 3;
            im.rows = rows; im.cols = cols; im.step = size_t(cols) * 3;
            rc = b200ocr_pool_submit(pool, i, &im, &ticket);
          }
          sub_us += (long long)(seconds_since(t0) * 1e6);
          if (rc != B200OCR_OK) { ++fails; continue; }
          pending.push_back(ticket);
          if (int(pending.size()) >= window) collect();
        }
        while (!pending.empty()) collect();
        tot->words += words; tot->fails += fails; tot->submit_us += sub_us; tot->wait_us += wait_us;
      });
    for (auto& t : ts) t.join();
  };

  Totals warm;
  for (int p = 0; p < warm_passes; ++p) stream(0, n_warm, &warm);
  char* st0 = nullptr;
  b200ocr_pool_status(pool, &st0);
  Totals tot;
  const auto t0 = Clock::now();
  stream(n_warm, n_items, &tot);
  const double dt = seconds_since(t0);
  char* st1 = nullptr;
  b200ocr_pool_status(pool, &st1);
  const int n_timed = n_items - n_warm;
  printf("{\"rate\":%.3f,\"seconds\":%.6f,\"items\":%d,\"words\":%lld,\"fails\":%lld,\"warm_fails\":%lld,"
         "\"submit_us_per_item\":%.2f,\"wait_us_per_item\":%.2f,\"feeders\":%d,\"window\":%d,"
         "\"status_before\":%s,\"status_after\":%s}\n",
         n_timed / dt, dt, n_timed, tot.words.load(), tot.fails.load(), warm.fails.load(),
         double(tot.submit_us.load()) / n_timed, double(tot.wait_us.load()) / n_timed, feeders, window,
         st0 ? st0 : "null", st1 ? st1 : "null");
  b200ocr_free(st0);
  b200ocr_free(st1);
  b200ocr_pool_destroy(pool);
  b200ocr_host_free(pinned);
  return 0;
}
