// Library-level entry points: version, error slot, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace b200st {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace b200st

extern "C" {
int b200st_version(void) { return 100; }   // 0.1.0
const char* b200st_last_error(void) { return b200st::g_err; }
int64_t b200st_launch_count(void) { return b200st::g_launches.load(std::memory_order_relaxed); }
}
