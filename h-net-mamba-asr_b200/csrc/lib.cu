// Library-wide state: thread-local error message and the launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace hnb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = !(getenv("HNB_PDL") && getenv("HNB_PDL")[0] == '0');
  return on;
}
}  // namespace hnb

extern "C" int hnb_version(void) { return 100; }
extern "C" const char* hnb_last_error(void) { return hnb::g_err; }
extern "C" long long hnb_launch_count(void) { return hnb::g_launches.load(); }
extern "C" void hnb_reset_launch_count(void) { hnb::g_launches.store(0); }
