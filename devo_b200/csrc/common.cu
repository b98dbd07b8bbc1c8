// common.cu -- error string, ABI version, launch counter
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace devo {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace devo

extern "C" {
int devo_abi_version(void) { return DEVO_B200_ABI_VERSION; }
const char* devo_last_error(void) { return devo::g_err; }
uint64_t devo_launch_count(void) { return devo::g_launches.load(std::memory_order_relaxed); }
}
