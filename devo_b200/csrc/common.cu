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

// device-to-device copy as a kernel (see devo_copy_bytes in the header)
template <typename V>
__global__ void __launch_bounds__(256) copy_kernel(V* __restrict__ dst, const V* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" {
int devo_copy_bytes(void* dst, const void* src, size_t nbytes, void* stream) {
  if (nbytes == 0) return DEVO_OK;
  DEVO_REQUIRE(dst != nullptr && src != nullptr, DEVO_EINVAL, "copy_bytes: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (((uintptr_t)dst | (uintptr_t)src | (uintptr_t)nbytes) & 15) == 0;
  const size_t n = vec ? nbytes / 16 : nbytes;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (vec) copy_kernel<uint4><<<(unsigned)blocks, 256, 0, s>>>((uint4*)dst, (const uint4*)src, n);
  else copy_kernel<unsigned char><<<(unsigned)blocks, 256, 0, s>>>((unsigned char*)dst, (const unsigned char*)src, n);
  DEVO_LAUNCH_CHECK("copy_bytes");
  return DEVO_OK;
}
int devo_abi_version(void) { return DEVO_B200_ABI_VERSION; }
const char* devo_last_error(void) { return devo::g_err; }
uint64_t devo_launch_count(void) { return devo::g_launches.load(std::memory_order_relaxed); }
}
