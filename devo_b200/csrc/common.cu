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

// debug (tools/gru_cold_probe.py): read `n` 16-byte words with a chosen cache policy, optionally holding dynamic shared memory
__global__ void __launch_bounds__(256) debug_read_kernel(const uint4* __restrict__ p, size_t n, int mode, unsigned* sink) {
  extern __shared__ unsigned dbg_smem[];
  unsigned acc = 0;
  if (mode == 3) {                         // no memory traffic: spin for n cycles
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)n) acc += 1;
    n = 0;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = (mode == 0) ? __ldca(p + i) : (mode == 1) ? __ldcg(p + i) : __ldcs(p + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) { dbg_smem[threadIdx.x] = acc; *sink = dbg_smem[(threadIdx.x + 1) & 255]; }
}

extern "C" {
int devo_debug_read(const void* p, size_t nbytes, int mode, int smem_bytes, int blocks, void* sink, void* stream) {
  if (smem_bytes > 48 * 1024) cudaFuncSetAttribute(debug_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  debug_read_kernel<<<blocks, 256, smem_bytes < 1024 ? 1024 : smem_bytes, (cudaStream_t)stream>>>((const uint4*)p, mode == 3 ? nbytes : nbytes / 16, mode, (unsigned*)sink);
  return (int)cudaGetLastError();
}
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
