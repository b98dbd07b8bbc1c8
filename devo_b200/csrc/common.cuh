// common.cuh -- shared helpers for libdevo_b200 (error reporting, launch accounting, dtype traits)
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/devo_b200.h"

namespace devo {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DEVO_REQUIRE(cond, code, ...)                    \
  do {                                                   \
    if (!(cond)) {                                       \
      devo::set_error(__VA_ARGS__);                      \
      return (code);                                     \
    }                                                    \
  } while (0)

// call after every kernel launch: records launch-configuration errors
#define DEVO_LAUNCH_CHECK(name)                                              \
  do {                                                                       \
    devo::count_launch();                                                    \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      devo::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e)); \
      return (int)_e;                                                        \
    }                                                                        \
  } while (0)

#define DEVO_CUDA(call)                                                       \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      devo::set_error("%s failed: %s", #call, cudaGetErrorString(_e));        \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

template <typename T> struct ElemTraits;
template <> struct ElemTraits<__half> {
  static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
  using acc_t = float;
};
template <> struct ElemTraits<__nv_bfloat16> {
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
  using acc_t = float;
};
template <> struct ElemTraits<float> {
  static __device__ __forceinline__ float to_float(float v) { return v; }
  static __device__ __forceinline__ float from_float(float v) { return v; }
  using acc_t = float;
};
template <> struct ElemTraits<double> {
  static __device__ __forceinline__ double to_float(double v) { return v; }
  static __device__ __forceinline__ double from_float(double v) { return v; }
  using acc_t = double;
};

static inline int elem_size(int dtype) {
  switch (dtype) {
    case DEVO_F16: case DEVO_BF16: return 2;
    case DEVO_F32: return 4;
    case DEVO_F64: return 8;
  }
  return 0;
}

// Launch with programmatic stream serialisation (PDL): the grid may start while its predecessor in the stream is still
// running; the kernel must execute griddepcontrol.wait before it touches anything the predecessor produced.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ... the same with the grid partitioned into thread-block clusters of `cluster_x` CTAs (gridDim.x must be a multiple)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, size_t smem,
                                             cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DEVO_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define DEVO_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")

// csrc/segment.cu: devo_segment_softmax_sum with one more promise from the caller (see the kernel)
int segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart, const int32_t* ngroups,
                        int max_groups, void* y_out, int dtype, int n_rows, int dim, void* stream, int plan_is_older);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember, per device ordinal, the largest size
// already configured for one kernel (a process-wide flag would leave every device but the first unconfigured).
struct SmemConfig {
  size_t hw[64] = {};
  // true when the attribute has to be (re)set on the current device to cover `smem` bytes
  bool need(size_t smem) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (smem > hw[d]) { hw[d] = smem; return true; }
    return false;
  }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace devo
