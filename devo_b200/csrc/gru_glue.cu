// gru_glue.cu -- fused element-wise glue of the update operator's GRU.
//
// The dense layers of `Update` (devo/enet.py:32-99) stay cuBLAS; everything between them -- residual adds,
// LayerNorms, neighbour gathers with masks, gated residuals, ReLU/cast pairs -- is ~50 separate ATen launches
// per iteration in the reference (each a full pass over an [E,384] tensor).  These kernels fuse each chain into
// one pass.  Rounding points follow torch.autocast's dtype flow exactly (Linear outputs are half, LayerNorm
// outputs float32, element-wise ops round to their promoted type), so results match the unfused path to
// LayerNorm-reduction-order noise.  T = __half or __nv_bfloat16 (the autocast dtype).
#include "common.cuh"

namespace {
using devo::ElemTraits;

template <typename T> __device__ __forceinline__ float rnd(float v) { return ElemTraits<T>::to_float(ElemTraits<T>::from_float(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kMaxPerLane = 16;   // dim <= 512

// one warp per row: y = LayerNorm(x) * gamma + beta in fp32 (two-pass, biased variance, eps inside the sqrt)
// MODE 0: x = float(half(half(a+b)+c))   (a,b,c in T; c may be null)     -> out32
// MODE 1: x = x32                                                        -> out32 and out16 = T(y)
// MODE 2: x = float(x16)                                                 -> out16 = T(relu(y))
template <typename T, int MODE>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                        const T* __restrict__ c, const float* __restrict__ x32,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float* __restrict__ out32, T* __restrict__ out16,
                                                        int rows, int dim) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t base = (size_t)row * dim;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) {
    const int i = lane + 32 * k;
    float x = 0.f;
    if (i < dim) {
      if (MODE == 0) {
        float t = rnd<T>(ElemTraits<T>::to_float(a[base + i]) + ElemTraits<T>::to_float(b[base + i]));
        if (c) t = rnd<T>(t + ElemTraits<T>::to_float(c[base + i]));
        x = t;
      } else if (MODE == 1) {
        x = x32[base + i];
      } else {
        x = ElemTraits<T>::to_float(a[base + i]);
      }
    }
    v[k] = x;
    s += x;
  }
  const float mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) {
    const int i = lane + 32 * k;
    if (i < dim) { const float d = v[k] - mean; q += d * d; }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)dim + eps);
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) {
    const int i = lane + 32 * k;
    if (i < dim) {
      float y = (v[k] - mean) * rstd * gamma[i] + beta[i];
      if (MODE == 2) {
        out16[base + i] = ElemTraits<T>::from_float(fmaxf(y, 0.f));
      } else {
        out32[base + i] = y;
        if (MODE == 1 && out16) out16[base + i] = ElemTraits<T>::from_float(y);
      }
    }
  }
}

// ---- vectorised row kernels: blockDim = (dim/4, rows_per_cta); one thread = 4 consecutive channels
// (16-byte float4 loads/stores, 8-byte loads/stores of 4 T).  dim % 4 == 0.
template <typename T> struct alignas(8) Vec4 { T v[4]; };

template <typename T> __device__ __forceinline__ float4 ld4(const T* p) {
  const Vec4<T> r = *reinterpret_cast<const Vec4<T>*>(p);
  return make_float4(ElemTraits<T>::to_float(r.v[0]), ElemTraits<T>::to_float(r.v[1]), ElemTraits<T>::to_float(r.v[2]), ElemTraits<T>::to_float(r.v[3]));
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 f) {
  Vec4<T> r;
  r.v[0] = ElemTraits<T>::from_float(f.x); r.v[1] = ElemTraits<T>::from_float(f.y);
  r.v[2] = ElemTraits<T>::from_float(f.z); r.v[3] = ElemTraits<T>::from_float(f.w);
  *reinterpret_cast<Vec4<T>*>(p) = r;
}

// out[e,:] = idx[e] >= 0 ? T(x32[idx[e],:]) : 0      (mask * net[:, ix] then the Linear's input cast)
template <typename T>
__global__ void gather_mask_cast_kernel(const float* __restrict__ x32, const int64_t* __restrict__ idx,
                                        T* __restrict__ out, int rows, int dim) {
  const int e = blockIdx.x * blockDim.y + threadIdx.y, c = threadIdx.x * 4;
  if (e >= rows || c >= dim) return;
  const long long j = idx[e];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j >= 0) v = *reinterpret_cast<const float4*>(x32 + (size_t)j * dim + c);
  st4<T>(out + (size_t)e * dim + c, v);
}

// net32[e,:] += float(y16[g,:]) with g = gid ? gid[e] : e   (residual add, optionally through a group gather)
template <typename T>
__global__ void residual_add_kernel(float* __restrict__ net32, const T* __restrict__ y16, const int32_t* __restrict__ gid,
                                    T* __restrict__ out16, int rows, int dim) {
  const int e = blockIdx.x * blockDim.y + threadIdx.y, c = threadIdx.x * 4;
  if (e >= rows || c >= dim) return;
  const size_t at = (size_t)e * dim + c;
  const size_t src = gid ? (size_t)gid[e] * dim + c : at;
  float4 v = *reinterpret_cast<const float4*>(net32 + at);
  const float4 y = ld4<T>(y16 + src);
  v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
  *reinterpret_cast<float4*>(net32 + at) = v;
  if (out16) st4<T>(out16 + at, v);                    // the next Linear's input cast, for free
}

// out32 = x32 + float( T( T(sigmoid(gate16)) * res16 ) )      (GatedResidual: x + gate(x) * res(x))
template <typename T>
__global__ void gated_residual_kernel(const float* __restrict__ x32, const T* __restrict__ gate_pre,
                                      const T* __restrict__ res, float* __restrict__ out32, long long total) {
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= total) return;
  float4 x = *reinterpret_cast<const float4*>(x32 + q);
  const float4 g = ld4<T>(gate_pre + q), r = ld4<T>(res + q);
  x.x += rnd<T>(rnd<T>(1.0f / (1.0f + expf(-g.x))) * r.x);
  x.y += rnd<T>(rnd<T>(1.0f / (1.0f + expf(-g.y))) * r.y);
  x.z += rnd<T>(rnd<T>(1.0f / (1.0f + expf(-g.z))) * r.z);
  x.w += rnd<T>(rnd<T>(1.0f / (1.0f + expf(-g.w))) * r.w);
  *reinterpret_cast<float4*>(out32 + q) = x;
}

// out16 = T(relu(x32))
template <typename T>
__global__ void relu_cast_kernel(const float* __restrict__ x32, T* __restrict__ out16, long long total, int relu) {
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= total) return;
  float4 v = *reinterpret_cast<const float4*>(x32 + q);
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  st4<T>(out16 + q, v);
}

// heads: delta = T(W_d . T(relu(x32)) + b_d),  weight = T(sigmoid(T(W_w . T(relu(x32)) + b_w)))   (enet.py:68-77)
// one warp per row; W16 = [4, dim] (rows: d.x, d.y, w.x, w.y), fp32 accumulation like the GEMM it replaces
template <typename T>
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ x32, const T* __restrict__ W16,
                                                    const T* __restrict__ b16, T* __restrict__ delta,
                                                    T* __restrict__ weight, int rows, int dim) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane * 4; c < dim; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x32 + (size_t)row * dim + c);
    const float h0 = rnd<T>(fmaxf(v.x, 0.f)), h1 = rnd<T>(fmaxf(v.y, 0.f)), h2 = rnd<T>(fmaxf(v.z, 0.f)), h3 = rnd<T>(fmaxf(v.w, 0.f));
#pragma unroll
    for (int o = 0; o < 4; o++) {
      const float4 w = ld4<T>(W16 + (size_t)o * dim + c);
      acc[o] += h0 * w.x + h1 * w.y + h2 * w.z + h3 * w.w;
    }
  }
#pragma unroll
  for (int o = 0; o < 4; o++) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    const float d0 = rnd<T>(acc[0] + ElemTraits<T>::to_float(b16[0])), d1 = rnd<T>(acc[1] + ElemTraits<T>::to_float(b16[1]));
    const float w0 = rnd<T>(acc[2] + ElemTraits<T>::to_float(b16[2])), w1 = rnd<T>(acc[3] + ElemTraits<T>::to_float(b16[3]));
    delta[(size_t)row * 2 + 0] = ElemTraits<T>::from_float(d0);
    delta[(size_t)row * 2 + 1] = ElemTraits<T>::from_float(d1);
    weight[(size_t)row * 2 + 0] = ElemTraits<T>::from_float(1.0f / (1.0f + expf(-w0)));
    weight[(size_t)row * 2 + 1] = ElemTraits<T>::from_float(1.0f / (1.0f + expf(-w1)));
  }
}

}  // namespace

#define GLUE_DISPATCH(dtype, NAME, ...)                                                            \
  if (dtype == DEVO_F16) { using T = __half; __VA_ARGS__; }                                        \
  else if (dtype == DEVO_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }                           \
  else { DEVO_REQUIRE(false, DEVO_EINVAL, NAME ": dtype must be f16 or bf16"); }                   \
  DEVO_LAUNCH_CHECK(NAME);                                                                         \
  return DEVO_OK;

extern "C" {

int devo_glue_layernorm(int mode, int dtype, const void* a, const void* b, const void* c, const float* x32,
                        const float* gamma, const float* beta, float eps, float* out32, void* out16, int rows,
                        int dim, void* stream) {
  DEVO_REQUIRE(dim > 0 && dim <= 32 * kMaxPerLane, DEVO_ECAPACITY, "glue_layernorm: dim %d unsupported", dim);
  DEVO_REQUIRE(mode >= 0 && mode <= 2, DEVO_EINVAL, "glue_layernorm: bad mode");
  if (rows <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (rows + 7) / 8;
  GLUE_DISPATCH(dtype, "glue_layernorm", {
    if (mode == 0) layernorm_kernel<T, 0><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (const T*)c, x32, gamma, beta, eps, out32, (T*)out16, rows, dim);
    else if (mode == 1) layernorm_kernel<T, 1><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (const T*)c, x32, gamma, beta, eps, out32, (T*)out16, rows, dim);
    else layernorm_kernel<T, 2><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (const T*)c, x32, gamma, beta, eps, out32, (T*)out16, rows, dim);
  })
}

int devo_glue_gather_mask_cast(int dtype, const float* x32, const int64_t* idx, void* out16, int rows, int dim, void* stream) {
  if (rows <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(dim % 4 == 0 && dim <= 4096, DEVO_EINVAL, "glue_gather_mask_cast: dim must be a multiple of 4");
  const int tx = dim / 4, ty = tx >= 256 ? 1 : 256 / tx;
  dim3 block(tx, ty), grid((rows + ty - 1) / ty);
  GLUE_DISPATCH(dtype, "glue_gather_mask_cast",
                (gather_mask_cast_kernel<T><<<grid, block, 0, s>>>(x32, idx, (T*)out16, rows, dim)))
}

int devo_glue_residual_add(int dtype, float* net32, const void* y16, const int32_t* gid, void* out16, int rows, int dim, void* stream) {
  if (rows <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(dim % 4 == 0 && dim <= 4096, DEVO_EINVAL, "glue_residual_add: dim must be a multiple of 4");
  const int tx = dim / 4, ty = tx >= 256 ? 1 : 256 / tx;
  dim3 block(tx, ty), grid((rows + ty - 1) / ty);
  GLUE_DISPATCH(dtype, "glue_residual_add",
                (residual_add_kernel<T><<<grid, block, 0, s>>>(net32, (const T*)y16, gid, (T*)out16, rows, dim)))
}

int devo_glue_gated_residual(int dtype, const float* x32, const void* gate_pre, const void* res, float* out32, int64_t total, void* stream) {
  if (total <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(total % 4 == 0, DEVO_EINVAL, "glue_gated_residual: element count must be a multiple of 4");
  GLUE_DISPATCH(dtype, "glue_gated_residual",
                (gated_residual_kernel<T><<<(int)((total / 4 + 255) / 256), 256, 0, s>>>(x32, (const T*)gate_pre, (const T*)res, out32, total)))
}

int devo_glue_relu_cast(int dtype, const float* x32, void* out16, int64_t total, int relu, void* stream) {
  if (total <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(total % 4 == 0, DEVO_EINVAL, "glue_relu_cast: element count must be a multiple of 4");
  GLUE_DISPATCH(dtype, "glue_relu_cast", (relu_cast_kernel<T><<<(int)((total / 4 + 255) / 256), 256, 0, s>>>(x32, (T*)out16, total, relu)))
}

int devo_glue_heads(int dtype, const float* x32, const void* W16, const void* b16, void* delta, void* weight, int rows, int dim, void* stream) {
  DEVO_REQUIRE(dim % 4 == 0, DEVO_EINVAL, "glue_heads: dim must be a multiple of 4");
  if (rows <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  GLUE_DISPATCH(dtype, "glue_heads",
                (heads_kernel<T><<<(rows + 7) / 8, 256, 0, s>>>(x32, (const T*)W16, (const T*)b16, (T*)delta, (T*)weight, rows, dim)))
}

}  // extern "C"
