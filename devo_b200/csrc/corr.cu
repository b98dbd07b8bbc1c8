// corr.cu -- generic sparse patch correlation lookup (forward + backward) and patch gather
// for planar feature maps of any element type (half / bf16 / float / double), any P, radius.
//
// Reference behaviour: devo/altcorr/correlation_kernel.cu:82-136 (one thread per window
// dot product, 2*C strided scalar loads each), :193-233 (~19 ATen launches for the bilinear
// blend + permute), :139-190/:236-286 (backward: 2*C global atomics per thread + 4 zero
// volumes), :16-80/:288-333 (patchify).
// Here: one CTA per edge.  The patch's C x P^2 features are staged once in shared memory,
// the (2r+2)^2 window volume is kept in shared memory (never written to HBM), accumulation is
// fp32 (fp64 for double inputs) and the bilinear blend + output permutation are fused.
// The B200 fast path for the inference configuration (fp16/bf16, C=128, P=3, r=3) is in
// corr_fast.cu (TMA-staged pixel-major tiles + tensor-core MMA); this file is the
// shape/dtype-generic path and the fp32/fp64 training path.
#include "common.cuh"

namespace {

using devo::ElemTraits;

constexpr int kCorrThreads = 256;

template <typename T> __device__ __forceinline__ void atomic_add_elem(T* p, typename ElemTraits<T>::acc_t v);
template <> __device__ __forceinline__ void atomic_add_elem<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_elem<double>(double* p, double v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_elem<__half>(__half* p, float v) { atomicAdd(p, __float2half_rn(v)); }
template <> __device__ __forceinline__ void atomic_add_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------------- forward
// grid = (E, B).  smem: f1s[C*PP] acc_t, V[PP*D*D] acc_t, geom[PP*4] (fx, fy as int; dx, dy)
template <typename T>
__global__ void __launch_bounds__(kCorrThreads) corr_forward_kernel(
    const T* __restrict__ fmap1, const T* __restrict__ fmap2, const float* __restrict__ coords,
    const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, T* __restrict__ out,
    int Np, int Nf, int C, int H, int W, int E, int P, int R) {
  using acc_t = typename ElemTraits<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PP = P * P, D = 2 * R + 2, DD = D * D, Dm = D - 1;
  acc_t* f1s = reinterpret_cast<acc_t*>(smem_raw);          // [C][PP]
  acc_t* V = f1s + (size_t)C * PP;                          // [PP][D][D]
  int* gx = reinterpret_cast<int*>(V + (size_t)PP * DD);    // [PP] floor(x)
  int* gy = gx + PP;
  float* fdx = reinterpret_cast<float*>(gy + PP);
  float* fdy = fdx + PP;

  const int e = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int ix = (int)ii[e], jx = (int)jj[e];
  const T* f1 = fmap1 + ((size_t)b * Np + ix) * C * PP;
  const T* f2 = fmap2 + ((size_t)b * Nf + jx) * C * H * W;
  const float* co = coords + ((size_t)b * E + e) * 2 * PP;

  for (int q = tid; q < C * PP; q += kCorrThreads) f1s[q] = (acc_t)ElemTraits<T>::to_float(f1[q]);
  if (tid < PP) {
    const float x = co[tid], y = co[PP + tid];
    const float flx = floorf(x), fly = floorf(y);
    gx[tid] = (int)flx; gy[tid] = (int)fly;
    fdx[tid] = x - flx; fdy[tid] = y - fly;
  }
  __syncthreads();

  const size_t HW = (size_t)H * W;
  for (int q = tid; q < PP * DD; q += kCorrThreads) {
    const int bb = q % D, a = (q / D) % D, p = q / DD;     // b' fastest => adjacent threads read adjacent pixels
    const int i1 = gy[p] + a - R, j1 = gx[p] + bb - R;
    acc_t s = 0;
    if (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) {
      const T* src = f2 + (size_t)i1 * W + j1;
      const acc_t* fp = f1s + p;
      acc_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int c = 0;
      for (; c + 4 <= C; c += 4) {
        const acc_t v0 = (acc_t)ElemTraits<T>::to_float(src[(size_t)(c + 0) * HW]);
        const acc_t v1 = (acc_t)ElemTraits<T>::to_float(src[(size_t)(c + 1) * HW]);
        const acc_t v2 = (acc_t)ElemTraits<T>::to_float(src[(size_t)(c + 2) * HW]);
        const acc_t v3 = (acc_t)ElemTraits<T>::to_float(src[(size_t)(c + 3) * HW]);
        s0 += fp[(c + 0) * PP] * v0; s1 += fp[(c + 1) * PP] * v1;
        s2 += fp[(c + 2) * PP] * v2; s3 += fp[(c + 3) * PP] * v3;
      }
      for (; c < C; c++) s0 += fp[c * PP] * (acc_t)ElemTraits<T>::to_float(src[(size_t)c * HW]);
      s = (s0 + s1) + (s2 + s3);
    }
    V[q] = s;
  }
  __syncthreads();

  // bilinear blend (:221-230) and permuted store (:232): out[b,e,xo=b',yo=a,i0,j0]
  T* o = out + ((size_t)b * E + e) * Dm * Dm * PP;
  for (int q = tid; q < Dm * Dm * PP; q += kCorrThreads) {
    const int p = q % PP, yo = (q / PP) % Dm, xo = q / (PP * Dm);
    const acc_t dx = (acc_t)fdx[p], dy = (acc_t)fdy[p];
    const acc_t* v = V + (size_t)p * DD + yo * D + xo;
    const acc_t r = (1 - dx) * (1 - dy) * v[0] + dx * (1 - dy) * v[1] + (1 - dx) * dy * v[D] + dx * dy * v[D + 1];
    o[q] = ElemTraits<T>::from_float(r);
  }
}

// ---------------------------------------------------------------------------------- backward
// grid = (E, B).  smem: gV[PP*D*D] float, f1s[C*PP] acc, geometry
template <typename T>
__global__ void __launch_bounds__(kCorrThreads) corr_backward_kernel(
    const T* __restrict__ fmap1, const T* __restrict__ fmap2, const float* __restrict__ coords,
    const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, const float* __restrict__ grad,
    T* __restrict__ g1, T* __restrict__ g2, int Np, int Nf, int C, int H, int W, int E, int P, int R) {
  using acc_t = typename ElemTraits<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PP = P * P, D = 2 * R + 2, DD = D * D, Dm = D - 1;
  acc_t* f1s = reinterpret_cast<acc_t*>(smem_raw);          // [C][PP]
  acc_t* gV = f1s + (size_t)C * PP;                         // [PP][D][D]
  int* gx = reinterpret_cast<int*>(gV + (size_t)PP * DD);
  int* gy = gx + PP;
  float* fdx = reinterpret_cast<float*>(gy + PP);
  float* fdy = fdx + PP;
  __shared__ int s_box[4];   // x0, y0, bw, bh

  const int e = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int ix = (int)ii[e], jx = (int)jj[e];
  const T* f1 = fmap1 + ((size_t)b * Np + ix) * C * PP;
  const T* f2 = fmap2 + ((size_t)b * Nf + jx) * C * H * W;
  T* o1 = g1 + ((size_t)b * Np + ix) * C * PP;
  T* o2 = g2 + ((size_t)b * Nf + jx) * C * H * W;
  const float* co = coords + ((size_t)b * E + e) * 2 * PP;
  const float* g = grad + ((size_t)b * E + e) * Dm * Dm * PP;   // [xo][yo][p]

  for (int q = tid; q < C * PP; q += kCorrThreads) f1s[q] = (acc_t)ElemTraits<T>::to_float(f1[q]);
  if (tid < PP) {
    const float x = co[tid], y = co[PP + tid];
    const float flx = floorf(x), fly = floorf(y);
    gx[tid] = (int)flx; gy[tid] = (int)fly;
    fdx[tid] = x - flx; fdy[tid] = y - fly;
  }
  __syncthreads();
  if (tid == 0) {
    int x0 = gx[0], x1 = gx[0], y0 = gy[0], y1 = gy[0];
    for (int p = 1; p < PP; p++) {
      x0 = min(x0, gx[p]); x1 = max(x1, gx[p]);
      y0 = min(y0, gy[p]); y1 = max(y1, gy[p]);
    }
    s_box[0] = x0 - R; s_box[1] = y0 - R; s_box[2] = x1 - x0 + D; s_box[3] = y1 - y0 + D;
  }
  // transposed bilinear blend (:252-269): gV[p][a][b'] gathers the (up to) four outputs it fed
  for (int q = tid; q < PP * DD; q += kCorrThreads) {
    const int bb = q % D, a = (q / D) % D, p = q / DD;
    const acc_t dx = (acc_t)fdx[p], dy = (acc_t)fdy[p];
    acc_t s = 0;
    // output (yo,xo) used V[yo+{0,1}][xo+{0,1}]
    if (a < Dm && bb < Dm)  s += (1 - dx) * (1 - dy) * (acc_t)g[((size_t)bb * Dm + a) * PP + p];
    if (a < Dm && bb >= 1)  s += dx * (1 - dy) * (acc_t)g[((size_t)(bb - 1) * Dm + a) * PP + p];
    if (a >= 1 && bb < Dm)  s += (1 - dx) * dy * (acc_t)g[((size_t)bb * Dm + (a - 1)) * PP + p];
    if (a >= 1 && bb >= 1)  s += dx * dy * (acc_t)g[((size_t)(bb - 1) * Dm + (a - 1)) * PP + p];
    gV[q] = s;
  }
  __syncthreads();

  const size_t HW = (size_t)H * W;
  // fmap1 gradient: one (c,p) per thread, sum over the window, then a single atomic per (c,p)
  for (int q = tid; q < C * PP; q += kCorrThreads) {
    const int p = q % PP, c = q / PP;
    const T* src = f2 + (size_t)c * HW;
    acc_t s = 0;
    for (int a = 0; a < D; a++) {
      const int i1 = gy[p] + a - R;
      if (i1 < 0 || i1 >= H) continue;
      for (int bb = 0; bb < D; bb++) {
        const int j1 = gx[p] + bb - R;
        if (j1 < 0 || j1 >= W) continue;
        s += gV[(size_t)p * DD + a * D + bb] * (acc_t)ElemTraits<T>::to_float(src[(size_t)i1 * W + j1]);
      }
    }
    atomic_add_elem<T>(o1 + q, s);
  }
  // fmap2 gradient: contributions of the PP patch pixels are pre-summed per bounding-box pixel
  const int x0 = s_box[0], y0 = s_box[1], bw = s_box[2], bh = s_box[3];
  if ((long long)bw * bh <= 1024) {
    const int npx = bw * bh;
    for (int q = tid; q < C * npx; q += kCorrThreads) {
      const int px = q % npx, c = q / npx;
      const int u = px / bw, v = px % bw;
      const int i1 = y0 + u, j1 = x0 + v;
      if (i1 < 0 || i1 >= H || j1 < 0 || j1 >= W) continue;
      acc_t s = 0;
      for (int p = 0; p < PP; p++) {
        const int a = i1 - (gy[p] - R), bb = j1 - (gx[p] - R);
        if (a >= 0 && a < D && bb >= 0 && bb < D) s += gV[(size_t)p * DD + a * D + bb] * f1s[c * PP + p];
      }
      atomic_add_elem<T>(o2 + (size_t)c * HW + (size_t)i1 * W + j1, s);
    }
  } else {   // degenerate geometry (patch pixels far apart): per-window atomics
    for (int q = tid; q < C * PP * DD; q += kCorrThreads) {
      const int bb = q % D, a = (q / D) % D, p = (q / DD) % PP, c = q / (DD * PP);
      const int i1 = gy[p] + a - R, j1 = gx[p] + bb - R;
      if (i1 < 0 || i1 >= H || j1 < 0 || j1 >= W) continue;
      atomic_add_elem<T>(o2 + (size_t)c * HW + (size_t)i1 * W + j1, gV[(size_t)p * DD + a * D + bb] * f1s[c * PP + p]);
    }
  }
}

// ---------------------------------------------------------------------------------- patchify
template <typename T>
__global__ void patchify_forward_kernel(const T* __restrict__ net, const float* __restrict__ coords,
                                        T* __restrict__ patches, int B, int C, int H, int W, int M, int R) {
  const int D = 2 * R + 2;
  const long long total = (long long)B * M * C * D * D;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)blockDim.x * gridDim.x) {
    long long t = q;
    const int bb = (int)(t % D); t /= D;
    const int a = (int)(t % D); t /= D;
    const int c = (int)(t % C); t /= C;
    const int m = (int)(t % M); t /= M;
    const int b = (int)t;
    const float x = coords[((size_t)b * M + m) * 2 + 0], y = coords[((size_t)b * M + m) * 2 + 1];
    const int i = (int)floorf(y) + a - R, j = (int)floorf(x) + bb - R;
    T v = ElemTraits<T>::from_float(0);
    if (i >= 0 && i < H && j >= 0 && j < W) v = net[(((size_t)b * C + c) * H + i) * W + j];
    patches[q] = v;
  }
}

template <typename T>
__global__ void patchify_backward_kernel(const T* __restrict__ pg, const float* __restrict__ coords,
                                         T* __restrict__ net_grad, int B, int C, int H, int W, int M, int R) {
  const int D = 2 * R + 2;
  const long long total = (long long)B * M * C * D * D;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)blockDim.x * gridDim.x) {
    long long t = q;
    const int bb = (int)(t % D); t /= D;
    const int a = (int)(t % D); t /= D;
    const int c = (int)(t % C); t /= C;
    const int m = (int)(t % M); t /= M;
    const int b = (int)t;
    const float x = coords[((size_t)b * M + m) * 2 + 0], y = coords[((size_t)b * M + m) * 2 + 1];
    const int i = (int)floorf(y) + a - R, j = (int)floorf(x) + bb - R;
    if (i >= 0 && i < H && j >= 0 && j < W)
      atomic_add_elem<T>(net_grad + (((size_t)b * C + c) * H + i) * W + j,
                         (typename ElemTraits<T>::acc_t)ElemTraits<T>::to_float(pg[q]));
  }
}

template <typename T>
static int corr_fwd_launch(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii,
                           const int64_t* jj, void* out, int B, int Np, int Nf, int C, int H, int W, int E,
                           int P, int R, cudaStream_t s) {
  using acc_t = typename ElemTraits<T>::acc_t;
  const int PP = P * P, D = 2 * R + 2;
  const size_t smem = ((size_t)C * PP + (size_t)PP * D * D) * sizeof(acc_t) + (size_t)PP * 16;
  DEVO_REQUIRE(smem <= 200 * 1024, DEVO_ECAPACITY, "corr_forward: C*P*P too large for shared memory");
  static devo::SmemConfig configured;
  if (configured.need(smem)) {
    DEVO_CUDA(cudaFuncSetAttribute(corr_forward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  corr_forward_kernel<T><<<dim3(E, B), kCorrThreads, smem, s>>>((const T*)fmap1, (const T*)fmap2, coords, ii, jj,
                                                                (T*)out, Np, Nf, C, H, W, E, P, R);
  DEVO_LAUNCH_CHECK("corr_forward");
  return DEVO_OK;
}

template <typename T>
static int corr_bwd_launch(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii,
                           const int64_t* jj, const float* grad, void* g1, void* g2, int B, int Np, int Nf,
                           int C, int H, int W, int E, int P, int R, cudaStream_t s) {
  using acc_t = typename ElemTraits<T>::acc_t;
  const int PP = P * P, D = 2 * R + 2;
  DEVO_CUDA(cudaMemsetAsync(g1, 0, (size_t)B * Np * C * PP * sizeof(T), s));
  DEVO_CUDA(cudaMemsetAsync(g2, 0, (size_t)B * Nf * C * H * W * sizeof(T), s));
  if (E == 0) return DEVO_OK;
  const size_t smem = ((size_t)C * PP + (size_t)PP * D * D) * sizeof(acc_t) + (size_t)PP * 16;
  DEVO_REQUIRE(smem <= 200 * 1024, DEVO_ECAPACITY, "corr_backward: C*P*P too large for shared memory");
  static devo::SmemConfig configured;
  if (configured.need(smem)) {
    DEVO_CUDA(cudaFuncSetAttribute(corr_backward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  corr_backward_kernel<T><<<dim3(E, B), kCorrThreads, smem, s>>>((const T*)fmap1, (const T*)fmap2, coords, ii, jj,
                                                                 grad, (T*)g1, (T*)g2, Np, Nf, C, H, W, E, P, R);
  DEVO_LAUNCH_CHECK("corr_backward");
  return DEVO_OK;
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

#define DISPATCH_ELEM(dtype, NAME, CALL)                                         \
  switch (dtype) {                                                               \
    case DEVO_F16: { using T = __half; return CALL; }                            \
    case DEVO_BF16: { using T = __nv_bfloat16; return CALL; }                    \
    case DEVO_F32: { using T = float; return CALL; }                             \
    case DEVO_F64: { using T = double; return CALL; }                            \
    default: DEVO_REQUIRE(false, DEVO_EINVAL, NAME ": unsupported dtype %d", dtype); \
  }

extern "C" {

int devo_corr_forward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii,
                      const int64_t* jj, void* out, int dtype, int B, int Np, int Nf, int C, int H, int W,
                      int E, int P, int radius, void* stream) {
  DEVO_REQUIRE(B >= 0 && E >= 0 && C > 0 && P > 0 && radius >= 0, DEVO_EINVAL, "corr_forward: bad sizes");
  if (E == 0 || B == 0) return DEVO_OK;
  DEVO_REQUIRE(B <= 65535, DEVO_EINVAL, "corr_forward: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_ELEM(dtype, "corr_forward",
                corr_fwd_launch<T>(fmap1, fmap2, coords, ii, jj, out, B, Np, Nf, C, H, W, E, P, radius, s));
}

int devo_corr_backward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii,
                       const int64_t* jj, const float* grad, void* fmap1_grad, void* fmap2_grad, int dtype,
                       int B, int Np, int Nf, int C, int H, int W, int E, int P, int radius, void* stream) {
  DEVO_REQUIRE(B >= 0 && E >= 0 && C > 0 && P > 0 && radius >= 0, DEVO_EINVAL, "corr_backward: bad sizes");
  DEVO_REQUIRE(B <= 65535, DEVO_EINVAL, "corr_backward: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_ELEM(dtype, "corr_backward",
                corr_bwd_launch<T>(fmap1, fmap2, coords, ii, jj, grad, fmap1_grad, fmap2_grad, B, Np, Nf, C, H,
                                   W, E, P, radius, s));
}

int devo_patchify_forward(const void* net, const float* coords, void* patches, int dtype, int B, int C, int H,
                          int W, int M, int radius, void* stream) {
  const int D = 2 * radius + 2;
  const long long total = (long long)B * M * C * D * D;
  if (total <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define PF(T) (patchify_forward_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)net, coords, (T*)patches, B, C, H, W, M, radius))
  switch (dtype) {
    case DEVO_F16: PF(__half); break;
    case DEVO_BF16: PF(__nv_bfloat16); break;
    case DEVO_F32: PF(float); break;
    case DEVO_F64: PF(double); break;
    default: DEVO_REQUIRE(false, DEVO_EINVAL, "patchify_forward: unsupported dtype %d", dtype);
  }
#undef PF
  DEVO_LAUNCH_CHECK("patchify_forward");
  return DEVO_OK;
}

int devo_patchify_backward(const void* patch_grad, const float* coords, void* net_grad, int dtype, int B, int C,
                           int H, int W, int M, int radius, void* stream) {
  const int D = 2 * radius + 2;
  const long long total = (long long)B * M * C * D * D;
  cudaStream_t s = (cudaStream_t)stream;
  const int es = devo::elem_size(dtype);
  DEVO_REQUIRE(es > 0, DEVO_EINVAL, "patchify_backward: unsupported dtype %d", dtype);
  DEVO_CUDA(cudaMemsetAsync(net_grad, 0, (size_t)B * C * H * W * es, s));
  if (total <= 0) return DEVO_OK;
#define PB(T) (patchify_backward_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)patch_grad, coords, (T*)net_grad, B, C, H, W, M, radius))
  switch (dtype) {
    case DEVO_F16: PB(__half); break;
    case DEVO_BF16: PB(__nv_bfloat16); break;
    case DEVO_F32: PB(float); break;
    case DEVO_F64: PB(double); break;
  }
#undef PB
  DEVO_LAUNCH_CHECK("patchify_backward");
  return DEVO_OK;
}

}  // extern "C"
