// segment.cu -- fused segment softmax + weighted segment sum for the update operator's SoftAgg
// (devo/blocks.py:40-48:  w = scatter_softmax(g(x), jx);  y = scatter_sum(f(x) * w, jx)),
// the role torch_scatter 2.0.9 plays in the reference.  Groups come from the graph plan
// (graph_plan.cu): rows of a group are perm[gstart[g] .. gstart[g+1]).
//
// grid = (groups, channel chunks of 128); threads = 128 channels (coalesced) x up to 8 row-parts, single pass with an online (running max) softmax: every g/f element is read exactly
// once; nothing but the [n_groups, dim] result is written.  fp32 arithmetic regardless of storage type.
#include <cstdlib>
#include "common.cuh"

namespace {
using devo::ElemTraits;

// 16-byte loads of 16-bit types: blockDim = (16 chunk-threads, parts); a CTA owns 128 channels (16 chunks of 8) of one
// group; part p scans rows s0+p, s0+p+parts, ... with a running-max softmax per channel, 4 rows (8 independent 16-byte
// loads) in flight per thread; the parts are merged through shared memory.
template <typename T> __device__ __forceinline__ void unpack8f(uint4 u, float* v);
template <> __device__ __forceinline__ void unpack8f<__half>(uint4 u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int k = 0; k < 4; k++) { const float2 f = __half22float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8f<__nv_bfloat16>(uint4 u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; k++) { const float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
// blockDim = (16 chunk-threads, P row-parts, GPC groups): a CTA owns 128 channels of GPC consecutive groups; part p of a
// group scans its rows s0+p, s0+p+P, ... -- P is chosen so that a part has <= 4 rows, i.e. ALL loads of a thread (perm,
// then 8 independent 16-byte g / f loads) are in flight at once: the kernel is two L2 round trips plus a shared-memory merge.
template <typename T>
__global__ void __launch_bounds__(512) segment_softmax_sum_vec_kernel(const T* __restrict__ g, const T* __restrict__ f,
                                                                      const int32_t* __restrict__ perm, const int32_t* __restrict__ gstart,
                                                                      const int32_t* __restrict__ ngroups, T* __restrict__ y, int dim, int max_groups,
                                                                      int plan_is_older) {
  extern __shared__ float red[];   // [GPC][P][3][128]
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int P = blockDim.y, part = threadIdx.y, gl = threadIdx.z;
  const int grp = blockIdx.x * blockDim.z + gl;
  const int c0 = (blockIdx.y * 16 + threadIdx.x) * 8;        // first of this thread's 8 channels
  const bool in_range = grp < max_groups;
  // `plan_is_older`: the grouping (ngroups, gstart, perm) was complete before the PREDECESSOR of this launch started
  // (inside devo_gru_update: it is an input of the whole update), so the chain of three dependent index loads is issued
  // before the wait for the predecessor's g / f rows instead of after it
  int G = 0, s0 = 0, s1 = 0;
  int rows0[4] = {-1, -1, -1, -1};
  if (plan_is_older) {
    G = *ngroups;
    if (in_range && grp < G) {
      s0 = gstart[grp]; s1 = gstart[grp + 1];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int ss = s0 + part + u * P; rows0[u] = ss < s1 ? perm[ss] : -1; }
    }
  }
  DEVO_PDL_WAIT();
  if (!plan_is_older) G = *ngroups;
  const bool live = in_range && grp < G;
  float m[8], den[8], num[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { m[k] = -INFINITY; den[k] = 0.f; num[k] = 0.f; }
  if (live) {
    if (!plan_is_older) { s0 = gstart[grp]; s1 = gstart[grp + 1]; }
    for (int s = s0 + part; s < s1; s += 4 * P) {
      uint4 gq[4], fq[4];
      int rows[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ss = s + u * P;
        rows[u] = (plan_is_older && s == s0 + part) ? rows0[u] : (ss < s1 ? perm[ss] : -1);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (rows[u] >= 0) {
          const size_t r = (size_t)rows[u] * dim + c0;
          gq[u] = *reinterpret_cast<const uint4*>(g + r);
          fq[u] = *reinterpret_cast<const uint4*>(f + r);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (rows[u] < 0) continue;
        float gv[8], fv[8];
        unpack8f<T>(gq[u], gv);
        unpack8f<T>(fq[u], fv);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (gv[k] > m[k]) {
            const float sc = __expf(m[k] - gv[k]);   // exp(-inf) = 0 on the first row
            den[k] *= sc; num[k] *= sc; m[k] = gv[k];
          }
          const float e = __expf(gv[k] - m[k]);
          den[k] += e;
          num[k] += e * fv[k];
        }
      }
    }
  }
  float* rg = red + (size_t)gl * P * 3 * 128;
  const int t = threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    rg[(part * 3 + 0) * 128 + t + k] = m[k];
    rg[(part * 3 + 1) * 128 + t + k] = den[k];
    rg[(part * 3 + 2) * 128 + t + k] = num[k];
  }
  __syncthreads();
  // merge the parts (fixed order).  With P > 8 in two levels: thread (part, lane) folds the run of 8 parts number part/8
  // for channel (part % 8) * 16 + lane, then the threads of parts 0..7 fold the P/8 run results -- no thread walks more
  // than 8 + P/8 partials (one level: 128 threads walked all 32 partials of a 96-row group twice, ~1 us)
  const int ch = (part & 7) * 16 + threadIdx.x;            // the channel this thread merges (parts 0..7 cover all 128)
  auto fold = [&](int p0, int p1, float& mm, float& d, float& n) {
    mm = -INFINITY;
    for (int p = p0; p < p1; p++) mm = fmaxf(mm, rg[(p * 3 + 0) * 128 + ch]);
    d = 0.f; n = 0.f;
    for (int p = p0; p < p1; p++) {
      const float mp = rg[(p * 3 + 0) * 128 + ch];
      const float sc = (mp > -INFINITY) ? __expf(mp - mm) : 0.f;
      d += sc * rg[(p * 3 + 1) * 128 + ch];
      n += sc * rg[(p * 3 + 2) * 128 + ch];
    }
  };
  if (P > 8) {
    const int runs = P >> 3, sub = part >> 3;              // P is 16 or 32: 2 or 4 runs of 8 parts
    float mm, d, n;
    fold(sub * 8, sub * 8 + 8, mm, d, n);
    __syncthreads();
    rg[(sub * 3 + 0) * 128 + ch] = mm;
    rg[(sub * 3 + 1) * 128 + ch] = d;
    rg[(sub * 3 + 2) * 128 + ch] = n;
    __syncthreads();
    if (!in_range || part >= 8) return;
    float out = 0.f;
    if (live) {
      fold(0, runs, mm, d, n);
      out = d > 0.f ? n / d : 0.f;
    }
    y[(size_t)grp * dim + blockIdx.y * 128 + ch] = ElemTraits<T>::from_float(out);
    return;
  }
  if (!in_range) return;
  for (int c2 = part * 16 + threadIdx.x; c2 < 128; c2 += 16 * P) {
    float out = 0.f;
    if (live) {
      float mm = -INFINITY;
      for (int p = 0; p < P; p++) mm = fmaxf(mm, rg[(p * 3 + 0) * 128 + c2]);
      float d = 0.f, n = 0.f;
      for (int p = 0; p < P; p++) {
        const float mp = rg[(p * 3 + 0) * 128 + c2];
        const float sc = (mp > -INFINITY) ? __expf(mp - mm) : 0.f;
        d += sc * rg[(p * 3 + 1) * 128 + c2];
        n += sc * rg[(p * 3 + 2) * 128 + c2];
      }
      out = d > 0.f ? n / d : 0.f;
    }
    y[(size_t)grp * dim + blockIdx.y * 128 + c2] = ElemTraits<T>::from_float(out);     // padding groups: zeros
  }
}

// blockDim = (dim_threads, parts): part p of a group scans rows s0+p, s0+p+parts, ... with a running-max
// softmax; the parts are merged through shared memory.  Loads are issued 4 rows ahead of use.
template <typename T>
__global__ void segment_softmax_sum_kernel(const T* __restrict__ g, const T* __restrict__ f,
                                           const int32_t* __restrict__ perm, const int32_t* __restrict__ gstart,
                                           const int32_t* __restrict__ ngroups, T* __restrict__ y, int dim) {
  extern __shared__ float red[];   // [parts][3][blockDim.x]
  // PDL: a dependent kernel launched with programmatic stream serialisation (the h GEMM of the update operator) may
  // begin its set-up now; it waits (griddepcontrol.wait) for this grid's completion before reading y
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  DEVO_PDL_WAIT();                 // launched with PDL itself: g / f are the previous kernel's outputs
  const int grp = blockIdx.x;
  const int G = *ngroups;
  const int parts = blockDim.y, part = threadIdx.y;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;    // one channel per thread, channel chunks across blockIdx.y
  T* yo = y + (size_t)grp * dim;
  if (grp >= G) {   // padding rows of the fixed-size output
    if (part == 0 && c < dim) yo[c] = ElemTraits<T>::from_float(0.f);
    return;
  }
  const int s0 = gstart[grp], s1 = gstart[grp + 1];
  if (c < dim) {
    float m = -INFINITY, den = 0.f, num = 0.f;
    for (int s = s0 + part; s < s1; s += 4 * parts) {
      float gv[4], fv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ss = s + u * parts;
        if (ss < s1) {
          const size_t r = (size_t)perm[ss] * dim + c;
          gv[u] = (float)ElemTraits<T>::to_float(g[r]);
          fv[u] = (float)ElemTraits<T>::to_float(f[r]);
        } else {
          gv[u] = -INFINITY; fv[u] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (gv[u] > m) {
          const float sc = __expf(m - gv[u]);   // exp(-inf) = 0 on the first row
          den *= sc; num *= sc; m = gv[u];
        }
        if (gv[u] > -INFINITY) {
          const float e = __expf(gv[u] - m);
          den += e;
          num += e * fv[u];
        }
      }
    }
    red[(part * 3 + 0) * blockDim.x + threadIdx.x] = m;
    red[(part * 3 + 1) * blockDim.x + threadIdx.x] = den;
    red[(part * 3 + 2) * blockDim.x + threadIdx.x] = num;
  }
  __syncthreads();
  if (part == 0 && c < dim) {
    float m = -INFINITY;
    for (int p = 0; p < parts; p++) m = fmaxf(m, red[(p * 3 + 0) * blockDim.x + threadIdx.x]);
    float den = 0.f, num = 0.f;
    for (int p = 0; p < parts; p++) {
      const float mp = red[(p * 3 + 0) * blockDim.x + threadIdx.x];
      const float sc = (mp > -INFINITY) ? __expf(mp - m) : 0.f;
      den += sc * red[(p * 3 + 1) * blockDim.x + threadIdx.x];
      num += sc * red[(p * 3 + 2) * blockDim.x + threadIdx.x];
    }
    yo[c] = ElemTraits<T>::from_float(den > 0.f ? num / den : 0.f);
  }
}
}  // namespace

// `plan_is_older`: see segment_softmax_sum_vec_kernel (only devo_gru_update may claim it)
int devo::segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                              const int32_t* ngroups, int max_groups, void* y_out, int dtype, int n_rows,
                              int dim, void* stream, int plan_is_older) {
  if (max_groups <= 0 || dim <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int tx = 128;                                        // channels per CTA
  int parts = 8;                                             // row-parts per channel
  const int avg_rows = (n_rows > 0 && max_groups > 0) ? n_rows / max_groups : 1;
  while (parts > 1 && avg_rows < 2 * parts) parts >>= 1;     // small groups: fewer parts
  dim3 block(tx, parts);
  dim3 grid(max_groups, (dim + tx - 1) / tx);
  const size_t smem = (size_t)parts * 3 * tx * sizeof(float);
#define SEG(T) DEVO_CUDA(devo::launch_pdl(segment_softmax_sum_kernel<T>, grid, block, smem, s, (const T*)g, (const T*)f, perm, gstart, ngroups, (T*)y_out, dim))
  if ((dtype == DEVO_F16 || dtype == DEVO_BF16) && dim % 128 == 0 &&
      ((((uintptr_t)g | (uintptr_t)f | (uintptr_t)y_out) & 15) == 0)) {
    // row-parts per group: 4 rows per part (every load of a thread in flight at once) when that still gives at most one
    // CTA per SM, else 8 rows per part -- measured at S8 (tools/segment_timing.py): the launch of more than ~148
    // 512-thread CTAs costs more than a second round of loads (96-row groups: 7.7 us with 16 parts / 96 CTAs, 9.0 us
    // with 32 parts / 192 CTAs; 8-row groups: 6.0 us with 2 parts / 144 CTAs, 7.8 us with 4 parts / 288 CTAs)
    int P = 1;
    while (P < 32 && P * 4 < avg_rows) P <<= 1;
    while (P > 1 && (P / 2) * 8 >= avg_rows && ((max_groups + 32 / P - 1) / (32 / P)) * (dim / 128) > 148) P >>= 1;
    const int GPC = 32 / P;                                 // groups per CTA: 512 threads
    dim3 vblock(16, P, GPC), vgrid((max_groups + GPC - 1) / GPC, dim / 128);
    const size_t vsmem = (size_t)32 * 3 * 128 * sizeof(float);
    if (dtype == DEVO_F16)
      DEVO_CUDA(devo::launch_pdl(segment_softmax_sum_vec_kernel<__half>, vgrid, vblock, vsmem, s, (const __half*)g, (const __half*)f, perm, gstart, ngroups, (__half*)y_out, dim, max_groups, plan_is_older));
    else
      DEVO_CUDA(devo::launch_pdl(segment_softmax_sum_vec_kernel<__nv_bfloat16>, vgrid, vblock, vsmem, s, (const __nv_bfloat16*)g, (const __nv_bfloat16*)f, perm, gstart, ngroups, (__nv_bfloat16*)y_out, dim, max_groups, plan_is_older));
    DEVO_LAUNCH_CHECK("segment_softmax_sum");
    return DEVO_OK;
  }
  switch (dtype) {
    case DEVO_F16: SEG(__half); break;
    case DEVO_BF16: SEG(__nv_bfloat16); break;
    case DEVO_F32: SEG(float); break;
    default: DEVO_REQUIRE(false, DEVO_EINVAL, "segment_softmax_sum: unsupported dtype %d", dtype);
  }
#undef SEG
  DEVO_LAUNCH_CHECK("segment_softmax_sum");
  return DEVO_OK;
}

extern "C" int devo_segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                                        const int32_t* ngroups, int max_groups, void* y_out, int dtype, int n_rows,
                                        int dim, void* stream) {
  return devo::segment_softmax_sum(g, f, perm, gstart, ngroups, max_groups, y_out, dtype, n_rows, dim, stream, 0);
}
