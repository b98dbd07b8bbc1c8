// segment.cu -- fused segment softmax + weighted segment sum for the update operator's SoftAgg
// (devo/blocks.py:40-48:  w = scatter_softmax(g(x), jx);  y = scatter_sum(f(x) * w, jx)),
// the role torch_scatter 2.0.9 plays in the reference.  Groups come from the graph plan
// (graph_plan.cu): rows of a group are perm[gstart[g] .. gstart[g+1]).
//
// One CTA per group; threads cover the channels (coalesced along the channel axis) times up to 4
// row-parts, single pass with an online (running max) softmax: every g/f element is read exactly
// once; nothing but the [n_groups, dim] result is written.  fp32 arithmetic regardless of storage type.
#include "common.cuh"

namespace {
using devo::ElemTraits;

// blockDim = (dim_threads, parts): part p of a group scans rows s0+p, s0+p+parts, ... with a running-max
// softmax; the parts are merged through shared memory.  Loads are issued 4 rows ahead of use.
template <typename T>
__global__ void segment_softmax_sum_kernel(const T* __restrict__ g, const T* __restrict__ f,
                                           const int32_t* __restrict__ perm, const int32_t* __restrict__ gstart,
                                           const int32_t* __restrict__ ngroups, T* __restrict__ y, int dim) {
  extern __shared__ float red[];   // [parts][3][dim]
  const int grp = blockIdx.x;
  const int G = *ngroups;
  const int parts = blockDim.y, part = threadIdx.y;
  T* yo = y + (size_t)grp * dim;
  if (grp >= G) {   // padding rows of the fixed-size output
    if (part == 0)
      for (int c = threadIdx.x; c < dim; c += blockDim.x) yo[c] = ElemTraits<T>::from_float(0.f);
    return;
  }
  const int s0 = gstart[grp], s1 = gstart[grp + 1];
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
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
    red[(part * 3 + 0) * dim + c] = m;
    red[(part * 3 + 1) * dim + c] = den;
    red[(part * 3 + 2) * dim + c] = num;
  }
  __syncthreads();
  if (part == 0) {
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      float m = -INFINITY;
      for (int p = 0; p < parts; p++) m = fmaxf(m, red[(p * 3 + 0) * dim + c]);
      float den = 0.f, num = 0.f;
      for (int p = 0; p < parts; p++) {
        const float mp = red[(p * 3 + 0) * dim + c];
        const float sc = (mp > -INFINITY) ? __expf(mp - m) : 0.f;
        den += sc * red[(p * 3 + 1) * dim + c];
        num += sc * red[(p * 3 + 2) * dim + c];
      }
      yo[c] = ElemTraits<T>::from_float(den > 0.f ? num / den : 0.f);
    }
  }
}
}  // namespace

extern "C" int devo_segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                                        const int32_t* ngroups, int max_groups, void* y_out, int dtype, int n_rows,
                                        int dim, void* stream) {
  if (max_groups <= 0 || dim <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int tx = dim >= 256 ? 256 : ((dim + 31) / 32) * 32;
  int parts = 1024 / tx;
  if (parts > 4) parts = 4;
  if (n_rows > 0 && max_groups > 0 && n_rows / max_groups < 2 * parts) parts = 1;   // tiny groups: no split
  dim3 block(tx, parts);
  const size_t smem = (size_t)parts * 3 * dim * sizeof(float);
  DEVO_REQUIRE(smem <= 48 * 1024, DEVO_ECAPACITY, "segment_softmax_sum: dim too large");
#define SEG(T) segment_softmax_sum_kernel<T><<<max_groups, block, smem, s>>>((const T*)g, (const T*)f, perm, gstart, ngroups, (T*)y_out, dim)
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
