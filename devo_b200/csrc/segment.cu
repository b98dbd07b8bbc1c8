// segment.cu -- fused segment softmax + weighted segment sum for the update operator's SoftAgg
// (devo/blocks.py:40-48:  w = scatter_softmax(g(x), jx);  y = scatter_sum(f(x) * w, jx)),
// the role torch_scatter 2.0.9 plays in the reference.  Groups come from the graph plan
// (graph_plan.cu): rows of a group are perm[gstart[g] .. gstart[g+1]).
//
// One CTA per group, one thread per channel, single pass with an online (running max) softmax:
// every g/f element is read exactly once, coalesced along the channel axis; nothing but the
// [n_groups, dim] result is written.  fp32 arithmetic regardless of the storage type.
#include "common.cuh"

namespace {
using devo::ElemTraits;

template <typename T>
__global__ void segment_softmax_sum_kernel(const T* __restrict__ g, const T* __restrict__ f,
                                           const int32_t* __restrict__ perm, const int32_t* __restrict__ gstart,
                                           const int32_t* __restrict__ ngroups, T* __restrict__ y, int dim) {
  const int grp = blockIdx.x;
  const int G = *ngroups;
  T* yo = y + (size_t)grp * dim;
  if (grp >= G) {   // padding rows of the fixed-size output
    for (int c = threadIdx.x; c < dim; c += blockDim.x) yo[c] = ElemTraits<T>::from_float(0.f);
    return;
  }
  const int s0 = gstart[grp], s1 = gstart[grp + 1];
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float m = -INFINITY, den = 0.f, num = 0.f;
    for (int s = s0; s < s1; s++) {
      const size_t r = (size_t)perm[s] * dim + c;
      const float gv = (float)ElemTraits<T>::to_float(g[r]);
      const float fv = (float)ElemTraits<T>::to_float(f[r]);
      if (gv > m) {
        const float sc = __expf(m - gv);   // exp(-inf) = 0 on the first row
        den *= sc; num *= sc; m = gv;
      }
      const float e = __expf(gv - m);
      den += e;
      num += e * fv;
    }
    yo[c] = ElemTraits<T>::from_float(den > 0.f ? num / den : 0.f);
  }
}
}  // namespace

extern "C" int devo_segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                                        const int32_t* ngroups, int max_groups, void* y_out, int dtype, int n_rows,
                                        int dim, void* stream) {
  (void)n_rows;
  if (max_groups <= 0 || dim <= 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int threads = dim >= 512 ? 512 : ((dim + 31) / 32) * 32;
#define SEG(T) segment_softmax_sum_kernel<T><<<max_groups, threads, 0, s>>>((const T*)g, (const T*)f, perm, gstart, ngroups, (T*)y_out, dim)
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
