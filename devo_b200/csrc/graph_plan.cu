// graph_plan.cu -- edge-graph analysis shared by neighbors / BA / segment ops.
//
// Replaces, on the GPU and without any host round trip:
//   * cuda_ba.neighbors          devo/fastba/ba.cpp:104-149 (torch::_unique -> D2H -> host stable_sort -> H2D)
//   * torch::_unique(kk)         devo/fastba/ba_cuda.cu:435-437
//   * torch.unique(return_inverse) in SoftAgg, devo/blocks.py:41
//
// Edges are sorted by (ka, kb, edge index).  For E <= 16384 the whole analysis is ONE kernel
// launch: a single CTA radix-sorts (key,index) pairs in shared memory (cub::BlockRadixSort on
// just the significant bits, found by an in-kernel max-reduce), then derives head flags, a
// prefix sum (dense group ids), group starts, unique keys and the prev/next links.  Larger
// graphs take a multi-kernel path around cub::DeviceRadixSort / DeviceScan.
#include <cub/cub.cuh>
#include "common.cuh"

namespace {

__device__ __forceinline__ int bits_needed(unsigned long long v) {   // v = max value; bits to hold it (>=1)
  int b = 64 - __clzll(v);
  return b < 1 ? 1 : b;
}

template <int NT, int IPT>
struct PlanSmall {
  using Sort = cub::BlockRadixSort<unsigned long long, NT, IPT, int>;
  using Sort32 = cub::BlockRadixSort<unsigned int, NT, IPT, int, 5>;   // keys that fit 32 bits: half the exchange traffic
  using Reduce = cub::BlockReduce<unsigned long long, NT>;
  using Scan = cub::BlockScan<int, NT>;
  union Temp {
    typename Sort::TempStorage sort;
    typename Sort32::TempStorage sort32;
    typename Reduce::TempStorage reduce;
    typename Scan::TempStorage scan;
  };
};

template <int NT, int IPT>
__global__ void __launch_bounds__(NT) plan_small_kernel(
    const int64_t* __restrict__ ka, const int64_t* __restrict__ kb, int E,
    int32_t* __restrict__ perm, int32_t* __restrict__ gid, int32_t* __restrict__ gstart,
    int64_t* __restrict__ gkey, int32_t* __restrict__ ngroups, int64_t* __restrict__ ix,
    int64_t* __restrict__ jx, int64_t* __restrict__ sorted_a /* workspace i64[E] */,
    int32_t* __restrict__ perm_ws /* workspace i32[E] (used when perm == NULL) */,
    int hint_bits_a, int hint_bits_b /* significant bits of ka / kb when the caller gave both bounds, else -1 */) {
  using P = PlanSmall<NT, IPT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename P::Temp& temp = *reinterpret_cast<typename P::Temp*>(smem_raw);
  __shared__ unsigned long long s_max[2];
  __shared__ int s_total;

  const int tid = threadIdx.x;
  unsigned long long a[IPT], b[IPT];
  unsigned long long ma = 0, mb = 0;
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    int idx = tid * IPT + k;
    if (idx < E) {
      a[k] = (unsigned long long)ka[idx];
      b[k] = (unsigned long long)kb[idx];
      ma = a[k] > ma ? a[k] : ma;
      mb = b[k] > mb ? b[k] : mb;
    } else {
      a[k] = 0; b[k] = 0;
    }
  }
  int bits_a, bits_b;
  // both exclusive bounds known on the host: no max-reductions -- unless a key breaks its bound (one block-wide OR), in which
  // case the bounds are ignored: the hint can make the plan faster, never wrong
  bool hinted = (hint_bits_a > 0 && hint_bits_b > 0);
  if (hinted) hinted = !__syncthreads_or((int)(((ma >> hint_bits_a) | (mb >> hint_bits_b)) != 0ull));
  if (hinted) {
    bits_a = hint_bits_a; bits_b = hint_bits_b;
  } else {
    ma = typename P::Reduce(temp.reduce).Reduce(ma, cub::Max());
    __syncthreads();
    mb = typename P::Reduce(temp.reduce).Reduce(mb, cub::Max());
    if (tid == 0) { s_max[0] = ma; s_max[1] = mb; }
    __syncthreads();
    bits_b = bits_needed(s_max[1]);
    bits_a = bits_needed(s_max[0]);
  }
  if (bits_a + bits_b > 62) bits_a = 62 - bits_b;   // ids are tensor indices; cannot happen in practice
  const int bits = bits_a + bits_b;

  unsigned long long keys[IPT];
  int vals[IPT];
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    int idx = tid * IPT + k;
    vals[k] = idx;
    keys[k] = (idx < E) ? ((a[k] << bits_b) | b[k]) : (1ull << bits);   // padding sorts last
  }
  if (bits + 1 <= 32) {     // block-uniform
    unsigned int k32[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) k32[k] = (unsigned int)keys[k];
    typename P::Sort32(temp.sort32).Sort(k32, vals, 0, bits + 1);
#pragma unroll
    for (int k = 0; k < IPT; k++) keys[k] = k32[k];
  } else {
    typename P::Sort(temp.sort).Sort(keys, vals, 0, bits + 1);
  }
  __syncthreads();

  int32_t* pm = perm ? perm : perm_ws;
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    int s = tid * IPT + k;
    if (s < E) {
      pm[s] = vals[k];
      sorted_a[s] = (int64_t)(keys[k] >> bits_b);
    }
  }
  __syncthreads();   // global writes of this CTA are visible to the CTA after the barrier

  int head[IPT];
  int local = 0;
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    int s = tid * IPT + k;
    head[k] = 0;
    if (s < E) {
      long long me = (long long)(keys[k] >> bits_b);
      head[k] = (s == 0) || (sorted_a[s - 1] != me);
    }
    local += head[k];
  }
  int excl, total;
  typename P::Scan(temp.scan).ExclusiveSum(local, excl, total);
  if (tid == 0) s_total = total;
  int run = excl;
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    int s = tid * IPT + k;
    if (s < E) {
      run += head[k];
      const int g = run - 1;
      const int e = vals[k];
      const long long me = (long long)(keys[k] >> bits_b);
      if (gid) gid[e] = g;
      if (head[k]) {
        if (gstart) gstart[g] = s;
        if (gkey) gkey[g] = me;
      }
      if (ix) ix[e] = head[k] ? -1 : (int64_t)pm[s - 1];
      if (jx) jx[e] = (s + 1 < E && sorted_a[s + 1] == me) ? (int64_t)pm[s + 1] : -1;
    }
  }
  if (tid == 0) {
    if (gstart) gstart[total] = E;
    if (ngroups) ngroups[0] = total;
  }
}

// ---- large path kernels -----------------------------------------------------------------
__global__ void plan_make_keys(const int64_t* __restrict__ ka, const int64_t* __restrict__ kb, int E,
                               unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) {
    keys[i] = ((unsigned long long)ka[i] << 32) | ((unsigned long long)kb[i] & 0xffffffffull);
    vals[i] = i;
  }
}
__global__ void plan_heads(const unsigned long long* __restrict__ keys, int E, int* __restrict__ flag) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < E) flag[s] = (s == 0) || ((keys[s] >> 32) != (keys[s - 1] >> 32));
}
__global__ void plan_finalize(const unsigned long long* __restrict__ keys, const int* __restrict__ pm,
                              const int* __restrict__ incl, int E, int32_t* __restrict__ gid,
                              int32_t* __restrict__ gstart, int64_t* __restrict__ gkey,
                              int32_t* __restrict__ ngroups, int64_t* __restrict__ ix, int64_t* __restrict__ jx) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < E) {
    const unsigned long long me = keys[s] >> 32;
    const bool head = (s == 0) || ((keys[s - 1] >> 32) != me);
    const int g = incl[s] - 1;
    const int e = pm[s];
    if (gid) gid[e] = g;
    if (head) {
      if (gstart) gstart[g] = s;
      if (gkey) gkey[g] = (int64_t)me;
    }
    if (ix) ix[e] = head ? -1 : (int64_t)pm[s - 1];
    if (jx) jx[e] = (s + 1 < E && (keys[s + 1] >> 32) == me) ? (int64_t)pm[s + 1] : -1;
    if (s == E - 1) {
      if (gstart) gstart[g + 1] = E;
      if (ngroups) ngroups[0] = g + 1;
    }
  }
}

constexpr int kSmallMax = 16384;

struct LargeLayout {
  size_t keys_in, keys_out, vals_in, vals_out, flag, incl, cub_temp, cub_bytes, total;
};
static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static LargeLayout large_layout(int E) {
  LargeLayout L;
  size_t off = 0;
  L.keys_in = off;  off += align_up((size_t)E * 8);
  L.keys_out = off; off += align_up((size_t)E * 8);
  L.vals_in = off;  off += align_up((size_t)E * 4);
  L.vals_out = off; off += align_up((size_t)E * 4);
  L.flag = off;     off += align_up((size_t)E * 4);
  L.incl = off;     off += align_up((size_t)E * 4);
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, b1, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, E, 0, 64, (cudaStream_t)0);
  cub::DeviceScan::InclusiveSum((void*)nullptr, b2, (int*)nullptr, (int*)nullptr, E, (cudaStream_t)0);
  L.cub_bytes = b1 > b2 ? b1 : b2;
  L.cub_temp = off; off += align_up(L.cub_bytes);
  L.total = off;
  return L;
}

template <int NT, int IPT>
static int launch_small(const int64_t* ka, const int64_t* kb, int E, int32_t* perm, int32_t* gid,
                        int32_t* gstart, int64_t* gkey, int32_t* ngroups, int64_t* ix, int64_t* jx,
                        void* ws, cudaStream_t s, int hint_a, int hint_b) {
  using P = PlanSmall<NT, IPT>;
  static devo::SmemConfig configured;
  const int smem = (int)sizeof(typename P::Temp);
  if (configured.need((size_t)smem)) DEVO_CUDA(cudaFuncSetAttribute(plan_small_kernel<NT, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int64_t* sorted_a = (int64_t*)ws;
  int32_t* perm_ws = (int32_t*)((char*)ws + align_up((size_t)E * 8));
  plan_small_kernel<NT, IPT><<<1, NT, smem, s>>>(ka, kb, E, perm, gid, gstart, gkey, ngroups, ix, jx,
                                                 sorted_a, perm_ws, hint_a, hint_b);
  DEVO_LAUNCH_CHECK("graph_plan(small)");
  return DEVO_OK;
}

}  // namespace

extern "C" {

size_t devo_graph_plan_workspace(int E) {
  if (E <= 0) return 256;
  if (E <= kSmallMax) return align_up((size_t)E * 8) + align_up((size_t)E * 4);
  return large_layout(E).total;
}

int devo_graph_plan(const int64_t* ka, const int64_t* kb, int E, int64_t max_ka, int64_t max_kb,
                    int32_t* perm, int32_t* gid, int32_t* gstart, int64_t* gkey, int32_t* ngroups,
                    int64_t* ix, int64_t* jx, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(E >= 0, DEVO_EINVAL, "graph_plan: E < 0");
  if (E == 0) {
    if (ngroups) DEVO_CUDA(cudaMemsetAsync(ngroups, 0, 4, s));
    if (gstart) DEVO_CUDA(cudaMemsetAsync(gstart, 0, 4, s));
    return DEVO_OK;
  }
  DEVO_REQUIRE(workspace && workspace_bytes >= devo_graph_plan_workspace(E), DEVO_EWORKSPACE,
               "graph_plan: workspace too small (%zu < %zu)", workspace_bytes, devo_graph_plan_workspace(E));
  // both exclusive bounds given: the number of significant key bits is known here and the kernel skips its max-reductions
  int hint_a = -1, hint_b = -1;
  if (max_ka > 0 && max_kb > 0) {
    auto nbits = [](int64_t bound) { int b = 1; while (b < 62 && ((int64_t)1 << b) < bound) b++; return b; };
    hint_a = nbits(max_ka); hint_b = nbits(max_kb);
    if (hint_a + hint_b > 62) hint_a = hint_b = -1;
  }
  if (E <= 2048) return launch_small<512, 4>(ka, kb, E, perm, gid, gstart, gkey, ngroups, ix, jx, workspace, s, hint_a, hint_b);
  if (E <= 8192) return launch_small<1024, 8>(ka, kb, E, perm, gid, gstart, gkey, ngroups, ix, jx, workspace, s, hint_a, hint_b);
  if (E <= kSmallMax) return launch_small<512, 32>(ka, kb, E, perm, gid, gstart, gkey, ngroups, ix, jx, workspace, s, hint_a, hint_b);

  // large path: ids are assumed < 2^32 (they index tensors)
  LargeLayout L = large_layout(E);
  char* w = (char*)workspace;
  auto* keys_in = (unsigned long long*)(w + L.keys_in);
  auto* keys_out = (unsigned long long*)(w + L.keys_out);
  int* vals_in = (int*)(w + L.vals_in);
  int* vals_out = perm ? perm : (int*)(w + L.vals_out);
  int* flag = (int*)(w + L.flag);
  int* incl = (int*)(w + L.incl);
  const int nb = devo::cdiv(E, 256);
  plan_make_keys<<<nb, 256, 0, s>>>(ka, kb, E, keys_in, vals_in);
  DEVO_LAUNCH_CHECK("graph_plan(keys)");
  int end_bit = 64;
  if (max_ka > 0) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < max_ka) b++;
    end_bit = 32 + b;
  }
  size_t tb = L.cub_bytes;
  DEVO_CUDA(cub::DeviceRadixSort::SortPairs(w + L.cub_temp, tb, keys_in, keys_out, vals_in, vals_out, E, 0, end_bit, s));
  devo::count_launch(4);
  plan_heads<<<nb, 256, 0, s>>>(keys_out, E, flag);
  DEVO_LAUNCH_CHECK("graph_plan(heads)");
  tb = L.cub_bytes;
  DEVO_CUDA(cub::DeviceScan::InclusiveSum(w + L.cub_temp, tb, flag, incl, E, s));
  devo::count_launch(2);
  plan_finalize<<<nb, 256, 0, s>>>(keys_out, vals_out, incl, E, gid, gstart, gkey, ngroups, ix, jx);
  DEVO_LAUNCH_CHECK("graph_plan(finalize)");
  return DEVO_OK;
}

int devo_neighbors(const int64_t* ii, const int64_t* jj, int64_t* ix, int64_t* jx, int E,
                   void* workspace, size_t workspace_bytes, void* stream) {
  return devo_graph_plan(ii, jj, E, -1, -1, nullptr, nullptr, nullptr, nullptr, nullptr, ix, jx,
                         workspace, workspace_bytes, stream);
}

}  // extern "C"
