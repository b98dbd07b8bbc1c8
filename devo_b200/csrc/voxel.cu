// voxel.cu -- the front of the frame path: event stream -> voxel grid and voxel-grid normalisation.
//
//   devo_events_to_voxel   utils/event_utils.py:180-231 (to_voxel_grid): every event is spread over the 8 neighbouring
//                          (bin, y, x) cells with trilinear weights polarity * (1-|dx|) * (1-|dy|) * (1-|dt|), t rescaled
//                          to [0, bins-1] in float64 like the reference.  One thread per event, 8 float atomics (the
//                          reference: 8 masked index_add_ passes over the whole event list on the CPU).
//   devo_voxel_normalize   utils/voxel_utils.py:6-52 (std / rescale, training) and devo/devo.py:419-452 (inference):
//                          standardisation of the NON-ZERO entries (mean / std over non-zeros, zeros stay zero) or rescaling
//                          of positive / negative entries by their extrema.  Two launches: per-block partial statistics
//                          (fixed order => deterministic), then every block folds the partials and applies.  HBM-bound:
//                          2 reads + 1 write of the grid (the reference: ~12 ATen launches, 2 boolean-mask temporaries).
#include "common.cuh"

namespace {

constexpr int kStatBlocks = 256;     // partial-statistics blocks per group
constexpr int kStatThreads = 256;

struct Stats { float nnz, sum, sumsq, maxpos, minneg; };

__device__ __forceinline__ Stats combine(Stats a, Stats b) {
  Stats r;
  r.nnz = a.nnz + b.nnz; r.sum = a.sum + b.sum; r.sumsq = a.sumsq + b.sumsq;
  r.maxpos = fmaxf(a.maxpos, b.maxpos); r.minneg = fminf(a.minneg, b.minneg);
  return r;
}
__device__ __forceinline__ Stats warp_reduce(Stats s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stats t;
    t.nnz = __shfl_xor_sync(0xffffffffu, s.nnz, o); t.sum = __shfl_xor_sync(0xffffffffu, s.sum, o);
    t.sumsq = __shfl_xor_sync(0xffffffffu, s.sumsq, o); t.maxpos = __shfl_xor_sync(0xffffffffu, s.maxpos, o);
    t.minneg = __shfl_xor_sync(0xffffffffu, s.minneg, o);
    s = combine(s, t);
  }
  return s;
}

// grid = (kStatBlocks, groups): block b of group g scans its contiguous slice of the group's `n` elements
__global__ void __launch_bounds__(kStatThreads) voxel_stats_kernel(const float* __restrict__ x, long long n, float* __restrict__ partials) {
  __shared__ Stats sh[kStatThreads / 32];
  const float* xg = x + (size_t)blockIdx.y * n;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per, hi = min(n, lo + per);
  Stats s = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = lo + threadIdx.x; i < hi; i += kStatThreads) {
    const float v = xg[i];
    if (v != 0.f) { s.nnz += 1.f; s.sum += v; s.sumsq += v * v; }
    s.maxpos = fmaxf(s.maxpos, v);
    s.minneg = fminf(s.minneg, v);
  }
  s = warp_reduce(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    Stats t = sh[0];
    for (int w = 1; w < kStatThreads / 32; w++) t = combine(t, sh[w]);
    float* p = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 5;
    p[0] = t.nnz; p[1] = t.sum; p[2] = t.sumsq; p[3] = t.maxpos; p[4] = t.minneg;
  }
}

// mode 0: std over non-zeros (only if EVERY group has a non-zero, voxel_utils.py:18);  mode 1: rescale pos / neg
__global__ void __launch_bounds__(kStatThreads) voxel_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                                   const float* __restrict__ partials, int nblocks_stats, int groups,
                                                                   int mode, float* __restrict__ stats_out) {
  __shared__ float s_par[4];
  __shared__ int s_all;
  if (threadIdx.x == 0) {
    // fold the partials of this group in block order (every block does the same small sum: deterministic, no extra launch)
    int all_nonempty = 1;
    Stats mine = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < groups; g++) {
      Stats t = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int b = 0; b < nblocks_stats; b++) {
        const float* p = partials + ((size_t)g * nblocks_stats + b) * 5;
        Stats q = {p[0], p[1], p[2], p[3], p[4]};
        t = combine(t, q);
      }
      if (t.nnz <= 0.f) all_nonempty = 0;
      if (g == (int)blockIdx.y) mine = t;
    }
    s_all = all_nonempty;
    if (mode == 0) {
      const float mean = mine.sum / mine.nnz;
      s_par[0] = mean;
      s_par[1] = sqrtf(mine.sumsq / mine.nnz - mean * mean);
    } else {
      s_par[2] = mine.maxpos > 0.f ? mine.maxpos : 1e-5f;       // voxel_utils.py:41-42: 1e-5 when a polarity is absent
      s_par[3] = mine.minneg < 0.f ? mine.minneg : 1e-5f;
    }
    if (stats_out && blockIdx.x == 0) {
      float* o = stats_out + (size_t)blockIdx.y * 5;
      o[0] = mine.nnz; o[1] = mine.sum; o[2] = mine.sumsq; o[3] = mine.maxpos; o[4] = mine.minneg;
    }
  }
  __syncthreads();
  const float* xg = x + (size_t)blockIdx.y * n;
  float* yg = y + (size_t)blockIdx.y * n;
  const bool active = (mode != 0) || s_all;
  const float mean = s_par[0], stdv = s_par[1], mx = s_par[2], mn = s_par[3];
  for (long long i = (long long)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kStatThreads) {
    const float v = xg[i];
    float o = v;
    if (active) {
      if (mode == 0) o = (v != 0.f) ? (v - mean) / stdv : 0.f * ((v - mean) / stdv);   // mask * (x - mean) / std
      else o = v > 0.f ? v / mx : (v < 0.f ? v / -mn : v);
    }
    yg[i] = o;
  }
}

__global__ void events_to_voxel_kernel(const float* __restrict__ xs, const float* __restrict__ ys, const double* __restrict__ ts,
                                       const float* __restrict__ ps, long long n, float* __restrict__ grid, int B, int H, int W) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double t0 = ts[0], dur = ts[n - 1] - ts[0];
  const double t = (ts[e] - t0) * (double)(B - 1) / dur;
  const float x = xs[e], y = ys[e];
  const float pol = ps[e] == 0.f ? -1.f : ps[e];             // polarity 0 means negative (event_utils.py:198-199)
  const float lx = floorf(x), ly = floorf(y);
  const double lt = floor(t);
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const float cx = lx + (float)(c & 1), cy = ly + (float)((c >> 1) & 1);
    const double ct = lt + (double)(c >> 2);
    if (cx < 0.f || cy < 0.f || ct < 0.0 || cx > (float)(W - 1) || cy > (float)(H - 1) || ct > (double)(B - 1)) continue;
    // the reference stacks (x, y, t, p) into ONE float64 array (event_utils.py:204), so the weight is a float64 product
    const double w = (double)pol * (1.0 - fabs((double)cx - (double)x)) * (1.0 - fabs((double)cy - (double)y)) * (1.0 - fabs(ct - t));
    atomicAdd(&grid[((size_t)ct * H + (size_t)cy) * W + (size_t)cx], (float)w);
  }
}

}  // namespace

extern "C" {

size_t devo_voxel_workspace(int groups) { return (size_t)(groups > 0 ? groups : 1) * kStatBlocks * 5 * sizeof(float); }

int devo_voxel_normalize(const float* x, float* y, long long n_per_group, int groups, int mode, float* stats_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  DEVO_REQUIRE(x && y && n_per_group >= 0 && groups >= 1 && (mode == 0 || mode == 1), DEVO_EINVAL, "voxel_normalize: bad arguments");
  DEVO_REQUIRE(workspace && workspace_bytes >= devo_voxel_workspace(groups), DEVO_EWORKSPACE, "voxel_normalize: workspace too small");
  if (n_per_group == 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  voxel_stats_kernel<<<dim3(kStatBlocks, groups), kStatThreads, 0, s>>>(x, n_per_group, (float*)workspace);
  DEVO_LAUNCH_CHECK("voxel_stats");
  const int blocks = (int)((n_per_group + 4 * kStatThreads - 1) / (4 * kStatThreads));
  voxel_apply_kernel<<<dim3(blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks), groups), kStatThreads, 0, s>>>(
      x, y, n_per_group, (const float*)workspace, kStatBlocks, groups, mode, stats_out);
  DEVO_LAUNCH_CHECK("voxel_apply");
  return DEVO_OK;
}

int devo_events_to_voxel(const float* xs, const float* ys, const double* ts, const float* ps, long long n_events,
                         float* grid, int bins, int H, int W, void* stream) {
  DEVO_REQUIRE(grid && bins >= 2 && H > 0 && W > 0 && n_events >= 0, DEVO_EINVAL, "events_to_voxel: bad arguments");
  if (n_events == 0) return DEVO_OK;
  DEVO_REQUIRE(xs && ys && ts && ps, DEVO_EINVAL, "events_to_voxel: NULL event arrays");
  events_to_voxel_kernel<<<(unsigned)((n_events + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xs, ys, ts, ps, n_events, grid, bins, H, W);
  DEVO_LAUNCH_CHECK("events_to_voxel");
  return DEVO_OK;
}

}  // extern "C"
