// corr_bwd_pm.cu -- backward of the patch correlation lookup on PIXEL-MAJOR float32 buffers (training path).
//
// What it replaces: corr_backward_kernel of the reference (devo/altcorr/correlation_kernel.cu:139-190, launched from
// :236-286): one thread per (edge, window position, patch pixel) issuing 2*C scalar atomics into planar [C][H][W] volumes
// -- every atomic its own 32-byte sector.  The generic kernel of this library (csrc/corr.cu) pre-sums the nine patch pixels
// per bounding-box pixel but keeps the planar layout, so its C * 121 atomics per edge are still one sector each; at the
// training shape (E = 21600, C = 128, two levels) that was 9.7 ms of a 26 ms iteration.
//
// Here both feature volumes are pixel-major ([frame][y][x][C], as the forward lookup reads them) and one WARP owns one
// pixel of the edge's bounding box with lane <-> C/32 consecutive channels:
//   * the frame pixel is ONE coalesced C*4-byte load, its gradient ONE vector reduction (red.global.add.v4.f32) per lane:
//     C/4 16-byte reductions per box pixel instead of C scalar ones, each sector touched once;
//   * the patch features live in registers (9 x C/32 per lane), so does the patch-feature gradient, which is summed over
//     the box pixels in registers, over the CTA's warps in shared memory and leaves as 9*C/4 vector reductions per edge;
//   * which of the nine windows cover a box pixel, and with which weight, is tabulated once per edge in shared memory.
// The transposed bilinear blend (grad of the 7x7 outputs -> grad of the 8x8 window values) is the same arithmetic as in
// csrc/corr.cu.  Host side (cuda_corr.backward): pixel-major copies in / planar copies out are layout plumbing.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int R = 3, D = 2 * R + 2, Dm = D - 1, PP = 9, DD = D * D;
constexpr int kMaxBoxArea = 20 * 20;      // beyond that (patch pixels far apart) every patch pixel walks its own window

template <int VEC> struct Vec;
template <> struct Vec<4> { using type = float4; };
template <> struct Vec<2> { using type = float2; };

__device__ __forceinline__ void red_add(float* p, const float (&v)[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void red_add(float* p, const float (&v)[2]) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}
// bulk reduction shared -> global (TMA unit, `cp.reduce.async.bulk ... add.f32`): ONE instruction adds a pixel's C floats
__device__ __forceinline__ void bulk_red_add(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void store_vec(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store_vec(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
constexpr int kSlots = 4;                 // per-warp ring of staged pixels for the bulk reductions

__device__ __forceinline__ void load_vec(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_vec(const float* p, float (&v)[2]) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  v[0] = t.x; v[1] = t.y;
}

// grid = E.  f1: planar [Np][C][9]; f2pm: [Nf][H][W][C]; grad: [E][7 (x-off)][7 (y-off)][9]; g1pm: [Np][9][C]; g2pm like f2pm.
template <int VEC, bool BULK>
__global__ void __launch_bounds__(kThreads) corr_backward_pm_kernel(
    const float* __restrict__ f1, const float* __restrict__ f2pm, const float* __restrict__ coords,
    const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, const float* __restrict__ grad,
    float* __restrict__ g1pm, float* __restrict__ g2pm, int Np, int Nf, int H, int W) {
  constexpr int C = 32 * VEC;
  __shared__ float gV[PP * DD];            // [p][a (row)][b (col)]: gradient of the 8x8 window values
  __shared__ float f1s[C * PP];            // planar copy of the patch features, later reused as the g1 accumulator [p][C]
  __shared__ __align__(16) float wtab[kMaxBoxArea * 12];   // per box pixel: nine window weights + a "covered" flag
  __shared__ __align__(16) float ring[BULK ? kWarps * kSlots * C : 4];   // BULK: staged pixel gradients (source of the bulk reductions)
  __shared__ int gx[PP], gy[PP];
  __shared__ float fdx[PP], fdy[PP];
  __shared__ int s_box[4];

  const int e = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ix = ii[e], jx = jj[e];
  if (ix < 0 || ix >= Np || jx < 0 || jx >= Nf) return;
  const float* co = coords + (size_t)e * 2 * PP;
  const float* g = grad + (size_t)e * Dm * Dm * PP;

  for (int q = tid; q < C * PP; q += kThreads) f1s[q] = f1[(size_t)ix * C * PP + q];
  if (tid < PP) {
    const float x = co[tid], y = co[PP + tid];
    // (saturated far outside any image, so that box arithmetic cannot overflow; such a window is out of bounds anyway)
    const float flx = fminf(fmaxf(floorf(x), -1.0e6f), 1.0e6f), fly = fminf(fmaxf(floorf(y), -1.0e6f), 1.0e6f);
    gx[tid] = (int)flx; gy[tid] = (int)fly;
    fdx[tid] = x - floorf(x); fdy[tid] = y - floorf(y);
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
  // transposed bilinear blend (correlation_kernel.cu:252-269): window value (a, b) fed up to four outputs
  for (int q = tid; q < PP * DD; q += kThreads) {
    const int bb = q % D, a = (q / D) % D, p = q / DD;
    const float dx = fdx[p], dy = fdy[p];
    float s = 0.f;
    if (a < Dm && bb < Dm) s += (1 - dx) * (1 - dy) * g[(bb * Dm + a) * PP + p];
    if (a < Dm && bb >= 1) s += dx * (1 - dy) * g[((bb - 1) * Dm + a) * PP + p];
    if (a >= 1 && bb < Dm) s += (1 - dx) * dy * g[(bb * Dm + (a - 1)) * PP + p];
    if (a >= 1 && bb >= 1) s += dx * dy * g[((bb - 1) * Dm + (a - 1)) * PP + p];
    gV[q] = s;
  }
  __syncthreads();

  // this lane's channels of the nine patch pixels, and their gradient
  float f1r[PP][VEC], acc[PP][VEC];
#pragma unroll
  for (int p = 0; p < PP; p++)
#pragma unroll
    for (int v = 0; v < VEC; v++) { f1r[p][v] = f1s[(lane * VEC + v) * PP + p]; acc[p][v] = 0.f; }
  __syncthreads();                         // f1s is free: it becomes the CTA's g1 accumulator
  for (int q = tid; q < C * PP; q += kThreads) f1s[q] = 0.f;

  const float* f2 = f2pm + (size_t)jx * H * W * C + lane * VEC;
  float* o2 = g2pm + (size_t)jx * H * W * C + lane * VEC;
  const int x0 = s_box[0], y0 = s_box[1], bw = s_box[2], bh = s_box[3];
  if ((long long)bw * bh <= kMaxBoxArea) {
    const int npx = bw * bh;
    // weight table of the box: wtab[q] = {w_0 .. w_8, covered, -, -} for box pixel q, w_p = gV[p][.][.] where pixel q lies in
    // patch pixel p's 8x8 window, else 0; `covered` = 0 for a pixel outside the image or outside all nine windows.  Built
    // once by the whole CTA: the per-pixel loop below is 3 shared-memory loads + 72 FMAs, no index arithmetic per window.
    for (int idx = tid; idx < npx * 10; idx += kThreads) {
      const int q = idx / 10, p = idx - q * 10;
      const int i1 = y0 + q / bw, j1 = x0 + q % bw;
      const bool in_img = (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W);
      float w = 0.f;
      if (p < PP) {
        const int a = i1 - (gy[p] - R), bb = j1 - (gx[p] - R);
        if (in_img && a >= 0 && a < D && bb >= 0 && bb < D) w = gV[p * DD + a * D + bb];
      } else if (in_img) {
#pragma unroll
        for (int pp = 0; pp < PP; pp++) {
          const int a = i1 - (gy[pp] - R), bb = j1 - (gx[pp] - R);
          if (a >= 0 && a < D && bb >= 0 && bb < D) w = 1.f;
        }
      }
      wtab[q * 12 + p] = w;
    }
    __syncthreads();
    // one box pixel per warp and round; the pixels of the next two rounds are requested before this round's arithmetic
    float cur[VEC] = {}, nx1[VEC] = {}, nx2[VEC] = {};
    auto request = [&](int q, float (&dst)[VEC]) {
      if (q < npx && wtab[q * 12 + 9] != 0.f) load_vec(f2 + ((size_t)(y0 + q / bw) * W + (x0 + q % bw)) * C, dst);
    };
    request(warp, cur);
    request(warp + kWarps, nx1);
    int slot = 0;
    for (int q = warp; q < npx; q += kWarps) {
      request(q + 2 * kWarps, nx2);
      const float4* wt = reinterpret_cast<const float4*>(wtab + q * 12);
      const float4 wa = wt[0], wb = wt[1], wc = wt[2];
      if (wc.y != 0.f) {                                      // warp-uniform
        const float w[PP] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) o[v] = 0.f;
#pragma unroll
        for (int p = 0; p < PP; p++)
#pragma unroll
          for (int v = 0; v < VEC; v++) { o[v] = fmaf(w[p], f1r[p][v], o[v]); acc[p][v] = fmaf(w[p], cur[v], acc[p][v]); }
        float* gdst = o2 + ((size_t)(y0 + q / bw) * W + (x0 + q % bw)) * C;
        if constexpr (BULK) {
          // stage the pixel in this warp's ring slot; lane 0 hands the 4*C bytes to the TMA unit as one reduction.  The slot
          // was the source of the bulk group issued kSlots pixels ago: at most kSlots - 1 younger groups may still be reading.
          float* sl = ring + (warp * kSlots + slot) * C;
          if (lane == 0) bulk_wait_read<kSlots - 1>();
          __syncwarp();
          store_vec(sl + lane * VEC, o);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { bulk_red_add(gdst - lane * VEC, sl, C * 4); bulk_commit(); }
          slot = (slot + 1) % kSlots;
        } else {
          red_add(gdst, o);
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; v++) { cur[v] = nx1[v]; nx1[v] = nx2[v]; }
    }
  } else {
    // degenerate geometry: each patch pixel walks its own 8x8 window
    for (int q = warp; q < PP * DD; q += kWarps) {
      const int bb = q % D, a = (q / D) % D, p = q / DD;
      const int i1 = gy[p] + a - R, j1 = gx[p] + bb - R;
      if (i1 < 0 || i1 >= H || j1 < 0 || j1 >= W) continue;
      const float w = gV[q];
      float cur[VEC], o[VEC];
      load_vec(f2 + ((size_t)i1 * W + j1) * C, cur);
#pragma unroll
      for (int pp = 0; pp < PP; pp++)
        if (pp == p) {
#pragma unroll
          for (int v = 0; v < VEC; v++) { o[v] = w * f1r[pp][v]; acc[pp][v] = fmaf(w, cur[v], acc[pp][v]); }
        }
      red_add(o2 + ((size_t)i1 * W + j1) * C, o);
    }
  }
  __syncthreads();                         // the zeroing of the accumulator is complete
#pragma unroll
  for (int p = 0; p < PP; p++)
#pragma unroll
    for (int v = 0; v < VEC; v++) atomicAdd(&f1s[p * C + lane * VEC + v], acc[p][v]);
  __syncthreads();
  float* o1 = g1pm + (size_t)ix * PP * C;
  if constexpr (BULK) {
    // the CTA's [9][C] patch-feature gradient is contiguous in shared memory and in g1pm: one bulk reduction
    fence_async_smem();
    __syncthreads();
    if (tid == 0) { bulk_red_add(o1, f1s, PP * C * 4); bulk_commit(); }
    if (lane == 0) bulk_wait_read<0>();       // every issuing lane: its sources must outlive the reads
    return;
  }
  for (int q = tid; q < PP * 32; q += kThreads) {
    float o[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) o[v] = f1s[q * VEC + v];
    red_add(o1 + q * VEC, o);
  }
}

}  // namespace

extern "C" int devo_corr_backward_pm(const float* fmap1, const float* fmap2_pm, const float* coords, const int64_t* ii,
                                     const int64_t* jj, const float* grad, float* fmap1_grad_pm, float* fmap2_grad_pm,
                                     int Np, int Nf, int C, int H, int W, int E, void* stream) {
  DEVO_REQUIRE(E >= 0 && Np >= 0 && Nf >= 0 && H > 0 && W > 0, DEVO_EINVAL, "corr_backward_pm: bad sizes");
  DEVO_REQUIRE(C == 64 || C == 128, DEVO_EUNSUPPORTED, "corr_backward_pm: C must be 64 or 128 (got %d)", C);
  DEVO_REQUIRE((((uintptr_t)fmap2_pm | (uintptr_t)fmap1_grad_pm | (uintptr_t)fmap2_grad_pm) & 15) == 0, DEVO_EINVAL,
               "corr_backward_pm: pixel-major buffers must be 16-byte aligned");
  if (E == 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  static const int bulk = [] { const char* e = getenv("DEVO_CORR_BWD_BULK"); return e ? atoi(e) : 0; }();
#define BWD(VEC, BULK) corr_backward_pm_kernel<VEC, BULK><<<E, kThreads, 0, s>>>(fmap1, fmap2_pm, coords, ii, jj, grad, fmap1_grad_pm, fmap2_grad_pm, Np, Nf, H, W)
  if (C == 128) { if (bulk) BWD(4, true); else BWD(4, false); }
  else { if (bulk) BWD(2, true); else BWD(2, false); }
#undef BWD
  DEVO_LAUNCH_CHECK("corr_backward_pm");
  return DEVO_OK;
}
