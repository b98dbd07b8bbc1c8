// patch_gather.cu -- the tail of the reference's Patchifier (devo/enet.py:179-191) as ONE launch.
//
// Reference: per arriving frame three `altcorr.patchify` calls (a gather kernel + ~12 ATen launches each for the bilinear
// blend of four shifted windows: imap at radius 0, gmap at radius P/2, the (x, y, inverse depth) grid at radius P/2, after
// building that grid with `coords_grid_with_index`), then the engine re-packs gmap to pixel-major for the lookup.  Here
// one kernel evaluates the same bilinear blend (float arithmetic: the reference multiplies the half windows by float32
// offsets, i.e. it blends in float32 too) for all three and writes
//   gmap  planar [N*M, C, P, P] (the reference layout, optional) and / or pixel-major [N*M, P*P, C] (what
//         devo_corr_lookup_fused reads: no separate gmap_pack),
//   imap  [N*M, D],
//   patches [N*M, 3, P, P] float32 = (x, y, disps[y, x]) at the window positions -- the grid is never materialised.
// Out-of-image taps contribute zero, like the reference's zero-initialised patch buffer.
#include "common.cuh"

namespace {
using devo::ElemTraits;

template <typename T>
__device__ __forceinline__ float tap(const T* __restrict__ plane, int H, int W, int i, int j) {
  return (i >= 0 && i < H && j >= 0 && j < W) ? ElemTraits<T>::to_float(plane[(size_t)i * W + j]) : 0.f;
}

// one CTA per patch; threads sweep channels (coalescing is along the channel-strided planar input either way: each
// value is a separate 2/4-byte gather, the patch pixels of one channel share at most three 32-byte sectors)
template <typename T>
__global__ void __launch_bounds__(256) patch_gather_kernel(const T* __restrict__ fmap, const T* __restrict__ imap,
                                                           const float* __restrict__ disps, const float* __restrict__ coords,
                                                           T* __restrict__ gmap_planar, T* __restrict__ gmap_pm,
                                                           T* __restrict__ imap_out, float* __restrict__ patches, int C, int D,
                                                           int H, int W, int M, int P) {
  const int m = blockIdx.x, n = blockIdx.y;
  const size_t pid = (size_t)n * M + m;
  const float x = coords[pid * 2 + 0], y = coords[pid * 2 + 1];
  const float fx = floorf(x), fy = floorf(y);
  const int j0 = (int)fx, i0 = (int)fy;
  const float dx = x - fx, dy = y - fy;
  const float w00 = (1.f - dy) * (1.f - dx), w01 = (1.f - dy) * dx, w10 = dy * (1.f - dx), w11 = dy * dx;
  const int R = P / 2, PP = P * P;
  const size_t HW = (size_t)H * W;
  // gmap: C x P x P
  if (fmap != nullptr) {
    const T* src = fmap + (size_t)n * C * HW;
    for (int q = threadIdx.x; q < C * PP; q += blockDim.x) {
      const int c = q / PP, p = q - c * PP;
      const int a = p / P, b = p - a * P;                 // window row / column
      const int i = i0 + a - R, j = j0 + b - R;
      const T* pl = src + (size_t)c * HW;
      const float v = w00 * tap(pl, H, W, i, j) + w01 * tap(pl, H, W, i, j + 1) + w10 * tap(pl, H, W, i + 1, j) +
                      w11 * tap(pl, H, W, i + 1, j + 1);
      const T o = ElemTraits<T>::from_float(v);
      if (gmap_planar != nullptr) gmap_planar[pid * C * PP + q] = o;
      if (gmap_pm != nullptr) gmap_pm[(pid * PP + p) * C + c] = o;
    }
  }
  // imap: D x 1 x 1
  if (imap != nullptr) {
    const T* src = imap + (size_t)n * D * HW;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const T* pl = src + (size_t)c * HW;
      const float v = w00 * tap(pl, H, W, i0, j0) + w01 * tap(pl, H, W, i0, j0 + 1) + w10 * tap(pl, H, W, i0 + 1, j0) +
                      w11 * tap(pl, H, W, i0 + 1, j0 + 1);
      imap_out[pid * D + c] = ElemTraits<T>::from_float(v);
    }
  }
  // patches: (x, y, inverse depth) of the window positions, blended like any other channel of the reference's grid
  if (patches != nullptr && threadIdx.x < 3 * PP) {
    const int ch = threadIdx.x / PP, p = threadIdx.x - ch * PP;
    const int a = p / P, b = p - a * P;
    const int i = i0 + a - R, j = j0 + b - R;
    auto grid = [&](int ii, int jj) -> float {
      if (ii < 0 || ii >= H || jj < 0 || jj >= W) return 0.f;
      if (ch == 0) return (float)jj;
      if (ch == 1) return (float)ii;
      return disps != nullptr ? disps[(size_t)n * HW + (size_t)ii * W + jj] : 1.0f;
    };
    patches[(pid * 3 + ch) * PP + p] = w00 * grid(i, j) + w01 * grid(i, j + 1) + w10 * grid(i + 1, j) + w11 * grid(i + 1, j + 1);
  }
}
}  // namespace

extern "C" int devo_patch_gather(const void* fmap, const void* imap, const float* disps, const float* coords,
                                 void* gmap_planar, void* gmap_pm, void* imap_out, float* patches, int dtype, int N, int C,
                                 int D, int H, int W, int M, int P, void* stream) {
  DEVO_REQUIRE(coords != nullptr, DEVO_EINVAL, "patch_gather: coords is NULL");
  DEVO_REQUIRE(P >= 1 && P <= 7 && (P & 1), DEVO_EINVAL, "patch_gather: patch size %d unsupported (odd, <= 7)", P);
  DEVO_REQUIRE(3 * P * P <= 256, DEVO_EINVAL, "patch_gather: patch too large");
  DEVO_REQUIRE(N >= 0 && M >= 0 && H > 0 && W > 0, DEVO_EINVAL, "patch_gather: bad sizes");
  DEVO_REQUIRE(N <= 65535, DEVO_EINVAL, "patch_gather: too many frames");
  DEVO_REQUIRE(fmap == nullptr || gmap_planar != nullptr || gmap_pm != nullptr, DEVO_EINVAL, "patch_gather: fmap without a gmap output");
  DEVO_REQUIRE(imap == nullptr || imap_out != nullptr, DEVO_EINVAL, "patch_gather: imap without an output");
  if (N == 0 || M == 0) return DEVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(M, N);
#define PG(T) (patch_gather_kernel<T><<<grid, 256, 0, s>>>((const T*)fmap, (const T*)imap, disps, coords, (T*)gmap_planar, \
                                                           (T*)gmap_pm, (T*)imap_out, patches, C, D, H, W, M, P))
  switch (dtype) {
    case DEVO_F16: PG(__half); break;
    case DEVO_BF16: PG(__nv_bfloat16); break;
    case DEVO_F32: PG(float); break;
    default: DEVO_REQUIRE(false, DEVO_EINVAL, "patch_gather: unsupported dtype %d", dtype);
  }
#undef PG
  DEVO_LAUNCH_CHECK("patch_gather");
  return DEVO_OK;
}
