// transform_grad.cu -- backward of the fused projective transform (devo/projective_ops.py:53-105) for training:
// gradients of coords [E,P,P,2] AND of the centre-pixel Jacobians Ji, Jj [E,2,6], Jz [E,2] with respect to the poses
// (lietorch's convention: the gradient of a group element is the gradient with respect to a LEFT tangent perturbation
// G <- Exp(xi) G, stored in the first 6 of the 7 slots), the patches (x, y, inverse depth of every pixel) -- one launch
// instead of the ~60-node autograd graph the composed path records (lietorch Inv / Mul / Act4 / AdjT backward kernels plus
// ~40 ATen element-wise nodes), called 6-8 times per training iteration (enet.py:341,363-372) and twice inside ba.BA.
//
//   patch gradients of coords : analytic, every pixel:  d pi / d X1 (with the clamp(Z, 0.1) of proj, :41) times [R | t]
//   everything else           : the forward expressions (Gj * Gi^-1, act4, proj; Ji, Jj, Jz) are re-evaluated on DUAL NUMBERS
//                   (value + one derivative), once per input direction (6 + 6 left pose tangents; x, y, d of the centre
//                   pixel for the Jacobian outputs); lie.cuh is scalar-generic, so the derivative runs through the very
//                   code the forward kernel instantiates with float and cannot drift from it (normalisations, the
//                   translation-only variant, Adj conventions included).  15 evaluations x ~1 kFLOP per edge: negligible.
// Accumulation into the per-pose / per-patch gradients uses float atomics (as ATen's index_add in the composed path does).
#include "common.cuh"
#include "lie.cuh"

namespace {

// ---- first-order dual number -------------------------------------------------------------------------------------
struct Dual {
  float v, d;
  __host__ __device__ Dual() : v(0.f), d(0.f) {}
  __host__ __device__ Dual(float a) : v(a), d(0.f) {}
  __host__ __device__ Dual(double a) : v((float)a), d(0.f) {}
  __host__ __device__ Dual(int a) : v((float)a), d(0.f) {}
  __host__ __device__ Dual(float a, float b) : v(a), d(b) {}
};
#define DHD __host__ __device__ __forceinline__
DHD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
DHD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
DHD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
DHD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
DHD Dual operator/(Dual a, Dual b) { const float r = 1.0f / b.v; return Dual(a.v * r, (a.d - a.v * r * b.d) * r); }
DHD Dual operator*(int a, Dual b) { return Dual((float)a * b.v, (float)a * b.d); }
DHD Dual operator*(float a, Dual b) { return Dual(a * b.v, a * b.d); }
DHD Dual operator*(Dual b, float a) { return Dual(a * b.v, a * b.d); }
DHD Dual operator-(int a, Dual b) { return Dual((float)a - b.v, -b.d); }
DHD Dual operator+(Dual a, float b) { return Dual(a.v + b, a.d); }
DHD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
DHD Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
DHD Dual& operator/=(Dual& a, Dual b) { a = a / b; return a; }
DHD Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
DHD bool operator<(Dual a, Dual b) { return a.v < b.v; }
DHD bool operator>(Dual a, Dual b) { return a.v > b.v; }
DHD Dual sqrt(Dual a) { const float s = sqrtf(a.v); return Dual(s, 0.5f * a.d / s); }
DHD Dual fabs(Dual a) { return a.v < 0.f ? -a : a; }
DHD float val(float a) { return a; }
DHD float val(Dual a) { return a.v; }

// ---- the forward expressions of the centre-pixel Jacobians, scalar-generic ---------------------------------------------
// out: Jj[12] (row-major 2x6), Ji[12], Jz[2]   -- the same formulas as transform_kernel (csrc/ba.cu), projective_ops.py:73-100
template <typename T>
__device__ __forceinline__ void centre_jacobians(const lie::SE3<T>& Gi, const lie::SE3<T>& Gj, T px, T py, T pd,
                                                 const float* Ki, const float* Kj, int tonly, T* out) {
  lie::SE3<T> Gij = Gj * Gi.inv();
  if (tonly) Gij.R.q = lie::Quat<T>{T(0.f), T(0.f), T(0.f), T(1.f)};
  T X0[4] = {(px - T(Ki[2])) / T(Ki[0]), (py - T(Ki[3])) / T(Ki[1]), T(1.0f), pd};
  T X1[4];
  Gij.act4(X0, X1);
  const T fx = T(Kj[0]), fy = T(Kj[1]);
  const T X = X1[0], Y = X1[1], Z = X1[2], H = X1[3];
  const bool far = fabsf(val(Z)) > 0.2f;
  const T d = far ? T(1.0f) / Z : T(0.0f);
  const T a0 = fx * d, a2 = -(fx * X * d * d), b1 = fy * d, b2 = -(fy * Y * d * d);
  T J[2][6];
  J[0][0] = a0 * H; J[0][1] = T(0.f); J[0][2] = a2 * H; J[0][3] = a2 * Y;             J[0][4] = a0 * Z - a2 * X; J[0][5] = -(a0 * Y);
  J[1][0] = T(0.f); J[1][1] = b1 * H; J[1][2] = b2 * H; J[1][3] = b2 * Y - b1 * Z;    J[1][4] = -(b2 * X);       J[1][5] = b1 * X;
  lie::Mat<T, 6, 6> A = Gij.Adj();
#pragma unroll
  for (int r = 0; r < 2; r++) {
    T o[6];
    lie::matTvec(A, J[r], o);
#pragma unroll
    for (int c = 0; c < 6; c++) { out[r * 6 + c] = J[r][c]; out[12 + r * 6 + c] = -o[c]; }
  }
  out[24] = a0 * Gij.t[0] + a2 * Gij.t[2];
  out[25] = b1 * Gij.t[1] + b2 * Gij.t[2];
}

// left perturbation of a pose by the k-th tangent direction, to first order: Exp(eps e_k) G
__device__ __forceinline__ lie::SE3<Dual> perturbed(const float* P, int k) {
  Dual d[7];
#pragma unroll
  for (int c = 0; c < 7; c++) d[c] = Dual(P[c]);
  lie::SE3<Dual> G = lie::SE3<Dual>::load(d);
  if (k < 0) return G;
  lie::SE3<Dual> dG;                    // Exp(eps e_k) = (t = eps tau, q = (eps phi / 2, 1)) + O(eps^2)
  dG.t[0] = Dual(0.f, k == 0 ? 1.f : 0.f); dG.t[1] = Dual(0.f, k == 1 ? 1.f : 0.f); dG.t[2] = Dual(0.f, k == 2 ? 1.f : 0.f);
  dG.R.q = lie::Quat<Dual>{Dual(0.f, k == 3 ? 0.5f : 0.f), Dual(0.f, k == 4 ? 0.5f : 0.f), Dual(0.f, k == 5 ? 0.5f : 0.f), Dual(1.f, 0.f)};
  return dG * G;
}

__global__ void transform_backward_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                                          const float* __restrict__ intrinsics, const int64_t* __restrict__ ii,
                                          const int64_t* __restrict__ jj, const int64_t* __restrict__ kk,
                                          const float* __restrict__ g_coords, const float* __restrict__ g_Ji,
                                          const float* __restrict__ g_Jj, const float* __restrict__ g_Jz,
                                          float* __restrict__ grad_poses, float* __restrict__ grad_patches,
                                          int E, int P, int layout, int tonly) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= E) return;
  const int PP = P * P;
  const int i = (int)ii[n], j = (int)jj[n], k = (int)kk[n];
  float Pi[7], Pj[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { Pi[c] = poses[(size_t)i * 7 + c]; Pj[c] = poses[(size_t)j * 7 + c]; }
  const float* Ki = intrinsics + (size_t)i * 4;
  const float* Kj = intrinsics + (size_t)j * 4;
  const float* pk = patches + (size_t)k * 3 * PP;
  float gi6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gj6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // ---- patch gradients of the coords (every pixel, analytic): d pi / d X1 (with proj's clamp(Z, 0.1), :41) * [R | t]
  if (g_coords != nullptr) {
    lie::SE3<float> Gi = lie::SE3<float>::load(Pi), Gj = lie::SE3<float>::load(Pj);
    lie::SE3<float> Gij = Gj * Gi.inv();
    if (tonly) Gij.R.q = lie::Quat<float>{0.f, 0.f, 0.f, 1.f};
    const lie::Mat<float, 3, 3> R = Gij.R.q.matrix();
    const float fx = Kj[0], fy = Kj[1];
    for (int p = 0; p < PP; p++) {
      float gx, gy;
      if (layout == 0) { gx = g_coords[((size_t)n * PP + p) * 2]; gy = g_coords[((size_t)n * PP + p) * 2 + 1]; }
      else { gx = g_coords[(size_t)n * 2 * PP + p]; gy = g_coords[(size_t)n * 2 * PP + PP + p]; }
      if (gx == 0.f && gy == 0.f) continue;
      const float X0[4] = {(pk[p] - Ki[2]) / Ki[0], (pk[PP + p] - Ki[3]) / Ki[1], 1.0f, pk[2 * PP + p]};
      float X1[4];
      Gij.act4(X0, X1);
      const bool clamped = X1[2] < 0.1f;
      const float dd = 1.0f / fmaxf(X1[2], 0.1f);
      const float gX = gx * fx * dd, gY = gy * fy * dd;
      const float gZ = clamped ? 0.f : -(gx * fx * X1[0] + gy * fy * X1[1]) * dd * dd;
      const float g0x = R(0, 0) * gX + R(1, 0) * gY + R(2, 0) * gZ;
      const float g0y = R(0, 1) * gX + R(1, 1) * gY + R(2, 1) * gZ;
      const float gd = Gij.t[0] * gX + Gij.t[1] * gY + Gij.t[2] * gZ;
      atomicAdd(&grad_patches[(size_t)k * 3 * PP + p], g0x / Ki[0]);
      atomicAdd(&grad_patches[(size_t)k * 3 * PP + PP + p], g0y / Ki[1]);
      atomicAdd(&grad_patches[(size_t)k * 3 * PP + 2 * PP + p], gd);
    }
  }

  // ---- dual-number passes: the pose gradients of the coords (all pixels) and every gradient of the Jacobian outputs
  const bool need_jac = (g_Jj != nullptr) || (g_Ji != nullptr) || (g_Jz != nullptr);
  const bool coords_dual = (g_coords != nullptr);
  if (need_jac || coords_dual) {
    const int centre = (P / 2) * P + (P / 2);
    float gout[26];
#pragma unroll
    for (int q = 0; q < 12; q++) { gout[q] = g_Jj ? g_Jj[(size_t)n * 12 + q] : 0.f; gout[12 + q] = g_Ji ? g_Ji[(size_t)n * 12 + q] : 0.f; }
    gout[24] = g_Jz ? g_Jz[(size_t)n * 2] : 0.f;
    gout[25] = g_Jz ? g_Jz[(size_t)n * 2 + 1] : 0.f;
    for (int dir = 0; dir < 15; dir++) {
      lie::SE3<Dual> Gi = perturbed(Pi, dir < 6 ? dir : -1);
      lie::SE3<Dual> Gj = perturbed(Pj, (dir >= 6 && dir < 12) ? dir - 6 : -1);
      float acc = 0.f;
      if (need_jac) {
        Dual px(pk[centre], dir == 12 ? 1.f : 0.f), py(pk[PP + centre], dir == 13 ? 1.f : 0.f), pd(pk[2 * PP + centre], dir == 14 ? 1.f : 0.f);
        Dual out[26];
        centre_jacobians<Dual>(Gi, Gj, px, py, pd, Ki, Kj, tonly, out);
#pragma unroll
        for (int q = 0; q < 26; q++) acc += gout[q] * out[q].d;
      }
      if (coords_dual && dir < 12) {
        lie::SE3<Dual> Gij = Gj * Gi.inv();
        if (tonly) Gij.R.q = lie::Quat<Dual>{Dual(0.f), Dual(0.f), Dual(0.f), Dual(1.f)};
        for (int p = 0; p < PP; p++) {
          float gx, gy;
          if (layout == 0) { gx = g_coords[((size_t)n * PP + p) * 2]; gy = g_coords[((size_t)n * PP + p) * 2 + 1]; }
          else { gx = g_coords[(size_t)n * 2 * PP + p]; gy = g_coords[(size_t)n * 2 * PP + PP + p]; }
          if (gx == 0.f && gy == 0.f) continue;
          Dual X0[4] = {Dual((pk[p] - Ki[2]) / Ki[0]), Dual((pk[PP + p] - Ki[3]) / Ki[1]), Dual(1.0f), Dual(pk[2 * PP + p])};
          Dual X1[4];
          Gij.act4(X0, X1);
          const Dual dd = (X1[2].v < 0.1f) ? Dual(10.0f) : Dual(1.0f) / X1[2];
          acc += gx * (Kj[0] * (dd * X1[0])).d + gy * (Kj[1] * (dd * X1[1])).d;
        }
      }
      if (dir < 6) gi6[dir] += acc;
      else if (dir < 12) gj6[dir - 6] += acc;
      else if (acc != 0.f) atomicAdd(&grad_patches[(size_t)k * 3 * PP + (dir - 12) * PP + centre], acc);
    }
  }
#pragma unroll
  for (int c = 0; c < 6; c++) {
    if (gi6[c] != 0.f) atomicAdd(&grad_poses[(size_t)i * 7 + c], gi6[c]);
    if (gj6[c] != 0.f) atomicAdd(&grad_poses[(size_t)j * 7 + c], gj6[c]);
  }
}

}  // namespace

extern "C" int devo_transform_backward(const float* poses, const float* patches, const float* intrinsics,
                                       const int64_t* ii, const int64_t* jj, const int64_t* kk,
                                       const float* g_coords, const float* g_Ji, const float* g_Jj, const float* g_Jz,
                                       float* grad_poses, float* grad_patches, int E, int P, int layout, int tonly,
                                       void* stream) {
  DEVO_REQUIRE(poses && patches && intrinsics && ii && jj && kk && grad_poses && grad_patches, DEVO_EINVAL,
               "transform_backward: NULL argument");
  DEVO_REQUIRE(E >= 0 && P >= 1 && P <= 7, DEVO_EINVAL, "transform_backward: bad sizes");
  if (E == 0) return DEVO_OK;
  transform_backward_kernel<<<(E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      poses, patches, intrinsics, ii, jj, kk, g_coords, g_Ji, g_Jj, g_Jz, grad_poses, grad_patches, E, P, layout, tonly);
  DEVO_LAUNCH_CHECK("transform_backward");
  return DEVO_OK;
}
