// ba.cu -- in-place Gauss-Newton bundle adjustment (cuda_ba.forward), reprojection
// (cuda_ba.reproject) and the fused projective transform (projective_ops.transform).
//
// Reference behaviour: devo/fastba/ba_cuda.cu:214-365 (per-edge residuals/Jacobians, ~340
// global float atomics per edge into B,E,C,v,u), :461-537 (~20 ATen launches per iteration:
// Schur complement via cuBLAS, cuSOLVER potrf with a host sync, back-substitution),
// :160-211 (retractions).  Same normal equations here, organised for B200:
//
//   plan   (graph_plan.cu, 1 launch)   edges grouped by patch: perm / gstart / gkey
//   accum  (1 launch per iteration)    CTA c owns a contiguous range of patches.  Per-edge
//                                      Jacobians in fp32 with the reference's formulas; each
//                                      residual row is the sparse vector g = [-Ji @ i', +Jj @ j']
//                                      (so B = sum w g g^T, v = sum w r g, E_k = sum w Jz g).
//                                      Rows + the per-patch vectors E_k (coefficient -Q_k) are
//                                      staged densely in shared memory (fp64) and the CTA's
//                                      partial of the *reduced* system  S = B - E Q E^T,
//                                      y = v - E Q u  is accumulated in registers, one fixed
//                                      set of matrix entries per thread, fixed summation order:
//                                      no floating-point atomics, bitwise reproducible.  An entry
//                                      only visits the edges listed for its pair of pose blocks
//                                      (bitsets in shared memory); the terms every edge of a patch
//                                      shares (its source frame's diagonal block / residual column)
//                                      are pre-summed per patch by a warp.
//   reduce (same launch)               clusters of 8 CTAs: partials parked in shared memory, slice r
//                                      of the 8 added by CTA r through DSMEM in rank order; one ticket
//                                      per CTA elects the last CTA of the grid.
//   solve  (same launch)               that CTA: cluster partials added in index order, damping,
//                                      block-6 LDL^T of [S|y] held in registers (2 barriers per pose
//                                      block), back-substitution, SE3 retraction.
//   depth update                       dZ_k = Q_k (u_k - E_k . dX); fused into the prologue of the
//                                      next iteration's accum launch (same patch ownership).
//
// => iterations+1 launches (+1 for the plan unless shared), no host sync, no floating-point atomics,
//    CUDA-graph capturable.
// Deliberate deviations from the reference (documented in DESIGN.md):
//   * the 6N x 6N system is accumulated and solved in fp64 (the reference: fp32 atomics,
//     run-to-run non-deterministic); inputs/outputs and per-edge Jacobians stay fp32.
//   * an edge that is masked out contributes exactly zero (the reference multiplies by a 0
//     weight, which turns Z==0 into NaN: ba_cuda.cu:268-281).
//   * pose block indices >= t1 are treated as fixed (the reference would write out of bounds).
//   * a non-positive-definite system sets *status = iteration+1 and skips the remaining
//     iterations instead of throwing from inside the launch sequence.
#include <cooperative_groups.h>
#include "common.cuh"

extern "C" size_t devo_graph_plan_workspace(int E);

namespace cg = cooperative_groups;
namespace {

constexpr int kAccThreads = 512;
constexpr int kSolveThreads = kAccThreads;   // the solve runs inside the last accumulate CTA
constexpr int kMaxN6 = 150;            // 25 free poses
constexpr size_t kAccSmemBudget = 200 * 1024;
constexpr int kMaxGroupsPerCta = 1023;   // group starts cached in shared memory (falls back to global beyond)

// ---- fastba's own SE3 helpers (un-normalised quaternions; ba_cuda.cu:18-156) ----------------
__device__ __forceinline__ void rot(const float* q, const float* X, float* Y) {
  float uv0 = 2.0f * (q[1] * X[2] - q[2] * X[1]);
  float uv1 = 2.0f * (q[2] * X[0] - q[0] * X[2]);
  float uv2 = 2.0f * (q[0] * X[1] - q[1] * X[0]);
  Y[0] = X[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1);
  Y[1] = X[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2);
  Y[2] = X[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0);
}
// Gij = Gj * Gi^-1 without renormalisation
__device__ __forceinline__ void rel_pose(const float* Pi, const float* Pj, float* tij, float* qij) {
  const float* ti = Pi; const float* qi = Pi + 3;
  const float* tj = Pj; const float* qj = Pj + 3;
  qij[0] = -qj[3] * qi[0] + qj[0] * qi[3] - qj[1] * qi[2] + qj[2] * qi[1];
  qij[1] = -qj[3] * qi[1] + qj[1] * qi[3] - qj[2] * qi[0] + qj[0] * qi[2];
  qij[2] = -qj[3] * qi[2] + qj[2] * qi[3] - qj[0] * qi[1] + qj[1] * qi[0];
  qij[3] = qj[3] * qi[3] + qj[0] * qi[0] + qj[1] * qi[1] + qj[2] * qi[2];
  float r[3];
  rot(qij, ti, r);
  tij[0] = tj[0] - r[0]; tij[1] = tj[1] - r[1]; tij[2] = tj[2] - r[2];
}
// Y = Ad(G)^T X
__device__ __forceinline__ void adjT(const float* t, const float* q, const float* X, float* Y) {
  float qinv[4] = {-q[0], -q[1], -q[2], q[3]};
  rot(qinv, X, Y);
  rot(qinv, X + 3, Y + 3);
  float u[3] = {t[2] * X[1] - t[1] * X[2], t[0] * X[2] - t[2] * X[0], t[1] * X[0] - t[0] * X[1]};
  float v[3];
  rot(qinv, u, v);
  Y[3] += v[0]; Y[4] += v[1]; Y[5] += v[2];
}
__device__ __forceinline__ void exp_se3(const float* xi, float* t, float* q) {
  const float* phi = xi + 3;
  float theta_sq = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  float theta_p4 = theta_sq * theta_sq;
  float theta = sqrtf(theta_sq);
  float imag, real;
  if (theta_sq < 1e-8) {   // double literals on purpose: the reference evaluates these series in double
    imag = (float)(0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_p4);
    real = (float)(1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_p4);
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
  }
  q[0] = imag * phi[0]; q[1] = imag * phi[1]; q[2] = imag * phi[2]; q[3] = real;
  float tau[3] = {xi[0], xi[1], xi[2]};
  t[0] = tau[0]; t[1] = tau[1]; t[2] = tau[2];
  if (theta > 1e-4) {
    float a = (1 - cosf(theta)) / theta_sq;
    float c[3] = {phi[1] * tau[2] - phi[2] * tau[1], phi[2] * tau[0] - phi[0] * tau[2], phi[0] * tau[1] - phi[1] * tau[0]};
    t[0] += a * c[0]; t[1] += a * c[1]; t[2] += a * c[2];
    float b = (theta - sinf(theta)) / (theta * theta_sq);
    float c2[3] = {phi[1] * c[2] - phi[2] * c[1], phi[2] * c[0] - phi[0] * c[2], phi[0] * c[1] - phi[1] * c[0]};
    t[0] += b * c2[0]; t[1] += b * c2[1]; t[2] += b * c2[2];
  }
}
// T <- Exp(xi) * T
__device__ __forceinline__ void retract_pose(const float* xi, float* P) {
  float dt[3], dq[4];
  exp_se3(xi, dt, dq);
  const float* q = P + 3;
  float q1[4];
  q1[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
  q1[1] = dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2];
  q1[2] = dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0];
  q1[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
  float t1[3];
  rot(dq, P, t1);
  P[0] = t1[0] + dt[0]; P[1] = t1[1] + dt[1]; P[2] = t1[2] + dt[2];
  P[3] = q1[0]; P[4] = q1[1]; P[5] = q1[2]; P[6] = q1[3];
}

struct EdgeTerms {
  float r[2], w[2], Jz[2];
  float Ji[2][6], Jj[2][6];
  bool active;
};

// reprojection_residuals_and_hessian :240-332, one edge
__device__ __forceinline__ void edge_terms(const float* __restrict__ poses, const float* __restrict__ patches,
                                           float fx, float fy, float cx, float cy,
                                           const float* __restrict__ target, const float* __restrict__ weight,
                                           int i, int j, int k, int n, int PP, int centre, EdgeTerms& T) {
  float Pi[7], Pj[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { Pi[c] = poses[(size_t)i * 7 + c]; Pj[c] = poses[(size_t)j * 7 + c]; }
  const float* pk = patches + (size_t)k * 3 * PP;
  float Xi[3] = {(pk[centre] - cx) / fx, (pk[PP + centre] - cy) / fy, 1.0f};
  const float dinv = pk[2 * PP + centre];
  float tij[3], qij[4], Xj[3];
  rel_pose(Pi, Pj, tij, qij);
  rot(qij, Xi, Xj);
  Xj[0] += dinv * tij[0]; Xj[1] += dinv * tij[1]; Xj[2] += dinv * tij[2];
  const float X = Xj[0], Y = Xj[1], Z = Xj[2], W = dinv;
  const float d = (Z >= 0.2f) ? 1.0f / Z : 0.0f;
  const float d2 = d * d;
  const float x1 = fx * (X / Z) + cx;
  const float y1 = fy * (Y / Z) + cy;
  const float rx = target[(size_t)n * 2 + 0] - x1;
  const float ry = target[(size_t)n * 2 + 1] - y1;
  const bool inb = (sqrtf(rx * rx + ry * ry) < 128) && (Z > 0.2f) && (x1 > -64) && (y1 > -64) &&
                   (x1 < 2 * cx + 64) && (y1 < 2 * cy + 64);
  T.active = inb;
  T.r[0] = rx; T.r[1] = ry;
  T.w[0] = inb ? weight[(size_t)n * 2 + 0] : 0.0f;
  T.w[1] = inb ? weight[(size_t)n * 2 + 1] : 0.0f;
  T.Jz[0] = fx * (tij[0] * d - tij[2] * (X * d2));
  T.Jz[1] = fy * (tij[1] * d - tij[2] * (Y * d2));
  T.Jj[0][0] = fx * W * d; T.Jj[0][1] = 0; T.Jj[0][2] = fx * -X * W * d2;
  T.Jj[0][3] = fx * -X * Y * d2; T.Jj[0][4] = fx * (1 + X * X * d2); T.Jj[0][5] = fx * -Y * d;
  T.Jj[1][0] = 0; T.Jj[1][1] = fy * W * d; T.Jj[1][2] = fy * -Y * W * d2;
  T.Jj[1][3] = fy * (-1 - Y * Y * d2); T.Jj[1][4] = fy * (X * Y * d2); T.Jj[1][5] = fy * X * d;
  adjT(tij, qij, T.Jj[0], T.Ji[0]);
  adjT(tij, qij, T.Jj[1], T.Ji[1]);
}

// ---- workspace ------------------------------------------------------------------------------
struct BaLayout {
  size_t perm, gstart, gkey, ngroups, plan_ws, plan_bytes, Q, U, Ek, partials, dX, ticket, total;
  int grid, nent, n6;
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

static int acc_grid(int E) {
  int g = (E + 63) / 64;
  if (g < 1) g = 1;
  g = (g + 7) / 8 * 8;         // whole clusters of kGroupCtas CTAs (the partial systems are reduced through DSMEM)
  if (g > 144) g = 144;
  return g;
}

static BaLayout ba_layout(int E, int nfree) {
  BaLayout L;
  const int n6 = 6 * (nfree > 0 ? nfree : 0);
  L.n6 = n6;
  L.nent = (n6 + 1) * (n6 + 2) / 2;
  L.grid = acc_grid(E);
  size_t off = 0;
  const size_t Em = (size_t)(E > 0 ? E : 1);
  L.perm = off;    off += al(Em * 4);
  L.gstart = off;  off += al((Em + 1) * 4);
  L.gkey = off;    off += al(Em * 8);
  L.ngroups = off; off += al(16);
  L.plan_bytes = devo_graph_plan_workspace(E);
  L.plan_ws = off; off += al(L.plan_bytes);
  L.Q = off;       off += al(Em * 8);
  L.U = off;       off += al(Em * 8);
  L.Ek = off;      off += al(Em * (size_t)(n6 > 0 ? n6 : 1) * 8);
  L.partials = off; off += al((size_t)((L.grid + 7) / 8) * L.nent * 8);   // one partial system per cluster of 8 CTAs
  L.dX = off;      off += al((size_t)(n6 > 0 ? n6 : 1) * 8);
  L.ticket = off;  off += al(4 * 32);   // [0]: how many CTAs of the running launch have delivered their slice
  L.total = off;
  return L;
}

// shared memory of the accumulate CTA: RCAP = 2 EB + GB dense rows of LD doubles, coef[RCAP], zj[RCAP], the per-patch
// pre-sums Hg[GB][28], the selection bitsets (one per pair of pose blocks, EB bits each), blk[EB], gbi[GB]
constexpr int kPre = 28;    // 21 entries of the source frame's diagonal block, 6 of its residual column, the y^T y corner
static bool acc_shape(int n6, int nfree, int& EB, int& GB, size_t& smem) {
  const int LD = n6 + 1;
  const size_t nblk = (size_t)(nfree > 0 ? nfree : 0) + 1;
  const size_t npairs = nblk * (nblk + 1) / 2;
  const size_t sel_max = npairs * (size_t)(kAccThreads / 32) * 4;
  const size_t per_edge = 2 * (size_t)(LD + 2) * 8 + 4, per_group = (size_t)(LD + 2 + kPre) * 8 + 4;
  const size_t avail = kAccSmemBudget - sel_max;
  EB = (int)(avail / (per_edge + per_group / 2));
  if (EB > kAccThreads) EB = kAccThreads;
  GB = (int)((avail - (size_t)EB * per_edge) / per_group);
  if (GB > EB) GB = EB;
  smem = (size_t)EB * per_edge + (size_t)GB * per_group + npairs * (size_t)((EB + 31) / 32) * 4 + 16;
  return EB >= 8 && GB >= 1;
}

// upper-triangular (row-major, a<=b) linear index -> (a,b) for an m x m matrix
__device__ __forceinline__ void tri_decode(int idx, int m, int& a, int& b) {
  // row a starts at a*m - a*(a-1)/2
  float fm = (float)(2 * m + 1);
  int aa = (int)floorf((fm - sqrtf(fm * fm - 8.0f * (float)idx)) * 0.5f);
  if (aa < 0) aa = 0;
  if (aa > m - 1) aa = m - 1;
  while (aa > 0 && aa * m - aa * (aa - 1) / 2 > idx) aa--;
  while (aa + 1 < m && (aa + 1) * m - (aa + 1) * aa / 2 <= idx) aa++;
  a = aa;
  b = idx - (aa * m - aa * (aa - 1) / 2) + aa;
}

// ---- solve (executed by the last accumulate CTA to finish) -----------------------------------
// smem: A packed lower-triangular n(n+1)/2 doubles (A(i,j) at i(i+1)/2 + j, j<=i), y[n+1], rinv[n] doubles.
// Square-root-free elimination of the augmented system [S | y] with ONE barrier per pivot:
//   step k:  r = 1/A_kk ;  A_ij -= A_ik A_jk r  (i>=j>k) ;  y_i -= A_ik y_k r  (i>k)
// reads touch column k only, writes touch columns > k only, so no second barrier is needed.
// Afterwards A_ik (i>k) = L_ik d_k and y = L^-1 y; warp 0 finishes x = L^-T D^-1 y with shuffles.
// S positive definite  <=>  every pivot d_k = A_kk > 0 (same acceptance test as Cholesky).
#ifdef DEVO_BA_TIMING
__device__ long long g_ba_clk[24];
__device__ __forceinline__ long long ba_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define BA_STAMP(i) do { if (threadIdx.x == 0) g_ba_clk[i] = ba_now(); } while (0)   /* ns, comparable across SMs */
#else
#define BA_STAMP(i)
#endif
// Solve S dX = y for the reduced system held as `nparts` partial sums (fixed summation order), then retract
// the poses.  smem: A packed lower-triangular n(n+1)/2 doubles, y[n+1], rinv[n], col[2][n+1] doubles.
//
// Square-root-free elimination of the augmented matrix [S | y] kept in REGISTERS: thread t owns the entries
// e = t, t+T, ... (row gi >= column gj; row gi == n is y^T) for the whole factorisation.  At pivot step k
// every entry with gj > k is updated as  v -= c[gi] c[gj] / d_k  where c = column k.  The owners of column
// k+1 publish their freshly updated values into a double-buffered shared column, so there is exactly ONE
// barrier per pivot and two shared loads + two DFMA-class instructions per entry and step.  The thread that
// finalises the next pivot also stores 1/d (the only fp64 division of the step).  Afterwards
// A_ik = L_ik d_k, y = L^-1 y; warp 0 finishes x = L^-T D^-1 y.  S positive definite <=> every pivot > 0
// (the same acceptance test as Cholesky).
// 1/d for a pivot.  It sits on the critical path of every elimination step, so its latency was measured on B200
// (tools/fp64_probe.cu): single-precision seed + two Newton steps 144 cycles (the two F2F conversions dominate), IEEE
// division 80 cycles, `rcp.approx.ftz.f64` (MUFU.RCP64H, ~20 good bits) + two Newton steps in fp64 (relative error
// ~1e-16 for normal d; the callers reject d <= 0 and non-finite d before using it) is what is used here.
__device__ __forceinline__ double pivot_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}

template <int KENT>
__device__ void ba_solve_device(unsigned char* smem_raw, float* poses, const double* partials, double* dX,
                                int32_t* status, int nparts, int t0, int nfree, int itr) {
  BA_STAMP(1);
  const int n = 6 * nfree;
  const int LD = n + 1;
  const int nent = (n + 1) * (n + 2) / 2;
  double* A = reinterpret_cast<double*>(smem_raw);
  double* y = A + n * (n + 1) / 2;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  if (n == 0) return;
  if (tid == 0) s_fail = 0;
  // (another CTA of this launch may have flagged a failure: the load is issued here and consumed after the reduction
  //  below, so its L2 round trip overlaps the partial sums')
  const int st_in = *(volatile int32_t*)status;
  // the pose this thread will retract at the very end (tid < nfree <= 25): its L2 round trip is paid here, not there
  float Pret[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
  if (tid < nfree) {
#pragma unroll
    for (int c = 0; c < 7; c++) Pret[c] = __ldcg(poses + (size_t)(t0 + tid) * 7 + c);
  }

  // fixed-order reduction of the partial sums; entry (a,b), a<=b of the (n+1)x(n+1) augmented matrix.
  // 8 independent accumulators keep 8 L2 loads in flight; combined in a fixed tree => deterministic.
  for (int idx = tid; idx < nent; idx += kSolveThreads) {
    // (<= 19 group partials of a 148-CTA grid: all loads of an entry are issued at once, predicated, and summed in a
    //  fixed tree)
    double s8[8];
#pragma unroll
    for (int u = 0; u < 8; u++) s8[u] = 0.0;
    for (int p = 0; p < nparts; p += 24) {
      double t[24];
#pragma unroll
      for (int u = 0; u < 24; u++) t[u] = (p + u < nparts) ? __ldcg(&partials[(size_t)(p + u) * nent + idx]) : 0.0;
#pragma unroll
      for (int u = 0; u < 24; u++) s8[u & 7] += t[u];
    }
    double s = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
    int a, b;
    tri_decode(idx, LD, a, b);
    if (b == n) {
      if (a < n) y[a] = s;                                // y = v - E Q u
    } else {
      if (a == b) s += 1e-4 * s + 1.0;                    // S += I o (1e-4 S + 1)   (:517-518)
      A[b * (b + 1) / 2 + a] = s;                         // symmetric: store as lower (b,a)
    }
  }
  if (st_in != 0) return;                                 // uniform
  __syncthreads();
  BA_STAMP(2);

  // ---- block-6 LDL^T of the augmented matrix [S | y] (one 6x6 pose block per pivot step) -----------------------
  // Thread t owns the entries e = t, t+T, ... of the lower triangle (row gi >= column gj; row n is y^T) in REGISTERS
  // for the whole factorisation.  Step k (pivot block K = [6k, 6k+6)):
  //   * the owners of column block K have published it as C[(n+1) x 6] (double-buffered by k)
  //   * each thread t < #rows factors the 6x6 block C[K,:] = L D L^T in registers (56 DFMA + 6 reciprocals; the
  //     pivots are the scalar LDL^T pivots: positive  <=>  S positive definite) and solves one row
  //     W_i = C_i (L D L^T)^-1  (36 DFMA)                                                  -> barrier
  //   * every entry right of the block:  v -= W[gi] . C[gj]  (6 DFMA); the freshly final entries of column block K+1
  //     are published for the next step; the entries of column block K are replaced by W (= L_iK; row n: D^-1 z)  -> barrier
  // 2 barriers per pose block instead of 1 per scalar pivot: 14 instead of 42 at 7 free poses, and the pivot
  // reciprocals leave the critical path of the trailing update.
  const int naug = (n + 1) * (n + 2) / 2 - 1;            // lower triangle of the augmented matrix without (n,n)
  double* Cb = y + n + 1;                                 // [2][(n+1)][6] published column blocks
  double* Wb = Cb + 2 * (n + 1) * 6;                      // [(n+1)][6]
  double v[KENT];
  int ent[KENT];                                          // (gi << 16) | gj, or -1
#pragma unroll
  for (int q = 0; q < KENT; q++) {
    const int e = tid + q * kSolveThreads;
    ent[q] = -1;
    v[q] = 0.0;
    if (e < naug) {
      int gi = (int)floorf((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while (gi * (gi + 1) / 2 > e) gi--;
      while ((gi + 1) * (gi + 2) / 2 <= e) gi++;
      const int gj = e - gi * (gi + 1) / 2;
      ent[q] = (gi << 16) | gj;
      v[q] = (gi < n) ? A[gi * (gi + 1) / 2 + gj] : y[gj];
      if (gj < 6) {                                       // column block 0
        Cb[gi * 6 + gj] = v[q];
        if (gi < 6) Cb[gj * 6 + gi] = v[q];               // symmetric fill of the pivot block
      }
    }
  }
  __syncthreads();
  BA_STAMP(3);
  for (int k = 0; k < nfree; k++) {
    const int K0 = 6 * k;
    const double* C = Cb + (k & 1) * (n + 1) * 6;
    double* Cn = Cb + ((k + 1) & 1) * (n + 1) * 6;
    // The threads that own a row below the pivot block (and the y row) factor the 6x6 block C[K,:] = L D L^T in
    // registers (unit-lower L below the diagonal of Sm, reciprocal pivots rp) and solve their row W_i = C_i (L D L^T)^-1
    {
      const int i = K0 + 6 + tid;
      if (i <= n) {
        double Sm[6][6], rp[6];
        bool okp = true;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int c = 0; c <= a; c++) Sm[a][c] = C[(K0 + a) * 6 + c];
#pragma unroll
        for (int pcol = 0; pcol < 6; pcol++) {
          const double d = Sm[pcol][pcol];
          okp = okp && (d > 0.0) && isfinite(d);
          rp[pcol] = pivot_rcp(d);
#pragma unroll
          for (int a = pcol + 1; a < 6; a++) {
            const double l = Sm[a][pcol] * rp[pcol];
#pragma unroll
            for (int c = pcol + 1; c <= a; c++) Sm[a][c] -= l * Sm[c][pcol];     // column pcol still unscaled here
          }
#pragma unroll
          for (int a = pcol + 1; a < 6; a++) Sm[a][pcol] *= rp[pcol];            // now L
        }
        if (tid == 0 && !okp) s_fail = 1;                 // thread 0 always owns a row (at least the y row)
        double w[6];
#pragma unroll
        for (int a = 0; a < 6; a++) w[a] = C[i * 6 + a];
#pragma unroll
        for (int a = 1; a < 6; a++)
#pragma unroll
          for (int c = 0; c < a; c++) w[a] -= Sm[a][c] * w[c];           // L u = c
#pragma unroll
        for (int a = 0; a < 6; a++) w[a] *= rp[a];                       // D^-1
#pragma unroll
        for (int a = 4; a >= 0; a--)
#pragma unroll
          for (int c = a + 1; c < 6; c++) w[a] -= Sm[c][a] * w[c];       // L^T w = u
#pragma unroll
        for (int a = 0; a < 6; a++) Wb[i * 6 + a] = w[a];
      }
    }
    __syncthreads();
    if (s_fail) break;                                    // uniform (written before the barrier)
#pragma unroll
    for (int q = 0; q < KENT; q++) {
      if (ent[q] < 0) continue;
      const int gi = ent[q] >> 16, gj = ent[q] & 0xffff;
      if (gj < K0) continue;
      if (gj < K0 + 6) {                                  // column block K: final, becomes the factor block
        if (gi >= K0 + 6) v[q] = Wb[gi * 6 + (gj - K0)];
        continue;
      }
      const double* wi = Wb + gi * 6;
      const double* cj = C + gj * 6;
      double acc = v[q];
#pragma unroll
      for (int a = 0; a < 6; a++) acc -= wi[a] * cj[a];
      v[q] = acc;
      if (gj < K0 + 12) {                                 // next pivot block's columns
        Cn[gi * 6 + (gj - K0 - 6)] = acc;
        if (gi < K0 + 12) Cn[gj * 6 + (gi - K0 - 6)] = acc;
      }
    }
    __syncthreads();
  }
  // write the factor back: A_iK = L_iK (rows below each pivot block), y = D^-1 L^-1 y
#pragma unroll
  for (int q = 0; q < KENT; q++) {
    if (ent[q] < 0) continue;
    const int gi = ent[q] >> 16, gj = ent[q] & 0xffff;
    if (gi < n) A[gi * (gi + 1) / 2 + gj] = v[q];
    else y[gj] = v[q];
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicCAS(status, 0, itr + 1);
    return;
  }
  BA_STAMP(4);
  // back substitution  x_K = w_K - sum_{i >= 6k+6} L[i,K]^T x_i, pose blocks in descending order, by warp 0: lane
  // (part, a) = (lane / 6, lane % 6) sums every 5th row of column a (a chain of <= ceil(n/5) DFMAs), the 5 parts are added
  // in a fixed order through shared memory.  (Lanes over rows with a 5-level double-precision shuffle tree per column
  // took ~1000 cycles per pose block: 3.1 us of the solve at 7 free poses.)
  if (tid < 32) {
    double* part = Wb;                                    // [5][6] scratch (the factor blocks are no longer needed)
    const int a = tid % 6, pr = tid / 6;
    for (int k = nfree - 2; k >= 0; k--) {
      const int K0 = 6 * k;
      if (tid < 30) {
        double acc = 0.0;
        for (int base = K0 + 6 + pr; base < n; base += 40) {
          double l8[8], x8[8];
#pragma unroll
          for (int t = 0; t < 8; t++) {                   // all 16 loads in flight, then the ordered chain
            const int i = base + 5 * t;
            const bool in = i < n;
            l8[t] = in ? A[i * (i + 1) / 2 + K0 + a] : 0.0;
            x8[t] = in ? y[i] : 0.0;
          }
#pragma unroll
          for (int t = 0; t < 8; t++) acc += l8[t] * x8[t];
        }
        part[pr * 6 + a] = acc;
      }
      __syncwarp();
      if (tid < 6) y[K0 + tid] -= (((part[tid] + part[6 + tid]) + (part[12 + tid] + part[18 + tid])) + part[24 + tid]);
      __syncwarp();
    }
  }
  __syncthreads();
  BA_STAMP(5);
  bool bad = false;
  for (int i = tid; i < n; i += kSolveThreads) {
    dX[i] = y[i];
    if (!isfinite(y[i])) bad = true;
  }
  if (__syncthreads_or(bad)) {
    if (tid == 0) atomicCAS(status, 0, itr + 1);
    return;
  }
  // pose retraction  T <- Exp(dX) T   (:160-188)
  if (tid < nfree) {                                      // (nfree <= kMaxN6 / 6 < kSolveThreads)
    float xi[6];
#pragma unroll
    for (int c = 0; c < 6; c++) xi[c] = (float)y[6 * tid + c];
    float* dst = poses + (size_t)(t0 + tid) * 7;
    retract_pose(xi, Pret);
#pragma unroll
    for (int c = 0; c < 7; c++) dst[c] = Pret[c];
  }
}

// ---- order-preserving reduction of the per-CTA partials ---------------------------------------------------
// The grid runs as thread-block clusters of kGroupCtas CTAs (see the end of ba_accumulate_kernel): which CTA does
// the work depends on timing, what is summed in which order does not.
constexpr int kGroupCtas = 8;

// ---- accumulate kernel ----------------------------------------------------------------------
// smem: X[R][LD] doubles, coef[R] doubles, zj[R] doubles (Jz per row), per-patch pre-sums, block-pair bitsets, batch
// bookkeeping (acc_shape() is the host-side mirror of the carve-up)
template <int EPT>
__global__ void __launch_bounds__(kAccThreads, 1) ba_accumulate_kernel(
    float* __restrict__ poses_rw, float* __restrict__ patches, const float* __restrict__ intrinsics,
    const float* __restrict__ target, const float* __restrict__ weight, const float* __restrict__ lmbda,
    const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, const int64_t* __restrict__ kk,
    const int32_t* __restrict__ perm, const int32_t* __restrict__ gstart, const int64_t* __restrict__ gkey,
    const int32_t* __restrict__ ngroups_p,
    double* __restrict__ Qg, double* __restrict__ Ug, double* __restrict__ Ekg,
    double* __restrict__ partials, double* dX, int32_t* __restrict__ ticket,
    int32_t* __restrict__ status, int E, int PP, int centre, int t0, int nfree, int n_poses,
    int EB, int GB, int apply_update, int do_accumulate, int itr, double* __restrict__ sys_out,
    int plan_is_older, int32_t* __restrict__ status_or) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const float* poses = poses_rw;
  const int n6 = 6 * nfree;
  const int LD = n6 + 1;
  const int RCAP = 2 * EB + GB;
  double* X = reinterpret_cast<double*>(smem_raw);
  double* coef = X + (size_t)RCAP * LD;
  double* zj = coef + RCAP;
  // per (pose block A <= pose block B) pair -- the residual column counts as block `nfree` --: the bitset of the batch's
  // edges whose two residual rows are non-zero in both blocks
  double* Hg = zj + RCAP;                                             // [GB][kPre] per-patch pre-sums (see below)
  unsigned int* sel = reinterpret_cast<unsigned int*>(Hg + (size_t)GB * kPre);
  const int nblk = nfree + 1;
  const int npairs = nblk * (nblk + 1) / 2;
  int* s_blk = reinterpret_cast<int*>(sel + (size_t)npairs * ((EB + 31) >> 5));   // [EB] (bi+1) | (bj+1) << 8 | active << 16
  int* s_gbi = s_blk + EB;                                           // [GB] the patch's free source block, or -1
  __shared__ int s_batch[3];   // gs, ge, bad
  __shared__ __align__(16) float s_intr[4];   // (aligned: read with one LDS.128 that must not straddle s_batch)
  __shared__ int s_gstart[kMaxGroupsPerCta + 1];

  const int tid = threadIdx.x;
  const int nent = (n6 + 1) * (n6 + 2) / 2;
  // ---- who sums what (pure index arithmetic: done before the wait for the previous launch).  Every entry (a,b) of the upper triangle of [S|y] is summed by NC adjacent lanes ("chunks"): lane
  // c takes the c-th part of the batch's edge slots and every NC-th per-patch row; the NC partial sums are combined in
  // a fixed tree at the end, so the result does not depend on timing.  Entries are dealt PAIR-MAJOR (all entries of
  // block pair (0,0), then (0,1), ...): a warp holds 32/NC entries of one or two block pairs, i.e. its lanes walk the same
  // edge lists below.  Measured at S8 (tools/ba_timing.py): NC = 1 -> the accumulate phase takes 5.3 us (the lists of
  // the CTA's own source frame, 64 edges, are a serial chain in one warp); NC = 4 balances the lanes but pays the
  // per-slot set-up four times and most rounds of four edges run half empty: 3.2 us for the lists yet a slower
  // iteration overall (43.8 vs 42.0 us).  NC = 1 it is.
  constexpr int NC = 1;
  constexpr int SL = EPT * NC;                          // (entry, chunk) slots per thread
  constexpr int EPS = kAccThreads / NC;                 // entries per round of slots
  const int chunk = tid & (NC - 1);
  double acc[SL];
  // a | b << 8 | hcode << 16 | (a / 6) << 24 ; hcode: 0, or 1 + the entry's index in a patch's pre-sums (when the entry
  // lies in a diagonal block, in the residual column, or is the corner).  An empty slot points at (0,0), never stored.
  unsigned int ab[SL];
  unsigned int live = 0u;                               // bit q: slot q holds an entry
#pragma unroll
  for (int q = 0; q < SL; q++) {
    acc[q] = 0.0;
    ab[q] = 0u;
    const int idx = tid / NC + q * EPS;
    if (idx < nent) {
      // block row A: a diagonal pair (21 entries), nfree-1-A off-diagonal pairs (36) and the residual-column pair (6)
      int a, b, hcode = 0, A = 0, off = 0;
      while (A < nfree && off + 27 + 36 * (nfree - 1 - A) <= idx) { off += 27 + 36 * (nfree - 1 - A); A++; }
      int r = idx - off;
      if (A == nfree) { a = n6; b = n6; hcode = kPre; }                    // the y^T y corner
      else if (r < 21) {                                                   // diagonal block: x <= y
        hcode = 1 + r;
        int x = 0;
        while (r >= 6 - x) { r -= 6 - x; x++; }
        a = 6 * A + x; b = 6 * A + x + r;
      } else if (r < 21 + 36 * (nfree - 1 - A)) {
        r -= 21;
        a = 6 * A + (r % 36) / 6; b = 6 * (A + 1 + r / 36) + (r % 6);
      } else { a = 6 * A + (r - 21 - 36 * (nfree - 1 - A)); b = n6; hcode = 22 + (a - 6 * A); }
      ab[q] = (unsigned)a | ((unsigned)b << 8) | ((unsigned)hcode << 16) | ((unsigned)A << 24);
      live |= 1u << q;
    }
  }
  // Launched with programmatic stream serialisation.  What the previous Gauss-Newton iteration writes (poses, depths, dX,
  // Q/u/E_k, partials, tickets, status) may only be touched after the wait; the edge list and its grouping are inputs of
  // the whole call, so from the second iteration on (`itr > 0`: the predecessor is this kernel) -- or from the first when
  // the caller vouches that the plan is older than the predecessor (`plan_is_older`, devo_ba_forward_prepared) -- the
  // chain of dependent index loads perm -> ii/jj/kk is issued BEFORE the wait and overlaps the predecessor's tail.
  int pre_n = -1, pre_i = 0, pre_j = 0, pre_k = 0, pre_G = 0;
  long long pre_key = -1;                                // patch of the first group this warp updates
  bool pre = false;
  if ((itr > 0 || plan_is_older) && do_accumulate) {
    pre = true;
    pre_G = *ngroups_p;
    const int gpc_ = (pre_G + gridDim.x - 1) / gridDim.x;
    const int g0_ = blockIdx.x * gpc_, g1_ = min(pre_G, g0_ + gpc_);
    if (g0_ < g1_) {
      const int eb_ = gstart[g0_];
      if (eb_ + tid < gstart[g1_]) {          // (a superset of the first batch; unused values are simply dropped)
        pre_n = perm[eb_ + tid];
        pre_i = (int)ii[pre_n]; pre_j = (int)jj[pre_n]; pre_k = (int)kk[pre_n];
      }
    }
    if (tid < 4) s_intr[tid] = intrinsics[tid];
    if (apply_update && g0_ + (tid >> 5) < g1_) pre_key = gkey[g0_ + (tid >> 5)];
  }
  DEVO_PDL_WAIT();
  DEVO_PDL_TRIGGER();
#ifdef DEVO_BA_TIMING
  if (blockIdx.x == (gridDim.x >> 1) && tid == 0) g_ba_clk[0] = ba_now();
  if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && do_accumulate) g_ba_clk[16] = ba_now();
#endif
  const int st = *(volatile int32_t*)status;               // (consumed below, after the loads of the depth update are out)
  // the last launch of a call folds the call's status into the caller's sticky word (devo_ba_forward_prepared)
  if (!do_accumulate && status_or != nullptr && blockIdx.x == 0 && tid == 0 && st != 0) atomicOr(status_or, st);
  if (!pre && tid < 4) s_intr[tid] = intrinsics[tid];   // only intrinsics[0] is used (:232-238)
  const int G = pre ? pre_G : *ngroups_p;
  const int gpc = (G + gridDim.x - 1) / gridDim.x;
  const int g0 = blockIdx.x * gpc;
  const int g1u = min(G, g0 + gpc);
  const float lm = lmbda[0];

  // ---- prologue: apply the previous iteration's depth update  dZ_k = Q_k (u_k - E_k . dX)  to the patches this CTA
  // owns: one warp per patch, lanes over the columns of E_k (fixed-order shuffle tree); the loads do not wait for the
  // status word, only the store does
  if (apply_update) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int g = g0 + warp; g < g1u; g += kAccThreads / 32) {
      const double* ek = Ekg + (size_t)g * (n6 > 0 ? n6 : 1);
      double part = 0.0;
      for (int c = lane; c < n6; c += 32) part += __ldcg(ek + c) * __ldcg(dX + c);
      const double ug = __ldcg(Ug + g), qg = __ldcg(Qg + g);
      const long long key = (pre_key >= 0 && g == g0 + warp) ? pre_key : gkey[g];
      float* pk = patches + (size_t)key * 3 * PP + 2 * PP;
      float d = pk[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (st == 0) {
        d = d + (float)(qg * (ug - part));
        d = (d > 20) ? 1.0f : d;
        d = fmaxf(d, 1e-4f);
        for (int c = lane; c < PP; c += 32) pk[c] = d;
      }
    }
  }
  if (st != 0 && (!do_accumulate || nfree <= 0)) return;
  // An earlier iteration (or, for a CTA that starts late, another CTA of this launch) failed: the reference would have
  // thrown.  The CTA does no work but still takes part in the cluster reduction below -- its peers wait for it -- with a
  // zero partial; the solving CTA sees the status and leaves the poses alone.
  const bool skip = (st != 0);
  if (skip && sys_out && blockIdx.x == 0) {   // sharded form: peers must see a system and this rank's failure
    for (int q = tid; q < nent; q += kAccThreads) sys_out[q] = 0.0;
    if (tid == 0) sys_out[nent] = 1.0;
  }
  const int g1 = skip ? g0 : g1u;
  const bool gs_cached = (g1 - g0) <= kMaxGroupsPerCta;
  if (gs_cached)
    for (int q = tid; q <= g1 - g0; q += kAccThreads) s_gstart[q] = gstart[g0 + q];
  auto GS = [&](int g) { return gs_cached ? s_gstart[g - g0] : gstart[g]; };
  if (!do_accumulate) return;
#ifdef DEVO_BA_TIMING
  if (blockIdx.x == (gridDim.x >> 1) && tid == 0) g_ba_clk[17] = ba_now();
#endif
  __syncthreads();   // depth updates of this CTA's patches are visible to its own threads; s_intr ready
  const float fx = s_intr[0], fy = s_intr[1], cx = s_intr[2], cy = s_intr[3];

  const int SW = (EB + 31) >> 5;                        // bitset words per pair

  int gs = g0;
  while (gs < g1) {
    // ---- choose a batch of whole groups: <= EB edges, <= GB groups
    if (tid == 0) {
      const int base = GS(gs);
      int ge = gs;
      while (ge < g1 && (ge - gs) < GB && (GS(ge + 1) - base) <= EB) ge++;
      s_batch[0] = gs; s_batch[1] = ge; s_batch[2] = (ge == gs);
    }
    __syncthreads();
    const int ge = s_batch[1];
    if (s_batch[2]) {        // a single patch has more edges than a batch can hold
      if (tid == 0) atomicCAS(status, 0, DEVO_ECAPACITY);
      break;
    }
    const int ebase = GS(gs);
    const int ne = GS(ge) - ebase;
    const int ng = ge - gs;
    const int R = 2 * ne + ng;

    // ---- one thread per edge: residual and Jacobians in registers (a chain of dependent global loads: perm -> ii/jj/kk
    //      -> poses/patch), issued BEFORE the rows are zeroed so that the zeroing by all threads hides that latency
    EdgeTerms T;
    int bi = 0, bj = 0;
    if (tid < ne) {
      int n, i, j, k;
      if (pre && gs == g0 && pre_n >= 0) { n = pre_n; i = pre_i; j = pre_j; k = pre_k; }
      else { n = perm[ebase + tid]; i = (int)ii[n]; j = (int)jj[n]; k = (int)kk[n]; }
      edge_terms(poses, patches, fx, fy, cx, cy, target, weight, i, j, k, n, PP, centre, T);
      bi = i - t0; bj = j - t0;
    }
    // zero the dense rows and the selection bitsets
    for (int q = tid; q < R * LD; q += kAccThreads) X[q] = 0.0;
    for (int q = tid; q < npairs * SW; q += kAccThreads) sel[q] = 0u;
    __syncthreads();

    // ---- two dense rows per edge
    if (tid < ne) {
      const bool fi = (bi >= 0 && bi < nfree), fj = (bj >= 0 && bj < nfree);
#pragma unroll
      for (int rho = 0; rho < 2; rho++) {
        const int r = 2 * tid + rho;
        double* row = X + (size_t)r * LD;
        const bool on = T.active;
        if (on) {
          if (fi) {
#pragma unroll
            for (int c = 0; c < 6; c++) row[6 * bi + c] -= (double)T.Ji[rho][c];
          }
          if (fj) {
#pragma unroll
            for (int c = 0; c < 6; c++) row[6 * bj + c] += (double)T.Jj[rho][c];
          }
          row[n6] = (double)T.r[rho];
        }
        coef[r] = on ? (double)T.w[rho] : 0.0;
        zj[r] = on ? (double)T.Jz[rho] : 0.0;
      }
      s_blk[tid] = (fi ? bi + 1 : 0) | ((fj ? bj + 1 : 0) << 8) | (T.active ? (1 << 16) : 0);
    }
    __syncthreads();

#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && gs == g0) g_ba_clk[8] = ba_now();
#endif
    // ---- per patch, two warp roles:
    //  role 0: C_k, u_k, Q_k and the dense vector E_k (lanes over columns)
    //  role 1: the patch's PRE-SUMS.  In the patch graph every edge of a patch leaves from the patch's own frame, so the
    //          diagonal block and the residual column of that frame get a term from EVERY edge of the patch -- at S8 a
    //          list of all 64 edges of the CTA, summed by the one warp that owns those 27 entries (5 us, the critical
    //          path of the launch).  Here a warp adds them up per patch (28 lanes: 21 + 6 entries and the y^T y corner,
    //          <= 2 x edges-of-the-patch terms each); the entry owners below then add one number per patch.  The warp
    //          also files the patch's edges into the block-pair lists -- except under the pairs it has pre-summed.  A
    //          patch whose edges leave from different (or fixed) frames gets no pre-sum but the corner and keeps its lists.
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int item = warp; item < 2 * ng; item += kAccThreads / 32) {
        const int gi = item >> 1;
        const int g = gs + gi;
        const int e0 = GS(g) - ebase, e1 = GS(g + 1) - ebase;
        const int r0 = 2 * e0, r1 = 2 * e1;
        if ((item & 1) == 0) {
          double C = 0.0, u = 0.0;
          for (int r = r0; r < r1; r++) {
            const double wz = coef[r] * zj[r];
            C += wz * zj[r];
            u += wz * X[(size_t)r * LD + n6];
          }
          const double Q = 1.0 / (C + (double)lm);
          double* erow = X + (size_t)(2 * ne + gi) * LD;
          for (int c = lane; c < n6; c += 32) {
            double e = 0.0;
            for (int r = r0; r < r1; r++) e += coef[r] * zj[r] * X[(size_t)r * LD + c];
            erow[c] = e;
            Ekg[(size_t)g * n6 + c] = e;
          }
          if (lane == 0) {
            erow[n6] = u;
            coef[2 * ne + gi] = -Q;
            Qg[g] = Q;
            Ug[g] = u;
          }
        } else {
          const int f = (s_blk[e0] & 0xff) - 1;             // free source block of the first edge, or -1
          bool same = true;
          for (int e = e0 + lane; e < e1; e += 32) same = same && (((s_blk[e] & 0xff) - 1) == f);
          const int gb = (__all_sync(0xffffffffu, same) && f >= 0) ? f : -1;
          if (lane == 0) s_gbi[gi] = gb;
          if (lane < kPre && (gb >= 0 || lane == kPre - 1)) {
            int ca, cb;
            if (lane < 21) {
              int x = 0, r = lane;
              while (r >= 6 - x) { r -= 6 - x; x++; }
              ca = 6 * gb + x; cb = ca + r;
            } else if (lane < 27) { ca = 6 * gb + (lane - 21); cb = n6; }
            else { ca = n6; cb = n6; }
            double h = 0.0;
            for (int r = r0; r < r1; r++) {
              const double* row = X + (size_t)r * LD;
              h += coef[r] * row[ca] * row[cb];
            }
            Hg[gi * kPre + lane] = h;
          }
          for (int e = e0 + lane; e < e1; e += 32) {
            const int blk = s_blk[e];
            if (!(blk >> 16)) continue;                     // an inactive edge has zero coefficients: listed nowhere
            // the blocks this edge's rows touch, ascending and distinct (+ the residual column); every pair of them gets
            // the edge's bit, unless the patch's pre-sums cover the pair
            int bl[3];
            int nb = 0;
            const int b0 = (blk & 0xff) - 1, b1 = ((blk >> 8) & 0xff) - 1;
            if (b0 >= 0 && b1 >= 0 && b0 != b1) { bl[nb++] = min(b0, b1); bl[nb++] = max(b0, b1); }
            else if (b0 >= 0) bl[nb++] = b0;
            else if (b1 >= 0) bl[nb++] = b1;
            bl[nb++] = nfree;
            const unsigned int bit = 1u << (e & 31);
            for (int x = 0; x < nb; x++)
              for (int y2 = x; y2 < nb; y2++) {
                const int A = bl[x], B = bl[y2];
                if (A == nfree) continue;                                   // the corner is always pre-summed
                if (A == gb && (B == gb || B == nfree)) continue;
                atomicOr(&sel[(A * nblk - A * (A - 1) / 2 + (B - A)) * SW + (e >> 5)], bit);
              }
          }
        }
      }
    }
    __syncthreads();

#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && gs == g0) g_ba_clk[9] = ba_now();
#endif
    // ---- reduced-system partial: acc(a,b) += coef_r * X[r][a] * X[r][b]
    // An edge row only touches the pose blocks of its two frames, so most (row, entry) pairs are exact zeros: each slot
    // walks the bitset of its block pair, restricted to its chunk of the edge slots -- only edges that do contribute, in
    // ascending order.  (A mask test per edge and entry made this loop instruction-issue bound: 7.8 us at S8.)
    const int cw = (ne + NC - 1) / NC;
    const int lo = chunk * cw, hi = min(ne, lo + cw);
#pragma unroll
    for (int q = 0; q < SL; q++) {
      if (!((live >> q) & 1u) || lo >= hi) continue;
      const int a = ab[q] & 0xff, b = (ab[q] >> 8) & 0xff;
      const int A = ab[q] >> 24, B = b / 6;               // b == n6: block nfree (the residual column)
      if (A == nfree) continue;                           // the corner has no list (pre-summed per patch)
      const unsigned int* sp = sel + (A * nblk - A * (A - 1) / 2 + (B - A)) * SW;
      double s_ = acc[q];
      for (int w = lo >> 5; w <= (hi - 1) >> 5; w++) {
        unsigned int bits = sp[w];
        if ((w << 5) < lo) bits &= 0xffffffffu << (lo & 31);
        if (((w + 1) << 5) > hi) bits &= 0xffffffffu >> (32 - (hi & 31));      // (here hi & 31 != 0)
        while (bits) {
          // four listed edges per round: all 24 shared loads are issued before the (ordered) chain of 8 DFMAs
          int e4[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            ok[u] = bits != 0u;
            e4[u] = ok[u] ? (w << 5) + __ffs(bits) - 1 : 0;
            bits &= bits - 1u;                            // (0 & 0xffffffff == 0)
          }
          double c0[4], c1[4], x0a[4], x0b[4], x1a[4], x1b[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const double* row0 = X + (size_t)(2 * e4[u]) * LD;
            c0[u] = coef[2 * e4[u]]; c1[u] = coef[2 * e4[u] + 1];
            x0a[u] = row0[a]; x0b[u] = row0[b]; x1a[u] = row0[LD + a]; x1b[u] = row0[LD + b];
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (ok[u] && c0[u] != 0.0) s_ += c0[u] * x0a[u] * x0b[u];
            if (ok[u] && c1[u] != 0.0) s_ += c1[u] * x1a[u] * x1b[u];
          }
        }
      }
      acc[q] = s_;
    }
#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && gs == g0) g_ba_clk[12] = ba_now();
    if (blockIdx.x == (gridDim.x >> 1) && tid == 433 && gs == g0) g_ba_clk[14] = ba_now();
#endif
    for (int gi = chunk; gi < ng; gi += NC) {          // per patch: the dense row E_k (coefficient -Q_k) and the pre-sums
      const int r = 2 * ne + gi;
      const double c = coef[r];
      const double* row = X + (size_t)r * LD;
      const double* hg = Hg + gi * kPre;
      const int gb = s_gbi[gi];
#pragma unroll
      for (int q = 0; q < SL; q++) {
        const int a = ab[q] & 0xff, b = (ab[q] >> 8) & 0xff;
        const int hcode = (ab[q] >> 16) & 0xff, A = ab[q] >> 24;
        acc[q] += c * row[a] * row[b];
        if (hcode != 0 && (A == gb || hcode == kPre)) acc[q] += hg[hcode - 1];
      }
    }
#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && gs == g0) g_ba_clk[13] = ba_now();
    if (blockIdx.x == (gridDim.x >> 1) && tid == 433 && gs == g0) g_ba_clk[15] = ba_now();
#endif
    __syncthreads();
#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0 && gs == g0) g_ba_clk[7] = ba_now();
#endif
    gs = ge;
  }

  // ---- reduction of the per-CTA partials + solve + retraction in the same launch ---------------------------------
  // The grid is launched as clusters of kGroupCtas CTAs.  Every CTA parks its partial system in its own shared memory;
  // after a cluster barrier CTA r of a cluster adds slice r of the 8 partials in rank order with DSMEM loads and writes
  // it to the cluster's partial in global memory; one atomic ticket per CTA then elects the last CTA of the grid, which
  // adds the cluster partials in index order and solves.  What is summed in which order never depends on timing.
  // (Before: partials through L2, a ticket per group of 8 and a second ticket over the groups -- two fence + atomic +
  // reload round trips, ~10 us from the first CTA's ticket to the assembled system at S8.)
  (void)n_poses; (void)E;
  if (nfree <= 0) return;
  double* Ps = reinterpret_cast<double*>(smem_raw);       // [nent]; the dense rows are dead (barrier at the loop end)
#pragma unroll
  for (int q = 0; q < SL; q++) {
    double v = acc[q];
    if (NC >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);          // (c0 + c1), (c2 + c3): IEEE addition commutes, so both
    if (NC >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);          // lanes of a pair hold the same bits
    if (chunk == 0 && ((live >> q) & 1u)) {
      const int a = ab[q] & 0xff, b = (ab[q] >> 8) & 0xff;
      Ps[a * LD - a * (a - 1) / 2 + (b - a)] = v;
    }
  }
#ifdef DEVO_BA_TIMING
  if (blockIdx.x == (gridDim.x >> 1) && tid == 0) g_ba_clk[10] = ba_now();
#endif
  {
    cg::cluster_group cl = cg::this_cluster();
    const int ngrp = (int)gridDim.x / kGroupCtas;
    const int grp = (int)blockIdx.x / kGroupCtas;
    const int rank = (int)cl.block_rank();
    double* gpart = partials;                                // [ngrp][nent]
    cl.sync();                                               // every partial of the cluster is in place (release/acquire)
    const int slice = (nent + kGroupCtas - 1) / kGroupCtas;
    for (int q = tid; q < slice; q += kAccThreads) {
      const int idx = rank * slice + q;
      if (idx < nent) {
        double t8[kGroupCtas];
#pragma unroll
        for (int u = 0; u < kGroupCtas; u++) t8[u] = cl.map_shared_rank(Ps, u)[idx];
        gpart[(size_t)grp * nent + idx] = ((t8[0] + t8[1]) + (t8[2] + t8[3])) + ((t8[4] + t8[5]) + (t8[6] + t8[7]));
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_batch[1] = atomicAdd(&ticket[0], 1);
    __syncthreads();
#ifdef DEVO_BA_TIMING
    if (blockIdx.x == (gridDim.x >> 1) && tid == 0) g_ba_clk[11] = ba_now();
#endif
    // nobody may leave (or reuse its shared memory) while a peer still reads its partial: a CTA that has its ticket has
    // finished reading, so the last CTA of the grid may go on at once; the others wait for their cluster at the end
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    {
      if (s_batch[1] == (int)gridDim.x - 1) {                // last CTA of the grid: solve
        if (tid == 0) ticket[0] = 0;                         // armed for the next launch
        __threadfence();
        if (sys_out) {
          // edge-sharded form: publish this rank's partial of [S|y] (fixed summation order, no
          // damping yet) + its status word; the all-reduce over ranks and devo_ba_sharded_solve follow.
          for (int idx = tid; idx < nent; idx += kAccThreads) {
            double sacc = 0.0;
            for (int p = 0; p < ngrp; p++) sacc += __ldcg(&gpart[(size_t)p * nent + idx]);
            sys_out[idx] = sacc;
          }
          if (tid == 0) sys_out[nent] = (*(volatile int32_t*)status != 0) ? 1.0 : 0.0;
        } else if (n6 <= 44) ba_solve_device<2>(smem_raw, poses_rw, gpart, dX, status, ngrp, t0, nfree, itr);
        else if (n6 <= 88) ba_solve_device<8>(smem_raw, poses_rw, gpart, dX, status, ngrp, t0, nfree, itr);
        else ba_solve_device<(((kMaxN6 + 1) * (kMaxN6 + 2) / 2) + kSolveThreads - 1) / kSolveThreads>(
            smem_raw, poses_rw, gpart, dX, status, ngrp, t0, nfree, itr);
        BA_STAMP(6);
      }
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// ---- stand-alone solve for the edge-sharded form: `sys` = [S|y] already summed over ranks ------------------
template <int KENT>
__global__ void __launch_bounds__(kSolveThreads, 1) ba_solve_kernel(float* poses, const double* __restrict__ sys,
                                                                    double* dX, int32_t* status, int t0, int nfree, int itr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n6 = 6 * nfree;
  const int nent = (n6 + 1) * (n6 + 2) / 2;
  if (sys[nent] != 0.0) {                  // some rank failed in accumulate: every rank stops here, identically
    if (threadIdx.x == 0) atomicCAS(status, 0, DEVO_ECAPACITY);
    return;
  }
  ba_solve_device<KENT>(smem_raw, poses, sys, dX, status, 1, t0, nfree, itr);
}

// ---- the same solve with the all-reduce FUSED in, over NVLink peer memory ----------------------------------------------
// Every rank keeps its partial system in a symmetric (peer-mapped) buffer: [2][nsys] doubles (double-buffered by
// iteration parity) + one 64-bit epoch flag per parity.  The kernel publishes this rank's flag, waits until every
// peer's flag has reached the epoch (ld.acquire.sys over NVLink), adds the partial systems IN RANK ORDER with peer
// loads (so every rank forms bitwise the same sum), and solves -- no NCCL launch, no separate reduction kernel.
// Double buffering is enough: a rank can only reach iteration k+2 after every peer has published k+1, i.e. after it
// finished reading iteration k.  A peer that never arrives trips a time-out (status DEVO_ECAPACITY) instead of a hang.
template <int KENT>
__global__ void __launch_bounds__(kSolveThreads, 1) ba_solve_peer_kernel(
    float* poses, const unsigned long long* __restrict__ peer_ptrs, int world, int rank, unsigned long long epoch,
    int parity, size_t nsys, double* sum_buf, double* dX, int32_t* status, int t0, int nfree, int itr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_timeout;
  const int tid = threadIdx.x;
  const int n6 = 6 * nfree;
  const int nent = (n6 + 1) * (n6 + 2) / 2;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid == 0) {                          // this rank's partial (written by the previous kernel of the stream) is complete
    unsigned long long* my_flag = reinterpret_cast<unsigned long long*>(peer_ptrs[rank]) + 2 * nsys + parity;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(my_flag), "l"(epoch) : "memory");
  }
  if (tid < world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(peer_ptrs[tid]) + 2 * nsys + parity;
    long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    unsigned long long v = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      long long t_now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
      if (t_now - t_start > 2000000000LL) { s_timeout = 1; break; }       // 2 s: a peer is gone
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (s_timeout) {
    if (tid == 0) atomicCAS(status, 0, DEVO_ECAPACITY);
    return;
  }
  for (int idx = tid; idx <= nent; idx += kSolveThreads) {                // rank order => identical bits on every rank
    double acc = 0.0;
    for (int p = 0; p < world; p++) {
      const double* src = reinterpret_cast<const double*>(peer_ptrs[p]) + (size_t)parity * nsys;
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + idx) : "memory");
      acc += v;
    }
    sum_buf[idx] = acc;
  }
  __threadfence();
  __syncthreads();
  if (sum_buf[nent] != 0.0) {              // some rank failed in accumulate: every rank stops here, identically
    if (tid == 0) atomicCAS(status, 0, DEVO_ECAPACITY);
    return;
  }
  ba_solve_device<KENT>(smem_raw, poses, sum_buf, dX, status, 1, t0, nfree, itr);
}

}  // namespace

namespace {
// ---- reproject (:368-418) ---------------------------------------------------------------------
__global__ void reproject_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                                 const float* __restrict__ intrinsics, const int64_t* __restrict__ ii,
                                 const int64_t* __restrict__ jj, const int64_t* __restrict__ kk,
                                 float* __restrict__ coords, int E, int PP) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * PP) return;
  const int n = t / PP, p = t - n * PP;
  const float fx = intrinsics[0], fy = intrinsics[1], cx = intrinsics[2], cy = intrinsics[3];
  const int i = (int)ii[n], j = (int)jj[n], k = (int)kk[n];
  float Pi[7], Pj[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { Pi[c] = poses[(size_t)i * 7 + c]; Pj[c] = poses[(size_t)j * 7 + c]; }
  float tij[3], qij[4];
  rel_pose(Pi, Pj, tij, qij);
  const float* pk = patches + (size_t)k * 3 * PP;
  float Xi[3] = {(pk[p] - cx) / fx, (pk[PP + p] - cy) / fy, 1.0f};
  const float d = pk[2 * PP + p];
  float Xj[3];
  rot(qij, Xi, Xj);
  Xj[0] += d * tij[0]; Xj[1] += d * tij[1]; Xj[2] += d * tij[2];
  coords[(size_t)n * 2 * PP + p] = fx * (Xj[0] / Xj[2]) + cx;
  coords[(size_t)n * 2 * PP + PP + p] = fy * (Xj[1] / Xj[2]) + cy;
}

}  // namespace

// =============================================================================================
#include "lie.cuh"
namespace {
// fused projective_ops.transform forward (devo/projective_ops.py:53-105), lietorch semantics
// (quaternions renormalised on every construction), one thread per (edge, pixel)
__global__ void transform_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                                 const float* __restrict__ intrinsics, const int64_t* __restrict__ ii,
                                 const int64_t* __restrict__ jj, const int64_t* __restrict__ kk,
                                 float* __restrict__ coords, float* __restrict__ valid, float* __restrict__ Ji_o,
                                 float* __restrict__ Jj_o, float* __restrict__ Jz_o, int E, int P, int layout, int tonly) {
  const int PP = P * P;
  DEVO_PDL_TRIGGER();   // the lookup kernel that follows may set up its barriers / TMEM while this grid runs
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * PP) return;
  const int n = t / PP, p = t - n * PP;
  const int i = (int)ii[n], j = (int)jj[n], k = (int)kk[n];
  float di[7], dj[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { di[c] = poses[(size_t)i * 7 + c]; dj[c] = poses[(size_t)j * 7 + c]; }
  lie::SE3<float> Gi = lie::SE3<float>::load(di), Gj = lie::SE3<float>::load(dj);
  lie::SE3<float> Gij = Gj * Gi.inv();
  if (tonly) { Gij.R.q = lie::Quat<float>{0.f, 0.f, 0.f, 1.f}; }
  const float* Ki = intrinsics + (size_t)i * 4;
  const float* Kj = intrinsics + (size_t)j * 4;
  const float* pk = patches + (size_t)k * 3 * PP;
  float X0[4] = {(pk[p] - Ki[2]) / Ki[0], (pk[PP + p] - Ki[3]) / Ki[1], 1.0f, pk[2 * PP + p]};
  float X1[4];
  Gij.act4(X0, X1);
  const float fx = Kj[0], fy = Kj[1], cx = Kj[2], cy = Kj[3];
  const float dd = 1.0f / fmaxf(X1[2], 0.1f);
  const float x = fx * (dd * X1[0]) + cx;
  const float y = fy * (dd * X1[1]) + cy;
  if (layout == 0) {
    coords[((size_t)n * PP + p) * 2 + 0] = x;
    coords[((size_t)n * PP + p) * 2 + 1] = y;
  } else {
    coords[(size_t)n * 2 * PP + p] = x;
    coords[(size_t)n * 2 * PP + PP + p] = y;
  }
  const int centre = (P / 2) * P + (P / 2);
  if (p == centre) {
    const float X = X1[0], Y = X1[1], Z = X1[2], H = X1[3];
    if (valid) valid[n] = (Z > 0.2f) ? 1.0f : 0.0f;
    if (Jj_o) {
      const float d = (fabsf(Z) > 0.2f) ? 1.0f / Z : 0.0f;
      // Jj = Jp * Ja  (2x4 * 4x6)
      const float a0 = fx * d, a2 = -fx * X * d * d, b1 = fy * d, b2 = -fy * Y * d * d;
      float J[2][6];
      J[0][0] = a0 * H; J[0][1] = 0.f;    J[0][2] = a2 * H; J[0][3] = a2 * Y;           J[0][4] = a0 * Z - a2 * X; J[0][5] = -a0 * Y;
      J[1][0] = 0.f;    J[1][1] = b1 * H; J[1][2] = b2 * H; J[1][3] = -b1 * Z + b2 * Y; J[1][4] = -b2 * X;         J[1][5] = b1 * X;
      lie::Mat<float, 6, 6> A = Gij.Adj();
#pragma unroll
      for (int r = 0; r < 2; r++) {
        float o[6];
        lie::matTvec(A, J[r], o);
#pragma unroll
        for (int c = 0; c < 6; c++) {
          Jj_o[((size_t)n * 2 + r) * 6 + c] = J[r][c];
          if (Ji_o) Ji_o[((size_t)n * 2 + r) * 6 + c] = -o[c];
        }
      }
      if (Jz_o) {
        Jz_o[(size_t)n * 2 + 0] = a0 * Gij.t[0] + a2 * Gij.t[2];
        Jz_o[(size_t)n * 2 + 1] = b1 * Gij.t[1] + b2 * Gij.t[2];
      }
    }
  }
}
}  // namespace

template <int EPT>
static int launch_accumulate(const BaLayout& L, char* w, float* poses, float* patches, const float* intrinsics,
                             const float* target, const float* weight, const float* lmbda, const int64_t* ii,
                             const int64_t* jj, const int64_t* kk, int32_t* status, int E, int PP, int centre,
                             int t0, int nfree, int n_poses, int EB, int GB, size_t smem, int apply_update,
                             int do_accumulate, int itr, cudaStream_t s, const int32_t* perm_p, const int32_t* gstart_p,
                             const int64_t* gkey_p, const int32_t* ngroups_p, double* sys_out = nullptr,
                             int plan_is_older = 0, int32_t* status_or = nullptr) {
  static devo::SmemConfig configured;
  if (configured.need(smem)) {
    DEVO_CUDA(cudaFuncSetAttribute(ba_accumulate_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  DEVO_CUDA(devo::launch_pdl_cluster(ba_accumulate_kernel<EPT>, dim3(L.grid), dim3(kAccThreads), kGroupCtas, smem, s,
      poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, perm_p, gstart_p, gkey_p, ngroups_p,
      (double*)(w + L.Q), (double*)(w + L.U), (double*)(w + L.Ek), (double*)(w + L.partials),
      (double*)(w + L.dX), (int32_t*)(w + L.ticket), status, E, PP, centre, t0, nfree, n_poses, EB, GB, apply_update, do_accumulate, itr, sys_out,
      plan_is_older, status_or));
  DEVO_LAUNCH_CHECK("ba_accumulate");
  return DEVO_OK;
}

extern "C" {

#ifdef DEVO_BA_TIMING
int devo_ba_debug_clocks(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_ba_clk, sizeof(long long) * 24); }
#endif

size_t devo_ba_workspace(int E, int n_free_poses) { return ba_layout(E, n_free_poses).total; }

// The two clears are ONE 32-thread kernel, not memset nodes: inside a captured step a memset node runs on a copy engine,
// and when the caller uploads the next step's inputs meanwhile (one 5 MB H2D copy, ~100 us) the 4-byte memset queued
// behind that copy and held the update operator back by 35-60 us per step (tools/e2e_probe.py).
__global__ void ba_prepare_kernel(int32_t* __restrict__ status, int32_t* __restrict__ ticket) {
  if (threadIdx.x == 0) *status = 0;
  if (ticket != nullptr) ticket[threadIdx.x] = 0;
}


static int ba_forward_impl(float* poses, float* patches, const float* intrinsics, const float* target,
                           const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                           const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1,
                           int iterations, void* workspace, size_t workspace_bytes, int32_t* status, void* stream,
                           const int32_t* ext_perm, const int32_t* ext_gstart, const int64_t* ext_gkey,
                           const int32_t* ext_ngroups, bool prepared = false, int32_t* status_or = nullptr) {
  cudaStream_t s = (cudaStream_t)stream;
  DEVO_REQUIRE(status != nullptr, DEVO_EINVAL, "ba_forward: status pointer is NULL");
  if (E <= 0 || iterations <= 0) {
    if (!prepared) { ba_prepare_kernel<<<1, 32, 0, s>>>(status, nullptr); DEVO_LAUNCH_CHECK("ba_prepare"); }
    return DEVO_OK;
  }
  int nfree = t1 - t0;
  if (nfree < 0) nfree = 0;
  DEVO_REQUIRE(P >= 1 && P <= 8, DEVO_EINVAL, "ba_forward: patch size %d unsupported", P);
  DEVO_REQUIRE(6 * nfree <= kMaxN6, DEVO_ECAPACITY,
               "ba_forward: %d free poses exceed the in-shared-memory solver capacity (%d)", nfree, kMaxN6 / 6);
  DEVO_REQUIRE(t0 >= 0 && t1 <= n_poses, DEVO_EINVAL, "ba_forward: pose window [%d,%d) outside [0,%d)", t0, t1, n_poses);
  BaLayout L = ba_layout(E, nfree);
  DEVO_REQUIRE(workspace && workspace_bytes >= L.total, DEVO_EWORKSPACE,
               "ba_forward: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  char* w = (char*)workspace;
  const int PP = P * P;
  const int centre = (P == 3) ? 4 : (1 * P + 1 < PP ? 1 * P + 1 : 0);   // the reference hard-codes [1][1]
  // patches grouped by kk (== torch::_unique(kk), ba_cuda.cu:435-437), edges of a patch contiguous
  int rc = DEVO_OK;
  const int32_t* perm_p = (const int32_t*)(w + L.perm);
  const int32_t* gstart_p = (const int32_t*)(w + L.gstart);
  const int64_t* gkey_p = (const int64_t*)(w + L.gkey);
  const int32_t* ngroups_p = (const int32_t*)(w + L.ngroups);
  if (ext_perm) {   // the caller already analysed this edge list (devo_graph_plan on (kk, jj))
    perm_p = ext_perm; gstart_p = ext_gstart; gkey_p = ext_gkey; ngroups_p = ext_ngroups;
  } else {
    rc = devo_graph_plan(kk, jj, E, n_patches, n_poses, (int32_t*)(w + L.perm), nullptr,
                         (int32_t*)(w + L.gstart), (int64_t*)(w + L.gkey), (int32_t*)(w + L.ngroups),
                         nullptr, nullptr, w + L.plan_ws, L.plan_bytes, stream);
    if (rc != DEVO_OK) return rc;
  }

  const int n6 = L.n6;
  int EB = 0, GB = 0;
  size_t smem_acc = 0;
  DEVO_REQUIRE(acc_shape(n6, nfree, EB, GB, smem_acc), DEVO_ECAPACITY, "ba_forward: system too large for shared memory");
  const int ept = (L.nent + kAccThreads - 1) / kAccThreads;

#define ACC(EPT_, APPLY, DOACC, ITR)                                                                         \
  launch_accumulate<EPT_>(L, w, poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, status, E, PP, \
                          centre, t0, nfree, n_poses, EB, GB, smem_acc, APPLY, DOACC, ITR, s, perm_p, gstart_p, gkey_p, ngroups_p, \
                          nullptr, prepared ? 1 : 0, status_or)
#define ACC_DISPATCH(APPLY, DOACC, ITR)                   \
  do {                                                    \
    if (ept <= 2) rc = ACC(2, APPLY, DOACC, ITR);         \
    else if (ept <= 4) rc = ACC(4, APPLY, DOACC, ITR);    \
    else if (ept <= 8) rc = ACC(8, APPLY, DOACC, ITR);    \
    else if (ept <= 16) rc = ACC(16, APPLY, DOACC, ITR);  \
    else rc = ACC(24, APPLY, DOACC, ITR);                 \
    if (rc != DEVO_OK) return rc;                         \
  } while (0)

  DEVO_REQUIRE(ept <= 24, DEVO_ECAPACITY, "ba_forward: system too large (%d entries)", L.nent);
  const size_t smem_solve = ((size_t)n6 * (n6 + 1) / 2 + (n6 + 1) + 18 * (n6 + 1) + 8) * 8;   // A, y, 2 column blocks + W
  DEVO_REQUIRE(smem_solve <= smem_acc, DEVO_ECAPACITY, "ba_forward: solver does not fit the accumulate CTA's shared memory");
  if (!prepared) {     // status word + ticket area: one small kernel (no copy-engine work, see devo_ba_prepare)
    ba_prepare_kernel<<<1, 32, 0, s>>>(status, (int32_t*)(w + L.ticket));
    DEVO_LAUNCH_CHECK("ba_prepare");
  }
  for (int itr = 0; itr < iterations; itr++) {
    ACC_DISPATCH(itr > 0 ? 1 : 0, 1, itr);   // accumulate + (last CTA) solve + retraction
  }
  ACC_DISPATCH(1, 0, iterations);   // final depth update only
#undef ACC
#undef ACC_DISPATCH
  return DEVO_OK;
}

// ---- edge-sharded BA: one frame graph split over ranks by owning patch ---------------------------
// Depth blocks C,u and the columns of E are rank-local, so the Schur complement distributes:
//   S = sum_g (B_g - E_g Q_g E_g^T),  y = sum_g (v_g - E_g Q_g u_g).
// Per Gauss-Newton iteration: devo_ba_sharded_accumulate (local edges -> sys_out, fp64) -> ONE all-reduce of
// devo_ba_system_doubles(nfree) doubles over the ranks (NCCL, host side) -> devo_ba_sharded_solve (identical on every
// rank: damping, LDL^T, retraction of the replicated poses).  The depth update of the local patches is applied in the
// prologue of the next accumulate call (flags bit 0) or by a final call with flags = 1 (no accumulate).
size_t devo_ba_system_doubles(int n_free_poses) {
  const int n6 = 6 * (n_free_poses > 0 ? n_free_poses : 0);
  return (size_t)(n6 + 1) * (n6 + 2) / 2 + 1;   // upper triangle of [S|y] (+ the y^T y corner) + a status word
}

int devo_ba_sharded_accumulate(float* poses, float* patches, const float* intrinsics, const float* target,
                               const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                               const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1, int itr,
                               int flags, double* sys_out, void* workspace, size_t workspace_bytes, int32_t* status,
                               void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int apply_update = flags & 1, do_accumulate = (flags >> 1) & 1, replan = (flags >> 2) & 1;
  DEVO_REQUIRE(status != nullptr, DEVO_EINVAL, "ba_sharded_accumulate: status pointer is NULL");
  const int nfree = t1 - t0;
  DEVO_REQUIRE(nfree > 0, DEVO_EINVAL, "ba_sharded_accumulate: needs at least one free pose (structure-only BA has no exchange)");
  DEVO_REQUIRE(P >= 1 && P <= 8, DEVO_EINVAL, "ba_sharded_accumulate: patch size %d unsupported", P);
  DEVO_REQUIRE(6 * nfree <= kMaxN6, DEVO_ECAPACITY, "ba_sharded_accumulate: %d free poses exceed the solver capacity (%d)",
               nfree, kMaxN6 / 6);
  DEVO_REQUIRE(t0 >= 0 && t1 <= n_poses, DEVO_EINVAL, "ba_sharded_accumulate: pose window [%d,%d) outside [0,%d)", t0, t1, n_poses);
  DEVO_REQUIRE(!do_accumulate || sys_out != nullptr, DEVO_EINVAL, "ba_sharded_accumulate: sys_out is NULL");
  BaLayout L = ba_layout(E, nfree);
  if (replan) DEVO_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  if (E <= 0) {   // a rank without edges contributes a zero system
    if (do_accumulate) DEVO_CUDA(cudaMemsetAsync(sys_out, 0, devo_ba_system_doubles(nfree) * 8, s));
    return DEVO_OK;
  }
  DEVO_REQUIRE(workspace && workspace_bytes >= L.total, DEVO_EWORKSPACE,
               "ba_sharded_accumulate: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  char* w = (char*)workspace;
  const int PP = P * P;
  const int centre = (P == 3) ? 4 : (1 * P + 1 < PP ? 1 * P + 1 : 0);
  int rc = DEVO_OK;
  if (replan) {
    rc = devo_graph_plan(kk, jj, E, n_patches, n_poses, (int32_t*)(w + L.perm), nullptr, (int32_t*)(w + L.gstart),
                         (int64_t*)(w + L.gkey), (int32_t*)(w + L.ngroups), nullptr, nullptr, w + L.plan_ws,
                         L.plan_bytes, stream);
    if (rc != DEVO_OK) return rc;
    DEVO_CUDA(cudaMemsetAsync(w + L.ticket, 0, 4 * 32, s));
  }
  const int n6 = L.n6;
  int EB = 0, GB = 0;
  size_t smem_acc = 0;
  DEVO_REQUIRE(acc_shape(n6, nfree, EB, GB, smem_acc), DEVO_ECAPACITY, "ba_sharded_accumulate: system too large for shared memory");
  const int ept = (L.nent + kAccThreads - 1) / kAccThreads;
  DEVO_REQUIRE(ept <= 24, DEVO_ECAPACITY, "ba_sharded_accumulate: system too large (%d entries)", L.nent);
#define ACCS(EPT_)                                                                                                \
  launch_accumulate<EPT_>(L, w, poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, status, E, PP, centre, \
                          t0, nfree, n_poses, EB, GB, smem_acc, apply_update, do_accumulate, itr, s,                \
                          (const int32_t*)(w + L.perm), (const int32_t*)(w + L.gstart), (const int64_t*)(w + L.gkey), \
                          (const int32_t*)(w + L.ngroups), sys_out)
  if (ept <= 2) rc = ACCS(2);
  else if (ept <= 4) rc = ACCS(4);
  else if (ept <= 8) rc = ACCS(8);
  else if (ept <= 16) rc = ACCS(16);
  else rc = ACCS(24);
#undef ACCS
  return rc;
}

int devo_ba_sharded_solve(float* poses, const double* sys, int E, int n_poses, int t0, int t1, int itr, void* workspace,
                          size_t workspace_bytes, int32_t* status, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int nfree = t1 - t0;
  DEVO_REQUIRE(status != nullptr && sys != nullptr, DEVO_EINVAL, "ba_sharded_solve: NULL pointer");
  DEVO_REQUIRE(nfree > 0 && 6 * nfree <= kMaxN6, DEVO_ECAPACITY, "ba_sharded_solve: %d free poses unsupported", nfree);
  DEVO_REQUIRE(t0 >= 0 && t1 <= n_poses, DEVO_EINVAL, "ba_sharded_solve: pose window [%d,%d) outside [0,%d)", t0, t1, n_poses);
  BaLayout L = ba_layout(E, nfree);   // same E as the accumulate calls: dX is read back by their depth-update prologue
  DEVO_REQUIRE(workspace && workspace_bytes >= L.total, DEVO_EWORKSPACE,
               "ba_sharded_solve: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  const int n6 = 6 * nfree;
  const size_t smem = ((size_t)n6 * (n6 + 1) / 2 + (n6 + 1) + 18 * (n6 + 1) + 8) * 8;   // A, y, 2 column blocks + W
  double* dX = (double*)((char*)workspace + L.dX);
#define SOLVE(K_)                                                                                              \
  do {                                                                                                         \
    static devo::SmemConfig configured;                                                                        \
    if (configured.need(smem)) {                                                                               \
      DEVO_CUDA(cudaFuncSetAttribute(ba_solve_kernel<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    }                                                                                                          \
    ba_solve_kernel<K_><<<1, kSolveThreads, smem, s>>>(poses, sys, dX, status, t0, nfree, itr);                \
  } while (0)
  if (n6 <= 44) SOLVE(2);
  else if (n6 <= 88) SOLVE(8);
  else SOLVE((((kMaxN6 + 1) * (kMaxN6 + 2) / 2) + kSolveThreads - 1) / kSolveThreads);
#undef SOLVE
  DEVO_LAUNCH_CHECK("ba_sharded_solve");
  return DEVO_OK;
}

int devo_ba_sharded_solve_peer(float* poses, const void* peer_ptrs_dev, int world, int rank, uint64_t epoch, int E,
                               int n_poses, int t0, int t1, int itr, void* workspace, size_t workspace_bytes,
                               int32_t* status, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int nfree = t1 - t0;
  DEVO_REQUIRE(status != nullptr && peer_ptrs_dev != nullptr, DEVO_EINVAL, "ba_sharded_solve_peer: NULL pointer");
  DEVO_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, DEVO_EINVAL, "ba_sharded_solve_peer: bad rank/world");
  DEVO_REQUIRE(nfree > 0 && 6 * nfree <= kMaxN6, DEVO_ECAPACITY, "ba_sharded_solve_peer: %d free poses unsupported", nfree);
  DEVO_REQUIRE(t0 >= 0 && t1 <= n_poses, DEVO_EINVAL, "ba_sharded_solve_peer: pose window [%d,%d) outside [0,%d)", t0, t1, n_poses);
  BaLayout L = ba_layout(E, nfree);
  DEVO_REQUIRE(workspace && workspace_bytes >= L.total, DEVO_EWORKSPACE,
               "ba_sharded_solve_peer: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  const int n6 = 6 * nfree;
  const size_t nsys = devo_ba_system_doubles(nfree);
  const size_t smem = ((size_t)n6 * (n6 + 1) / 2 + (n6 + 1) + 18 * (n6 + 1) + 8) * 8;
  double* dX = (double*)((char*)workspace + L.dX);
  double* sum_buf = (double*)((char*)workspace + L.partials);   // the per-CTA partials are dead once the system is exported
  const int parity = (int)(epoch & 1);
#define SOLVEP(K_)                                                                                              \
  do {                                                                                                          \
    static devo::SmemConfig configured;                                                                         \
    if (configured.need(smem)) {                                                                                \
      DEVO_CUDA(cudaFuncSetAttribute(ba_solve_peer_kernel<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    }                                                                                                           \
    ba_solve_peer_kernel<K_><<<1, kSolveThreads, smem, s>>>(poses, (const unsigned long long*)peer_ptrs_dev, world, rank, \
                                                            (unsigned long long)epoch, parity, nsys, sum_buf, dX, status, t0, nfree, itr); \
  } while (0)
  if (n6 <= 44) SOLVEP(2);
  else if (n6 <= 88) SOLVEP(8);
  else SOLVEP((((kMaxN6 + 1) * (kMaxN6 + 2) / 2) + kSolveThreads - 1) / kSolveThreads);
#undef SOLVEP
  DEVO_LAUNCH_CHECK("ba_sharded_solve_peer");
  return DEVO_OK;
}

int devo_ba_forward(float* poses, float* patches, const float* intrinsics, const float* target,
                    const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                    const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1,
                    int iterations, void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, E, n_poses, n_patches, P, t0, t1,
                         iterations, workspace, workspace_bytes, status, stream, nullptr, nullptr, nullptr, nullptr);
}

int devo_ba_forward_planned(float* poses, float* patches, const float* intrinsics, const float* target,
                            const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                            const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1,
                            int iterations, const int32_t* perm, const int32_t* gstart, const int64_t* gkey,
                            const int32_t* ngroups, void* workspace, size_t workspace_bytes, int32_t* status,
                            void* stream) {
  DEVO_REQUIRE(perm && gstart && gkey && ngroups, DEVO_EINVAL, "ba_forward_planned: plan pointers must not be NULL");
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, E, n_poses, n_patches, P, t0, t1,
                         iterations, workspace, workspace_bytes, status, stream, perm, gstart, gkey, ngroups);
}

// ---- the same call with its housekeeping moved off the critical path -------------------------------------------------
// Between the update operator and the first Gauss-Newton launch devo_ba_forward_planned issues two memsets (status,
// ticket) -- ~4 us of serialised DMA nodes that also break the programmatic launch chain -- and the engine followed the
// call with a one-thread kernel that ORs the status into its sticky word.  devo_ba_prepare does the two memsets whenever
// the caller likes (the engine: on the side stream that analyses the graph, while reprojection / lookup / update operator
// run); devo_ba_forward_prepared then launches nothing but the iterations, reads the plan ahead of the programmatic wait
// from the first iteration on (the caller vouches the plan is older than the preceding kernel of the stream) and folds
// the status into `status_or` (may be NULL) in its last launch.
int devo_ba_prepare(void* workspace, size_t workspace_bytes, int E, int n_free_poses, int32_t* status, void* stream) {
  DEVO_REQUIRE(status != nullptr, DEVO_EINVAL, "ba_prepare: status pointer is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* ticket = nullptr;
  if (E > 0) {
    BaLayout L = ba_layout(E, n_free_poses > 0 ? n_free_poses : 0);
    DEVO_REQUIRE(workspace && workspace_bytes >= L.total, DEVO_EWORKSPACE, "ba_prepare: workspace too small (%zu < %zu)",
                 workspace_bytes, L.total);
    ticket = (int32_t*)((char*)workspace + L.ticket);
  }
  ba_prepare_kernel<<<1, 32, 0, s>>>(status, ticket);
  DEVO_LAUNCH_CHECK("ba_prepare");
  return DEVO_OK;
}

int devo_ba_forward_prepared(float* poses, float* patches, const float* intrinsics, const float* target,
                             const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                             const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1,
                             int iterations, const int32_t* perm, const int32_t* gstart, const int64_t* gkey,
                             const int32_t* ngroups, void* workspace, size_t workspace_bytes, int32_t* status,
                             int32_t* status_or, void* stream) {
  DEVO_REQUIRE(perm && gstart && gkey && ngroups, DEVO_EINVAL, "ba_forward_prepared: plan pointers must not be NULL");
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, E, n_poses, n_patches, P, t0, t1,
                         iterations, workspace, workspace_bytes, status, stream, perm, gstart, gkey, ngroups, true, status_or);
}

int devo_reproject(const float* poses, const float* patches, const float* intrinsics, const int64_t* ii,
                   const int64_t* jj, const int64_t* kk, float* coords, int E, int P, void* stream) {
  if (E <= 0) return DEVO_OK;
  const int PP = P * P;
  reproject_kernel<<<devo::cdiv((long long)E * PP, 256), 256, 0, (cudaStream_t)stream>>>(
      poses, patches, intrinsics, ii, jj, kk, coords, E, PP);
  DEVO_LAUNCH_CHECK("reproject");
  return DEVO_OK;
}

int devo_transform_forward(const float* poses, const float* patches, const float* intrinsics,
                           const int64_t* ii, const int64_t* jj, const int64_t* kk, float* coords_out,
                           float* valid_out, float* Ji, float* Jj, float* Jz, int E, int P, int layout,
                           int tonly, void* stream) {
  if (E <= 0) return DEVO_OK;
  DEVO_REQUIRE(coords_out != nullptr, DEVO_EINVAL, "transform_forward: coords_out is NULL");
  DEVO_REQUIRE(!(Ji && !Jj), DEVO_EINVAL, "transform_forward: Ji requires Jj");
  transform_kernel<<<devo::cdiv((long long)E * P * P, 256), 256, 0, (cudaStream_t)stream>>>(
      poses, patches, intrinsics, ii, jj, kk, coords_out, valid_out, Ji, Jj, Jz, E, P, layout, tonly);
  DEVO_LAUNCH_CHECK("transform_forward");
  return DEVO_OK;
}

}  // extern "C"
