// lie_ops.cu -- the 19 batched Lie-group ops of `lietorch_backends`
// (reference interface: devo/lietorch/src/lietorch.cpp:286-316; op semantics
// lietorch_gpu.cu:20-294) for SO3 / RxSO3 / SE3 / Sim3 x {float,double}.
//
// One group element per thread, whole element in registers (lie.cuh), grid-stride loop
// sized to a multiple of the SM count.  Elements are 4..8 scalars (16..64 B), so the ops
// are pure streaming: the launch is latency/HBM bound and the arithmetic is free.
#include "common.cuh"
#include "lie.cuh"

namespace {

using namespace lie;

constexpr int kThreads = 256;

static int lie_grid(int64_t n) {
  int64_t blocks = (n + kThreads - 1) / kThreads;
  const int64_t cap = 148 * 8;   // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

#define LIE_LOOP(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)blockDim.x * gridDim.x)

template <typename T, int L> __device__ __forceinline__ void ld(const T* p, T* r) {
#pragma unroll
  for (int k = 0; k < L; k++) r[k] = p[k];
}
template <typename T, int L> __device__ __forceinline__ void st(T* p, const T* r) {
#pragma unroll
  for (int k = 0; k < L; k++) p[k] = r[k];
}
// gradient w.r.t. a group element: K values then zeros up to N
template <typename G, typename T> __device__ __forceinline__ void st_grad(T* p, const T* r) {
#pragma unroll
  for (int k = 0; k < G::K; k++) p[k] = r[k];
#pragma unroll
  for (int k = G::K; k < G::N; k++) p[k] = T(0);
}

template <typename G, typename T> __global__ void k_exp(const T* a_, T* X_, int64_t n) {
  LIE_LOOP(i, n) { T a[G::K]; ld<T, G::K>(a_ + i * G::K, a); T X[G::N]; G::Exp(a).store(X); st<T, G::N>(X_ + i * G::N, X); }
}
template <typename G, typename T> __global__ void k_exp_bwd(const T* g_, const T* a_, T* da_, int64_t n) {
  LIE_LOOP(i, n) {
    T a[G::K], g[G::K], da[G::K];
    ld<T, G::K>(a_ + i * G::K, a); ld<T, G::K>(g_ + i * G::N, g);
    vecmat(g, G::left_jacobian(a), da);
    st<T, G::K>(da_ + i * G::K, da);
  }
}
template <typename G, typename T> __global__ void k_log(const T* X_, T* a_, int64_t n) {
  LIE_LOOP(i, n) { T X[G::N]; ld<T, G::N>(X_ + i * G::N, X); T a[G::K]; G::load(X).Log(a); st<T, G::K>(a_ + i * G::K, a); }
}
template <typename G, typename T> __global__ void k_log_bwd(const T* g_, const T* X_, T* dX_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], g[G::K], a[G::K], d[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(g_ + i * G::K, g);
    G::load(X).Log(a);
    vecmat(g, G::left_jacobian_inverse(a), d);
    st_grad<G, T>(dX_ + i * G::N, d);
  }
}
template <typename G, typename T> __global__ void k_inv(const T* X_, T* Y_, int64_t n) {
  LIE_LOOP(i, n) { T X[G::N]; ld<T, G::N>(X_ + i * G::N, X); T Y[G::N]; G::load(X).inv().store(Y); st<T, G::N>(Y_ + i * G::N, Y); }
}
template <typename G, typename T> __global__ void k_inv_bwd(const T* g_, const T* X_, T* dX_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], g[G::K], d[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(g_ + i * G::N, g);
    vecmat(g, G::load(X).inv().Adj(), d);
#pragma unroll
    for (int k = 0; k < G::K; k++) d[k] = -d[k];
    st_grad<G, T>(dX_ + i * G::N, d);
  }
}
template <typename G, typename T> __global__ void k_mul(const T* X_, const T* Y_, T* Z_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], Y[G::N], Z[G::N];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::N>(Y_ + i * G::N, Y);
    (G::load(X) * G::load(Y)).store(Z);
    st<T, G::N>(Z_ + i * G::N, Z);
  }
}
template <typename G, typename T> __global__ void k_mul_bwd(const T* g_, const T* X_, const T* Y_, T* dX_, T* dY_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], g[G::K], d[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(g_ + i * G::N, g);
    st_grad<G, T>(dX_ + i * G::N, g);
    vecmat(g, G::load(X).Adj(), d);
    st_grad<G, T>(dY_ + i * G::N, d);
  }
}
template <typename G, typename T> __global__ void k_adj(const T* X_, const T* a_, T* b_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], a[G::K], b[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(a_ + i * G::K, a);
    matvec(G::load(X).Adj(), a, b);
    st<T, G::K>(b_ + i * G::K, b);
  }
}
template <typename G, typename T> __global__ void k_adj_bwd(const T* g_, const T* X_, const T* a_, T* dX_, T* da_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], a[G::K], g[G::K], b[G::K], da[G::K], dX[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(a_ + i * G::K, a); ld<T, G::K>(g_ + i * G::K, g);
    auto A = G::load(X).Adj();
    matvec(A, a, b);
    vecmat(g, A, da);
    vecmat(g, G::adj(b), dX);
#pragma unroll
    for (int k = 0; k < G::K; k++) dX[k] = -dX[k];
    st<T, G::K>(da_ + i * G::K, da);
    st_grad<G, T>(dX_ + i * G::N, dX);
  }
}
template <typename G, typename T> __global__ void k_adjT(const T* X_, const T* a_, T* b_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], a[G::K], b[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(a_ + i * G::K, a);
    matTvec(G::load(X).Adj(), a, b);
    st<T, G::K>(b_ + i * G::K, b);
  }
}
template <typename G, typename T> __global__ void k_adjT_bwd(const T* g_, const T* X_, const T* a_, T* dX_, T* da_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], a[G::K], g[G::K], Ag[G::K], dX[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(a_ + i * G::K, a); ld<T, G::K>(g_ + i * G::K, g);
    matvec(G::load(X).Adj(), g, Ag);
    vecmat(a, G::adj(Ag), dX);
#pragma unroll
    for (int k = 0; k < G::K; k++) dX[k] = -dX[k];
    st<T, G::K>(da_ + i * G::K, Ag);
    st_grad<G, T>(dX_ + i * G::N, dX);
  }
}
template <typename G, typename T> __global__ void k_act(const T* X_, const T* p_, T* q_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], p[3], q[3];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, 3>(p_ + i * 3, p);
    G::load(X).act(p, q);
    st<T, 3>(q_ + i * 3, q);
  }
}
template <typename G, typename T> __global__ void k_act_bwd(const T* g_, const T* X_, const T* p_, T* dX_, T* dp_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], p[3], q[3], g[3], dp[3], dX[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, 3>(p_ + i * 3, p); ld<T, 3>(g_ + i * 3, g);
    G Xg = G::load(X);
    Xg.act(p, q);
    auto M = Xg.Matrix4();
#pragma unroll
    for (int c = 0; c < 3; c++) dp[c] = g[0] * M(0, c) + g[1] * M(1, c) + g[2] * M(2, c);
    vecmat(g, G::act_jacobian(q), dX);
    st<T, 3>(dp_ + i * 3, dp);
    st_grad<G, T>(dX_ + i * G::N, dX);
  }
}
template <typename G, typename T> __global__ void k_act4(const T* X_, const T* p_, T* q_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], p[4], q[4];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, 4>(p_ + i * 4, p);
    G::load(X).act4(p, q);
    st<T, 4>(q_ + i * 4, q);
  }
}
template <typename G, typename T> __global__ void k_act4_bwd(const T* g_, const T* X_, const T* p_, T* dX_, T* dp_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], p[4], q[4], g[4], dp[4], dX[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, 4>(p_ + i * 4, p); ld<T, 4>(g_ + i * 4, g);
    G Xg = G::load(X);
    Xg.act4(p, q);
    vecmat(g, Xg.Matrix4(), dp);
    vecmat(g, G::act4_jacobian(q), dX);
    st<T, 4>(dp_ + i * 4, dp);
    st_grad<G, T>(dX_ + i * G::N, dX);
  }
}
template <typename G, typename T> __global__ void k_matrix(const T* X_, T* M_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N]; ld<T, G::N>(X_ + i * G::N, X);
    auto M = G::load(X).Matrix4();
    st<T, 16>(M_ + i * 16, M.m);
  }
}
template <typename G, typename T> __global__ void k_projector(const T* X_, T* P_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N]; ld<T, G::N>(X_ + i * G::N, X);
    auto P = G::load(X).projector();
    st<T, G::N * G::N>(P_ + i * G::N * G::N, P.m);
  }
}
template <typename G, typename T> __global__ void k_jinv(const T* X_, const T* a_, T* b_, int64_t n) {
  LIE_LOOP(i, n) {
    T X[G::N], a[G::K], l[G::K], b[G::K];
    ld<T, G::N>(X_ + i * G::N, X); ld<T, G::K>(a_ + i * G::K, a);
    G::load(X).Log(l);
    matvec(G::left_jacobian_inverse(l), a, b);
    st<T, G::K>(b_ + i * G::K, b);
  }
}

// ---- dispatch over (group, dtype) ------------------------------------------------------
#define LIE_DISPATCH(NAME, KERNEL, ...)                                                          \
  do {                                                                                           \
    if (n <= 0) return DEVO_OK;                                                                  \
    cudaStream_t s = (cudaStream_t)stream;                                                       \
    int grid = lie_grid(n);                                                                      \
    if (dtype == DEVO_F32) {                                                                     \
      using T = float;                                                                           \
      switch (group) {                                                                           \
        case DEVO_SO3:   KERNEL<SO3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;        \
        case DEVO_RXSO3: KERNEL<RxSO3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;      \
        case DEVO_SE3:   KERNEL<SE3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;        \
        case DEVO_SIM3:  KERNEL<Sim3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;       \
        default: DEVO_REQUIRE(false, DEVO_EINVAL, NAME ": bad group id %d", group);              \
      }                                                                                          \
    } else if (dtype == DEVO_F64) {                                                              \
      using T = double;                                                                          \
      switch (group) {                                                                           \
        case DEVO_SO3:   KERNEL<SO3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;        \
        case DEVO_RXSO3: KERNEL<RxSO3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;      \
        case DEVO_SE3:   KERNEL<SE3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;        \
        case DEVO_SIM3:  KERNEL<Sim3<T>, T><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;       \
        default: DEVO_REQUIRE(false, DEVO_EINVAL, NAME ": bad group id %d", group);              \
      }                                                                                          \
    } else {                                                                                     \
      DEVO_REQUIRE(false, DEVO_EINVAL, NAME ": dtype must be f32 or f64");                       \
    }                                                                                            \
    DEVO_LAUNCH_CHECK(NAME);                                                                     \
    return DEVO_OK;                                                                              \
  } while (0)

#define CT(p) ((const T*)(p))
#define MT(p) ((T*)(p))

}  // namespace

extern "C" {
int devo_lie_expm(int group, int dtype, const void* a, void* X, int64_t n, void* stream) {
  LIE_DISPATCH("lie_expm", k_exp, CT(a), MT(X), n);
}
int devo_lie_expm_backward(int group, int dtype, const void* grad, const void* a, void* da, int64_t n, void* stream) {
  LIE_DISPATCH("lie_expm_backward", k_exp_bwd, CT(grad), CT(a), MT(da), n);
}
int devo_lie_logm(int group, int dtype, const void* X, void* a, int64_t n, void* stream) {
  LIE_DISPATCH("lie_logm", k_log, CT(X), MT(a), n);
}
int devo_lie_logm_backward(int group, int dtype, const void* grad, const void* X, void* dX, int64_t n, void* stream) {
  LIE_DISPATCH("lie_logm_backward", k_log_bwd, CT(grad), CT(X), MT(dX), n);
}
int devo_lie_inv(int group, int dtype, const void* X, void* Y, int64_t n, void* stream) {
  LIE_DISPATCH("lie_inv", k_inv, CT(X), MT(Y), n);
}
int devo_lie_inv_backward(int group, int dtype, const void* grad, const void* X, void* dX, int64_t n, void* stream) {
  LIE_DISPATCH("lie_inv_backward", k_inv_bwd, CT(grad), CT(X), MT(dX), n);
}
int devo_lie_mul(int group, int dtype, const void* X, const void* Y, void* Z, int64_t n, void* stream) {
  LIE_DISPATCH("lie_mul", k_mul, CT(X), CT(Y), MT(Z), n);
}
int devo_lie_mul_backward(int group, int dtype, const void* grad, const void* X, const void* Y, void* dX, void* dY, int64_t n, void* stream) {
  LIE_DISPATCH("lie_mul_backward", k_mul_bwd, CT(grad), CT(X), CT(Y), MT(dX), MT(dY), n);
}
int devo_lie_adj(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream) {
  LIE_DISPATCH("lie_adj", k_adj, CT(X), CT(a), MT(b), n);
}
int devo_lie_adj_backward(int group, int dtype, const void* grad, const void* X, const void* a, void* dX, void* da, int64_t n, void* stream) {
  LIE_DISPATCH("lie_adj_backward", k_adj_bwd, CT(grad), CT(X), CT(a), MT(dX), MT(da), n);
}
int devo_lie_adjT(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream) {
  LIE_DISPATCH("lie_adjT", k_adjT, CT(X), CT(a), MT(b), n);
}
int devo_lie_adjT_backward(int group, int dtype, const void* grad, const void* X, const void* a, void* dX, void* da, int64_t n, void* stream) {
  LIE_DISPATCH("lie_adjT_backward", k_adjT_bwd, CT(grad), CT(X), CT(a), MT(dX), MT(da), n);
}
int devo_lie_act(int group, int dtype, const void* X, const void* p, void* q, int64_t n, void* stream) {
  LIE_DISPATCH("lie_act", k_act, CT(X), CT(p), MT(q), n);
}
int devo_lie_act_backward(int group, int dtype, const void* grad, const void* X, const void* p, void* dX, void* dp, int64_t n, void* stream) {
  LIE_DISPATCH("lie_act_backward", k_act_bwd, CT(grad), CT(X), CT(p), MT(dX), MT(dp), n);
}
int devo_lie_act4(int group, int dtype, const void* X, const void* p, void* q, int64_t n, void* stream) {
  LIE_DISPATCH("lie_act4", k_act4, CT(X), CT(p), MT(q), n);
}
int devo_lie_act4_backward(int group, int dtype, const void* grad, const void* X, const void* p, void* dX, void* dp, int64_t n, void* stream) {
  LIE_DISPATCH("lie_act4_backward", k_act4_bwd, CT(grad), CT(X), CT(p), MT(dX), MT(dp), n);
}
int devo_lie_as_matrix(int group, int dtype, const void* X, void* T4x4, int64_t n, void* stream) {
  LIE_DISPATCH("lie_as_matrix", k_matrix, CT(X), MT(T4x4), n);
}
int devo_lie_projector(int group, int dtype, const void* X, void* PNxN, int64_t n, void* stream) {
  LIE_DISPATCH("lie_projector", k_projector, CT(X), MT(PNxN), n);
}
int devo_lie_jinv(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream) {
  LIE_DISPATCH("lie_jinv", k_jinv, CT(X), CT(a), MT(b), n);
}
}
