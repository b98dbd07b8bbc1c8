// lie.cuh -- Eigen-free SO3 / RxSO3 / SE3 / Sim3 math for device (and host) code.
//
// Semantics follow the reference headers devo/lietorch/include/{so3,rxso3,se3,sim3}.h
// (formulas, small-angle switches at EPS=1e-6 from common.h:7, quaternion order (x,y,z,w),
// renormalisation on every construction from memory or from a quaternion product), but
// the code is written from scratch around flat register arrays so that a whole element
// lives in registers and every op is a straight-line sequence of FMAs.
#pragma once
#include <math.h>

#define LIE_HD __host__ __device__ __forceinline__

namespace lie {

template <typename T> struct Eps { static constexpr T value = T(1e-6); };
template <typename T> LIE_HD T pi_v() { return T(3.14159265358979323846); }

// ---- tiny fixed-size row-major matrix -------------------------------------------------
template <typename T, int R, int C>
struct Mat {
  T m[R * C];
  LIE_HD T& operator()(int r, int c) { return m[r * C + c]; }
  LIE_HD const T& operator()(int r, int c) const { return m[r * C + c]; }
  LIE_HD void zero() {
#pragma unroll
    for (int i = 0; i < R * C; i++) m[i] = T(0);
  }
  LIE_HD void identity() {
    zero();
#pragma unroll
    for (int i = 0; i < (R < C ? R : C); i++) m[i * C + i] = T(1);
  }
};

template <typename T, int R, int K, int C>
LIE_HD Mat<T, R, C> matmul(const Mat<T, R, K>& A, const Mat<T, K, C>& B) {
  Mat<T, R, C> O;
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int c = 0; c < C; c++) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < K; k++) s += A(r, k) * B(k, c);
      O(r, c) = s;
    }
  return O;
}
// y = M x
template <typename T, int R, int C>
LIE_HD void matvec(const Mat<T, R, C>& M, const T* x, T* y) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    T s = T(0);
#pragma unroll
    for (int c = 0; c < C; c++) s += M(r, c) * x[c];
    y[r] = s;
  }
}
// y = M^T x
template <typename T, int R, int C>
LIE_HD void matTvec(const Mat<T, R, C>& M, const T* x, T* y) {
#pragma unroll
  for (int c = 0; c < C; c++) {
    T s = T(0);
#pragma unroll
    for (int r = 0; r < R; r++) s += M(r, c) * x[r];
    y[c] = s;
  }
}
// y = x^T M  (row vector times matrix) == M^T x
template <typename T, int R, int C>
LIE_HD void vecmat(const T* x, const Mat<T, R, C>& M, T* y) { matTvec(M, x, y); }

template <typename T>
LIE_HD Mat<T, 3, 3> hat3(const T* v) {
  Mat<T, 3, 3> H;
  H(0, 0) = T(0);  H(0, 1) = -v[2]; H(0, 2) = v[1];
  H(1, 0) = v[2];  H(1, 1) = T(0);  H(1, 2) = -v[0];
  H(2, 0) = -v[1]; H(2, 1) = v[0];  H(2, 2) = T(0);
  return H;
}
template <typename T>
LIE_HD void cross3(const T* a, const T* b, T* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T, int R, int C, int RR, int CC>
LIE_HD void set_block(Mat<T, R, C>& D, int r0, int c0, const Mat<T, RR, CC>& S, T scale = T(1)) {
#pragma unroll
  for (int r = 0; r < RR; r++)
#pragma unroll
    for (int c = 0; c < CC; c++) D(r0 + r, c0 + c) = scale * S(r, c);
}

// ---- unit quaternion (x,y,z,w) ----------------------------------------------------------
template <typename T>
struct Quat {
  T x, y, z, w;
  LIE_HD void normalize() {
    T n = sqrt(x * x + y * y + z * z + w * w);
    x /= n; y /= n; z /= n; w /= n;
  }
  LIE_HD Quat conj() const { return Quat{-x, -y, -z, w}; }
  LIE_HD Quat operator*(const Quat& b) const {   // Hamilton product
    return Quat{w * b.x + x * b.w + y * b.z - z * b.y,
                w * b.y + y * b.w + z * b.x - x * b.z,
                w * b.z + z * b.w + x * b.y - y * b.x,
                w * b.w - x * b.x - y * b.y - z * b.z};
  }
  LIE_HD void rotate(const T* p, T* out) const {  // so3.h:52-57
    T qv[3] = {x, y, z}, uv[3], c[3];
    cross3(qv, p, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(qv, uv, c);
    out[0] = p[0] + w * uv[0] + c[0];
    out[1] = p[1] + w * uv[1] + c[1];
    out[2] = p[2] + w * uv[2] + c[2];
  }
  LIE_HD Mat<T, 3, 3> matrix() const {           // Eigen toRotationMatrix
    T tx = 2 * x, ty = 2 * y, tz = 2 * z;
    T twx = tx * w, twy = ty * w, twz = tz * w;
    T txx = tx * x, txy = ty * x, txz = tz * x;
    T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat<T, 3, 3> R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
  }
};

template <typename T>
LIE_HD Quat<T> quat_exp(const T* phi) {   // so3.h:141-157 (+ normalising constructor)
  T theta2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  T theta = sqrt(theta2);
  T imag, real;
  if (theta < Eps<T>::value) {
    T theta4 = theta2 * theta2;
    imag = T(0.5) - T(1.0 / 48.0) * theta2 + T(1.0 / 3840.0) * theta4;
    real = T(1) - T(1.0 / 8.0) * theta2 + T(1.0 / 384.0) * theta4;
  } else {
    imag = sin(T(0.5) * theta) / theta;
    real = cos(T(0.5) * theta);
  }
  Quat<T> q{imag * phi[0], imag * phi[1], imag * phi[2], real};
  q.normalize();
  return q;
}

template <typename T>
LIE_HD void quat_log(const Quat<T>& q, T* phi) {   // so3.h:106-139
  T sq = q.x * q.x + q.y * q.y + q.z * q.z;
  T w = q.w;
  T f;
  if (sq < Eps<T>::value * Eps<T>::value) {
    T w2 = w * w;
    f = T(2) / w - T(2.0 / 3.0) * sq / (w * w2);
  } else {
    T n = sqrt(sq);
    if (fabs(w) < Eps<T>::value) {
      f = (w > T(0)) ? pi_v<T>() / n : -pi_v<T>() / n;
    } else {
      f = T(2) * atan(n / w) / n;
    }
  }
  phi[0] = f * q.x; phi[1] = f * q.y; phi[2] = f * q.z;
}

template <typename T>
LIE_HD Mat<T, 3, 3> so3_left_jacobian(const T* phi) {   // so3.h:159-177
  Mat<T, 3, 3> Phi = hat3(phi);
  Mat<T, 3, 3> Phi2 = matmul(Phi, Phi);
  T theta2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  T theta = sqrt(theta2);
  bool small = theta < Eps<T>::value;
  T c1 = small ? T(0.5) - T(1.0 / 24.0) * theta2 : (T(1) - cos(theta)) / theta2;
  T c2 = small ? T(1.0 / 6.0) - T(1.0 / 120.0) * theta2 : (theta - sin(theta)) / (theta2 * theta);
  Mat<T, 3, 3> J;
  J.identity();
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] += c1 * Phi.m[i] + c2 * Phi2.m[i];
  return J;
}

template <typename T>
LIE_HD Mat<T, 3, 3> so3_left_jacobian_inverse(const T* phi) {   // so3.h:179-195
  Mat<T, 3, 3> Phi = hat3(phi);
  Mat<T, 3, 3> Phi2 = matmul(Phi, Phi);
  T theta2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  T theta = sqrt(theta2);
  T half = T(0.5) * theta;
  T c2 = (theta < Eps<T>::value) ? T(1.0 / 12.0)
                                 : (T(1) - theta * cos(half) / (T(2) * sin(half))) / (theta * theta);
  Mat<T, 3, 3> J;
  J.identity();
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] += T(-0.5) * Phi.m[i] + c2 * Phi2.m[i];
  return J;
}

// =========================================================================================
// SO3
template <typename T>
struct SO3 {
  static constexpr int K = 3, N = 4;
  Quat<T> q;
  LIE_HD static SO3 load(const T* d) { SO3 g; g.q = Quat<T>{d[0], d[1], d[2], d[3]}; g.q.normalize(); return g; }
  LIE_HD void store(T* d) const { d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w; }
  LIE_HD static SO3 Exp(const T* a) { SO3 g; g.q = quat_exp(a); return g; }
  LIE_HD void Log(T* a) const { quat_log(q, a); }
  LIE_HD SO3 inv() const { SO3 g; g.q = q.conj(); g.q.normalize(); return g; }
  LIE_HD SO3 operator*(const SO3& o) const { SO3 g; g.q = q * o.q; g.q.normalize(); return g; }
  LIE_HD void act(const T* p, T* out) const { q.rotate(p, out); }
  LIE_HD void act4(const T* p, T* out) const { q.rotate(p, out); out[3] = p[3]; }
  LIE_HD Mat<T, 3, 3> Adj() const { return q.matrix(); }
  LIE_HD static Mat<T, 3, 3> adj(const T* a) { return hat3(a); }
  LIE_HD Mat<T, 4, 4> Matrix4() const { Mat<T, 4, 4> M; M.identity(); set_block(M, 0, 0, q.matrix()); return M; }
  LIE_HD static Mat<T, 3, 3> left_jacobian(const T* a) { return so3_left_jacobian(a); }
  LIE_HD static Mat<T, 3, 3> left_jacobian_inverse(const T* a) { return so3_left_jacobian_inverse(a); }
  LIE_HD Mat<T, 4, 4> projector() const {   // so3.h:73-82
    Mat<T, 4, 4> J; J.zero();
    T nq[3] = {-q.x, -q.y, -q.z};
    Mat<T, 3, 3> H = hat3(nq);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) J(r, c) = T(0.5) * ((r == c ? q.w : T(0)) + H(r, c));
    J(3, 0) = T(0.5) * nq[0]; J(3, 1) = T(0.5) * nq[1]; J(3, 2) = T(0.5) * nq[2];
    return J;
  }
  LIE_HD static Mat<T, 3, 3> act_jacobian(const T* p) { T n[3] = {-p[0], -p[1], -p[2]}; return hat3(n); }
  LIE_HD static Mat<T, 4, 3> act4_jacobian(const T* p) {
    Mat<T, 4, 3> J; J.zero(); T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 0, hat3(n)); return J;
  }
};

// =========================================================================================
// RxSO3   data = [q(4), s]
template <typename T>
struct RxSO3 {
  static constexpr int K = 4, N = 5;
  Quat<T> q; T s;
  LIE_HD static RxSO3 load(const T* d) { RxSO3 g; g.q = Quat<T>{d[0], d[1], d[2], d[3]}; g.q.normalize(); g.s = d[4]; return g; }
  LIE_HD void store(T* d) const { d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w; d[4] = s; }
  LIE_HD static RxSO3 Exp(const T* a) { RxSO3 g; g.q = quat_exp(a); g.s = exp(a[3]); return g; }
  LIE_HD void Log(T* a) const { quat_log(q, a); a[3] = log(s); }
  LIE_HD RxSO3 inv() const { RxSO3 g; g.q = q.conj(); g.q.normalize(); g.s = T(1.0) / s; return g; }
  LIE_HD RxSO3 operator*(const RxSO3& o) const { RxSO3 g; g.q = q * o.q; g.q.normalize(); g.s = s * o.s; return g; }
  LIE_HD void act(const T* p, T* out) const { q.rotate(p, out); out[0] *= s; out[1] *= s; out[2] *= s; }
  LIE_HD void act4(const T* p, T* out) const { act(p, out); out[3] = p[3]; }
  LIE_HD Mat<T, 3, 3> Rotation() const { return q.matrix(); }
  LIE_HD Mat<T, 3, 3> Matrix() const { Mat<T, 3, 3> R = q.matrix();
#pragma unroll
    for (int i = 0; i < 9; i++) R.m[i] *= s;
    return R; }
  LIE_HD Mat<T, 4, 4> Adj() const { Mat<T, 4, 4> A; A.identity(); set_block(A, 0, 0, q.matrix()); return A; }
  LIE_HD static Mat<T, 4, 4> adj(const T* a) { Mat<T, 4, 4> A; A.zero(); set_block(A, 0, 0, hat3(a)); return A; }
  LIE_HD Mat<T, 4, 4> Matrix4() const { Mat<T, 4, 4> M; M.identity(); set_block(M, 0, 0, Matrix()); return M; }
  LIE_HD static Mat<T, 4, 4> left_jacobian(const T* a) { Mat<T, 4, 4> J; J.identity(); set_block(J, 0, 0, so3_left_jacobian(a)); return J; }
  LIE_HD static Mat<T, 4, 4> left_jacobian_inverse(const T* a) { Mat<T, 4, 4> J; J.identity(); set_block(J, 0, 0, so3_left_jacobian_inverse(a)); return J; }
  LIE_HD Mat<T, 5, 5> projector() const {   // rxso3.h:84-100
    Mat<T, 5, 5> J; J.zero();
    T nq[3] = {-q.x, -q.y, -q.z};
    Mat<T, 3, 3> H = hat3(nq);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) J(r, c) = T(0.5) * ((r == c ? q.w : T(0)) + H(r, c));
    J(3, 0) = T(0.5) * nq[0]; J(3, 1) = T(0.5) * nq[1]; J(3, 2) = T(0.5) * nq[2];
    J(4, 3) = s;
    return J;
  }
  LIE_HD static Mat<T, 3, 4> act_jacobian(const T* p) {
    Mat<T, 3, 4> J; T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 0, hat3(n));
    J(0, 3) = p[0]; J(1, 3) = p[1]; J(2, 3) = p[2]; return J;
  }
  LIE_HD static Mat<T, 4, 4> act4_jacobian(const T* p) {
    Mat<T, 4, 4> J; J.zero(); T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 0, hat3(n));
    J(0, 3) = p[0]; J(1, 3) = p[1]; J(2, 3) = p[2]; return J;
  }
  LIE_HD static Mat<T, 3, 3> calcW(const T* a) {   // rxso3.h:198-241
    const T one(1), half(0.5);
    const T sigma = a[3];
    const T theta = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    Mat<T, 3, 3> Phi = hat3(a);
    Mat<T, 3, 3> Phi2 = matmul(Phi, Phi);
    const T scale = exp(sigma);
    T A, B, C;
    if (fabs(sigma) < Eps<T>::value) {
      C = one;
      if (fabs(theta) < Eps<T>::value) { A = half; B = T(1. / 6.); }
      else { T t2 = theta * theta; A = (one - cos(theta)) / t2; B = (theta - sin(theta)) / (t2 * theta); }
    } else {
      C = (scale - one) / sigma;
      if (fabs(theta) < Eps<T>::value) {
        T s2 = sigma * sigma;
        A = ((sigma - one) * scale + one) / s2;
        B = (scale * half * s2 + scale - one - sigma * scale) / (s2 * sigma);
      } else {
        T t2 = theta * theta;
        T a_ = scale * sin(theta), b_ = scale * cos(theta), c_ = t2 + sigma * sigma;
        A = (a_ * sigma + (one - b_) * theta) / (theta * c_);
        B = (C - ((b_ - one) * sigma + a_ * theta) / (c_)) * one / (t2);
      }
    }
    Mat<T, 3, 3> W;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) W(r, c) = A * Phi(r, c) + B * Phi2(r, c) + (r == c ? C : T(0));
    return W;
  }
};

// =========================================================================================
// SE3   data = [t(3), q(4)]
template <typename T>
struct SE3 {
  static constexpr int K = 6, N = 7;
  T t[3]; SO3<T> R;
  LIE_HD static SE3 load(const T* d) { SE3 g; g.t[0] = d[0]; g.t[1] = d[1]; g.t[2] = d[2]; g.R = SO3<T>::load(d + 3); return g; }
  LIE_HD void store(T* d) const { d[0] = t[0]; d[1] = t[1]; d[2] = t[2]; R.store(d + 3); }
  LIE_HD static SE3 Exp(const T* a) {   // se3.h:131-140
    SE3 g; g.R = SO3<T>::Exp(a + 3);
    matvec(so3_left_jacobian(a + 3), a, g.t);
    return g;
  }
  LIE_HD void Log(T* a) const {         // se3.h:120-129
    R.Log(a + 3);
    matvec(so3_left_jacobian_inverse(a + 3), t, a);
  }
  LIE_HD SE3 inv() const { SE3 g; g.R = R.inv(); T r[3]; g.R.act(t, r); g.t[0] = -r[0]; g.t[1] = -r[1]; g.t[2] = -r[2]; return g; }
  LIE_HD SE3 operator*(const SE3& o) const {
    SE3 g; g.R = R * o.R; T r[3]; R.act(o.t, r);
    g.t[0] = t[0] + r[0]; g.t[1] = t[1] + r[1]; g.t[2] = t[2] + r[2]; return g;
  }
  LIE_HD void act(const T* p, T* out) const { R.act(p, out); out[0] += t[0]; out[1] += t[1]; out[2] += t[2]; }
  LIE_HD void act4(const T* p, T* out) const {
    R.act(p, out); out[0] += t[0] * p[3]; out[1] += t[1] * p[3]; out[2] += t[2] * p[3]; out[3] = p[3];
  }
  LIE_HD Mat<T, 6, 6> Adj() const {     // se3.h:65-75
    Mat<T, 3, 3> Rm = R.q.matrix();
    Mat<T, 3, 3> tR = matmul(hat3(t), Rm);
    Mat<T, 6, 6> A; A.zero();
    set_block(A, 0, 0, Rm); set_block(A, 0, 3, tR); set_block(A, 3, 3, Rm);
    return A;
  }
  LIE_HD static Mat<T, 6, 6> adj(const T* a) {   // se3.h:102-114
    Mat<T, 6, 6> A; A.zero();
    Mat<T, 3, 3> Tau = hat3(a), Phi = hat3(a + 3);
    set_block(A, 0, 0, Phi); set_block(A, 0, 3, Tau); set_block(A, 3, 3, Phi);
    return A;
  }
  LIE_HD Mat<T, 4, 4> Matrix4() const {
    Mat<T, 4, 4> M; M.identity(); set_block(M, 0, 0, R.q.matrix());
    M(0, 3) = t[0]; M(1, 3) = t[1]; M(2, 3) = t[2]; return M;
  }
  LIE_HD static Mat<T, 3, 3> calcQ(const T* a) {   // se3.h:142-172
    Mat<T, 3, 3> Tau = hat3(a), Phi = hat3(a + 3);
    T theta = sqrt(a[3] * a[3] + a[4] * a[4] + a[5] * a[5]);
    T t2 = theta * theta, t4 = t2 * t2;
    bool small = theta < Eps<T>::value;
    T c1 = small ? T(1.0 / 6.0) - T(1.0 / 120.0) * t2 : (theta - sin(theta)) / (t2 * theta);
    T c2 = small ? T(1.0 / 24.0) - T(1.0 / 720.0) * t2 : (t2 + 2 * cos(theta) - 2) / (2 * t4);
    T c3 = small ? T(1.0 / 120.0) - T(1.0 / 2520.0) * t2
                 : (2 * theta - 3 * sin(theta) + theta * cos(theta)) / (2 * t4 * theta);
    Mat<T, 3, 3> PT = matmul(Phi, Tau), TP = matmul(Tau, Phi);
    Mat<T, 3, 3> PTP = matmul(PT, Phi), PP = matmul(Phi, Phi);
    Mat<T, 3, 3> PPT = matmul(PP, Tau), TPP = matmul(Tau, PP);
    Mat<T, 3, 3> PTPP = matmul(PTP, Phi), PPTP = matmul(Phi, PTP);
    Mat<T, 3, 3> Q;
#pragma unroll
    for (int i = 0; i < 9; i++)
      Q.m[i] = T(0.5) * Tau.m[i] + c1 * (PT.m[i] + TP.m[i] + PTP.m[i]) +
               c2 * (PPT.m[i] + TPP.m[i] - 3 * PTP.m[i]) + c3 * (PTPP.m[i] + PPTP.m[i]);
    return Q;
  }
  LIE_HD static Mat<T, 6, 6> left_jacobian(const T* a) {
    Mat<T, 3, 3> J = so3_left_jacobian(a + 3), Q = calcQ(a);
    Mat<T, 6, 6> O; O.zero();
    set_block(O, 0, 0, J); set_block(O, 0, 3, Q); set_block(O, 3, 3, J);
    return O;
  }
  LIE_HD static Mat<T, 6, 6> left_jacobian_inverse(const T* a) {
    Mat<T, 3, 3> Ji = so3_left_jacobian_inverse(a + 3), Q = calcQ(a);
    Mat<T, 3, 3> JQJ = matmul(matmul(Ji, Q), Ji);
    Mat<T, 6, 6> O; O.zero();
    set_block(O, 0, 0, Ji); set_block(O, 0, 3, JQJ, T(-1)); set_block(O, 3, 3, Ji);
    return O;
  }
  LIE_HD Mat<T, 7, 7> projector() const {   // se3.h:116-124
    Mat<T, 7, 7> J; J.zero();
    J(0, 0) = J(1, 1) = J(2, 2) = T(1);
    T nt[3] = {-t[0], -t[1], -t[2]};
    set_block(J, 0, 3, hat3(nt));
    set_block(J, 3, 3, R.projector());
    return J;
  }
  LIE_HD static Mat<T, 3, 6> act_jacobian(const T* p) {
    Mat<T, 3, 6> J; J.zero(); J(0, 0) = J(1, 1) = J(2, 2) = T(1);
    T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 3, hat3(n)); return J;
  }
  LIE_HD static Mat<T, 4, 6> act4_jacobian(const T* p) {
    Mat<T, 4, 6> J; J.zero(); J(0, 0) = J(1, 1) = J(2, 2) = p[3];
    T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 3, hat3(n)); return J;
  }
};

// =========================================================================================
// Sim3   data = [t(3), q(4), s]
template <typename T>
struct Sim3 {
  static constexpr int K = 7, N = 8;
  T t[3]; RxSO3<T> R;
  LIE_HD static Sim3 load(const T* d) { Sim3 g; g.t[0] = d[0]; g.t[1] = d[1]; g.t[2] = d[2]; g.R = RxSO3<T>::load(d + 3); return g; }
  LIE_HD void store(T* d) const { d[0] = t[0]; d[1] = t[1]; d[2] = t[2]; R.store(d + 3); }
  LIE_HD static Sim3 Exp(const T* a) {   // sim3.h:155-165
    Sim3 g; g.R = RxSO3<T>::Exp(a + 3);
    matvec(RxSO3<T>::calcW(a + 3), a, g.t);
    return g;
  }
  LIE_HD void Log(T* a) const {          // sim3.h:143-153  (W.inverse() * t, 3x3 cofactor inverse)
    R.Log(a + 3);
    Mat<T, 3, 3> W = RxSO3<T>::calcW(a + 3);
    T c00 = W(1, 1) * W(2, 2) - W(1, 2) * W(2, 1);
    T c01 = W(1, 2) * W(2, 0) - W(1, 0) * W(2, 2);
    T c02 = W(1, 0) * W(2, 1) - W(1, 1) * W(2, 0);
    T det = W(0, 0) * c00 + W(0, 1) * c01 + W(0, 2) * c02;
    T id = T(1) / det;
    Mat<T, 3, 3> Wi;
    Wi(0, 0) = c00 * id; Wi(0, 1) = (W(0, 2) * W(2, 1) - W(0, 1) * W(2, 2)) * id; Wi(0, 2) = (W(0, 1) * W(1, 2) - W(0, 2) * W(1, 1)) * id;
    Wi(1, 0) = c01 * id; Wi(1, 1) = (W(0, 0) * W(2, 2) - W(0, 2) * W(2, 0)) * id; Wi(1, 2) = (W(0, 2) * W(1, 0) - W(0, 0) * W(1, 2)) * id;
    Wi(2, 0) = c02 * id; Wi(2, 1) = (W(0, 1) * W(2, 0) - W(0, 0) * W(2, 1)) * id; Wi(2, 2) = (W(0, 0) * W(1, 1) - W(0, 1) * W(1, 0)) * id;
    matvec(Wi, t, a);
  }
  LIE_HD Sim3 inv() const { Sim3 g; g.R = R.inv(); T r[3]; g.R.act(t, r); g.t[0] = -r[0]; g.t[1] = -r[1]; g.t[2] = -r[2]; return g; }
  LIE_HD Sim3 operator*(const Sim3& o) const {
    Sim3 g; g.R = R * o.R; T r[3]; R.act(o.t, r);
    g.t[0] = t[0] + r[0]; g.t[1] = t[1] + r[1]; g.t[2] = t[2] + r[2]; return g;
  }
  LIE_HD void act(const T* p, T* out) const { R.act(p, out); out[0] += t[0]; out[1] += t[1]; out[2] += t[2]; }
  LIE_HD void act4(const T* p, T* out) const {
    R.act(p, out); out[0] += t[0] * p[3]; out[1] += t[1] * p[3]; out[2] += t[2] * p[3]; out[3] = p[3];
  }
  LIE_HD Mat<T, 7, 7> Adj() const {   // sim3.h:85-98
    Mat<T, 7, 7> A; A.identity();
    Mat<T, 3, 3> Rm = R.Rotation();
    set_block(A, 0, 0, R.Matrix());
    set_block(A, 0, 3, matmul(hat3(t), Rm));
    A(0, 6) = -t[0]; A(1, 6) = -t[1]; A(2, 6) = -t[2];
    set_block(A, 3, 3, Rm);
    return A;
  }
  LIE_HD static Mat<T, 7, 7> adj(const T* a) {   // sim3.h:124-141
    Mat<T, 7, 7> A; A.zero();
    Mat<T, 3, 3> Tau = hat3(a), Phi = hat3(a + 3);
    set_block(A, 0, 0, Phi);
    A(0, 0) += a[6]; A(1, 1) += a[6]; A(2, 2) += a[6];
    set_block(A, 0, 3, Tau);
    A(0, 6) = -a[0]; A(1, 6) = -a[1]; A(2, 6) = -a[2];
    set_block(A, 3, 3, Phi);
    return A;
  }
  LIE_HD Mat<T, 4, 4> Matrix4() const {
    Mat<T, 4, 4> M; M.identity(); set_block(M, 0, 0, R.Matrix());
    M(0, 3) = t[0]; M(1, 3) = t[1]; M(2, 3) = t[2]; return M;
  }
  LIE_HD static Mat<T, 7, 7> left_jacobian(const T* a) {   // sim3.h:167-179 (1/720 term dropped in the reference)
    Mat<T, 7, 7> Xi = adj(a);
    Mat<T, 7, 7> Xi2 = matmul(Xi, Xi);
    Mat<T, 7, 7> Xi3 = matmul(Xi, Xi2);
    Mat<T, 7, 7> Xi4 = matmul(Xi2, Xi2);
    Mat<T, 7, 7> J; J.identity();
#pragma unroll
    for (int i = 0; i < 49; i++)
      J.m[i] += T(1.0 / 2.0) * Xi.m[i] + T(1.0 / 6.0) * Xi2.m[i] + T(1.0 / 24.0) * Xi3.m[i] + T(1.0 / 120.0) * Xi4.m[i];
    return J;
  }
  LIE_HD static Mat<T, 7, 7> left_jacobian_inverse(const T* a) {   // sim3.h:181-191
    Mat<T, 7, 7> Xi = adj(a);
    Mat<T, 7, 7> Xi2 = matmul(Xi, Xi);
    Mat<T, 7, 7> Xi4 = matmul(Xi2, Xi2);
    Mat<T, 7, 7> J; J.identity();
#pragma unroll
    for (int i = 0; i < 49; i++)
      J.m[i] += T(-1.0 / 2.0) * Xi.m[i] + T(1.0 / 12.0) * Xi2.m[i] - T(1.0 / 720.0) * Xi4.m[i];
    return J;
  }
  LIE_HD Mat<T, 8, 8> projector() const {   // sim3.h:74-83
    Mat<T, 8, 8> J; J.zero();
    J(0, 0) = J(1, 1) = J(2, 2) = T(1);
    T nt[3] = {-t[0], -t[1], -t[2]};
    set_block(J, 0, 3, hat3(nt));
    J(0, 6) = t[0]; J(1, 6) = t[1]; J(2, 6) = t[2];
    set_block(J, 3, 3, R.projector());
    return J;
  }
  LIE_HD static Mat<T, 3, 7> act_jacobian(const T* p) {
    Mat<T, 3, 7> J; J.zero(); J(0, 0) = J(1, 1) = J(2, 2) = T(1);
    T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 3, hat3(n));
    J(0, 6) = p[0]; J(1, 6) = p[1]; J(2, 6) = p[2]; return J;
  }
  LIE_HD static Mat<T, 4, 7> act4_jacobian(const T* p) {
    Mat<T, 4, 7> J; J.zero(); J(0, 0) = J(1, 1) = J(2, 2) = p[3];
    T n[3] = {-p[0], -p[1], -p[2]}; set_block(J, 0, 3, hat3(n));
    J(0, 6) = p[0]; J(1, 6) = p[1]; J(2, 6) = p[2]; return J;
  }
};

}  // namespace lie
