// host_math.hpp -- small dense maths shared by host orchestration and the device-side
// Gauss-Newton solve kernel of libbadba_b200 (product code; independent of oracle/).
//
// Semantics follow the reference's host code so that results stay within the parity budget:
//   SE3f update  global_T_frame * exp(-x)      direct_ba_alternating.cc:214
//   Sophus SE3/SO3 exp, log, inverse, product  libvis/third_party/sophus/sophus/{se3,so3}.hpp
//   fp64 LDLT of the upper triangle            direct_ba_alternating.cc:206, kernel_opt_intrinsics.cc:171,272
//   convergence test                           convergence_analysis.h:45-52
// Pose layout: float[7] = {qx,qy,qz,qw,tx,ty,tz} (Sophus::SE3f::data()).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define BBA_HD __host__ __device__ __forceinline__
#else
#define BBA_HD inline
#endif

namespace bba {

constexpr float kSophusEpsilonF = 1e-5f;   // sophus/common.hpp:146-148

// Trigonometry evaluated in fp64 and rounded: these run once per keyframe per Gauss-Newton
// iteration, and must not degrade to the fast-math approximations the kernels are built with.
BBA_HD float PSin(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
BBA_HD float PCos(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
BBA_HD float PAtan(float x) { return static_cast<float>(atan(static_cast<double>(x))); }
BBA_HD float PSqrt(float x) { return static_cast<float>(sqrt(static_cast<double>(x))); }
BBA_HD float PDiv(float a, float b) { return static_cast<float>(static_cast<double>(a) / static_cast<double>(b)); }

struct Pose {  // global_T_frame or its inverse
  float q[4];  // x y z w
  float t[3];
};

BBA_HD void QuatRotate(const float q[4], const float v[3], float out[3]) {
  // Eigen::QuaternionBase::_transformVector
  const float ux = 2.f * (q[1] * v[2] - q[2] * v[1]);
  const float uy = 2.f * (q[2] * v[0] - q[0] * v[2]);
  const float uz = 2.f * (q[0] * v[1] - q[1] * v[0]);
  out[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  out[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  out[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

BBA_HD void QuatToMatrix(const float q[4], float R[9]) {
  // Eigen::QuaternionBase::toRotationMatrix
  const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.f - (txx + tyy);
}

BBA_HD Pose Inverse(const Pose& a) {   // se3.hpp:127-130
  Pose r;
  r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
  const float nt[3] = {-a.t[0], -a.t[1], -a.t[2]};
  QuatRotate(r.q, nt, r.t);
  return r;
}

BBA_HD Pose Compose(const Pose& a, const Pose& b) {   // se3.hpp:203-207, so3.hpp:215-232
  Pose r;
  float rt[3];
  QuatRotate(a.q, b.t, rt);
  r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
  const float ax = a.q[0], ay = a.q[1], az = a.q[2], aw = a.q[3];
  const float bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
  r.q[3] = aw * bw - ax * bx - ay * by - az * bz;
  r.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  r.q[1] = aw * by + ay * bw + az * bx - ax * bz;
  r.q[2] = aw * bz + az * bw + ax * by - ay * bx;
  const float sn = r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3];
  if (sn != 1.0f) {
    const float s = PDiv(2.0f, 1.0f + sn);
    r.q[0] *= s; r.q[1] *= s; r.q[2] *= s; r.q[3] *= s;
  }
  return r;
}

// Row-major 3x4 [R|t].
BBA_HD void ToMatrix3x4(const Pose& p, float M[12]) {
  float R[9];
  QuatToMatrix(p.q, R);
  M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = p.t[0];
  M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = p.t[1];
  M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = p.t[2];
}

BBA_HD void Hat(const float w[3], float O[9]) {
  O[0] = 0.f; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0.f; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0.f;
}

BBA_HD void Mat3Mul(const float A[9], const float B[9], float C[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

BBA_HD Pose Exp(const float a[6]) {   // se3.hpp:293-313 + so3.hpp:282-312
  Pose r;
  const float* om = a + 3;
  const float theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const float theta = PSqrt(theta_sq);
  float imag, real;
  if (theta < kSophusEpsilonF) {
    const float p4 = theta_sq * theta_sq;
    imag = 0.5f - (1.0f / 48.0f) * theta_sq + (1.0f / 3840.0f) * p4;
    real = 1.f - 0.5f * theta_sq + (1.0f / 384.0f) * p4;
  } else {
    const float h = 0.5f * theta;
    imag = PDiv(PSin(h), theta);
    real = PCos(h);
  }
  r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
  float O[9], O2[9], V[9];
  Hat(om, O);
  Mat3Mul(O, O, O2);
  if (theta < kSophusEpsilonF) {
    QuatToMatrix(r.q, V);
  } else {
    const float c1 = PDiv(1.f - PCos(theta), theta_sq);
    const float c2 = PDiv(theta - PSin(theta), theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = c1 * O[i] + c2 * O2[i];
    V[0] += 1.f; V[4] += 1.f; V[8] += 1.f;
  }
  for (int i = 0; i < 3; ++i) r.t[i] = V[i * 3] * a[0] + V[i * 3 + 1] * a[1] + V[i * 3 + 2] * a[2];
  return r;
}

BBA_HD void Log(const Pose& p, float out[6]) {   // se3.hpp:435-468 + so3.hpp:421-466
  const float sq_n = p.q[0] * p.q[0] + p.q[1] * p.q[1] + p.q[2] * p.q[2];
  const float n = PSqrt(sq_n);
  const float w = p.q[3];
  float f;
  if (n < kSophusEpsilonF) {
    f = PDiv(2.f, w) - PDiv(2.f * sq_n, w * w * w);
  } else if (fabsf(w) < kSophusEpsilonF) {
    f = PDiv(w > 0.f ? 3.14159265358979323846f : -3.14159265358979323846f, n);
  } else {
    f = PDiv(2.f * PAtan(PDiv(n, w)), n);
  }
  const float theta = f * n;
  const float om[3] = {f * p.q[0], f * p.q[1], f * p.q[2]};
  float O[9], O2[9];
  Hat(om, O);
  Mat3Mul(O, O, O2);
  float c2;
  if (fabsf(theta) < kSophusEpsilonF) {
    c2 = 1.f / 12.f;
  } else {
    const float h = 0.5f * theta;
    c2 = PDiv(1.f - PDiv(theta * PCos(h), 2.f * PSin(h)), theta * theta);
  }
  for (int i = 0; i < 3; ++i) {
    float v0 = -0.5f * O[i * 3] + c2 * O2[i * 3] + (i == 0 ? 1.f : 0.f);
    float v1 = -0.5f * O[i * 3 + 1] + c2 * O2[i * 3 + 1] + (i == 1 ? 1.f : 0.f);
    float v2 = -0.5f * O[i * 3 + 2] + c2 * O2[i * 3 + 2] + (i == 2 ? 1.f : 0.f);
    out[i] = v0 * p.t[0] + v1 * p.t[1] + v2 * p.t[2];
  }
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

BBA_HD bool IsScale1PoseEstimationConverged(const float x[6]) {   // convergence_analysis.h:45-52
  const float s = 1e-06f / 1e-07f;
  const float sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + (s * x[3]) * (s * x[3]) + (s * x[4]) * (s * x[4]) +
                   (s * x[5]) * (s * x[5]);
  return sq < 1e-06f;
}

// x = A^-1 b for the symmetric N x N matrix whose upper triangle is packed row-major in `upper`
// (N(N+1)/2 entries, the layout of gauss_newton.cuh:59-73).  fp64, symmetric pivoting on the
// largest diagonal entry (what Eigen's LDLT does).  Rank-deficient directions get x = 0.
template <int N>
BBA_HD void SolveLDLT(const double* upper, const double* b, double* x) {
  double M[N * N];
  int perm[N];
  int idx = 0;
  for (int r = 0; r < N; ++r) {
    perm[r] = r;
    for (int c = r; c < N; ++c) {
      M[r * N + c] = upper[idx];
      M[c * N + r] = upper[idx];
      ++idx;
    }
  }
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(M[k * N + k]);
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(M[i * N + i]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {
      for (int c = 0; c < N; ++c) { const double t = M[k * N + c]; M[k * N + c] = M[p * N + c]; M[p * N + c] = t; }
      for (int r = 0; r < N; ++r) { const double t = M[r * N + k]; M[r * N + k] = M[r * N + p]; M[r * N + p] = t; }
      const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    const double d = M[k * N + k];
    if (d == 0.0) {
      for (int i = k + 1; i < N; ++i) M[i * N + k] = 0.0;
      continue;
    }
    for (int i = k + 1; i < N; ++i) M[i * N + k] /= d;
    for (int i = k + 1; i < N; ++i)
      for (int j = k + 1; j <= i; ++j) {
        M[i * N + j] -= M[i * N + k] * d * M[j * N + k];
        M[j * N + i] = M[i * N + j];
      }
  }
  double y[N];
  for (int i = 0; i < N; ++i) y[i] = b[perm[i]];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < i; ++j) y[i] -= M[i * N + j] * y[j];
  for (int i = 0; i < N; ++i) {
    const double d = M[i * N + i];
    y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0;
  }
  for (int i = N - 1; i >= 0; --i)
    for (int j = i + 1; j < N; ++j) y[i] -= M[j * N + i] * y[j];
  for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
}

}  // namespace bba
