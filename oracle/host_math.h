/* oracle/host_math.h -- TEST INFRASTRUCTURE (never linked into the product).
 *
 * Eigen/Sophus-free restatement of the small host-side maths the reference's
 * BA loop uses.  Shared by the CPU oracle (badba_oracle.c) and by the driver
 * of the reference's own CUDA kernels (ref_driver.cu).  The product library
 * has its own independent implementation (badslam_b200/csrc/host_math.hpp), so
 * the two can be cross-checked.
 *
 * Follows (all paths relative to /root/reference):
 *   libvis/third_party/sophus/sophus/so3.hpp:282-312   SO3::expAndTheta
 *   libvis/third_party/sophus/sophus/so3.hpp:215-232   SO3::operator*= (renormalisation)
 *   libvis/third_party/sophus/sophus/so3.hpp:421-466   SO3::logAndTheta
 *   libvis/third_party/sophus/sophus/se3.hpp:127-130   SE3::inverse
 *   libvis/third_party/sophus/sophus/se3.hpp:203-207   SE3::operator*=
 *   libvis/third_party/sophus/sophus/se3.hpp:293-313   SE3::exp
 *   libvis/third_party/sophus/sophus/se3.hpp:435-468   SE3::log
 *   libvis/third_party/sophus/sophus/common.hpp:144-151 Constants<float>::epsilon = 1e-5
 *   applications/badslam/src/badslam/convergence_analysis.h:45-52
 *   applications/badslam/src/badslam/direct_ba_alternating.cc:206 (fp64 LDLT solve)
 *   libvis/src/libvis/camera_frustum.h:43-250 (co-visibility frustum test)
 *
 * Third-party arithmetic that is NOT under /root/reference: Eigen 3.3.7
 * (README.md:78).  Quaternion product / toRotationMatrix / LDLT follow Eigen's
 * published formulas; parity of the dense solves is unpinned by any reference
 * test and is anchored on the reference's call sites only.
 *
 * Pose layout everywhere: float[7] = {qx, qy, qz, qw, tx, ty, tz}, which is
 * exactly Sophus::SE3f::data().
 */
#ifndef BADBA_ORACLE_HOST_MATH_H
#define BADBA_ORACLE_HOST_MATH_H

#include <math.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_SOPHUS_EPS_F 1e-5f

/* ---- quaternion helpers (Eigen conventions) ---- */

/* Eigen::Quaternion product a*b.  q = {x,y,z,w}. */
static inline void hm_quat_mul(const float a[4], const float b[4], float out[4]) {
  float ax = a[0], ay = a[1], az = a[2], aw = a[3];
  float bx = b[0], by = b[1], bz = b[2], bw = b[3];
  float w = aw * bw - ax * bx - ay * by - az * bz;
  float x = aw * bx + ax * bw + ay * bz - az * by;
  float y = aw * by + ay * bw + az * bx - ax * bz;
  float z = aw * bz + az * bw + ax * by - ay * bx;
  out[0] = x; out[1] = y; out[2] = z; out[3] = w;
}

/* Eigen::QuaternionBase::toRotationMatrix, row-major 3x3. */
static inline void hm_quat_to_R(const float q[4], float R[9]) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  float twx = tx * w, twy = ty * w, twz = tz * w;
  float txx = tx * x, txy = ty * x, txz = tz * x;
  float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}

/* Eigen::QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 q.vec x v */
static inline void hm_quat_rotate(const float q[4], const float v[3], float out[3]) {
  float ux = 2.f * (q[1] * v[2] - q[2] * v[1]);
  float uy = 2.f * (q[2] * v[0] - q[0] * v[2]);
  float uz = 2.f * (q[0] * v[1] - q[1] * v[0]);
  out[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  out[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  out[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

/* ---- SE3 (float, Sophus semantics) ---- */

static inline void hm_se3_identity(float T[7]) {
  T[0] = T[1] = T[2] = 0.f; T[3] = 1.f; T[4] = T[5] = T[6] = 0.f;
}

/* se3.hpp:127-130 */
static inline void hm_se3_inverse(const float T[7], float out[7]) {
  float qi[4] = {-T[0], -T[1], -T[2], T[3]};
  float nt[3] = {-T[4], -T[5], -T[6]};
  float t[3];
  hm_quat_rotate(qi, nt, t);
  out[0] = qi[0]; out[1] = qi[1]; out[2] = qi[2]; out[3] = qi[3];
  out[4] = t[0]; out[5] = t[1]; out[6] = t[2];
}

/* se3.hpp:203-207 + so3.hpp:215-232: out = A * B */
static inline void hm_se3_mul(const float A[7], const float B[7], float out[7]) {
  float rt[3];
  hm_quat_rotate(A, B + 4, rt);
  float t0 = A[4] + rt[0], t1 = A[5] + rt[1], t2 = A[6] + rt[2];
  float q[4];
  hm_quat_mul(A, B, q);
  float sn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (sn != 1.0f) {
    float s = 2.0f / (1.0f + sn);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
  }
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = t0; out[5] = t1; out[6] = t2;
}

/* Row-major 3x4 [R|t] of T (se3.hpp:165-170). */
static inline void hm_se3_matrix3x4(const float T[7], float M[12]) {
  float R[9];
  hm_quat_to_R(T, R);
  M[0] = R[0]; M[1] = R[1]; M[2]  = R[2]; M[3]  = T[4];
  M[4] = R[3]; M[5] = R[4]; M[6]  = R[5]; M[7]  = T[5];
  M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = T[6];
}

/* se3.hpp:293-313, so3.hpp:282-312.  a = (upsilon, omega). */
static inline void hm_se3_exp(const float a[6], float out[7]) {
  float ox = a[3], oy = a[4], oz = a[5];
  float theta_sq = ox * ox + oy * oy + oz * oz;
  float theta = sqrtf(theta_sq);
  float half_theta = 0.5f * theta;
  float imag, real;
  if (theta < HM_SOPHUS_EPS_F) {
    float theta_po4 = theta_sq * theta_sq;
    imag = 0.5f - (float)(1.0 / 48.0) * theta_sq + (float)(1.0 / 3840.0) * theta_po4;
    real = 1.f - 0.5f * theta_sq + (float)(1.0 / 384.0) * theta_po4;
  } else {
    float s = sinf(half_theta);
    imag = s / theta;
    real = cosf(half_theta);
  }
  float q[4] = {imag * ox, imag * oy, imag * oz, real};
  /* Omega = hat(omega) */
  float O[9] = {0.f, -oz, oy,  oz, 0.f, -ox,  -oy, ox, 0.f};
  float O2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      O2[r * 3 + c] = O[r * 3 + 0] * O[0 * 3 + c] + O[r * 3 + 1] * O[1 * 3 + c] + O[r * 3 + 2] * O[2 * 3 + c];
  float V[9];
  if (theta < HM_SOPHUS_EPS_F) {
    hm_quat_to_R(q, V);
  } else {
    float c1 = (1.f - cosf(theta)) / theta_sq;
    float c2 = (theta - sinf(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = c1 * O[i] + c2 * O2[i];
    V[0] += 1.f; V[4] += 1.f; V[8] += 1.f;
  }
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = V[0] * a[0] + V[1] * a[1] + V[2] * a[2];
  out[5] = V[3] * a[0] + V[4] * a[1] + V[5] * a[2];
  out[6] = V[6] * a[0] + V[7] * a[1] + V[8] * a[2];
}

/* se3.hpp:435-468, so3.hpp:421-466 */
static inline void hm_se3_log(const float T[7], float out[6]) {
  float squared_n = T[0] * T[0] + T[1] * T[1] + T[2] * T[2];
  float n = sqrtf(squared_n);
  float w = T[3];
  float two_atan_nbyw_by_n;
  if (n < HM_SOPHUS_EPS_F) {
    float squared_w = w * w;
    two_atan_nbyw_by_n = 2.f / w - 2.f * squared_n / (w * squared_w);
  } else {
    if (fabsf(w) < HM_SOPHUS_EPS_F) {
      two_atan_nbyw_by_n = (w > 0.f ? (float)M_PI : -(float)M_PI) / n;
    } else {
      two_atan_nbyw_by_n = 2.f * atanf(n / w) / n;
    }
  }
  float theta = two_atan_nbyw_by_n * n;
  float ox = two_atan_nbyw_by_n * T[0], oy = two_atan_nbyw_by_n * T[1], oz = two_atan_nbyw_by_n * T[2];
  float O[9] = {0.f, -oz, oy,  oz, 0.f, -ox,  -oy, ox, 0.f};
  float O2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      O2[r * 3 + c] = O[r * 3 + 0] * O[0 * 3 + c] + O[r * 3 + 1] * O[1 * 3 + c] + O[r * 3 + 2] * O[2 * 3 + c];
  float Vi[9];
  float c2;
  if (fabsf(theta) < HM_SOPHUS_EPS_F) {
    c2 = (float)(1. / 12.);
  } else {
    float half_theta = 0.5f * theta;
    c2 = (1.f - theta * cosf(half_theta) / (2.f * sinf(half_theta))) / (theta * theta);
  }
  for (int i = 0; i < 9; ++i) Vi[i] = -0.5f * O[i] + c2 * O2[i];
  Vi[0] += 1.f; Vi[4] += 1.f; Vi[8] += 1.f;
  out[0] = Vi[0] * T[4] + Vi[1] * T[5] + Vi[2] * T[6];
  out[1] = Vi[3] * T[4] + Vi[4] * T[5] + Vi[5] * T[6];
  out[2] = Vi[6] * T[4] + Vi[7] * T[5] + Vi[8] * T[6];
  out[3] = ox; out[4] = oy; out[5] = oz;
}

/* convergence_analysis.h:45-52 */
static inline int hm_is_scale1_pose_converged(const float x[6]) {
  const float translation_threshold = 1e-06f;
  const float rotation_threshold = 1e-07f;
  const float s = translation_threshold / rotation_threshold;
  float sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] +
             (s * x[3]) * (s * x[3]) + (s * x[4]) * (s * x[4]) + (s * x[5]) * (s * x[5]);
  return sq < translation_threshold;
}

/* Solve the SPD system given by the UPPER triangle of the n x n matrix A
 * (row-major, n <= 6) in fp64 with a diagonally pivoted LDL^T, as
 * Eigen's selfadjointView<Upper>().ldlt().solve() does
 * (direct_ba_alternating.cc:206, kernel_opt_intrinsics.cc:171,272).
 * Returns 0 on success; x is all-zero if the matrix is exactly zero. */
static inline int hm_ldlt_solve(int n, const double* A_upper, const double* b, double* x) {
  double M[36];
  int perm[6];
  for (int r = 0; r < n; ++r) {
    perm[r] = r;
    for (int c = 0; c < n; ++c) M[r * n + c] = (c >= r) ? A_upper[r * n + c] : A_upper[c * n + r];
  }
  /* In-place pivoted LDL^T: P A P^T = L D L^T, L unit lower (stored below diag), D on diag. */
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(M[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(M[i * n + i]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {
      for (int c = 0; c < n; ++c) { double t = M[k * n + c]; M[k * n + c] = M[p * n + c]; M[p * n + c] = t; }
      for (int r = 0; r < n; ++r) { double t = M[r * n + k]; M[r * n + k] = M[r * n + p]; M[r * n + p] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double d = M[k * n + k];
    if (d == 0.0) {
      /* Eigen: remaining block treated as zero. */
      for (int i = k + 1; i < n; ++i) M[i * n + k] = 0.0;
      continue;
    }
    for (int i = k + 1; i < n; ++i) M[i * n + k] /= d;
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j <= i; ++j) {
        M[i * n + j] -= M[i * n + k] * d * M[j * n + k];
        M[j * n + i] = M[i * n + j];
      }
  }
  double y[6];
  for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) y[i] -= M[i * n + j] * y[j];
  for (int i = 0; i < n; ++i) {
    double d = M[i * n + i];
    /* Eigen's LDLT::solve uses a tolerance of 1/highest (pseudo-inverse of D). */
    y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) y[i] -= M[j * n + i] * y[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
  return 0;
}

/* ---- camera frustum co-visibility test (camera_frustum.h) ---- */

typedef struct {
  float p[8][3];
  float bmin[3], bmax[3];
} hm_frustum;

/* camera_frustum.h:149-178; K = {fx, fy, cx, cy} in pixel-corner convention
 * (UnprojectFromPixelCornerConv: ((x-cx)/fx, (y-cy)/fy, 1)). */
static inline void hm_frustum_create(hm_frustum* f, const float K[4], int width, int height,
                                     float min_depth, float max_depth, const float global_T_camera[7]) {
  float M[12];
  hm_se3_matrix3x4(global_T_camera, M);
  const float cxs[4] = {0.f, (float)width, 0.f, (float)width};
  const float cys[4] = {0.f, 0.f, (float)height, (float)height};
  for (int i = 0; i < 3; ++i) { f->bmin[i] = INFINITY; f->bmax[i] = -INFINITY; }
  for (int c = 0; c < 4; ++c) {
    float dx = (cxs[c] - K[2]) / K[0];
    float dy = (cys[c] - K[3]) / K[1];
    for (int d = 0; d < 2; ++d) {
      float depth = d == 0 ? min_depth : max_depth;
      float v[3] = {depth * dx, depth * dy, depth};
      float* o = f->p[2 * c + d];
      for (int r = 0; r < 3; ++r) {
        o[r] = M[r * 4 + 0] * v[0] + M[r * 4 + 1] * v[1] + M[r * 4 + 2] * v[2] + M[r * 4 + 3];
        if (o[r] < f->bmin[r]) f->bmin[r] = o[r];
        if (o[r] > f->bmax[r]) f->bmax[r] = o[r];
      }
    }
  }
}

static inline void hm_v3_sub(const float a[3], const float b[3], float o[3]) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void hm_v3_cross(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float hm_v3_dot(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* camera_frustum.h:180-218 */
static inline void hm_frustum_axes_planes(const hm_frustum* f, float axes[6][3], float pn[6][3], float pd[6]) {
  hm_v3_sub(f->p[7], f->p[6], axes[0]);
  hm_v3_sub(f->p[3], f->p[2], axes[1]);
  hm_v3_sub(f->p[5], f->p[4], axes[2]);
  hm_v3_sub(f->p[1], f->p[0], axes[3]);
  hm_v3_sub(f->p[2], f->p[6], axes[4]);
  hm_v3_sub(f->p[0], f->p[2], axes[5]);
  float fwd[3];
  hm_v3_cross(axes[5], axes[4], fwd);
  for (int i = 0; i < 3; ++i) { pn[0][i] = fwd[i]; pn[1][i] = -fwd[i]; }
  pd[0] = -hm_v3_dot(fwd, f->p[1]);
  pd[1] = hm_v3_dot(fwd, f->p[0]);
  hm_v3_cross(axes[0], axes[4], pn[2]); pd[2] = -hm_v3_dot(pn[2], f->p[6]);
  hm_v3_cross(axes[1], axes[5], pn[3]); pd[3] = -hm_v3_dot(pn[3], f->p[2]);
  hm_v3_cross(axes[4], axes[2], pn[4]); pd[4] = -hm_v3_dot(pn[4], f->p[4]);
  hm_v3_cross(axes[5], axes[0], pn[5]); pd[5] = -hm_v3_dot(pn[5], f->p[6]);
}

/* camera_frustum.h:73-143 */
static inline int hm_frustum_intersects(const hm_frustum* a, const hm_frustum* b) {
  /* Eigen AlignedBox intersection().isEmpty(): empty if any min > max. */
  for (int i = 0; i < 3; ++i) {
    float lo = a->bmin[i] > b->bmin[i] ? a->bmin[i] : b->bmin[i];
    float hi = a->bmax[i] < b->bmax[i] ? a->bmax[i] : b->bmax[i];
    if (lo > hi) return 0;
  }
  float axa[6][3], pna[6][3], pda[6], axb[6][3], pnb[6][3], pdb[6];
  hm_frustum_axes_planes(a, axa, pna, pda);
  for (int pl = 0; pl < 6; ++pl) {
    int v = 0;
    for (; v < 8; ++v) if (hm_v3_dot(pna[pl], b->p[v]) + pda[pl] < 0) break;
    if (v == 8) return 0;
  }
  hm_frustum_axes_planes(b, axb, pnb, pdb);
  for (int pl = 0; pl < 6; ++pl) {
    int v = 0;
    for (; v < 8; ++v) if (hm_v3_dot(pnb[pl], a->p[v]) + pdb[pl] < 0) break;
    if (v == 8) return 0;
  }
  /* NOTE: the reference crosses axes_[this_edge] with axes_[other_edge] of the
   * SAME frustum (camera_frustum.h:122: both from `this`); restated as is. */
  for (int e1 = 0; e1 < 6; ++e1) {
    for (int e2 = 0; e2 < 6; ++e2) {
      float dir[3];
      hm_v3_cross(axa[e1], axa[e2], dir);
      if (hm_v3_dot(dir, dir) < 1e-5f) continue;
      float amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
      for (int p = 0; p < 8; ++p) {
        float va = hm_v3_dot(dir, a->p[p]);
        float vb = hm_v3_dot(dir, b->p[p]);
        if (va < amin) amin = va;
        if (va > amax) amax = va;
        if (vb < bmin) bmin = vb;
        if (vb > bmax) bmax = vb;
      }
      if (amax <= bmin || amin >= bmax) return 0;
    }
  }
  return 1;
}

#ifdef __cplusplus
}
#endif
#endif
