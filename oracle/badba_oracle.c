/* oracle/badba_oracle.c -- TEST INFRASTRUCTURE ONLY (see badba_oracle.h).
 *
 * CPU restatement of the reference's direct-BA hot path.  Every function cites
 * the reference file:line it follows (paths relative to
 * /root/reference/applications/badslam/src/badslam/).  Arithmetic is fp32 like the
 * device code, sums are accumulated in fp64 (the reference sums in fp32 with a
 * non-deterministic atomic order, gauss_newton.cuh:63-91, so fp64 sums are the
 * value both GPU implementations scatter around).  -use_fast_math approximations
 * of the reference build are NOT mimicked.
 */
#include "badba_oracle.h"
#include "host_math.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants: cost_function.cuh:44-52,105-109,126 ; kernels.cuh:38-58 ---- */
#define K_DEPTH_RESIDUAL_WEIGHT 1.f
#define K_DEPTH_TUKEY 10.f
#define K_DEPTH_UNCERTAINTY_FACTOR 0.1f
#define K_DESC_RESIDUAL_WEIGHT 1e-2f
#define K_DESC_HUBER 10.f
#define K_TANGENT_SCALING 2.0f
#define K_INVALID_DEPTH_BIT 0x8000u
#define K_COS_NORMAL_COMPAT 0.76604f
#define K_SURFEL_ACTIVE_FLAG 1u

enum { ROW_X = 0, ROW_Y, ROW_Z, ROW_NORMAL, ROW_R2, ROW_COLOR, ROW_D1, ROW_D2, ROW_ACC0 };

static int g_tex_mode = 3;
void orc_set_tex_mode(int mode) { g_tex_mode = mode; }
int orc_get_tex_mode(void) { return g_tex_mode; }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

typedef struct { float x, y, z; } f3;
static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 scale3(f3 a, float s) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross3(f3 a, f3 b) {  /* cuda_util.cuh:76-80 */
  return mk3(a.y * b.z - b.y * a.z, b.x * a.z - a.x * b.z, a.x * b.y - b.x * a.y);
}
/* cuda_matrix.cuh:104-135 */
static inline f3 T_mul(const float T[12], f3 p) {
  return mk3(T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3],
             T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
             T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]);
}
static inline f3 T_rot(const float T[12], f3 p) {
  return mk3(T[0] * p.x + T[1] * p.y + T[2] * p.z,
             T[4] * p.x + T[5] * p.y + T[6] * p.z,
             T[8] * p.x + T[9] * p.y + T[10] * p.z);
}

/* robust_weighting.cuh:39-86 */
static inline float tukey_residual(float r, float p) {
  if (fabsf(r) < p) {
    float q = r / p, t = 1.f - q * q;
    return (1 / 6.f) * p * p * (1 - t * t * t);
  }
  return (1 / 6.f) * p * p;
}
static inline float tukey_weight(float r, float p) {
  if (fabsf(r) < p) {
    float q = r / p, t = 1.f - q * q;
    return t * t;
  }
  return 0.f;
}
static inline float huber_residual(float r, float p) {
  float a = fabsf(r);
  return (a < p) ? 0.5f * r * r : p * (a - 0.5f * p);
}
static inline float huber_weight(float r, float p) {
  float a = fabsf(r);
  return (a < p) ? 1.f : (p / a);
}
/* cost_function.cuh:91-98,177-185 */
static inline float depth_weight(float r) { return K_DEPTH_RESIDUAL_WEIGHT * tukey_weight(r, K_DEPTH_TUKEY); }
static inline float depth_cost(float r) { return K_DEPTH_RESIDUAL_WEIGHT * tukey_residual(r, K_DEPTH_TUKEY); }
static inline float desc_weight(float r) { return K_DESC_RESIDUAL_WEIGHT * huber_weight(r, K_DESC_HUBER); }
static inline float desc_cost(float r) { return K_DESC_RESIDUAL_WEIGHT * huber_residual(r, K_DESC_HUBER); }

/* util.cuh:62-69 */
static inline float raw_to_calibrated_depth(float a, float cfactor, float raw_to_float, uint16_t measured) {
  const float inv_depth = 1.0f / (raw_to_float * measured);
  return 1.f / (inv_depth + cfactor * expf(-a * inv_depth));
}
/* util.cuh:126-146 */
static inline f3 u16_to_image_space_normal(uint16_t v) {
  f3 r;
  r.x = (int8_t)(v & 0x00ff) * (1.0f / 127);
  r.y = (int8_t)((v & 0xff00) >> 8) * (1.0f / 127);
  r.z = 1 - r.x * r.x - r.y * r.y;
  r.z = -sqrtf((r.z > 0.f) ? r.z : 0.f);
  return r;
}
/* util_nvcc_only.cuh:67-95 */
static inline uint32_t small_float_to_ten_bit_signed(float value) {
  return 0x03ffu & (uint16_t)((int16_t)(value * 511 + ((value > 0) ? 0.5f : -0.5f)));
}
static inline float ten_bit_signed_to_small_float(uint32_t value) {
  uint16_t temp = (uint16_t)(((0x0200 & value) ? 0xfc00 : 0) | (0x03ff & value));
  return (int16_t)temp * (1.0f / 511);
}
static inline uint32_t pack_normal(f3 n) {
  return (small_float_to_ten_bit_signed(n.x) << 0) | (small_float_to_ten_bit_signed(n.y) << 10) |
         (small_float_to_ten_bit_signed(n.z) << 20);
}
static inline f3 unpack_normal(uint32_t v) {
  f3 n = mk3(ten_bit_signed_to_small_float(v >> 0), ten_bit_signed_to_small_float(v >> 10),
             ten_bit_signed_to_small_float(v >> 20));
  float factor = 1.0f / sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
  return scale3(n, factor);
}
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- per-keyframe view ---- */
typedef struct {
  const uint16_t* depth;
  const uint16_t* normals;
  const uint8_t* color;
  int w, h, cw, ch;
  /* surfel_projection.h:42-124 builders */
  float fx, fy, cx, cy;                 /* PixelCornerProjector (depth) */
  float fx_inv, fy_inv, cx_inv, cy_inv; /* PixelCenterUnprojector (depth) */
  float d2c_fx, d2c_fy, d2c_cx, d2c_cy; /* DepthToColorPixelCorner */
  float cfx, cfy, ccx, ccy;             /* colour PixelCornerProjector; PixelCenterProjector has cx-0.5 */
  float a, raw_to_float, baseline_fx;
  int cell, cf_w;
  const float* cfactor;
} kfview;

static void make_view(const orc_model* m, const orc_keyframes* kfs, int k, kfview* v) {
  size_t npx = (size_t)m->depth_w * m->depth_h;
  size_t ncpx = (size_t)m->color_w * m->color_h;
  v->depth = kfs->depth + npx * k;
  v->normals = kfs->normals + npx * k;
  v->color = kfs->color + ncpx * 4 * k;
  v->w = m->depth_w; v->h = m->depth_h; v->cw = m->color_w; v->ch = m->color_h;
  v->fx = m->depth_K[0]; v->fy = m->depth_K[1]; v->cx = m->depth_K[2]; v->cy = m->depth_K[3];
  v->fx_inv = 1.0f / v->fx;
  v->fy_inv = 1.0f / v->fy;
  v->cx_inv = -(v->cx - 0.5f) * v->fx_inv;
  v->cy_inv = -(v->cy - 0.5f) * v->fy_inv;
  v->cfx = m->color_K[0]; v->cfy = m->color_K[1]; v->ccx = m->color_K[2]; v->ccy = m->color_K[3];
  v->d2c_fx = v->cfx / v->fx;
  v->d2c_cx = -1 * v->cfx * v->cx / v->fx + v->ccx;
  v->d2c_fy = v->cfy / v->fy;
  v->d2c_cy = -1 * v->cfy * v->cy / v->fy + v->ccy;
  v->a = m->a; v->raw_to_float = m->raw_to_float_depth; v->baseline_fx = m->baseline_fx;
  v->cell = m->cell; v->cf_w = m->cf_w; v->cfactor = m->cfactor;
}

/* Emulates tex2D<float4>(color_texture, x, y).w for the texture of keyframe.cc:67-73
 * (clamp addressing, linear filter, normalized-float read, unnormalized coords). */
static inline float texel(const kfview* v, int i, int j) {
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  if (i > v->cw - 1) i = v->cw - 1;
  if (j > v->ch - 1) j = v->ch - 1;
  return v->color[((size_t)j * v->cw + i) * 4 + 3] * (1.0f / 255.0f);
}
static inline int texel_u8(const kfview* v, long i, long j) {
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  if (i > v->cw - 1) i = v->cw - 1;
  if (j > v->ch - 1) j = v->ch - 1;
  return v->color[((size_t)j * v->cw + i) * 4 + 3];
}
/* Modes 3/4 restate what the B200 texture unit was MEASURED to do (tools/tex_probe*.cu, profiles/texture_filter.md;
 * bit-exact on 2^20 random samples of a random 8-bit image, clamped borders included):
 *   - fractions a, b in 1.8 fixed point, round to nearest;
 *   - the weight of the far texel is the 8-bit rounded product  w11 = (a*b + 128) >> 8, the others follow by
 *     subtraction  w10 = a - w11, w01 = b - w11, w00 = 256 - w11 - w10 - w01  (so the weights always sum to 256);
 *   - texels are widened to unorm16 (t * 257), the weighted sum is rounded once ((s + 128) >> 8) and returned as r / 65535.
 * Mode 3 takes the fractions from the float coordinate, mode 4 converts the coordinate to fixed point first. */
static inline float tex_w_hw(const kfview* v, float x, float y, int fixed_first) {
  long i, j, a, b;
  if (fixed_first) {
    long x8 = lrintf(x * 256.f) - 128, y8 = lrintf(y * 256.f) - 128;
    i = x8 >> 8; j = y8 >> 8; a = x8 & 255; b = y8 & 255;
  } else {
    float xb = x - 0.5f, yb = y - 0.5f;
    float fi = floorf(xb), fj = floorf(yb);
    i = (long)fi; j = (long)fj;
    a = (long)floorf((xb - fi) * 256.f + 0.5f);
    b = (long)floorf((yb - fj) * 256.f + 0.5f);
  }
  long w11 = (a * b + 128) >> 8, w10 = a - w11, w01 = b - w11, w00 = 256 - w11 - w10 - w01;
  long sum = w00 * (texel_u8(v, i, j) * 257L) + w10 * (texel_u8(v, i + 1, j) * 257L) +
             w01 * (texel_u8(v, i, j + 1) * 257L) + w11 * (texel_u8(v, i + 1, j + 1) * 257L);
  return (float)((sum + 128) >> 8) / 65535.f;
}
static inline float tex_w(const kfview* v, float x, float y) {
  if (g_tex_mode >= 3) return tex_w_hw(v, x, y, g_tex_mode == 4);
  float xb = x - 0.5f, yb = y - 0.5f;
  float fi = floorf(xb), fj = floorf(yb);
  float al = xb - fi, be = yb - fj;
  if (g_tex_mode == 1) {
    al = floorf(al * 256.f + 0.5f) * (1.f / 256.f);
    be = floorf(be * 256.f + 0.5f) * (1.f / 256.f);
  } else if (g_tex_mode == 2) {
    al = floorf(al * 256.f) * (1.f / 256.f);
    be = floorf(be * 256.f) * (1.f / 256.f);
  }
  int i = (int)fi, j = (int)fj;
  float t00 = texel(v, i, j), t10 = texel(v, i + 1, j), t01 = texel(v, i, j + 1), t11 = texel(v, i + 1, j + 1);
  return (1.f - al) * (1.f - be) * t00 + al * (1.f - be) * t10 + (1.f - al) * be * t01 + al * be * t11;
}

float orc_tex_luma(const orc_model* m, const orc_keyframes* kfs, int k, float x, float y) {
  kfview v;
  make_view(m, kfs, k, &v);
  return tex_w(&v, x, y);
}

/* ---- projection + association ---- */
typedef struct {
  f3 gp;        /* surfel global position */
  f3 lp;        /* local position */
  f3 n;         /* global normal (unpacked, normalised) */
  float d;      /* calibrated depth of the pixel */
  int px, py;
  float pxf, pyf;
} assoc;

/* Stage reached (for the SURVEY 8d byte model): 0 culled, 1 in image, 2 depth ok (KF normal read), 3 associated.
 * surfel_projection_nvcc_only.cuh:48-127 (IsAssociatedWithPixel), :302-359 (SurfelProjectsToAssociatedPixel),
 * util.cuh:83-118 (ProjectSurfelToImage), cuda_matrix.cuh:115-124. */
static inline int project_associate(const kfview* v, const float T[12], f3 gp, uint32_t packed_normal, assoc* r) {
  r->gp = gp;
  r->lp.z = T[8] * gp.x + T[9] * gp.y + T[10] * gp.z + T[11];
  if (r->lp.z <= 0.f) return 0;
  r->lp.x = T[0] * gp.x + T[1] * gp.y + T[2] * gp.z + T[3];
  r->lp.y = T[4] * gp.x + T[5] * gp.y + T[6] * gp.z + T[7];
  r->pxf = v->fx * (r->lp.x / r->lp.z) + v->cx;
  r->pyf = v->fy * (r->lp.y / r->lp.z) + v->cy;
  /* static_cast<int> of a float out of int range is UB on the host; the device saturates.
   * Reject non-finite / huge values first (they fail the bounds test on the device as well). */
  if (!(r->pxf >= 0.f) || !(r->pyf >= 0.f) || !(r->pxf < 1e9f) || !(r->pyf < 1e9f)) return 0;
  r->px = (int)r->pxf;
  r->py = (int)r->pyf;
  if (r->px >= v->w || r->py >= v->h) return 0;

  uint16_t measured = v->depth[(size_t)r->py * v->w + r->px];
  if (measured & K_INVALID_DEPTH_BIT) return 1;
  r->d = raw_to_calibrated_depth(v->a, v->cfactor[(size_t)(r->py / v->cell) * v->cf_w + (r->px / v->cell)],
                                 v->raw_to_float, measured);
  r->n = unpack_normal(packed_normal);
  f3 ln = T_rot(T, r->n);
  /* cost_function.cuh:81-83 */
  float nx = v->fx_inv * r->px + v->cx_inv, ny = v->fy_inv * r->py + v->cy_inv;
  float stddev = (K_DEPTH_UNCERTAINTY_FACTOR * fabsf(ln.x * nx + ln.y * ny + ln.z) * (r->d * r->d)) / v->baseline_fx;
  float thr = K_DEPTH_TUKEY * stddev;
  if (fabsf(r->lp.z - r->d) > thr) return 1;
  float dist = sqrtf(dot3(r->lp, r->lp));
  float facing = (1.0f / dist) * dot3(r->lp, ln);
  if (facing > 0) return 1;
  f3 pn = u16_to_image_space_normal(v->normals[(size_t)r->py * v->w + r->px]);
  if (dot3(ln, pn) < K_COS_NORMAL_COMPAT) return 2;
  return 3;
}

/* surfel_projection.cuh:196-207 */
static inline int depth_to_color(const kfview* v, float pxf, float pyf, float* cx, float* cy) {
  *cx = v->d2c_fx * pxf + v->d2c_cx;
  *cy = v->d2c_fy * pyf + v->d2c_cy;
  return *cx >= 0 && *cy >= 0 && (int)*cx < v->cw && (int)*cy < v->ch;
}

/* cost_function.cuh:115-136 */
static inline void tangent_projections(const kfview* v, const float T[12], f3 gp, f3 n, float r2,
                                       float* t1x, float* t1y, float* t2x, float* t2y) {
  f3 t1 = cross3(n, (fabsf(n.x) > 0.9f) ? mk3(0, 1, 0) : mk3(1, 0, 0));
  t1 = scale3(t1, K_TANGENT_SCALING * sqrtf(r2 / fmaxf(1e-12f, dot3(t1, t1))));
  f3 p1 = T_mul(T, add3(gp, t1));
  *t1x = v->cfx * (p1.x / p1.z) + v->ccx;
  *t1y = v->cfy * (p1.y / p1.z) + v->ccy;
  f3 t2 = cross3(n, t1);
  t2 = scale3(t2, K_TANGENT_SCALING * sqrtf(r2 / fmaxf(1e-12f, dot3(t2, t2))));
  f3 p2 = T_mul(T, add3(gp, t2));
  *t2x = v->cfx * (p2.x / p2.z) + v->ccx;
  *t2y = v->cfy * (p2.y / p2.z) + v->ccy;
}

/* cost_function.cuh:191-254: finite-difference gradient at one sample point. */
static inline void point_gradient(const kfview* v, float x, float y, float* dx, float* dy) {
  float ax = fmaxf(0.f, x - 0.5f), ay = fmaxf(0.f, y - 0.5f);
  int ix = (ax < 2e9f) ? (int)ax : 2000000000;
  int iy = (ay < 2e9f) ? (int)ay : 2000000000;
  float tx = fmaxf(0.f, fminf(1.f, x - 0.5f - ix));
  float ty = fmaxf(0.f, fminf(1.f, y - 0.5f - iy));
  float tl = texel(v, ix, iy), tr = texel(v, ix + 1, iy), bl = texel(v, ix, iy + 1), br = texel(v, ix + 1, iy + 1);
  *dx = (br - bl) * ty + (tr - tl) * (1 - ty);
  *dy = (br - tr) * tx + (bl - tl) * (1 - tx);
}

typedef struct {
  float r1, r2;                  /* raw descriptor residuals */
  float gx1, gy1, gx2, gy2;      /* DescriptorJacobianWrtProjectedPosition outputs */
} desc_eval;

static inline void descriptor_eval(const kfview* v, float cx, float cy, float t1x, float t1y, float t2x, float t2y,
                                   float d1, float d2, desc_eval* e) {
  /* cost_function.cuh:140-156 */
  float intensity = tex_w(v, cx, cy);
  float t1i = tex_w(v, t1x, t1y), t2i = tex_w(v, t2x, t2y);
  e->r1 = (180.f * (t1i - intensity)) - d1;
  e->r2 = (180.f * (t2i - intensity)) - d2;
  float cdx, cdy, t1dx, t1dy, t2dx, t2dy;
  point_gradient(v, cx, cy, &cdx, &cdy);
  point_gradient(v, t1x, t1y, &t1dx, &t1dy);
  point_gradient(v, t2x, t2y, &t2dx, &t2dy);
  e->gx1 = 180.f * (t1dx - cdx);
  e->gy1 = 180.f * (t1dy - cdy);
  e->gx2 = 180.f * (t2dx - cdx);
  e->gy2 = 180.f * (t2dy - cdy);
}

/* kernel_opt_pose.cu:96-142 */
static inline void desc_pose_jacobian(const kfview* v, f3 ls, float gx, float gy, float J[6]) {
  gx *= v->cfx;
  gy *= v->cfy;
  float inv_z = 1.f / ls.z, z_sq = ls.z * ls.z, inv_z_sq = inv_z * inv_z, xy = ls.x * ls.y;
  J[0] = -gx * inv_z;
  J[1] = -gy * inv_z;
  J[2] = (ls.x * gx + ls.y * gy) * inv_z_sq;
  J[3] = ((ls.y * ls.y + z_sq) * gy + xy * gx) * inv_z_sq;
  J[4] = -((ls.x * ls.x + z_sq) * gx + xy * gy) * inv_z_sq;
  J[5] = -(ls.x * gy - ls.y * gx) * inv_z;
}

/* kernel_opt_pose.cu:45-94 ; returns raw residual */
static inline float depth_pose_residual_jacobian(const kfview* v, const assoc* r, f3 ln, float J[6], float* inv_stddev_out,
                                                 f3* unproj_out) {
  float nx = v->fx_inv * r->px + v->cx_inv, ny = v->fy_inv * r->py + v->cy_inv;
  /* cost_function.cuh:86-88 */
  float inv_stddev = v->baseline_fx / (K_DEPTH_UNCERTAINTY_FACTOR * fabsf(ln.x * nx + ln.y * ny + ln.z) * (r->d * r->d));
  f3 up = mk3(r->d * nx, r->d * ny, r->d);
  float raw = inv_stddev * dot3(ln, sub3(up, r->lp));
  if (J) {
    J[0] = inv_stddev * ln.x;
    J[1] = inv_stddev * ln.y;
    J[2] = inv_stddev * ln.z;
    J[3] = inv_stddev * (-ln.y * up.z + ln.z * up.y);
    J[4] = inv_stddev * (ln.x * up.z - ln.z * up.x);
    J[5] = inv_stddev * (-ln.x * up.y + ln.y * up.x);
  }
  if (inv_stddev_out) *inv_stddev_out = inv_stddev;
  if (unproj_out) *unproj_out = up;
  return raw;
}

static inline void accum_hb(double H[21], double b[6], const float J[6], float raw, float w) {
  int idx = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) H[idx++] += (double)(w * J[r] * J[c]);
  float wr = w * raw;
  for (int i = 0; i < 6; ++i) b[i] += (double)(wr * J[i]);
}

void orc_frame_T_global(const float global_T_frame[7], float out12[12]) {
  float inv[7];
  hm_se3_inverse(global_T_frame, inv);
  hm_se3_matrix3x4(inv, out12);
}

/* kernel_opt_pose.cu:251-383 */
void orc_pose_coeffs(const orc_model* m, const orc_keyframes* kfs, int k, const float T[12],
                     const float* surfels, int pitch, uint32_t n, orc_pose_stats* out) {
  kfview v;
  make_view(m, kfs, k, &v);
  memset(out, 0, sizeof(*out));
  out->n_pair = n;
  const int use_depth = m->use_depth_residuals, use_desc = m->use_descriptor_residuals;
#pragma omp parallel
  {
    orc_pose_stats loc;
    memset(&loc, 0, sizeof(loc));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      assoc r;
      f3 gp = mk3(surfels[ROW_X * (size_t)pitch + i], surfels[ROW_Y * (size_t)pitch + i], surfels[ROW_Z * (size_t)pitch + i]);
      int st = project_associate(&v, T, gp, f2u(surfels[ROW_NORMAL * (size_t)pitch + i]), &r);
      if (st >= 1) loc.n_inimg++;
      if (st >= 2) loc.n_depthok++;
      if (st < 3) continue;
      loc.n_assoc++;
      float J[6];
      if (use_depth) {
        f3 ln = T_rot(T, r.n);
        float raw = depth_pose_residual_jacobian(&v, &r, ln, J, NULL, NULL);
        accum_hb(loc.H, loc.b, J, raw, depth_weight(raw));
        loc.cost_depth += depth_cost(raw);
      }
      if (use_desc) {
        float ccx, ccy;
        if (depth_to_color(&v, r.pxf, r.pyf, &ccx, &ccy)) {
          loc.n_photo++;
          float t1x, t1y, t2x, t2y;
          tangent_projections(&v, T, r.gp, r.n, surfels[ROW_R2 * (size_t)pitch + i], &t1x, &t1y, &t2x, &t2y);
          desc_eval e;
          descriptor_eval(&v, ccx, ccy, t1x, t1y, t2x, t2y, surfels[ROW_D1 * (size_t)pitch + i],
                          surfels[ROW_D2 * (size_t)pitch + i], &e);
          desc_pose_jacobian(&v, r.lp, e.gx1, e.gy1, J);
          accum_hb(loc.H, loc.b, J, e.r1, desc_weight(e.r1));
          desc_pose_jacobian(&v, r.lp, e.gx2, e.gy2, J);
          accum_hb(loc.H, loc.b, J, e.r2, desc_weight(e.r2));
          loc.cost_desc1 += desc_cost(e.r1);
          loc.cost_desc2 += desc_cost(e.r2);
        }
      }
    }
#pragma omp critical
    {
      for (int j = 0; j < 21; ++j) out->H[j] += loc.H[j];
      for (int j = 0; j < 6; ++j) out->b[j] += loc.b[j];
      out->n_inimg += loc.n_inimg; out->n_depthok += loc.n_depthok;
      out->n_assoc += loc.n_assoc; out->n_photo += loc.n_photo;
      out->cost_depth += loc.cost_depth; out->cost_desc1 += loc.cost_desc1; out->cost_desc2 += loc.cost_desc2;
    }
  }
  if (!use_depth) out->cost_depth = 0;
}

int orc_pair_residuals_debug(const orc_model* m, const orc_keyframes* kfs, int k, const float T[12],
                             const float surfel[8], float r_out[3], float J_pose[18], float J_geom[9], float dbg[16]) {
  kfview v;
  make_view(m, kfs, k, &v);
  assoc r;
  int st = project_associate(&v, T, mk3(surfel[0], surfel[1], surfel[2]), f2u(surfel[3]), &r);
  memset(r_out, 0, 3 * sizeof(float));
  memset(J_pose, 0, 18 * sizeof(float));
  memset(J_geom, 0, 9 * sizeof(float));
  if (dbg) memset(dbg, 0, 16 * sizeof(float));
  if (st < 3) return 0;
  int flags = 1;
  f3 ln = T_rot(T, r.n);
  float inv_stddev;
  r_out[0] = depth_pose_residual_jacobian(&v, &r, ln, J_pose, &inv_stddev, NULL);
  J_geom[0] = -inv_stddev;  /* kernel_opt_geometry.cu:138 */
  if (dbg) { dbg[0] = (float)r.px; dbg[1] = (float)r.py; dbg[2] = r.d; dbg[3] = inv_stddev; }
  float ccx, ccy;
  if (depth_to_color(&v, r.pxf, r.pyf, &ccx, &ccy)) {
    flags |= 2;
    float t1x, t1y, t2x, t2y;
    tangent_projections(&v, T, r.gp, r.n, surfel[4], &t1x, &t1y, &t2x, &t2y);
    desc_eval e;
    descriptor_eval(&v, ccx, ccy, t1x, t1y, t2x, t2y, surfel[6], surfel[7], &e);
    r_out[1] = e.r1;
    r_out[2] = e.r2;
    desc_pose_jacobian(&v, r.lp, e.gx1, e.gy1, J_pose + 6);
    desc_pose_jacobian(&v, r.lp, e.gx2, e.gy2, J_pose + 12);
    /* kernel_opt_geometry.cu:176-181 */
    float term1 = -v.cfx * (ln.x * r.lp.z - ln.z * r.lp.x);
    float term2 = -v.cfy * (ln.y * r.lp.z - ln.z * r.lp.y);
    float term3 = 1.f / (r.lp.z * r.lp.z);
    J_geom[3] = -(e.gx1 * term1 + e.gy1 * term2) * term3;
    J_geom[6] = -(e.gx2 * term1 + e.gy2 * term2) * term3;
    J_geom[4] = -1.f;
    J_geom[8] = -1.f;
    if (dbg) {
      dbg[4] = ccx; dbg[5] = ccy; dbg[6] = t1x; dbg[7] = t1y; dbg[8] = t2x; dbg[9] = t2y;
      dbg[10] = e.gx1; dbg[11] = e.gy1; dbg[12] = e.gx2; dbg[13] = e.gy2;
    }
  }
  return flags;
}

int orc_pair_residuals(const orc_model* m, const orc_keyframes* kfs, int k, const float T[12],
                       const float surfel[8], float r_out[3], float J_pose[18], float J_geom[9]) {
  return orc_pair_residuals_debug(m, kfs, k, T, surfel, r_out, J_pose, J_geom, NULL);
}

/* direct_ba_alternating.cc:42-283 */
int orc_estimate_frame_pose(const orc_model* m, const orc_keyframes* kfs, int k, const float init[7],
                            const float* surfels, int pitch, uint32_t n, float out[7], int* converged_out,
                            int max_iterations) {
  float est[7];
  memcpy(est, init, sizeof(est));
  int converged = 0, iteration;
  for (iteration = 0; iteration < max_iterations; ++iteration) {
    float T[12];
    orc_frame_T_global(est, T);
    double Hd[36], bd[6], xd[6];
    memset(Hd, 0, sizeof(Hd));
    if (n == 0) {
      memset(bd, 0, sizeof(bd));
    } else {
      orc_pose_stats st;
      orc_pose_coeffs(m, kfs, k, T, surfels, pitch, n, &st);
      int idx = 0;
      /* the reference downloads fp32 H/b and casts to double (:173-179,206) */
      for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) Hd[r * 6 + c] = (double)(float)st.H[idx++];
      for (int i = 0; i < 6; ++i) bd[i] = (double)(float)st.b[i];
    }
    hm_ldlt_solve(6, Hd, bd, xd);
    float x[6], nx[6], e[7], next[7];
    for (int i = 0; i < 6; ++i) { x[i] = (float)xd[i]; nx[i] = -x[i]; }
    hm_se3_exp(nx, e);
    hm_se3_mul(est, e, next);
    memcpy(est, next, sizeof(est));
    converged = hm_is_scale1_pose_converged(x);
    if (converged) { ++iteration; break; }
  }
  memcpy(out, est, sizeof(est));
  if (converged_out) *converged_out = converged;
  return iteration;
}

/* kernel_surfel_activation.cu:38-79 + kernel_surfel_activation.cc:39-67 */
void orc_update_activation(const orc_model* m, const orc_keyframes* kfs, const float* surfels, int pitch,
                           uint32_t n, uint8_t* active) {
  if (n == 0) return;
  int K = kfs->K;
  float* Ts = (float*)malloc(sizeof(float) * 12 * (size_t)K);
  kfview* vs = (kfview*)malloc(sizeof(kfview) * (size_t)K);
  for (int k = 0; k < K; ++k) { orc_frame_T_global(kfs->global_T_frame + 7 * k, Ts + 12 * k); make_view(m, kfs, k, vs + k); }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    uint8_t flag = active[i] & (uint8_t)~K_SURFEL_ACTIVE_FLAG;
    f3 gp = mk3(surfels[ROW_X * (size_t)pitch + i], surfels[ROW_Y * (size_t)pitch + i], surfels[ROW_Z * (size_t)pitch + i]);
    uint32_t pn = f2u(surfels[ROW_NORMAL * (size_t)pitch + i]);
    for (int k = 0; k < K; ++k) {
      if (kfs->activation[k] != ORC_KF_ACTIVE) continue;
      assoc r;
      if (project_associate(vs + k, Ts + 12 * k, gp, pn, &r) == 3) { flag = K_SURFEL_ACTIVE_FLAG; break; }
    }
    active[i] = flag;
  }
  free(Ts);
  free(vs);
}

/* kernel_opt_geometry.cc:80-201 with kernels kernel_opt_geometry.cu (cited inline).
 * Surfel-major evaluation is exact: one thread owns one surfel in every reference kernel and
 * keyframes are visited in index order, so the fp32 accumulation order is identical. */
/* UpdateSurfelNormalsCUDA (kernel_opt_geometry.cc:39-77): accumulate :527-557, update :577-597. */
static void update_normals(const orc_keyframes* kfs, const float* Ts, const kfview* vs, float* surfels, size_t P, uint32_t n,
                           const uint8_t* active) {
  const int K = kfs->K;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    if (!(active[i] & K_SURFEL_ACTIVE_FLAG)) continue;
    f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
    uint32_t pn = f2u(surfels[ROW_NORMAL * P + i]);
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 0; k < K; ++k) {
      if (kfs->activation[k] == ORC_KF_INACTIVE) continue;
      assoc r;
      const float* T = Ts + 12 * k;
      if (project_associate(vs + k, T, gp, pn, &r) != 3) continue;
      f3 ln = u16_to_image_space_normal(vs[k].normals[(size_t)r.py * vs[k].w + r.px]);
      /* global_R_frame = (frame_T_global rotation)^T (keyframe.h: global_T_frame.rotationMatrix()) */
      a0 += T[0] * ln.x + T[4] * ln.y + T[8] * ln.z;
      a1 += T[1] * ln.x + T[5] * ln.y + T[9] * ln.z;
      a2 += T[2] * ln.x + T[6] * ln.y + T[10] * ln.z;
      a3 += 1.f;
    }
    surfels[(ROW_ACC0 + 0) * P + i] = a0; surfels[(ROW_ACC0 + 1) * P + i] = a1;
    surfels[(ROW_ACC0 + 2) * P + i] = a2; surfels[(ROW_ACC0 + 3) * P + i] = a3;
    if (a3 >= 1) {
      float inv = 1.f / a3;
      surfels[ROW_NORMAL * P + i] = u2f(pack_normal(mk3(inv * a0, inv * a1, inv * a2)));
    }
  }

}

void orc_optimize_geometry_iteration(const orc_model* m, const orc_keyframes* kfs, float* surfels, int pitch,
                                     uint32_t n, const uint8_t* active) {
  if (n == 0) return;
  int K = kfs->K;
  float* Ts = (float*)malloc(sizeof(float) * 12 * (size_t)K);
  kfview* vs = (kfview*)malloc(sizeof(kfview) * (size_t)K);
  for (int k = 0; k < K; ++k) { orc_frame_T_global(kfs->global_T_frame + 7 * k, Ts + 12 * k); make_view(m, kfs, k, vs + k); }
  const int use_depth = m->use_depth_residuals, use_desc = m->use_descriptor_residuals;
  const size_t P = (size_t)pitch;

  update_normals(kfs, Ts, vs, surfels, P, n, active);

  if (!use_desc) {
    /* --- position from depth residual only: :417-459 accumulate, :487-507 update --- */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      if (!(active[i] & K_SURFEL_ACTIVE_FLAG)) continue;
      f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
      uint32_t pn = f2u(surfels[ROW_NORMAL * P + i]);
      float H = 0, b = 0;
      for (int k = 0; k < K; ++k) {
        if (kfs->activation[k] == ORC_KF_INACTIVE) continue;
        assoc r;
        const float* T = Ts + 12 * k;
        if (project_associate(vs + k, T, gp, pn, &r) != 3) continue;
        f3 rn = T_rot(T, r.n);
        float inv_stddev;
        float raw = depth_pose_residual_jacobian(vs + k, &r, rn, NULL, &inv_stddev, NULL);
        float jac = -inv_stddev;
        float wj = depth_weight(raw) * jac;
        H += wj * jac;
        b += wj * raw;
      }
      surfels[(ROW_ACC0 + 0) * P + i] = H;
      surfels[(ROW_ACC0 + 1) * P + i] = b;
      if (H > 1e-6f) {
        float t = -1.f * b / H;
        f3 nrm = unpack_normal(pn);
        surfels[ROW_X * P + i] = gp.x + t * nrm.x;
        surfels[ROW_Y * P + i] = gp.y + t * nrm.y;
        surfels[ROW_Z * P + i] = gp.z + t * nrm.z;
      }
    }
  } else {
    /* --- position + descriptors jointly: :118-231 accumulate, :273-361 update --- */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      if (!(active[i] & K_SURFEL_ACTIVE_FLAG)) continue;
      f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
      uint32_t pn = f2u(surfels[ROW_NORMAL * P + i]);
      const float r2 = surfels[ROW_R2 * P + i];
      const float d1 = surfels[ROW_D1 * P + i], d2 = surfels[ROW_D2 * P + i];
      float A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < K; ++k) {
        if (kfs->activation[k] == ORC_KF_INACTIVE) continue;
        assoc r;
        const float* T = Ts + 12 * k;
        const kfview* v = vs + k;
        if (project_associate(v, T, gp, pn, &r) != 3) continue;
        f3 rn = T_rot(T, r.n);
        if (use_depth) {
          float inv_stddev;
          float raw = depth_pose_residual_jacobian(v, &r, rn, NULL, &inv_stddev, NULL);
          float jac = -inv_stddev;
          float w = depth_weight(raw);
          A[0] += w * jac * jac;
          A[6] += w * raw * jac;
        }
        float ccx, ccy;
        if (depth_to_color(v, r.pxf, r.pyf, &ccx, &ccy)) {
          float t1x, t1y, t2x, t2y;
          tangent_projections(v, T, r.gp, r.n, r2, &t1x, &t1y, &t2x, &t2y);
          desc_eval e;
          descriptor_eval(v, ccx, ccy, t1x, t1y, t2x, t2y, d1, d2, &e);
          float term1 = -v->cfx * (rn.x * r.lp.z - rn.z * r.lp.x);
          float term2 = -v->cfy * (rn.y * r.lp.z - rn.z * r.lp.y);
          float term3 = 1.f / (r.lp.z * r.lp.z);
          float j1 = -(e.gx1 * term1 + e.gy1 * term2) * term3;
          float j2 = -(e.gx2 * term1 + e.gy2 * term2) * term3;
          const float jd = -1.f;
          float w1 = desc_weight(e.r1), wr1 = w1 * e.r1;
          float w2 = desc_weight(e.r2), wr2 = w2 * e.r2;
          A[0] += w1 * j1 * j1 + w2 * j2 * j2;
          A[1] += w1 * j1 * jd;
          A[3] += w1 * jd * jd;
          A[6] += wr1 * j1 + wr2 * j2;
          A[7] += wr1 * jd;
          A[2] += w2 * j2 * jd;
          A[5] += w2 * jd * jd;
          A[8] += wr2 * jd;
        }
      }
      for (int j = 0; j < 9; ++j) surfels[(ROW_ACC0 + j) * P + i] = A[j];
      /* :273-361 */
      float H00 = A[0], H01 = A[1], H02 = A[2], H11 = A[3], H12 = A[4], H22 = A[5];
      const float kEps = 1e-6f;
      H00 += kEps; H11 += kEps; H22 += kEps;
      H00 = sqrtf(H00);
      H01 = H01 / H00;
      H11 = sqrtf(H11 - H01 * H01);
      H02 = H02 / H00;
      H12 = (H12 - H02 * H01) / H11;
      H22 = sqrtf(H22 - H02 * H02 - H12 * H12);
      float y0 = A[6] / H00;
      float y1 = (A[7] - H01 * y0) / H11;
      float y2 = (A[8] - H02 * y0 - H12 * y1) / H22;
      float x2 = y2 / H22;
      float x1 = (y1 - H12 * x2) / H11;
      float x0 = (y0 - H02 * x2 - H01 * x1) / H00;
      if (x0 != 0) {
        f3 nrm = unpack_normal(pn);
        surfels[ROW_X * P + i] = gp.x - x0 * nrm.x;
        surfels[ROW_Y * P + i] = gp.y - x0 * nrm.y;
        surfels[ROW_Z * P + i] = gp.z - x0 * nrm.z;
      }
      if (x1 != 0) surfels[ROW_D1 * P + i] = fmaxf(-180.f, fminf(180.f, d1 - x1));
      if (x2 != 0) surfels[ROW_D2 * P + i] = fmaxf(-180.f, fminf(180.f, d2 - x2));
    }
  }
  free(Ts);
  free(vs);
}

/* kernel_opt_intrinsics.cu:46-217 (accumulate), :265-347 (Schur intermediates), :374-424 (cfactor update)
 * and kernel_opt_intrinsics.cc:39-281 (host: prior on a, fp64 LDLT, parameter update). */
void orc_optimize_intrinsics(orc_model* m, const orc_keyframes* kfs, const float* surfels, int pitch, uint32_t n,
                             int opt_depth, int opt_color) {
  if (n == 0 || (!opt_depth && !opt_color)) return;
  const int K = kfs->K;
  const int Pn = m->cf_w * m->cf_h;
  const size_t P = (size_t)pitch;
  double A[15], b1[5], cH[10], cb[4];
  memset(A, 0, sizeof(A)); memset(b1, 0, sizeof(b1)); memset(cH, 0, sizeof(cH)); memset(cb, 0, sizeof(cb));
  double* B = (double*)calloc((size_t)5 * Pn, sizeof(double));
  double* D = (double*)calloc((size_t)Pn, sizeof(double));
  double* b2 = (double*)calloc((size_t)Pn, sizeof(double));
  uint32_t* obs = (uint32_t*)calloc((size_t)Pn, sizeof(uint32_t));

  for (int k = 0; k < K; ++k) {   /* ALL keyframes, also inactive ones (kernel_opt_intrinsics.cc:84-88) */
    kfview v;
    make_view(m, kfs, k, &v);
    float T[12];
    orc_frame_T_global(kfs->global_T_frame + 7 * k, T);
    /* serial over surfels: per-cell accumulators; fine for an oracle */
    for (uint32_t i = 0; i < n; ++i) {
      assoc r;
      f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
      if (project_associate(&v, T, gp, f2u(surfels[ROW_NORMAL * P + i]), &r) != 3) continue;
      float nx = v.fx_inv * r.px + v.cx_inv, ny = v.fy_inv * r.py + v.cy_inv;
      if (opt_depth) {
        int spx = r.px / v.cell, spy = r.py / v.cell;
        float cfactor = v.cfactor[(size_t)spy * v.cf_w + spx];
        float raw_inv_depth = 1.0f / (v.raw_to_float * v.depth[(size_t)r.py * v.w + r.px]);
        float exp_inv_depth = expf(-v.a * raw_inv_depth);
        float corrected_inv_depth = cfactor * exp_inv_depth + raw_inv_depth;
        if (fabsf(corrected_inv_depth) > 1e-4f) {
          f3 ln = T_rot(T, r.n);
          float dot = nx * ln.x + ny * ln.y + ln.z;
          float inv_stddev = v.baseline_fx / (K_DEPTH_UNCERTAINTY_FACTOR * fabsf(ln.x * nx + ln.y * ny + ln.z) * (r.d * r.d));
          float jac_base = inv_stddev * dot * exp_inv_depth / (corrected_inv_depth * corrected_inv_depth);
          float J[6];
          J[2] = inv_stddev * r.d * (r.n.x * T[0] + r.n.y * T[1] + r.n.z * T[2]);
          J[3] = inv_stddev * r.d * (r.n.x * T[4] + r.n.y * T[5] + r.n.z * T[6]);
          J[0] = r.px * J[2];
          J[1] = r.py * J[3];
          J[4] = cfactor * raw_inv_depth * jac_base;
          J[5] = -jac_base;
          f3 up = mk3(r.d * nx, r.d * ny, r.d);
          float raw = inv_stddev * dot3(ln, sub3(up, r.lp));
          int sp = spx + spy * v.cf_w;
          float w = depth_weight(raw);
          int idx = 0;
          for (int rr = 0; rr < 5; ++rr)
            for (int c = rr; c < 5; ++c) A[idx++] += (double)(w * J[rr] * J[c]);
          float wr = w * raw;
          for (int rr = 0; rr < 5; ++rr) b1[rr] += (double)(wr * J[rr]);
          for (int rr = 0; rr < 5; ++rr) B[(size_t)rr * Pn + sp] += (double)(w * J[rr] * J[5]);
          D[sp] += (double)(w * J[5] * J[5]);
          b2[sp] += (double)(w * raw * J[5]);
          obs[sp] += 1;
        }
      }
      if (opt_color) {
        float ccx, ccy;
        if (depth_to_color(&v, r.pxf, r.pyf, &ccx, &ccy)) {
          float t1x, t1y, t2x, t2y;
          tangent_projections(&v, T, r.gp, r.n, surfels[ROW_R2 * P + i], &t1x, &t1y, &t2x, &t2y);
          desc_eval e;
          descriptor_eval(&v, ccx, ccy, t1x, t1y, t2x, t2y, surfels[ROW_D1 * P + i], surfels[ROW_D2 * P + i], &e);
          float J1[4] = {e.gx1 * nx, e.gy1 * ny, e.gx1, e.gy1};
          float J2[4] = {e.gx2 * nx, e.gy2 * ny, e.gx2, e.gy2};
          const float* Js[2] = {J1, J2};
          float rs[2] = {e.r1, e.r2};
          for (int q = 0; q < 2; ++q) {
            if (rs[q] == 0) continue;   /* valid == (raw != 0), kernel_opt_intrinsics.cu:199,207 */
            float w = desc_weight(rs[q]);
            int idx = 0;
            for (int rr = 0; rr < 4; ++rr)
              for (int c = rr; c < 4; ++c) cH[idx++] += (double)(w * Js[q][rr] * Js[q][c]);
            float wr = w * rs[q];
            for (int rr = 0; rr < 4; ++rr) cb[rr] += (double)(wr * Js[q][rr]);
          }
        }
      }
    }
  }

  if (opt_depth) {
    /* buffers are fp32 on the device: round the accumulated values to float first */
    float* Bf = (float*)malloc(sizeof(float) * 5 * (size_t)Pn);
    float* Df = (float*)malloc(sizeof(float) * (size_t)Pn);
    for (int p = 0; p < Pn; ++p) { Df[p] = (float)D[p]; for (int r = 0; r < 5; ++r) Bf[(size_t)r * Pn + p] = (float)B[(size_t)r * Pn + p]; }
    double Ad[15], b1d[5];
    for (int i = 0; i < 15; ++i) Ad[i] = (double)(float)A[i];
    for (int i = 0; i < 5; ++i) b1d[i] = (double)(float)b1[i];
    for (int p = 0; p < Pn; ++p) {
      const float D_inverse = 1.0f / Df[p];
      if (!(D_inverse < 1e12f)) { Df[p] = NAN; continue; }
      float D_inv_b2 = D_inverse * (float)b2[p];
      Df[p] = D_inv_b2;
      int idx = 0;
      for (int r = 0; r < 5; ++r)
        for (int c = r; c < 5; ++c) Ad[idx++] -= (double)(Bf[(size_t)r * Pn + p] * D_inverse * Bf[(size_t)c * Pn + p]);
      for (int r = 0; r < 5; ++r) b1d[r] -= (double)(Bf[(size_t)r * Pn + p] * D_inv_b2);
      for (int r = 0; r < 5; ++r) Bf[(size_t)r * Pn + p] = D_inverse * Bf[(size_t)r * Pn + p];
    }
    double M[25], rhs[5], x1d[5];
    memset(M, 0, sizeof(M));
    int idx = 0;
    for (int r = 0; r < 5; ++r)
      for (int c = r; c < 5; ++c) M[r * 5 + c] = (double)(float)Ad[idx++];
    for (int r = 0; r < 5; ++r) rhs[r] = (double)(float)b1d[r];
    const float kAPriorWeight = 10;
    M[4 * 5 + 4] = (double)((float)M[4 * 5 + 4] + kAPriorWeight * kAPriorWeight);
    rhs[4] = (double)((float)rhs[4] + kAPriorWeight * kAPriorWeight * m->a);
    hm_ldlt_solve(5, M, rhs, x1d);
    float x1[5];
    for (int r = 0; r < 5; ++r) x1[r] = (float)x1d[r];
    float fx_inv = 1.0f / m->depth_K[0], fy_inv = 1.0f / m->depth_K[1];
    float cx_inv = -(m->depth_K[2] - 0.5f) * fx_inv, cy_inv = -(m->depth_K[3] - 0.5f) * fy_inv;
    float new_fx = 1.0f / (fx_inv - x1[0]);
    float new_fy = 1.0f / (fy_inv - x1[1]);
    float new_cx = -(new_fx * (cx_inv - x1[2])) + 0.5f;
    float new_cy = -(new_fy * (cy_inv - x1[3])) + 0.5f;
    m->depth_K[0] = new_fx; m->depth_K[1] = new_fy; m->depth_K[2] = new_cx; m->depth_K[3] = new_cy;
    m->a -= x1[4];
    for (int p = 0; p < Pn; ++p) {
      float offset = Df[p];
      if (isnan(offset)) offset = 0;
      else for (int r = 0; r < 5; ++r) offset -= Bf[(size_t)r * Pn + p] * x1[r];
      float cf = m->cfactor[p] - offset;
      if (obs[p] == 0) cf = 0;
      m->cfactor[p] = cf;
    }
    free(Bf); free(Df);
  }
  if (opt_color) {
    double M[16], rhs[4], x[4];
    memset(M, 0, sizeof(M));
    int idx = 0;
    for (int r = 0; r < 4; ++r)
      for (int c = r; c < 4; ++c) M[r * 4 + c] = (double)(float)cH[idx++];
    for (int r = 0; r < 4; ++r) rhs[r] = (double)(float)cb[r];
    hm_ldlt_solve(4, M, rhs, x);
    for (int r = 0; r < 4; ++r) m->color_K[r] = m->color_K[r] - (float)x[r];
  }
  free(B); free(D); free(b2); free(obs);
}

/* direct_ba.cc:231-249 applied for keyframes added in index order (AddKeyframe :197-205). */
void orc_compute_covisibility(const orc_model* m, orc_keyframes* kfs) {
  int K = kfs->K;
  hm_frustum* fr = (hm_frustum*)malloc(sizeof(hm_frustum) * (size_t)K);
  for (int k = 0; k < K; ++k)
    hm_frustum_create(fr + k, m->depth_K, m->depth_w, m->depth_h, kfs->min_depth[k], kfs->max_depth[k],
                      kfs->global_T_frame + 7 * k);
  memset(kfs->covis, 0, (size_t)K * K);
  for (int nk = 0; nk < K; ++nk)
    for (int k = 0; k < nk; ++k)
      if (hm_frustum_intersects(fr + nk, fr + k)) {
        kfs->covis[(size_t)nk * K + k] = 1;
        kfs->covis[(size_t)k * K + nk] = 1;
      }
  free(fr);
}

/* direct_ba.cc:549-564 */
static void determine_covisible_active(orc_keyframes* kfs) {
  int K = kfs->K;
  for (int k = 0; k < K; ++k) {
    if (kfs->activation[k] != ORC_KF_ACTIVE) continue;
    for (int o = 0; o < K; ++o)
      if (kfs->covis[(size_t)k * K + o] && kfs->activation[o] == ORC_KF_INACTIVE) kfs->activation[o] = ORC_KF_COVIS_ACTIVE;
  }
}

/* direct_ba_alternating.cc:285-738, including the do_surfel_updates branches (creation :399-430, merge + compaction :489-541). */
void orc_bundle_adjust(orc_model* m, orc_keyframes* kfs, float* surfels, int pitch, uint32_t n, uint8_t* active,
                       const orc_ba_options* opt, orc_ba_result* res) {
  memset(res, 0, sizeof(*res));
  const int K = kfs->K;
  const int fixed_window = opt->active_keyframe_window_start > 0 || opt->active_keyframe_window_end > 0;
  const int whole_window = !(opt->active_keyframe_window_start != 0 || opt->active_keyframe_window_end != K - 1);
  memset(active, 0, n);   /* :338 */
  int* with_new = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
  for (int iteration = 0; iteration < opt->max_iterations; ++iteration) {
    res->iterations_done++;
    if (fixed_window) {   /* :354-372 */
      for (int k = 0; k < K; ++k)
        kfs->activation[k] = (k >= opt->active_keyframe_window_start && k <= opt->active_keyframe_window_end)
                                 ? ORC_KF_ACTIVE : ORC_KF_INACTIVE;
      determine_covisible_active(kfs);
    }
    /* surfel creation, :399-430 */
    int n_with_new = 0;
    const uint32_t old_n = n;
    if (opt->optimize_geometry && opt->do_surfel_updates) {
      for (int k = 0; k < K; ++k) {
        if (kfs->activation[k] == ORC_KF_ACTIVE && opt->last_active_in_ba_iteration[k] != opt->ba_iteration_count) {
          opt->last_active_in_ba_iteration[k] = opt->ba_iteration_count;
          with_new[n_with_new++] = k;
        } else if (kfs->activation[k] == ORC_KF_COVIS_ACTIVE && opt->last_covis_in_ba_iteration[k] != opt->ba_iteration_count) {
          opt->last_covis_in_ba_iteration[k] = opt->ba_iteration_count;
        }
      }
      for (int q = 0; q < n_with_new; ++q)
        res->surfels_created += orc_create_surfels_for_keyframe(m, kfs, with_new[q], 1, opt->min_observation_count, surfels, pitch, &n,
                                                                opt->max_surfels);
    }
    /* :432-456: new surfels are active, the old ones are re-evaluated */
    if (opt->optimize_geometry && n > old_n) memset(active + old_n, K_SURFEL_ACTIVE_FLAG, n - old_n);
    if (!whole_window) memset(active, K_SURFEL_ACTIVE_FLAG, old_n);
    else orc_update_activation(m, kfs, surfels, pitch, old_n, active);
    /* :466-485 */
    if (opt->optimize_geometry) orc_optimize_geometry_iteration(m, kfs, surfels, pitch, n, active);
    /* surfel merge + compaction, :489-541 */
    if (opt->do_surfel_updates && n_with_new > 0) {
      uint32_t merged = 0;
      for (int q = 0; q < n_with_new; ++q)
        merged += orc_merge_surfels_for_keyframe(m, kfs, with_new[q], opt->surfel_merge_dist_factor, surfels, pitch, n);
      res->surfels_merged += merged;
      n = orc_compact_surfels(surfels, pitch, n, active);
    }
    /* :543-577 */
    int num_converged = 0;
    if (opt->optimize_poses) {
      res->n_assoc = res->n_photo = 0;
      res->cost = 0;
      for (int k = 0; k < K; ++k) {
        if (kfs->activation[k] == ORC_KF_INACTIVE) { ++num_converged; continue; }
        float* pose = kfs->global_T_frame + 7 * k;
        {  /* bookkeeping for the metric: counts at the pose step's starting state */
          float T[12];
          orc_pose_stats st;
          orc_frame_T_global(pose, T);
          orc_pose_coeffs(m, kfs, k, T, surfels, pitch, n, &st);
          res->n_assoc += st.n_assoc;
          res->n_photo += st.n_photo;
          res->cost += st.cost_depth + st.cost_desc1;
        }
        float est[7], ftg[7], diff[7], lg[6];
        int conv;
        res->pose_iterations_total += orc_estimate_frame_pose(m, kfs, k, pose, surfels, pitch, n, est, &conv,
                                                              opt->max_pose_iterations > 0 ? opt->max_pose_iterations : 30);
        hm_se3_inverse(pose, ftg);
        hm_se3_mul(ftg, est, diff);
        hm_se3_log(diff, lg);
        int moved = !hm_is_scale1_pose_converged(lg);
        memcpy(pose, est, sizeof(est));
        if (moved) kfs->activation[k] = ORC_KF_ACTIVE;
        else { kfs->activation[k] = ORC_KF_INACTIVE; ++num_converged; }
      }
    }
    /* :580-626 */
    if (opt->optimize_depth_intrinsics || opt->optimize_color_intrinsics)
      orc_optimize_intrinsics(m, kfs, surfels, pitch, n,
                              opt->optimize_depth_intrinsics && m->use_depth_residuals,
                              opt->optimize_color_intrinsics && m->use_descriptor_residuals);
    /* :693-701 */
    if (iteration >= opt->min_iterations - 1 && (num_converged == K || !opt->optimize_poses)) {
      res->converged = 1;
      break;
    }
    determine_covisible_active(kfs);   /* :711-717 */
  }
  free(with_new);
  res->surfels_size = n;
}


/* ------------------------------------------------------------------------------------------------------------------
 * PCG-based Gauss-Newton (DirectBA::BundleAdjustmentPCG, direct_ba_pcg.cc:43-819; kernels kernel_pcg.cu:179-1372).
 * Vectors are fp32 (PCGScalar = float, kernels.cuh:62); the sums that the device forms with block reductions + float
 * atomics (arbitrary order) are accumulated in fp64 here and rounded to fp32 once.
 * ------------------------------------------------------------------------------------------------------------------ */
#define PCG_DIAG_EPSILON 1e-8f   /* kernel_pcg.cu:44 */
#define PCG_A_PRIOR 10.f         /* kernel_pcg.cu:48 */

typedef struct {
  int opt_poses, opt_geometry, opt_di, opt_ci, use_depth, use_desc;
  int gauge;
  uint32_t surfel_start, stride, depth_start, color_start;
} pcg_layout;

/* One (surfel, keyframe) pair: the Jacobian rows of its up to three residuals (kernel_pcg.cu:204-505 / 670-1036). */
typedef struct {
  int depth_valid;      /* depth residual present */
  float raw, w, jg, jp[6];
  int di_valid; float jd[5], jcf; uint32_t cf_u;
  int desc_valid;       /* descriptor residuals present */
  float r1, r2, w1, w2, jg1, jg2, jp1[6], jp2[6], jc1[4], jc2[4];
} pcg_rows;

static void pcg_eval(const kfview* v, const float T[12], const float* surfels, size_t P, uint32_t i, const pcg_layout* L,
                     int init, pcg_rows* o) {
  memset(o, 0, sizeof(*o));
  assoc r;
  f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
  if (project_associate(v, T, gp, f2u(surfels[ROW_NORMAL * P + i]), &r) != 3) return;
  int visible = 1;
  f3 rn = T_rot(T, r.n);
  float nx = v->fx_inv * r.px + v->cx_inv, ny = v->fy_inv * r.py + v->cy_inv;
  if (L->use_depth) {
    float inv_stddev;
    f3 up;
    o->raw = depth_pose_residual_jacobian(v, &r, rn, o->jp, &inv_stddev, &up);
    o->w = depth_weight(o->raw);
    o->jg = -inv_stddev;
    o->depth_valid = 1;
    if (L->opt_di) {
      int spx = r.px / v->cell, spy = r.py / v->cell;
      float cfactor = v->cfactor[(size_t)spy * v->cf_w + spx];
      float raw_inv_depth = 1.0f / (v->raw_to_float * v->depth[(size_t)r.py * v->w + r.px]);
      float exp_inv_depth = expf(-v->a * raw_inv_depth);
      float corrected_inv_depth = cfactor * exp_inv_depth + raw_inv_depth;
      o->di_valid = !(fabsf(corrected_inv_depth) < 1e-4f);
      if (init && !o->di_valid) visible = 0;   /* kernel_pcg.cu:266-268 (only the init kernel clears `visible`) */
      float dot = nx * rn.x + ny * rn.y + rn.z;
      float jac_base = inv_stddev * dot * exp_inv_depth / (corrected_inv_depth * corrected_inv_depth);
      o->jd[2] = inv_stddev * r.d * (r.n.x * T[0] + r.n.y * T[1] + r.n.z * T[2]);
      o->jd[3] = inv_stddev * r.d * (r.n.x * T[4] + r.n.y * T[5] + r.n.z * T[6]);
      o->jd[0] = r.px * o->jd[2];
      o->jd[1] = r.py * o->jd[3];
      o->jd[4] = cfactor * raw_inv_depth * jac_base;
      o->jcf = -jac_base;
      o->cf_u = L->depth_start + 5u + (uint32_t)spx + (uint32_t)spy * (uint32_t)v->cf_w;
    }
  }
  if (L->use_desc && visible) {
    float ccx, ccy;
    if (!depth_to_color(v, r.pxf, r.pyf, &ccx, &ccy)) return;
    float t1x, t1y, t2x, t2y;
    tangent_projections(v, T, r.gp, r.n, surfels[ROW_R2 * P + i], &t1x, &t1y, &t2x, &t2y);
    desc_eval e;
    descriptor_eval(v, ccx, ccy, t1x, t1y, t2x, t2y, surfels[ROW_D1 * P + i], surfels[ROW_D2 * P + i], &e);
    o->desc_valid = 1;
    o->r1 = e.r1; o->r2 = e.r2;
    o->w1 = desc_weight(e.r1); o->w2 = desc_weight(e.r2);
    float gx1 = e.gx1 * v->cfx, gy1 = e.gy1 * v->cfy, gx2 = e.gx2 * v->cfx, gy2 = e.gy2 * v->cfy;
    float term1 = -(rn.x * r.lp.z - rn.z * r.lp.x);
    float term2 = -(rn.y * r.lp.z - rn.z * r.lp.y);
    float term3 = 1.f / (r.lp.z * r.lp.z);
    o->jg1 = -(gx1 * term1 + gy1 * term2) * term3;
    o->jg2 = -(gx2 * term1 + gy2 * term2) * term3;
    desc_pose_jacobian(v, r.lp, e.gx1, e.gy1, o->jp1);
    desc_pose_jacobian(v, r.lp, e.gx2, e.gy2, o->jp2);
    float g1x = gx1 / v->cfx, g1y = gy1 / v->cfy, g2x = gx2 / v->cfx, g2y = gy2 / v->cfy;   /* kernel_pcg.cu:454-457 */
    o->jc1[0] = g1x * nx; o->jc1[1] = g1y * ny; o->jc1[2] = g1x; o->jc1[3] = g1y;
    o->jc2[0] = g2x * nx; o->jc2[1] = g2y * ny; o->jc2[2] = g2x; o->jc2[3] = g2y;
  }
}

static inline uint32_t pcg_pose_index(int k, int gauge) { return (uint32_t)(6 * (k < gauge ? k : k - 1)); }

/* init = 1: out0 = r (-= J^T W F), out1 = M (+= diag J^T W J).  init = 0: out0 = g (+= J^T W J p), returns p^T J^T W J p. */
static double pcg_pass(const orc_model* m, const orc_keyframes* kfs, const float* surfels, size_t P, uint32_t n,
                       const pcg_layout* L, int init, const float* p, double* out0, double* out1) {
  const int K = kfs->K;
  double alpha_d = 0;
  for (int k = 0; k < K; ++k) {
    kfview v;
    make_view(m, kfs, k, &v);
    float T[12];
    orc_frame_T_global(kfs->global_T_frame + 7 * k, T);
    const int do_pose = L->opt_poses && k != L->gauge;
    const uint32_t pu = pcg_pose_index(k, L->gauge);
    double pose0[6] = {0}, pose1[6] = {0}, di0[5] = {0}, di1[5] = {0}, ci0[4] = {0}, ci1[4] = {0};
    double ad = 0;
#pragma omp parallel
    {
      double t_pose0[6] = {0}, t_pose1[6] = {0}, t_di0[5] = {0}, t_di1[5] = {0}, t_ci0[4] = {0}, t_ci1[4] = {0}, t_ad = 0;
#pragma omp for schedule(static)
      for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        const uint32_t i = (uint32_t)ii;
        pcg_rows o;
        pcg_eval(&v, T, surfels, P, i, L, init, &o);
        const uint32_t su = L->surfel_start + L->stride * i;
        if (o.depth_valid) {
          if (init) {
            const float wr = o.w * o.raw;
            if (L->opt_geometry) { out0[su] -= (double)(o.jg * wr); out1[su] += (double)(o.jg * o.w * o.jg); }
            if (do_pose) for (int c = 0; c < 6; ++c) { t_pose0[c] -= (double)(o.jp[c] * wr); t_pose1[c] += (double)(o.jp[c] * o.w * o.jp[c]); }
            if (L->opt_di && o.di_valid) {
              for (int c = 0; c < 5; ++c) { t_di0[c] -= (double)(o.jd[c] * wr); t_di1[c] += (double)(o.jd[c] * o.w * o.jd[c]); }
#pragma omp atomic
              out0[o.cf_u] -= (double)(o.jcf * wr);
#pragma omp atomic
              out1[o.cf_u] += (double)(o.jcf * o.w * o.jcf);
            }
          } else {
            float sum = 0;
            if (L->opt_geometry) sum += o.jg * p[su];
            if (do_pose) for (int c = 0; c < 6; ++c) sum += o.jp[c] * p[pu + c];
            if (L->opt_di && o.di_valid) {
              for (int c = 0; c < 5; ++c) sum += o.jd[c] * p[L->depth_start + c];
              sum += o.jcf * p[o.cf_u];
            }
            t_ad += (double)(sum * o.w * sum);
            sum *= o.w;
            if (L->opt_geometry) out0[su] += (double)(o.jg * sum);
            if (do_pose) for (int c = 0; c < 6; ++c) t_pose0[c] += (double)(o.jp[c] * sum);
            if (L->opt_di && o.di_valid) {
              for (int c = 0; c < 5; ++c) t_di0[c] += (double)(o.jd[c] * sum);
#pragma omp atomic
              out0[o.cf_u] += (double)(o.jcf * sum);
            }
          }
        }
        if (o.desc_valid) {
          if (init) {
            const float wr1 = o.w1 * o.r1, wr2 = o.w2 * o.r2;
            if (L->opt_geometry) {
              out0[su] -= (double)(o.jg1 * wr1 + o.jg2 * wr2);
              out1[su] += (double)(o.jg1 * o.w1 * o.jg1 + o.jg2 * o.w2 * o.jg2);
              out0[su + 1] += (double)wr1; out1[su + 1] += (double)o.w1;
              out0[su + 2] += (double)wr2; out1[su + 2] += (double)o.w2;
            }
            if (do_pose) for (int c = 0; c < 6; ++c) {
              t_pose0[c] -= (double)(o.jp1[c] * wr1 + o.jp2[c] * wr2);
              t_pose1[c] += (double)(o.jp1[c] * o.w1 * o.jp1[c] + o.jp2[c] * o.w2 * o.jp2[c]);
            }
            if (L->opt_ci) for (int c = 0; c < 4; ++c) {
              t_ci0[c] -= (double)(o.jc1[c] * wr1 + o.jc2[c] * wr2);
              t_ci1[c] += (double)(o.jc1[c] * o.w1 * o.jc1[c] + o.jc2[c] * o.w2 * o.jc2[c]);
            }
          } else {
            float s1 = 0, s2 = 0;
            if (L->opt_geometry) { s1 += o.jg1 * p[su] - p[su + 1]; s2 += o.jg2 * p[su] - p[su + 2]; }
            if (do_pose) for (int c = 0; c < 6; ++c) { s1 += o.jp1[c] * p[pu + c]; s2 += o.jp2[c] * p[pu + c]; }
            if (L->opt_ci) for (int c = 0; c < 4; ++c) { s1 += o.jc1[c] * p[L->color_start + c]; s2 += o.jc2[c] * p[L->color_start + c]; }
            t_ad += (double)(s1 * o.w1 * s1 + s2 * o.w2 * s2);
            s1 *= o.w1; s2 *= o.w2;
            if (L->opt_geometry) {
              out0[su] += (double)(o.jg1 * s1 + o.jg2 * s2);
              out0[su + 1] -= (double)s1;
              out0[su + 2] -= (double)s2;
            }
            if (do_pose) for (int c = 0; c < 6; ++c) t_pose0[c] += (double)(o.jp1[c] * s1 + o.jp2[c] * s2);
            if (L->opt_ci) for (int c = 0; c < 4; ++c) t_ci0[c] += (double)(o.jc1[c] * s1 + o.jc2[c] * s2);
          }
        }
      }
#pragma omp critical
      {
        for (int c = 0; c < 6; ++c) { pose0[c] += t_pose0[c]; pose1[c] += t_pose1[c]; }
        for (int c = 0; c < 5; ++c) { di0[c] += t_di0[c]; di1[c] += t_di1[c]; }
        for (int c = 0; c < 4; ++c) { ci0[c] += t_ci0[c]; ci1[c] += t_ci1[c]; }
        ad += t_ad;
      }
    }
    if (do_pose) for (int c = 0; c < 6; ++c) { out0[pu + c] += pose0[c]; if (init) out1[pu + c] += pose1[c]; }
    if (L->opt_di) for (int c = 0; c < 5; ++c) { out0[L->depth_start + c] += di0[c]; if (init) out1[L->depth_start + c] += di1[c]; }
    if (L->opt_ci) for (int c = 0; c < 4; ++c) { out0[L->color_start + c] += ci0[c]; if (init) out1[L->color_start + c] += ci1[c]; }
    alpha_d += ad;
  }
  return alpha_d;
}

static inline float pcg_diag_extra(uint32_t i, uint32_t a_index) {
  return PCG_DIAG_EPSILON + ((i == a_index) ? (PCG_A_PRIOR * PCG_A_PRIOR) : 0.f);
}

void orc_bundle_adjust_pcg(orc_model* m, orc_keyframes* kfs, float* surfels, int pitch, uint32_t n, uint8_t* active,
                           const orc_pcg_options* opt, orc_pcg_result* res) {
  memset(res, 0, sizeof(*res));
  const int K = kfs->K;
  const size_t P = (size_t)pitch;
  const uint32_t Pn = (uint32_t)(m->cf_w * m->cf_h);
  const uint32_t kInvalid = 0xffffffffu;
  const int max_inner = opt->max_inner_iterations > 0 ? opt->max_inner_iterations : 30;
  int* with_new = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
  int n_with_new = 0;
  for (int iteration = 0; iteration < opt->max_iterations; ++iteration) {
    ++res->iterations_done;
    /* surfel creation, direct_ba_pcg.cc:180-206 */
    n_with_new = 0;
    if (opt->optimize_geometry && opt->do_surfel_updates) {
      for (int k = 0; k < K; ++k) {
        if (kfs->activation[k] == ORC_KF_ACTIVE && opt->last_active_in_ba_iteration[k] != opt->ba_iteration_count) {
          opt->last_active_in_ba_iteration[k] = opt->ba_iteration_count;
          res->surfels_created += orc_create_surfels_for_keyframe(m, kfs, k, 1, opt->min_observation_count, surfels, pitch, &n,
                                                                  opt->max_surfels);
          with_new[n_with_new++] = k;
        } else if (kfs->activation[k] == ORC_KF_COVIS_ACTIVE && opt->last_covis_in_ba_iteration[k] != opt->ba_iteration_count) {
          opt->last_covis_in_ba_iteration[k] = opt->ba_iteration_count;
        }
      }
    }
    memset(active, K_SURFEL_ACTIVE_FLAG, n);   /* direct_ba_pcg.cc:209-212 */
    if (opt->optimize_geometry && n > 0) {     /* :215-227 */
      float* Ts = (float*)malloc(sizeof(float) * 12 * (size_t)K);
      kfview* vs = (kfview*)malloc(sizeof(kfview) * (size_t)K);
      for (int k = 0; k < K; ++k) { orc_frame_T_global(kfs->global_T_frame + 7 * k, Ts + 12 * k); make_view(m, kfs, k, vs + k); }
      update_normals(kfs, Ts, vs, surfels, P, n, active);
      free(Ts); free(vs);
    }
    pcg_layout L;
    L.use_depth = m->use_depth_residuals; L.use_desc = m->use_descriptor_residuals;
    L.opt_poses = opt->optimize_poses; L.opt_geometry = opt->optimize_geometry;
    L.opt_di = opt->optimize_depth_intrinsics && L.use_depth;      /* direct_ba.cc:427-434 */
    L.opt_ci = opt->optimize_color_intrinsics && L.use_desc;
    L.gauge = opt->gauge_keyframe;
    L.stride = L.use_desc ? 3u : 1u;
    uint32_t cur = 0, a_index = kInvalid;
    if (L.opt_poses) cur += 6u * (uint32_t)(K - 1);
    L.surfel_start = L.depth_start = L.color_start = kInvalid;
    if (L.opt_geometry) { L.surfel_start = cur; cur += L.stride * n; }
    if (L.opt_di) { L.depth_start = cur; cur += 5u + Pn; a_index = L.depth_start + 4u; }
    if (L.opt_ci) { L.color_start = cur; cur += 4u; }
    const uint32_t U = cur;
    int num_converged = L.opt_poses ? 1 : 0;
    if (U > 0) {
      double* acc0 = (double*)calloc(U, sizeof(double));
      double* acc1 = (double*)calloc(U, sizeof(double));
      float* r = (float*)malloc(sizeof(float) * U);
      float* M = (float*)malloc(sizeof(float) * U);
      float* delta = (float*)calloc(U, sizeof(float));
      float* g = (float*)calloc(U, sizeof(float));
      float* p = (float*)malloc(sizeof(float) * U);
      pcg_pass(m, kfs, surfels, P, n, &L, 1, NULL, acc0, acc1);
      double alpha_n = 0, beta_n = 0;
      for (uint32_t i = 0; i < U; ++i) {   /* PCGInit2CUDAKernel, kernel_pcg.cu:569-605 */
        r[i] = (float)acc0[i]; M[i] = (float)acc1[i];
        float r_value = r[i] + ((i == a_index) ? (-PCG_A_PRIOR * PCG_A_PRIOR * m->a) : 0.f);
        float p_value = r_value / (M[i] + pcg_diag_extra(i, a_index));
        p[i] = p_value;
        alpha_n += (double)(r_value * p_value);
      }
      float prev_r_norm = INFINITY;
      int without_improvement = 0;
      for (int step = 0; step < max_inner; ++step) {
        if (step > 0) alpha_n = beta_n;   /* :386 */
        memset(acc0, 0, sizeof(double) * U);
        double alpha_d = pcg_pass(m, kfs, surfels, P, n, &L, 0, p, acc0, NULL);
        /* AddAlphaDEpsilonTermsCUDAKernel runs after EVERY per-keyframe PCGStep1 launch (kernel_pcg.cu:1101-1112) */
        double eps = 0;
        for (uint32_t i = 0; i < U; ++i) eps += (double)(pcg_diag_extra(i, a_index) * p[i] * p[i]);
        alpha_d += eps * K;
        /* PCGStep2CUDAKernel, kernel_pcg.cu:1115-1166 */
        const float alpha = ((float)alpha_d >= 1e-35f) ? ((float)alpha_n / (float)alpha_d) : 0.f;
        beta_n = 0;
        for (uint32_t i = 0; i < U; ++i) {
          g[i] = (float)acc0[i];
          delta[i] += alpha * p[i];
          float r_value = r[i] - alpha * (g[i] + pcg_diag_extra(i, a_index) * p[i]);
          r[i] = r_value;
          float z = r_value / (M[i] + pcg_diag_extra(i, a_index));
          g[i] = z;
          beta_n += (double)(z * r_value);
        }
        ++res->inner_iterations_total;
        float r_norm = sqrtf((float)beta_n);
        res->last_r_norm = r_norm;
        if ((double)r_norm < (double)prev_r_norm - 1e-3) without_improvement = 0;
        else if (++without_improvement >= 3) break;
        prev_r_norm = r_norm;
        if (step < max_inner - 1) {   /* PCGStep3CUDAKernel, kernel_pcg.cu:1206-1224 */
          const float beta = ((float)alpha_n >= 1e-35f) ? ((float)beta_n / (float)alpha_n) : 0.f;
          for (uint32_t i = 0; i < U; ++i) p[i] = g[i] + beta * p[i];
        }
      }
      /* apply delta (direct_ba_pcg.cc:552-638) */
      if (L.opt_poses) {
        for (int k = 0; k < K; ++k) {
          if (k == L.gauge) continue;
          float d7[7], np[7], lg[6];
          hm_se3_exp(delta + pcg_pose_index(k, L.gauge), d7);
          hm_se3_mul(kfs->global_T_frame + 7 * k, d7, np);
          memcpy(kfs->global_T_frame + 7 * k, np, sizeof(np));
          hm_se3_log(d7, lg);
          if (hm_is_scale1_pose_converged(lg)) ++num_converged;
        }
      }
      if (L.opt_geometry) {   /* kernel_pcg.cu:1278-1308 */
        for (uint32_t i = 0; i < n; ++i) {
          const uint32_t su = L.surfel_start + L.stride * i;
          float t = delta[su];
          if (t != 0) {
            f3 nrm = unpack_normal(f2u(surfels[ROW_NORMAL * P + i]));
            surfels[ROW_X * P + i] += t * nrm.x;
            surfels[ROW_Y * P + i] += t * nrm.y;
            surfels[ROW_Z * P + i] += t * nrm.z;
          }
          if (L.use_desc) {
            surfels[ROW_D1 * P + i] = fmaxf(-180.f, fminf(180.f, surfels[ROW_D1 * P + i] + delta[su + 1]));
            surfels[ROW_D2 * P + i] = fmaxf(-180.f, fminf(180.f, surfels[ROW_D2 * P + i] + delta[su + 2]));
          }
        }
      }
      if (L.opt_di) {
        const float* b = delta + L.depth_start;
        double old_fx_inv = 1. / m->depth_K[0], old_fy_inv = 1. / m->depth_K[1];
        double old_cx_inv = -(m->depth_K[2] - 0.5) * old_fx_inv, old_cy_inv = -(m->depth_K[3] - 0.5) * old_fy_inv;
        double nfx = 1. / (old_fx_inv + b[0]), nfy = 1. / (old_fy_inv + b[1]);
        double ncx = -(nfx * (old_cx_inv + b[2])) + 0.5, ncy = -(nfy * (old_cy_inv + b[3])) + 0.5;
        m->depth_K[0] = (float)nfx; m->depth_K[1] = (float)nfy; m->depth_K[2] = (float)ncx; m->depth_K[3] = (float)ncy;
        m->a += b[4];
        for (uint32_t c = 0; c < Pn; ++c) m->cfactor[c] += b[5 + c];
      }
      if (L.opt_ci) for (int c = 0; c < 4; ++c) m->color_K[c] = (float)(m->color_K[c] + delta[L.color_start + c]);
      free(acc0); free(acc1); free(r); free(M); free(delta); free(g); free(p);
    }
    /* surfel merge + compaction for the keyframes that received new surfels, direct_ba_pcg.cc:644-690 */
    if (opt->do_surfel_updates && n_with_new > 0) {
      for (int q = 0; q < n_with_new; ++q)
        res->surfels_merged += orc_merge_surfels_for_keyframe(m, kfs, with_new[q], opt->surfel_merge_dist_factor, surfels, pitch, n);
      n = orc_compact_surfels(surfels, pitch, n, active);
    }
    if (iteration >= opt->min_iterations - 1 && (num_converged == K || !L.opt_poses)) {
      res->converged = 1;
      break;
    }
  }
  /* direct_ba_pcg.cc:775-815: without the end tasks, the keyframes of the LAST iteration's creation step are merged once more */
  if (!opt->increase_ba_iteration_count && opt->do_surfel_updates && n_with_new > 0) {
    for (int q = 0; q < n_with_new; ++q)
      res->surfels_merged += orc_merge_surfels_for_keyframe(m, kfs, with_new[q], opt->surfel_merge_dist_factor, surfels, pitch, n);
    n = orc_compact_surfels(surfels, pitch, n, active);
  }
  free(with_new);
  res->surfels_size = n;
}

/* Parity hook: r, M after the init pass; p = M^-1 r (with the prior on a); g after one J^T W J p sweep;
 * scalars = {alpha_n, alpha_d}.  Returns the unknown count (call with out_r = NULL to size the buffers). */
uint32_t orc_pcg_debug(const orc_model* m, const orc_keyframes* kfs, const float* surfels, int pitch, uint32_t n,
                       const orc_pcg_options* opt, float* out_r, float* out_M, float* out_p, float* out_g, double* out_scalars) {
  const int K = kfs->K;
  const size_t P = (size_t)pitch;
  const uint32_t Pn = (uint32_t)(m->cf_w * m->cf_h), kInvalid = 0xffffffffu;
  pcg_layout L;
  L.use_depth = m->use_depth_residuals; L.use_desc = m->use_descriptor_residuals;
  L.opt_poses = opt->optimize_poses; L.opt_geometry = opt->optimize_geometry;
  L.opt_di = opt->optimize_depth_intrinsics && L.use_depth;
  L.opt_ci = opt->optimize_color_intrinsics && L.use_desc;
  L.gauge = opt->gauge_keyframe;
  L.stride = L.use_desc ? 3u : 1u;
  uint32_t cur = 0, a_index = kInvalid;
  if (L.opt_poses) cur += 6u * (uint32_t)(K - 1);
  L.surfel_start = L.depth_start = L.color_start = kInvalid;
  if (L.opt_geometry) { L.surfel_start = cur; cur += L.stride * n; }
  if (L.opt_di) { L.depth_start = cur; cur += 5u + Pn; a_index = L.depth_start + 4u; }
  if (L.opt_ci) { L.color_start = cur; cur += 4u; }
  const uint32_t U = cur;
  if (!out_r) return U;
  double* acc0 = (double*)calloc(U, sizeof(double));
  double* acc1 = (double*)calloc(U, sizeof(double));
  pcg_pass(m, kfs, surfels, P, n, &L, 1, NULL, acc0, acc1);
  double alpha_n = 0;
  for (uint32_t i = 0; i < U; ++i) {
    out_r[i] = (float)acc0[i]; out_M[i] = (float)acc1[i];
    float r_value = out_r[i] + ((i == a_index) ? (-PCG_A_PRIOR * PCG_A_PRIOR * m->a) : 0.f);
    out_p[i] = r_value / (out_M[i] + pcg_diag_extra(i, a_index));
    alpha_n += (double)(r_value * out_p[i]);
  }
  memset(acc0, 0, sizeof(double) * U);
  double alpha_d = pcg_pass(m, kfs, surfels, P, n, &L, 0, out_p, acc0, NULL);
  double eps = 0;
  for (uint32_t i = 0; i < U; ++i) { out_g[i] = (float)acc0[i]; eps += (double)(pcg_diag_extra(i, a_index) * out_p[i] * out_p[i]); }
  out_scalars[0] = alpha_n;
  out_scalars[1] = alpha_d + eps * K;
  free(acc0); free(acc1);
  return U;
}


/* ------------------------------------------------------------------------------------------------------------------
 * DirectBA::PerformBASchemeEndTasks (direct_ba.cc:566-653) without the final merge:
 * DeleteSurfelsAndUpdateRadiiCUDA (kernel_delete_surfels.cc:40-98, .cu:42-164) + CompactSurfelsCUDA
 * (kernel_compact_surfels.cu:159-279).  Returns the number of deleted surfels; *n becomes the new surfels_size.
 * ------------------------------------------------------------------------------------------------------------------ */
static inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31u, man = h & 1023u;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {   /* subnormal */
      int e = -1;
      do { ++e; man <<= 1; } while (!(man & 1024u));
      bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 1023u) << 13;
    }
  } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
  else bits = sign | (exp + 112u) << 23 | man << 13;
  return u2f(bits);
}

uint32_t orc_end_tasks_with_merge(const orc_model* m, const orc_keyframes* kfs, float* surfels, int pitch, uint32_t* n_inout,
                                  int min_observation_count, const int32_t* last_active_in_ba_iteration, int ba_iteration_count,
                                  float surfel_merge_dist_factor) {
  /* direct_ba.cc:577-601: merge with every keyframe that was active in this BA iteration block, then the usual end tasks */
  uint32_t merged = 0;
  for (int k = 0; k < kfs->K; ++k)
    if (last_active_in_ba_iteration[k] == ba_iteration_count)
      merged += orc_merge_surfels_for_keyframe(m, kfs, k, surfel_merge_dist_factor, surfels, pitch, *n_inout);
  /* orc_end_tasks compacts every marked surfel (its own deletions and the merged ones) */
  return merged + orc_end_tasks(m, kfs, surfels, pitch, n_inout, min_observation_count);
}

uint32_t orc_end_tasks(const orc_model* m, const orc_keyframes* kfs, float* surfels, int pitch, uint32_t* n_inout,
                       int min_observation_count) {
  const uint32_t n = *n_inout;
  if (n == 0) return 0;
  const int K = kfs->K;
  const size_t P = (size_t)pitch;
  const size_t npx = (size_t)m->depth_w * m->depth_h;
  float* Ts = (float*)malloc(sizeof(float) * 12 * (size_t)(K > 0 ? K : 1));
  kfview* vs = (kfview*)malloc(sizeof(kfview) * (size_t)(K > 0 ? K : 1));
  for (int k = 0; k < K; ++k) { orc_frame_T_global(kfs->global_T_frame + 7 * k, Ts + 12 * k); make_view(m, kfs, k, vs + k); }
  uint32_t deleted = 0;
#pragma omp parallel for schedule(static) reduction(+ : deleted)
  for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
    const size_t i = (size_t)ii;
    float obs = 0, viol = 0, min_r2 = INFINITY;
    const float x = surfels[ROW_X * P + i];
    f3 gp = mk3(x, surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
    f3 nrm = unpack_normal(f2u(surfels[ROW_NORMAL * P + i]));
    for (int k = 0; k < K; ++k) {   /* every keyframe (kernel_delete_surfels.cc:69-80) */
      const kfview* v = vs + k;
      const float* T = Ts + 12 * k;
      /* SurfelProjectsToAssociatedPixel(..., SurfelProjectionResultXYFreeSpace*), surfel_projection_nvcc_only.cuh:482-511 */
      f3 lp;
      lp.z = T[8] * gp.x + T[9] * gp.y + T[10] * gp.z + T[11];
      if (!(lp.z > 0.f)) continue;
      lp.x = T[0] * gp.x + T[1] * gp.y + T[2] * gp.z + T[3];
      lp.y = T[4] * gp.x + T[5] * gp.y + T[6] * gp.z + T[7];
      float pxf = v->fx * (lp.x / lp.z) + v->cx, pyf = v->fy * (lp.y / lp.z) + v->cy;
      if (!(pxf >= 0.f) || !(pyf >= 0.f) || !(pxf < 1e9f) || !(pyf < 1e9f)) continue;
      int px = (int)pxf, py = (int)pyf;
      if (px >= v->w || py >= v->h) continue;
      uint16_t measured = v->depth[(size_t)py * v->w + px];
      if (measured & K_INVALID_DEPTH_BIT) continue;
      float d = raw_to_calibrated_depth(v->a, v->cfactor[(size_t)(py / v->cell) * v->cf_w + (px / v->cell)], v->raw_to_float, measured);
      f3 ln = T_rot(T, nrm);
      float nx = v->fx_inv * px + v->cx_inv, ny = v->fy_inv * py + v->cy_inv;
      float thr = K_DEPTH_TUKEY * ((K_DEPTH_UNCERTAINTY_FACTOR * fabsf(ln.x * nx + ln.y * ny + ln.z) * (d * d)) / v->baseline_fx);
      float diff = d - lp.z;
      if (diff > thr) { viol += 1.f; continue; }
      if (diff < -thr) continue;
      float dist = sqrtf(dot3(lp, lp));
      if ((1.0f / dist) * dot3(lp, ln) > 0) continue;
      if (dot3(ln, u16_to_image_space_normal(v->normals[(size_t)py * v->w + px])) < K_COS_NORMAL_COMPAT) continue;
      obs += 1.f;
      min_r2 = fminf(min_r2, half_to_float(kfs->radius[npx * k + (size_t)py * v->w + px]));
    }
    surfels[(ROW_ACC0 + 0) * P + i] = obs; surfels[(ROW_ACC0 + 1) * P + i] = viol; surfels[(ROW_ACC0 + 2) * P + i] = min_r2;
    if (obs < (float)min_observation_count || viol > obs) {   /* kernel_delete_surfels.cu:138-146 */
      if (f2u(x) != 0x7fffffffu) { surfels[ROW_X * P + i] = u2f(0x7fffffffu); ++deleted; }
    } else {
      surfels[ROW_R2 * P + i] = min_r2;
    }
  }
  free(Ts); free(vs);
  /* compaction of every marked surfel (also those a preceding merge marked): the r-th valid surfel from the end moves into the r-th free spot from the front if that lies in front of it */
  uint8_t* invalid = (uint8_t*)malloc(n);   /* validity BEFORE any move (the device kernel flags first, kernel_compact_surfels.cu:98-107) */
  uint32_t free_count = 0;
  for (uint32_t i = 0; i < n; ++i) { invalid[i] = f2u(surfels[ROW_X * P + i]) == 0x7fffffffu; free_count += invalid[i]; }
  uint32_t* free_list = (uint32_t*)malloc(sizeof(uint32_t) * (free_count ? free_count : 1));
  uint32_t nf = 0;
  for (uint32_t i = 0; i < n; ++i) if (invalid[i]) free_list[nf++] = i;
  uint32_t r = 0;
  for (uint32_t i = n; i-- > 0 && r < free_count;) {
    if (invalid[i]) continue;
    if (free_list[r] < i)
      for (int row = 0; row < ROW_ACC0; ++row) surfels[(size_t)row * P + free_list[r]] = surfels[(size_t)row * P + i];
    ++r;
  }
  free(invalid);
  free(free_list);
  *n_inout = n - free_count;
  return deleted;
}

/* ------------------------------------------------------------------------------------------------------------------
 * In-loop surfel lifecycle: DirectBA::CreateSurfelsForKeyframe (direct_ba.cc:340-405, kernel_create_surfels.cc:40-183,
 * kernel_create_surfels.cu:40-405) and DetermineSupportingSurfelsAndMergeSurfelsCUDA (kernel_supporting_surfels.cc:40-118,
 * kernel_supporting_surfels.cu:44-101).
 *
 * The reference resolves two races with atomicCAS "first come": which pixel of an unoccupied sparse cell seeds the new
 * surfel (kernel_create_surfels.cu:57-68) and which surfels become the supporting surfels of a cell
 * (kernel_supporting_surfels.cu:60-62).  This restatement (and the CUDA path it checks) uses the outcome of executing the
 * reference's threads in a FIXED order: the seed is the valid pixel with the smallest raster index of the cell; in the merge
 * the surfels arrive in the order of a bijective hash of their index (a pseudo-random order like the hardware's; plain index
 * order would always favour the oldest surfels).  With sparse_surfel_cell_size = 1 the creation is identical to the
 * reference's; otherwise the two agree in distribution (same cells seeded, same number of new surfels; a similar number
 * of merges between the same kind of neighbours).
 * ------------------------------------------------------------------------------------------------------------------ */
#define K_MERGE_BUFFER_COUNT 3   /* kernels.cuh:51 */

/* tex2D<float4>(color_texture, x, y) channel c with the measured B200 filter (see tex_w_hw) */
static inline float tex_channel_hw(const kfview* v, float x, float y, int c) {
  float xb = x - 0.5f, yb = y - 0.5f;
  float fi = floorf(xb), fj = floorf(yb);
  long i = (long)fi, j = (long)fj;
  long a = (long)floorf((xb - fi) * 256.f + 0.5f), b = (long)floorf((yb - fj) * 256.f + 0.5f);
  long w11 = (a * b + 128) >> 8, w10 = a - w11, w01 = b - w11, w00 = 256 - w11 - w10 - w01;
  long t[4];
  for (int q = 0; q < 4; ++q) {
    long ii = i + (q & 1), jj = j + (q >> 1);
    if (ii < 0) ii = 0;
    if (jj < 0) jj = 0;
    if (ii > v->cw - 1) ii = v->cw - 1;
    if (jj > v->ch - 1) jj = v->ch - 1;
    t[q] = v->color[((size_t)jj * v->cw + ii) * 4 + c] * 257L;
  }
  long sum = w00 * t[0] + w10 * t[1] + w01 * t[2] + w11 * t[3];
  return (float)((sum + 128) >> 8) / 65535.f;
}

static void mat34_from_pose(const float pose[7], float M[12]) { hm_se3_matrix3x4(pose, M); }

uint32_t orc_create_surfels_for_keyframe(const orc_model* m, const orc_keyframes* kfs, int k, int filter_new_surfels,
                                         int min_observation_count, float* surfels, int pitch, uint32_t* n_inout,
                                         uint32_t max_surfels) {
  const uint32_t n = *n_inout;
  const size_t P = (size_t)pitch;
  const int w = m->depth_w, h = m->depth_h, cell = m->cell, cf_w = m->cf_w, cf_h = m->cf_h;
  const size_t npx = (size_t)w * h;
  kfview v;
  make_view(m, kfs, k, &v);
  float T[12], G[12];   /* frame_T_global, global_T_frame */
  orc_frame_T_global(kfs->global_T_frame + 7 * k, T);
  mat34_from_pose(kfs->global_T_frame + 7 * k, G);
  /* DetermineSupportingSurfelsCUDA (merge_surfels = false): only "is the cell supported at all" is used (kernel_create_surfels.cc:68-80) */
  uint8_t* occupied = (uint8_t*)calloc((size_t)cf_w * cf_h, 1);
  for (uint32_t i = 0; i < n; ++i) {
    assoc r;
    f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
    if (project_associate(&v, T, gp, f2u(surfels[ROW_NORMAL * P + i]), &r) == 3) occupied[(size_t)(r.py / cell) * cf_w + r.px / cell] = 1;
  }
  /* seeds: CreateSurfelsForKeyframeCUDASerializingKernel (kernel_create_surfels.cu:40-73); of the valid pixels of the cell the
   * one with the smallest hashed raster index (fixed pseudo-random choice in place of the reference's atomicCAS winner) */
  uint8_t* flag = (uint8_t*)calloc(npx, 1);
  for (int cy = 0; cy < cf_h; ++cy)
    for (int cx = 0; cx < cf_w; ++cx) {
      if (occupied[(size_t)cy * cf_w + cx]) continue;
      uint32_t best_key = 0xffffffffu, best = 0xffffffffu;
      for (int y = cy * cell; y < (cy + 1) * cell && y < h; ++y)
        for (int x = cx * cell; x < (cx + 1) * cell && x < w; ++x) {
          if (x < 1 || y < 1 || x >= w - 1 || y >= h - 1) continue;
          if (v.depth[(size_t)y * w + x] & K_INVALID_DEPTH_BIT) continue;
          const uint32_t seq = (uint32_t)y * (uint32_t)w + (uint32_t)x;
          const uint32_t key = seq * 0x9E3779B1u;
          if (best == 0xffffffffu || key < best_key) { best_key = key; best = seq; }
        }
      if (best != 0xffffffffu) flag[best] = 1;
    }
  free(occupied);
  if (filter_new_surfels) {   /* kernel_create_surfels.cc:99-160 */
    for (size_t s = 0; s < npx; ++s) {
      if (!flag[s]) continue;
      const int y = (int)(s / w), x = (int)(s - (size_t)y * w);
      unsigned obs = 1, viol = 0;
      float d = raw_to_calibrated_depth(v.a, v.cfactor[(size_t)(y / cell) * cf_w + x / cell], v.raw_to_float, v.depth[s]);
      f3 p_in = mk3(d * (v.fx_inv * x + v.cx_inv), d * (v.fy_inv * y + v.cy_inv), d);
      f3 n_in = u16_to_image_space_normal(v.normals[s]);
      for (int c = 0; c < kfs->K; ++c) {   /* co_visibility_list (ascending ids, direct_ba.cc:231-249) */
        if (c == k || !kfs->covis[(size_t)k * kfs->K + c]) continue;
        kfview vc;
        make_view(m, kfs, c, &vc);
        /* covis_T_frame = covis.frame_T_global * keyframe.global_T_frame (direct_ba.cc:365-370) */
        float inv_c[7], rel[7], R[12];
        hm_se3_inverse(kfs->global_T_frame + 7 * c, inv_c);
        hm_se3_mul(inv_c, kfs->global_T_frame + 7 * k, rel);
        hm_se3_matrix3x4(rel, R);
        f3 lp;
        lp.z = R[8] * p_in.x + R[9] * p_in.y + R[10] * p_in.z + R[11];
        if (!(lp.z > 0.f)) continue;
        lp.x = R[0] * p_in.x + R[1] * p_in.y + R[2] * p_in.z + R[3];
        lp.y = R[4] * p_in.x + R[5] * p_in.y + R[6] * p_in.z + R[7];
        float pxf = vc.fx * (lp.x / lp.z) + vc.cx, pyf = vc.fy * (lp.y / lp.z) + vc.cy;
        if (!(pxf >= 0.f) || !(pyf >= 0.f) || !(pxf < 1e9f) || !(pyf < 1e9f)) continue;
        int px = (int)pxf, py = (int)pyf;
        if (px >= w || py >= h) continue;
        /* IsAssociatedWithPixel<true> for a pixel-defined surfel (surfel_projection_nvcc_only.cuh:130-236) */
        uint16_t measured = vc.depth[(size_t)py * w + px];
        if (measured & K_INVALID_DEPTH_BIT) continue;
        float pd = raw_to_calibrated_depth(vc.a, vc.cfactor[(size_t)(py / cell) * cf_w + px / cell], vc.raw_to_float, measured);
        f3 ln = T_rot(R, n_in);
        float nx = vc.fx_inv * px + vc.cx_inv, ny = vc.fy_inv * py + vc.cy_inv;
        float thr = K_DEPTH_TUKEY * ((K_DEPTH_UNCERTAINTY_FACTOR * fabsf(ln.x * nx + ln.y * ny + ln.z) * (pd * pd)) / vc.baseline_fx);
        float diff = pd - lp.z;
        if (diff > thr) { ++viol; continue; }
        if (diff < -thr) continue;
        float dist = sqrtf(dot3(lp, lp));
        if ((1.0f / dist) * dot3(lp, ln) > 0) continue;
        if (dot3(ln, u16_to_image_space_normal(vc.normals[(size_t)py * w + px])) < K_COS_NORMAL_COMPAT) continue;
        ++obs;
      }
      if (obs < (unsigned)min_observation_count || viol > obs) flag[s] = 0;   /* kernel_create_surfels.cu:318-334 */
    }
  }
  uint32_t new_count = 0;
  for (size_t s = 0; s < npx; ++s) new_count += flag[s];
  if (new_count == 0 || n + new_count > max_surfels) {   /* kernel_create_surfels.cc:163-166: error, nothing is created */
    free(flag);
    return 0;
  }
  uint32_t out = n;
  for (size_t s = 0; s < npx; ++s) {   /* CreateNewSurfel, kernel_create_surfels.cu:97-165; appended in raster order */
    if (!flag[s]) continue;
    const int y = (int)(s / w), x = (int)(s - (size_t)y * w);
    float d = raw_to_calibrated_depth(v.a, v.cfactor[(size_t)(y / cell) * cf_w + x / cell], v.raw_to_float, v.depth[s]);
    f3 gp = T_mul(G, mk3(d * (v.fx_inv * x + v.cx_inv), d * (v.fy_inv * y + v.cy_inv), d));
    f3 gn = T_rot(G, u16_to_image_space_normal(v.normals[s]));
    float r2 = half_to_float(kfs->radius[npx * k + s]);
    surfels[ROW_X * P + out] = gp.x; surfels[ROW_Y * P + out] = gp.y; surfels[ROW_Z * P + out] = gp.z;
    uint32_t packed = pack_normal(gn);
    surfels[ROW_NORMAL * P + out] = u2f(packed);
    surfels[ROW_R2 * P + out] = r2;
    float ccx = v.d2c_fx * (x + 0.5f) + v.d2c_cx, ccy = v.d2c_fy * (y + 0.5f) + v.d2c_cy;
    uint32_t col = 0;
    for (int c = 0; c < 3; ++c) col |= (uint32_t)(uint8_t)(255.f * tex_channel_hw(&v, ccx, ccy, c)) << (8 * c);
    surfels[ROW_COLOR * P + out] = u2f(col);
    /* the tangent projections use the UNPACKED float normal (kernel_create_surfels.cu:134-143) */
    float t1x, t1y, t2x, t2y;
    tangent_projections(&v, T, gp, gn, r2, &t1x, &t1y, &t2x, &t2y);
    desc_eval e;
    descriptor_eval(&v, ccx, ccy, t1x, t1y, t2x, t2y, 0.f, 0.f, &e);
    surfels[ROW_D1 * P + out] = e.r1;
    surfels[ROW_D2 * P + out] = e.r2;
    ++out;
  }
  free(flag);
  *n_inout = out;
  return new_count;
}

static int cmp_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

/* DetermineSupportingSurfelsAndMergeSurfelsCUDA for keyframe k; returns the number of surfels deleted (x = NaN pattern). */
uint32_t orc_merge_surfels_for_keyframe(const orc_model* m, const orc_keyframes* kfs, int k, float merge_dist_factor,
                                        float* surfels, int pitch, uint32_t n) {
  const size_t P = (size_t)pitch;
  const int cell = m->cell, cf_w = m->cf_w, cf_h = m->cf_h;
  kfview v;
  make_view(m, kfs, k, &v);
  float T[12];
  orc_frame_T_global(kfs->global_T_frame + 7 * k, T);
  const float cell_merge_dist_squared = (float)cell * cell * merge_dist_factor * merge_dist_factor;   /* kernel_supporting_surfels.cc:76-78 */
  const uint32_t kInvalid = 0xffffffffu;
  uint32_t* sup = (uint32_t*)malloc(sizeof(uint32_t) * K_MERGE_BUFFER_COUNT * (size_t)cf_w * cf_h);
  for (size_t c = 0; c < (size_t)K_MERGE_BUFFER_COUNT * cf_w * cf_h; ++c) sup[c] = kInvalid;
  uint32_t deleted = 0;
  /* arrival order: ascending key = index * 0x9E3779B1 mod 2^32 (a bijection) */
  uint64_t* order = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) order[i] = ((uint64_t)(uint32_t)(i * 0x9E3779B1u) << 32) | i;
  qsort(order, n, sizeof(uint64_t), cmp_u64);
  for (uint32_t q = 0; q < n; ++q) {
    const uint32_t i = (uint32_t)order[q];
    assoc r;
    f3 gp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);
    if (project_associate(&v, T, gp, f2u(surfels[ROW_NORMAL * P + i]), &r) != 3) continue;
    const size_t c = (size_t)(r.py / cell) * cf_w + r.px / cell;
    int del = 0;
    for (int b = 0; b < K_MERGE_BUFFER_COUNT; ++b) {   /* kernel_supporting_surfels.cu:59-87 (no break after a merge) */
      uint32_t* slot = sup + (size_t)b * cf_w * cf_h + c;
      if (*slot == kInvalid) { *slot = i; break; }
      const uint32_t s = *slot;
      f3 sn = unpack_normal(f2u(surfels[ROW_NORMAL * P + s])), tn = unpack_normal(f2u(surfels[ROW_NORMAL * P + i]));
      if (dot3(sn, tn) > K_COS_NORMAL_COMPAT) {
        f3 sp = mk3(surfels[ROW_X * P + s], surfels[ROW_Y * P + s], surfels[ROW_Z * P + s]);
        f3 tp = mk3(surfels[ROW_X * P + i], surfels[ROW_Y * P + i], surfels[ROW_Z * P + i]);   /* (NaN once deleted) */
        float min_r2 = fminf(surfels[ROW_R2 * P + s], surfels[ROW_R2 * P + i]);
        f3 dd = sub3(sp, tp);
        if (dot3(dd, dd) < min_r2 * cell_merge_dist_squared) {
          surfels[ROW_X * P + i] = u2f(0x7fffffffu);
          del = 1;
        }
      }
    }
    deleted += (uint32_t)del;
  }
  free(order);
  free(sup);
  return deleted;
}

/* CompactSurfelsCUDA (kernel_compact_surfels.cu:159-279); active may be NULL.  Returns the new surfels_size. */
uint32_t orc_compact_surfels(float* surfels, int pitch, uint32_t n, uint8_t* active) {
  const size_t P = (size_t)pitch;
  uint8_t* invalid = (uint8_t*)malloc(n ? n : 1);
  uint32_t free_count = 0;
  for (uint32_t i = 0; i < n; ++i) { invalid[i] = f2u(surfels[ROW_X * P + i]) == 0x7fffffffu; free_count += invalid[i]; }
  if (free_count) {
    uint32_t* free_list = (uint32_t*)malloc(sizeof(uint32_t) * free_count);
    uint32_t nf = 0;
    for (uint32_t i = 0; i < n; ++i) if (invalid[i]) free_list[nf++] = i;
    uint32_t r = 0;
    for (uint32_t i = n; i-- > 0 && r < free_count;) {
      if (invalid[i]) continue;
      if (free_list[r] < i) {
        for (int row = 0; row < ROW_ACC0; ++row) surfels[(size_t)row * P + free_list[r]] = surfels[(size_t)row * P + i];
        if (active) active[free_list[r]] = active[i];
      }
      ++r;
    }
    free(free_list);
  }
  free(invalid);
  return n - free_count;
}

void orc_se3_exp(const float a[6], float out[7]) { hm_se3_exp(a, out); }
void orc_se3_log(const float T[7], float out[6]) { hm_se3_log(T, out); }
void orc_se3_mul(const float A[7], const float B[7], float out[7]) { hm_se3_mul(A, B, out); }
void orc_se3_inverse(const float A[7], float out[7]) { hm_se3_inverse(A, out); }
