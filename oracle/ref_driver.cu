// oracle/ref_driver.cu -- TEST / BASELINE INFRASTRUCTURE ONLY (never loaded by the product path).
//
// Drives the reference's OWN, UNMODIFIED CUDA kernels (compiled by build_ref.sh from the sources where they
// lie under /root/reference) through their Call...Kernel launchers.  The reference's thin host wrappers
// (kernel_opt_pose.cc, kernel_opt_geometry.cc, kernel_surfel_activation.cc) and its BA loop
// (direct_ba_alternating.cc) need Eigen/Sophus, which this image does not have, so their control flow is
// restated here on top of oracle/host_math.h -- same launches, same order, same host<->device traffic
// (2 Clear launches + kernel + 2 D2H copies + stream sync per Gauss-Newton iteration and keyframe).
//
// This is (a) the authoritative parity oracle on the GPU box ("the reference's own CUDA DirectBA on identical
// synthetic RGB-D input", BASELINE.json north_star) and (b) the reference arm of bench.py.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "badslam/cuda_depth_processing.cuh"
#include "badslam/cuda_image_processing.cuh"
#include "badslam/kernel_create_surfels.h"
#include "badslam/kernel_delete_surfels.h"
#include "badslam/kernel_supporting_surfels.h"
#include "badslam/kernel_opt_geometry.h"
#include "badslam/kernel_opt_intrinsics.h"
#include "badslam/kernel_opt_pose.h"
#include "badslam/kernel_surfel_activation.h"
#include "badslam/kernels.cuh"
#include "badslam/surfel_projection.cuh"
#include "host_math.h"

using namespace vis;

// Launchers defined in the reference's kernel_pcg.cu.  They are declared in badslam/kernels.h, which cannot be included
// here (it pulls in Eigen through libvis/camera.h), so the prototypes are repeated (kernels.h:397-495).
namespace vis {
void PCGInitCUDA(cudaStream_t stream, const SurfelProjectionParameters& s, const DepthToColorPixelCorner& depth_to_color,
                 const PixelCenterUnprojector& depth_unprojector, const PixelCornerProjector& color_projector,
                 cudaTextureObject_t color_texture, u32 kf_pose_unknown_index, u32 surfel_unknown_start_index, bool optimize_poses,
                 bool optimize_geometry, bool use_depth_residuals, bool use_descriptor_residuals, bool optimize_depth_intrinsics,
                 bool optimize_color_intrinsics, u32 depth_intrinsics_unknown_start_index, u32 color_intrinsics_unknown_start_index,
                 CUDABuffer_<PCGScalar>* pcg_r, CUDABuffer_<PCGScalar>* pcg_M, u32 surfels_size);
void PCGInit2CUDA(cudaStream_t stream, u32 unknown_count, u32 a_unknown_index, float a, const CUDABuffer_<PCGScalar>& pcg_r,
                  const CUDABuffer_<PCGScalar>& pcg_M, CUDABuffer_<PCGScalar>* pcg_delta, CUDABuffer_<PCGScalar>* pcg_g,
                  CUDABuffer_<PCGScalar>* pcg_p, CUDABuffer_<PCGScalar>* pcg_alpha_n);
void PCGStep1CUDA(cudaStream_t stream, u32 unknown_count, const SurfelProjectionParameters& s,
                  const DepthToColorPixelCorner& depth_to_color, const PixelCenterUnprojector& depth_unprojector,
                  const PixelCornerProjector& color_projector, cudaTextureObject_t color_texture, u32 kf_pose_unknown_index,
                  u32 surfel_unknown_start_index, bool optimize_poses, bool optimize_geometry, bool use_depth_residuals,
                  bool use_descriptor_residuals, bool optimize_depth_intrinsics, bool optimize_color_intrinsics,
                  u32 depth_intrinsics_unknown_start_index, u32 a_unknown_index, u32 color_intrinsics_unknown_start_index,
                  CUDABuffer_<PCGScalar>* pcg_p, CUDABuffer_<PCGScalar>* pcg_g, CUDABuffer_<PCGScalar>* pcg_alpha_d, u32 surfels_size);
void PCGStep2CUDA(cudaStream_t stream, u32 unknown_count, u32 a_unknown_index, const CUDABuffer_<PCGScalar>& pcg_r,
                  const CUDABuffer_<PCGScalar>& pcg_M, CUDABuffer_<PCGScalar>* pcg_delta, CUDABuffer_<PCGScalar>* pcg_g,
                  CUDABuffer_<PCGScalar>* pcg_p, CUDABuffer_<PCGScalar>* pcg_alpha_n, CUDABuffer_<PCGScalar>* pcg_alpha_d,
                  CUDABuffer_<PCGScalar>* pcg_beta_n);
void PCGStep3CUDA(cudaStream_t stream, u32 unknown_count, CUDABuffer_<PCGScalar>* pcg_g, CUDABuffer_<PCGScalar>* pcg_p,
                  CUDABuffer_<PCGScalar>* pcg_alpha_n, CUDABuffer_<PCGScalar>* pcg_beta_n);
void UpdateSurfelsFromPCGDeltaCUDA(cudaStream_t stream, u32 surfels_size, CUDABuffer_<float>* surfels, bool use_descriptor_residuals,
                                   u32 surfel_unknown_start_index, const CUDABuffer_<PCGScalar>& pcg_delta);
void UpdateCFactorsFromPCGDeltaCUDA(cudaStream_t stream, CUDABuffer_<float>* cfactor_buffer, u32 cfactor_unknown_start_index,
                                    const CUDABuffer_<PCGScalar>& pcg_delta);
// kernel_downsample.cu (kernels.h:338-378)
void CalibrateDepthAndTransformColorToDepthCUDA(cudaStream_t stream, const DepthToColorPixelCorner& depth_to_color,
                                                const DepthParameters& depth_params, const CUDABuffer_<u16>& depth_buffer,
                                                cudaTextureObject_t color_texture, CUDABuffer_<float>* out_depth, CUDABuffer_<u8>* out_color);
void CalibrateDepthCUDA(cudaStream_t stream, const DepthParameters& depth_params, const CUDABuffer_<u16>& depth_buffer,
                        CUDABuffer_<float>* out_depth);
void CalibrateAndDownsampleImagesCUDA(cudaStream_t stream, bool downsample_color, const DepthParameters& depth_params,
                                      const CUDABuffer_<u16>& depth_buffer, const CUDABuffer_<u16>& normals_buffer,
                                      cudaTextureObject_t color_texture, CUDABuffer_<float>* downsampled_depth,
                                      CUDABuffer_<u16>* downsampled_normals, CUDABuffer_<u8>* downsampled_color, bool debug);
void DownsampleImagesCUDA(cudaStream_t stream, const CUDABuffer_<float>& depth_buffer, const CUDABuffer_<u16>& normals_buffer,
                          cudaTextureObject_t color_texture, CUDABuffer_<float>* downsampled_depth,
                          CUDABuffer_<u16>* downsampled_normals, CUDABuffer_<u8>* downsampled_color, bool debug);
void CompactSurfelsCUDA(cudaStream_t stream, void** free_spots_temp_storage, usize* free_spots_temp_storage_bytes, u32 surfel_count,
                        u32* surfels_size, CUDABuffer_<float>* surfels, CUDABuffer_<u8>* active_surfels);   // kernels.h:292-299
}  // namespace vis

namespace {

struct RefKeyframe {
  u16* depth = nullptr; size_t depth_pitch = 0;
  u16* normals = nullptr; size_t normals_pitch = 0;
  u16* radius = nullptr; size_t radius_pitch = 0;
  uchar4* color = nullptr; size_t color_pitch = 0;
  cudaTextureObject_t tex = 0;
  float pose[7];        // global_T_frame
  int activation = 0;   // 0 active, 1 covis-active, 2 inactive
  float min_depth = 0, max_depth = 0;
  std::vector<int> covis;
};

}  // namespace

struct ref_config {
  int depth_w, depth_h, color_w, color_h;
  float depth_K[4], color_K[4];
  float raw_to_float_depth, baseline_fx;
  int cell;
  int use_depth_residuals, use_descriptor_residuals;
};

struct ref_ba_options {
  int optimize_poses, optimize_geometry;
  int min_iterations, max_iterations;
  int active_keyframe_window_start, active_keyframe_window_end;
  int optimize_depth_intrinsics, optimize_color_intrinsics;
  int end_tasks;   // PerformBASchemeEndTasks at the end (increase_ba_iteration_count = true, direct_ba_alternating.cc:725-735)
};

struct ref_ba_result {
  int iterations_done, converged;
  unsigned long long n_count;   // debug residual_count summed over keyframes at the LAST iteration's pose step start
  double cost;                  // debug residual sum, same convention
  int pose_iterations_total;
  float ms_surfel_activation, ms_geometry_optimization, ms_pose_optimization;
  unsigned long long kernel_launches;
  unsigned int surfels_deleted, surfels_size;
  unsigned long long n_depth_count;   // count_residuals == 2: the debug count of the same launches with the descriptor residuals
                                      // off, i.e. the number of associated pairs (depth residuals) alone
};

struct ref_context {
  ref_config cfg;
  float a = 0.f;
  float* cfactor = nullptr; size_t cfactor_pitch = 0; int cf_w = 0, cf_h = 0;
  float* surfels = nullptr; size_t surfel_pitch = 0; u32 surfels_size = 0; u32 max_surfels = 0;
  u8* active = nullptr;
  // PoseEstimationHelperBuffers (kernels.h:47-58)
  u32* residual_count = nullptr; float* residual_sum = nullptr; float* H = nullptr; float* b = nullptr;
  std::vector<RefKeyframe> kfs;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4];
  unsigned long long launches = 0;
  // IntrinsicsOptimizationHelperBuffers (kernels.h:60-93), lazily allocated
  u32* intr_obs = nullptr; float* intr_A = nullptr; float* intr_B = nullptr; float* intr_D = nullptr;
  float* intr_b1 = nullptr; float* intr_b2 = nullptr; float* intr_H = nullptr; float* intr_b = nullptr;
  // PCG vectors r, M, delta, g, p (direct_ba_pcg.cc:256-266) + the three scalars
  float* pcg[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; size_t pcg_capacity = 0; float* pcg_scalars = nullptr;
  // PerformBASchemeEndTasks
  u32* deleted_count = nullptr; void* free_spots_temp = nullptr; usize free_spots_temp_bytes = 0;
  int min_observation_count[3] = {1, 2, 3};   // bad_slam_config.h:146,151,158
  // surfel creation / merge (direct_ba.cc:131-145)
  u32* sup[3] = {nullptr, nullptr, nullptr}; size_t sup_pitch = 0;
  u8* new_flag = nullptr; u32* new_indices = nullptr; void* new_temp = nullptr; usize new_temp_bytes = 0;
  float surfel_merge_dist_factor = 0.8f;      // bad_slam_config.h
  // image-pair odometry (PairwiseFrameTrackingBuffers + the RunOdometry inputs), lazily allocated by ref_track_frame_pairwise
  struct OdoImage { float* depth = nullptr; size_t depth_pitch = 0; u16* normals = nullptr; size_t normals_pitch = 0;
                    u8* color = nullptr; size_t color_pitch = 0; cudaTextureObject_t tex = 0; int w = 0, h = 0; bool owns_normals = false; };
  OdoImage odo[2][8];            // [0 base | 1 tracked][scale]
  u8* odo_gradmag[2] = {nullptr, nullptr}; size_t odo_gradmag_pitch[2] = {0, 0}; cudaTextureObject_t odo_gradmag_tex[2] = {0, 0};
  RefKeyframe odo_frame;         // the tracked frame's device images
  float* odo_debug = nullptr; size_t odo_debug_pitch = 0;   // debug_residual_image of the accumulate kernels (written when debug = true)
  int odo_scales = 0;
};

extern "C" unsigned int ref_end_tasks(ref_context* c);

namespace {

CUDAMatrix3x4 MakeFrameTGlobal(const float global_T_frame[7]) {
  float inv[7], M[12];
  hm_se3_inverse(global_T_frame, inv);
  hm_se3_matrix3x4(inv, M);
  CUDAMatrix3x4 r;
  r.row0 = make_float4(M[0], M[1], M[2], M[3]);
  r.row1 = make_float4(M[4], M[5], M[6], M[7]);
  r.row2 = make_float4(M[8], M[9], M[10], M[11]);
  return r;
}

CUDAMatrix3x3 MakeGlobalRFrame(const float global_T_frame[7]) {   // Keyframe::global_R_frame_cuda
  float R[9];
  hm_quat_to_R(global_T_frame, R);
  CUDAMatrix3x3 r;
  r.row0 = make_float3(R[0], R[1], R[2]);
  r.row1 = make_float3(R[3], R[4], R[5]);
  r.row2 = make_float3(R[6], R[7], R[8]);
  return r;
}

DepthParameters MakeDepthParams(ref_context* c) {   // direct_ba.cc:108-120
  DepthParameters d;
  d.cfactor_buffer = CUDABuffer_<float>(c->cfactor, c->cf_h, c->cf_w, c->cfactor_pitch);
  d.a = c->a;
  d.raw_to_float_depth = c->cfg.raw_to_float_depth;
  d.baseline_fx = c->cfg.baseline_fx;
  d.sparse_surfel_cell_size = c->cfg.cell;
  return d;
}

// surfel_projection.h:42-124
PixelCornerProjector CornerProjector(const float K[4]) { return PixelCornerProjector(K[0], K[1], K[2], K[3]); }
PixelCenterProjector CenterProjector(const float K[4]) { return PixelCenterProjector(K[0], K[1], K[2] - 0.5f, K[3] - 0.5f); }
PixelCenterUnprojector CenterUnprojector(const float K[4]) {
  const float fx_inv = 1.0f / K[0], fy_inv = 1.0f / K[1];
  return PixelCenterUnprojector(fx_inv, fy_inv, -(K[2] - 0.5f) * fx_inv, -(K[3] - 0.5f) * fy_inv);
}
DepthToColorPixelCorner DepthToColor(const ref_config& cfg) {
  DepthToColorPixelCorner r;
  r.width = cfg.color_w;
  r.height = cfg.color_h;
  r.fx = cfg.color_K[0] / cfg.depth_K[0];
  r.cx = -1 * cfg.color_K[0] * cfg.depth_K[2] / cfg.depth_K[0] + cfg.color_K[2];
  r.fy = cfg.color_K[1] / cfg.depth_K[1];
  r.cy = -1 * cfg.color_K[1] * cfg.depth_K[3] / cfg.depth_K[1] + cfg.color_K[3];
  return r;
}

SurfelProjectionParameters MakeProjection(ref_context* c, const RefKeyframe& kf, const CUDAMatrix3x4& frame_T_global,
                                          u32 surfels_size) {
  return SurfelProjectionParameters(
      CUDABuffer_<float>(c->surfels, kSurfelAttributeCount, c->max_surfels, c->surfel_pitch),
      CUDABuffer_<u16>(kf.depth, c->cfg.depth_h, c->cfg.depth_w, kf.depth_pitch),
      CUDABuffer_<u16>(kf.normals, c->cfg.depth_h, c->cfg.depth_w, kf.normals_pitch), MakeDepthParams(c),
      CornerProjector(c->cfg.depth_K), CenterUnprojector(c->cfg.depth_K), frame_T_global, surfels_size);
}

CUDABuffer_<float> SurfelBuf(ref_context* c) { return CUDABuffer_<float>(c->surfels, kSurfelAttributeCount, c->max_surfels, c->surfel_pitch); }
CUDABuffer_<u8> ActiveBuf(ref_context* c) { return CUDABuffer_<u8>(c->active, 1, c->max_surfels, c->max_surfels); }

// kernel_opt_pose.cc:39-97 (debug = true)
void AccumulatePoseEstimationCoeffs(ref_context* c, int k, const float global_T_frame[7], bool debug, u32* residual_count,
                                    float* residual_sum, float* H, float* b) {
  cudaStream_t s = c->stream;
  CUDABuffer_<u32> count_buf(c->residual_count, 1, 1, sizeof(u32));
  CUDABuffer_<float> sum_buf(c->residual_sum, 1, 1, sizeof(float));
  CUDABuffer_<float> H_buf(c->H, 1, 21, 21 * sizeof(float));
  CUDABuffer_<float> b_buf(c->b, 1, 6, 6 * sizeof(float));
  if (debug) {
    count_buf.Clear(0, s);
    sum_buf.Clear(0, s);
    c->launches += 2;
  }
  H_buf.Clear(0, s);
  b_buf.Clear(0, s);
  const RefKeyframe& kf = c->kfs[k];
  CallAccumulatePoseEstimationCoeffsCUDAKernel(
      s, debug, c->cfg.use_depth_residuals != 0, c->cfg.use_descriptor_residuals != 0,
      MakeProjection(c, kf, MakeFrameTGlobal(global_T_frame), c->surfels_size), DepthToColor(c->cfg),
      CenterProjector(c->cfg.color_K), CornerProjector(c->cfg.color_K), CenterUnprojector(c->cfg.depth_K), kf.tex, count_buf,
      sum_buf, H_buf, b_buf);
  c->launches += 3;
  if (debug) {
    cudaMemcpyAsync(residual_count, c->residual_count, sizeof(u32), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(residual_sum, c->residual_sum, sizeof(float), cudaMemcpyDeviceToHost, s);
  }
  cudaMemcpyAsync(H, c->H, 21 * sizeof(float), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(b, c->b, 6 * sizeof(float), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
}

// direct_ba_alternating.cc:42-283
int EstimateFramePose(ref_context* c, int k, const float init[7], float out[7], int* converged_out, bool debug_first,
                      u32* first_count, float* first_sum, u32* first_depth_count = nullptr) {
  float est[7];
  std::memcpy(est, init, sizeof(est));
  int converged = 0, iteration;
  for (iteration = 0; iteration < 30; ++iteration) {
    float H[21], b[6];
    u32 cnt = 0;
    float sum = 0;
    double Hd[36], bd[6], xd[6];
    std::memset(Hd, 0, sizeof(Hd));
    if (c->surfels_size == 0) {
      std::memset(bd, 0, sizeof(bd));
    } else {
      const bool dbg = debug_first && iteration == 0;
      AccumulatePoseEstimationCoeffs(c, k, est, dbg, &cnt, &sum, H, b);
      if (dbg) { *first_count = cnt; *first_sum = sum; }
      if (dbg && first_depth_count && c->cfg.use_depth_residuals) {
        // the same launch with the descriptor residuals switched off: its debug counter is the number of depth residuals
        // (kernel_opt_pose.cu:312-320), which separates n_assoc from the n_assoc + n_photo the combined counter reports
        float H2[21], b2[6], sum2 = 0;
        const int desc = c->cfg.use_descriptor_residuals;
        c->cfg.use_descriptor_residuals = 0;
        AccumulatePoseEstimationCoeffs(c, k, est, true, first_depth_count, &sum2, H2, b2);
        c->cfg.use_descriptor_residuals = desc;
      }
      int idx = 0;
      for (int r = 0; r < 6; ++r)
        for (int cc = r; cc < 6; ++cc) Hd[r * 6 + cc] = H[idx++];
      for (int i = 0; i < 6; ++i) bd[i] = b[i];
    }
    hm_ldlt_solve(6, Hd, bd, xd);
    float x[6], nx[6], e[7], next[7];
    for (int i = 0; i < 6; ++i) { x[i] = static_cast<float>(xd[i]); nx[i] = -x[i]; }
    hm_se3_exp(nx, e);
    hm_se3_mul(est, e, next);
    std::memcpy(est, next, sizeof(est));
    converged = hm_is_scale1_pose_converged(x);
    if (converged) { ++iteration; break; }
  }
  std::memcpy(out, est, sizeof(est));
  if (converged_out) *converged_out = converged;
  return iteration;
}

// kernel_surfel_activation.cc:39-67
void UpdateSurfelActivation(ref_context* c) {
  if (c->surfels_size == 0) return;
  CallSetSurfelInactiveKernel(c->stream, c->surfels_size, ActiveBuf(c));
  ++c->launches;
  for (const RefKeyframe& kf : c->kfs) {
    if (kf.activation != 0) continue;
    CallDetermineActiveSurfelsKernel(c->stream, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size), ActiveBuf(c));
    ++c->launches;
  }
}

// kernel_opt_geometry.cc:39-77
void UpdateSurfelNormals(ref_context* c) {
  if (c->surfels_size == 0) return;
  cudaStream_t s = c->stream;
  CallResetSurfelAccum0to3CUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
  ++c->launches;
  for (const RefKeyframe& kf : c->kfs) {
    if (kf.activation == 2) continue;
    CallAccumulateSurfelNormalOptimizationCoeffsCUDAKernel(s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size),
                                                           MakeGlobalRFrame(kf.pose), ActiveBuf(c));
    ++c->launches;
  }
  CallUpdateSurfelNormalCUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
  ++c->launches;
}

// kernel_opt_geometry.cc:80-201
void OptimizeGeometryIteration(ref_context* c) {
  if (c->surfels_size == 0) return;
  cudaStream_t s = c->stream;
  CallResetSurfelAccum0to3CUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
  ++c->launches;
  for (const RefKeyframe& kf : c->kfs) {
    if (kf.activation == 2) continue;
    CallAccumulateSurfelNormalOptimizationCoeffsCUDAKernel(s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size),
                                                           MakeGlobalRFrame(kf.pose), ActiveBuf(c));
    ++c->launches;
  }
  CallUpdateSurfelNormalCUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
  ++c->launches;
  if (!c->cfg.use_descriptor_residuals) {
    CallResetSurfelAccum0to1CUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
    ++c->launches;
    for (const RefKeyframe& kf : c->kfs) {
      if (kf.activation == 2) continue;
      CallAccumulateSurfelPositionOptimizationCoeffsFromDepthResidualCUDAKernel(
          s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size), CenterUnprojector(c->cfg.depth_K),
          DepthToColor(c->cfg), c->cfg.color_K[0], c->cfg.color_K[1], kf.tex, ActiveBuf(c));
      ++c->launches;
    }
    CallUpdateSurfelPositionCUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
    ++c->launches;
  } else {
    CallResetSurfelAccumCUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
    ++c->launches;
    for (const RefKeyframe& kf : c->kfs) {
      if (kf.activation == 2) continue;
      AccumulateSurfelPositionAndDescriptorOptimizationCoeffsCUDAKernel(
          s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size), CenterUnprojector(c->cfg.depth_K),
          DepthToColor(c->cfg), CornerProjector(c->cfg.color_K), kf.tex, ActiveBuf(c), c->cfg.use_depth_residuals != 0);
      ++c->launches;
    }
    CallUpdateSurfelPositionAndDescriptorCUDAKernel(s, c->surfels_size, SurfelBuf(c), ActiveBuf(c));
    ++c->launches;
  }
}

// kernel_opt_intrinsics.cc:39-281 (thin host wrapper around the reference's three kernels; Eigen's fp64 LDLT -> hm_ldlt_solve)
void OptimizeIntrinsics(ref_context* c, bool opt_depth, bool opt_color) {
  if (c->surfels_size == 0) return;
  constexpr int kARows = 5;
  cudaStream_t s = c->stream;
  const int P = c->cf_w * c->cf_h;
  if (!c->intr_obs) {
    cudaMalloc(&c->intr_obs, sizeof(u32) * P);
    cudaMalloc(&c->intr_A, sizeof(float) * 15);
    cudaMalloc(&c->intr_B, sizeof(float) * kARows * P);
    cudaMalloc(&c->intr_D, sizeof(float) * P);
    cudaMalloc(&c->intr_b1, sizeof(float) * kARows);
    cudaMalloc(&c->intr_b2, sizeof(float) * P);
    cudaMalloc(&c->intr_H, sizeof(float) * 10);
    cudaMalloc(&c->intr_b, sizeof(float) * 4);
  }
  CUDABuffer_<u32> obs(c->intr_obs, 1, P, sizeof(u32) * P);
  CUDABuffer_<float> A(c->intr_A, 1, 15, sizeof(float) * 15), B(c->intr_B, kARows, P, sizeof(float) * P),
      D(c->intr_D, 1, P, sizeof(float) * P), b1(c->intr_b1, 1, kARows, sizeof(float) * kARows),
      b2(c->intr_b2, 1, P, sizeof(float) * P), H(c->intr_H, 1, 10, sizeof(float) * 10), b(c->intr_b, 1, 4, sizeof(float) * 4);
  if (opt_depth) {
    cudaMemsetAsync(c->intr_obs, 0, sizeof(u32) * P, s);
    cudaMemsetAsync(c->intr_A, 0, sizeof(float) * 15, s);
    cudaMemsetAsync(c->intr_B, 0, sizeof(float) * kARows * P, s);
    cudaMemsetAsync(c->intr_D, 0, sizeof(float) * P, s);
    cudaMemsetAsync(c->intr_b1, 0, sizeof(float) * kARows, s);
    cudaMemsetAsync(c->intr_b2, 0, sizeof(float) * P, s);
  }
  if (opt_color) {
    cudaMemsetAsync(c->intr_H, 0, sizeof(float) * 10, s);
    cudaMemsetAsync(c->intr_b, 0, sizeof(float) * 4, s);
  }
  const PixelCenterUnprojector unproj = CenterUnprojector(c->cfg.depth_K);
  for (const RefKeyframe& kf : c->kfs) {
    CallAccumulateIntrinsicsCoefficientsCUDAKernel(s, opt_color, opt_depth,
                                                   MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size),
                                                   DepthToColor(c->cfg), CornerProjector(c->cfg.color_K), unproj,
                                                   c->cfg.color_K[0], c->cfg.color_K[1], kf.tex, obs, A, B, D, b1, b2, H, b);
    ++c->launches;
  }
  if (opt_depth) {
    CallComputeIntrinsicsIntermediateMatricesCUDAKernel(s, P, A, B, D, b1, b2);
    ++c->launches;
    float A_cpu[15], rhs[kARows];
    cudaMemcpyAsync(A_cpu, c->intr_A, sizeof(A_cpu), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(rhs, c->intr_b1, sizeof(rhs), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    constexpr float kAPriorWeight = 10;
    A_cpu[14] += kAPriorWeight * kAPriorWeight;
    rhs[4] += kAPriorWeight * kAPriorWeight * c->a;
    double Ad[kARows * kARows] = {0}, bd[kARows], xd[kARows];   // hm_ldlt_solve reads the upper triangle of a full matrix
    for (int r = 0, i = 0; r < kARows; ++r)
      for (int col = r; col < kARows; ++col) Ad[r * kARows + col] = A_cpu[i++];
    for (int i = 0; i < kARows; ++i) bd[i] = rhs[i];
    hm_ldlt_solve(kARows, Ad, bd, xd);
    float x1[kARows];
    for (int i = 0; i < kARows; ++i) x1[i] = static_cast<float>(xd[i]);
    const float new_fx = 1.0f / (unproj.fx_inv - x1[0]);
    const float new_fy = 1.0f / (unproj.fy_inv - x1[1]);
    const float new_cx = -(new_fx * (unproj.cx_inv - x1[2])) + 0.5f;
    const float new_cy = -(new_fy * (unproj.cy_inv - x1[3])) + 0.5f;
    cudaMemcpyAsync(c->intr_b1, x1, sizeof(x1), cudaMemcpyHostToDevice, s);
    CallSolveForPixelIntrinsicsUpdateCUDAKernel(s, P, obs, B, D, b1,
                                                CUDABuffer_<float>(c->cfactor, c->cf_h, c->cf_w, c->cfactor_pitch));
    ++c->launches;
    cudaStreamSynchronize(s);
    c->cfg.depth_K[0] = new_fx; c->cfg.depth_K[1] = new_fy; c->cfg.depth_K[2] = new_cx; c->cfg.depth_K[3] = new_cy;
    c->a -= x1[4];
  }
  if (opt_color) {   // kernel_opt_intrinsics.cc:256-280 (system accumulated above, before any update)
    float H_cpu[10], rhs4[4];
    cudaMemcpyAsync(H_cpu, c->intr_H, sizeof(H_cpu), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(rhs4, c->intr_b, sizeof(rhs4), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    double Hd[16] = {0}, b4[4], x4[4];
    for (int r = 0, i = 0; r < 4; ++r)
      for (int col = r; col < 4; ++col) Hd[r * 4 + col] = H_cpu[i++];
    for (int i = 0; i < 4; ++i) b4[i] = rhs4[i];
    hm_ldlt_solve(4, Hd, b4, x4);
    for (int i = 0; i < 4; ++i) c->cfg.color_K[i] -= static_cast<float>(x4[i]);
  }
}

void DetermineCovisibleActive(ref_context* c) {   // direct_ba.cc:549-564
  for (RefKeyframe& kf : c->kfs) {
    if (kf.activation != 0) continue;
    for (int o : kf.covis)
      if (c->kfs[o].activation == 2) c->kfs[o].activation = 1;
  }
}

}  // namespace

extern "C" {

ref_context* ref_create(const ref_config* cfg, unsigned int max_surfels) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return nullptr;
  ref_context* c = new ref_context();
  c->cfg = *cfg;
  c->cf_w = (cfg->depth_w - 1) / cfg->cell + 1;
  c->cf_h = (cfg->depth_h - 1) / cfg->cell + 1;
  // libvis CUDABuffer = cudaMallocPitch (cuda_buffer_inl.h:36-41)
  cudaMallocPitch(reinterpret_cast<void**>(&c->cfactor), &c->cfactor_pitch, c->cf_w * sizeof(float), c->cf_h);
  cudaMemset2D(c->cfactor, c->cfactor_pitch, 0, c->cf_w * sizeof(float), c->cf_h);
  c->max_surfels = max_surfels;
  cudaMallocPitch(reinterpret_cast<void**>(&c->surfels), &c->surfel_pitch, static_cast<size_t>(max_surfels) * sizeof(float),
                  kSurfelAttributeCount);
  cudaMemset2D(c->surfels, c->surfel_pitch, 0, static_cast<size_t>(max_surfels) * sizeof(float), kSurfelAttributeCount);
  cudaMalloc(&c->active, max_surfels);
  cudaMemset(c->active, 0, max_surfels);
  cudaMalloc(&c->residual_count, sizeof(u32));
  cudaMalloc(&c->residual_sum, sizeof(float));
  cudaMalloc(&c->H, 21 * sizeof(float));
  cudaMalloc(&c->b, 6 * sizeof(float));
  cudaStreamCreate(&c->stream);
  for (auto& e : c->ev) cudaEventCreate(&e);
  if (cudaDeviceSynchronize() != cudaSuccess) { delete c; return nullptr; }
  return c;
}

void ref_destroy(ref_context* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (RefKeyframe& kf : c->kfs) {
    if (kf.tex) cudaDestroyTextureObject(kf.tex);
    cudaFree(kf.depth); cudaFree(kf.normals); cudaFree(kf.radius); cudaFree(kf.color);
  }
  cudaFree(c->cfactor); cudaFree(c->surfels); cudaFree(c->active);
  cudaFree(c->residual_count); cudaFree(c->residual_sum); cudaFree(c->H); cudaFree(c->b);
  // image-pair odometry buffers
  for (int f = 0; f < 2; ++f) {
    if (c->odo_gradmag_tex[f]) cudaDestroyTextureObject(c->odo_gradmag_tex[f]);
    cudaFree(c->odo_gradmag[f]);
    for (int sc = 0; sc < c->odo_scales; ++sc) {
      ref_context::OdoImage& im = c->odo[f][sc];
      if (im.tex) cudaDestroyTextureObject(im.tex);
      cudaFree(im.depth);
      if (im.owns_normals && sc >= 1) cudaFree(im.normals);
      cudaFree(im.color);
    }
  }
  if (c->odo_frame.tex) cudaDestroyTextureObject(c->odo_frame.tex);
  cudaFree(c->odo_frame.depth); cudaFree(c->odo_frame.normals); cudaFree(c->odo_frame.color);
  cudaFree(c->odo_debug);
  for (auto& e : c->ev) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  delete c;
}

int ref_set_surfels(ref_context* c, const float* host, size_t host_pitch_bytes, unsigned int n) {
  if (n > c->max_surfels) return 1;
  c->surfels_size = n;
  if (n == 0) return 0;
  cudaMemcpy2D(c->surfels, c->surfel_pitch, host, host_pitch_bytes, static_cast<size_t>(n) * 4, 8, cudaMemcpyHostToDevice);
  return cudaGetLastError() != cudaSuccess;
}
int ref_get_surfels(ref_context* c, float* host, size_t host_pitch_bytes, int rows) {
  cudaMemcpy2D(host, host_pitch_bytes, c->surfels, c->surfel_pitch, static_cast<size_t>(c->surfels_size) * 4, rows,
               cudaMemcpyDeviceToHost);
  return cudaGetLastError() != cudaSuccess;
}
int ref_get_active(ref_context* c, unsigned char* host) {
  cudaMemcpy(host, c->active, c->surfels_size, cudaMemcpyDeviceToHost);
  return cudaGetLastError() != cudaSuccess;
}
int ref_set_active(ref_context* c, const unsigned char* host) {
  cudaMemcpy(c->active, host, c->surfels_size, cudaMemcpyHostToDevice);
  return cudaGetLastError() != cudaSuccess;
}
void ref_set_depth_params(ref_context* c, float a, const float* cfactor_dense) {
  c->a = a;
  if (cfactor_dense)
    cudaMemcpy2D(c->cfactor, c->cfactor_pitch, cfactor_dense, c->cf_w * sizeof(float), c->cf_w * sizeof(float), c->cf_h,
                 cudaMemcpyHostToDevice);
}
void ref_set_intrinsics(ref_context* c, const float* depth_K, const float* color_K) {
  if (depth_K) std::memcpy(c->cfg.depth_K, depth_K, sizeof(float) * 4);
  if (color_K) std::memcpy(c->cfg.color_K, color_K, sizeof(float) * 4);
}

int ref_add_keyframe(ref_context* c, const unsigned short* depth, const unsigned short* normals, const unsigned short* radius,
                     const unsigned char* color_rgba, const float pose[7], float min_depth, float max_depth) {
  RefKeyframe kf;
  const int w = c->cfg.depth_w, h = c->cfg.depth_h, cw = c->cfg.color_w, ch = c->cfg.color_h;
  cudaMallocPitch(reinterpret_cast<void**>(&kf.depth), &kf.depth_pitch, w * 2, h);
  cudaMallocPitch(reinterpret_cast<void**>(&kf.normals), &kf.normals_pitch, w * 2, h);
  cudaMallocPitch(reinterpret_cast<void**>(&kf.radius), &kf.radius_pitch, w * 2, h);
  cudaMallocPitch(reinterpret_cast<void**>(&kf.color), &kf.color_pitch, cw * 4, ch);
  cudaMemcpy2D(kf.depth, kf.depth_pitch, depth, w * 2, w * 2, h, cudaMemcpyHostToDevice);
  cudaMemcpy2D(kf.normals, kf.normals_pitch, normals, w * 2, w * 2, h, cudaMemcpyHostToDevice);
  if (radius) cudaMemcpy2D(kf.radius, kf.radius_pitch, radius, w * 2, w * 2, h, cudaMemcpyHostToDevice);
  cudaMemcpy2D(kf.color, kf.color_pitch, color_rgba, cw * 4, cw * 4, ch, cudaMemcpyHostToDevice);
  // CUDABuffer<uchar4>::CreateTextureObject as called in keyframe.cc:67-73 (cuda_buffer_inl.h:188-214)
  cudaResourceDesc res;
  std::memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = kf.color;
  res.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
  res.res.pitch2D.width = cw;
  res.res.pitch2D.height = ch;
  res.res.pitch2D.pitchInBytes = kf.color_pitch;
  cudaTextureDesc tex;
  std::memset(&tex, 0, sizeof(tex));
  tex.addressMode[0] = cudaAddressModeClamp;
  tex.addressMode[1] = cudaAddressModeClamp;
  tex.filterMode = cudaFilterModeLinear;
  tex.readMode = cudaReadModeNormalizedFloat;
  tex.normalizedCoords = 0;
  if (cudaCreateTextureObject(&kf.tex, &res, &tex, nullptr) != cudaSuccess) return -1;
  std::memcpy(kf.pose, pose, sizeof(kf.pose));
  kf.activation = 0;
  kf.min_depth = min_depth;
  kf.max_depth = max_depth;
  const int id = static_cast<int>(c->kfs.size());
  hm_frustum fn;
  hm_frustum_create(&fn, c->cfg.depth_K, w, h, min_depth, max_depth, pose);
  for (int k = 0; k < id; ++k) {   // direct_ba.cc:231-249
    hm_frustum fo;
    hm_frustum_create(&fo, c->cfg.depth_K, w, h, c->kfs[k].min_depth, c->kfs[k].max_depth, c->kfs[k].pose);
    if (hm_frustum_intersects(&fn, &fo)) {
      kf.covis.push_back(k);
      c->kfs[k].covis.push_back(id);
      if (c->kfs[k].activation == 2) c->kfs[k].activation = 1;
    }
  }
  c->kfs.push_back(kf);
  return id;
}

void ref_get_pose(ref_context* c, int k, float pose[7]) { std::memcpy(pose, c->kfs[k].pose, sizeof(float) * 7); }
void ref_set_pose(ref_context* c, int k, const float pose[7]) { std::memcpy(c->kfs[k].pose, pose, sizeof(float) * 7); }
int ref_get_activation(ref_context* c, int k) { return c->kfs[k].activation; }
void ref_set_activation(ref_context* c, int k, int a) { c->kfs[k].activation = a; }
unsigned long long ref_launch_count(ref_context* c) { return c->launches; }

void ref_pose_coeffs(ref_context* c, int k, const float pose[7], float H[21], float b[6], unsigned int* count, float* cost) {
  AccumulatePoseEstimationCoeffs(c, k, pose, true, count, cost, H, b);
}

int ref_estimate_frame_pose(ref_context* c, int k, const float init[7], float out[7], int* converged) {
  return EstimateFramePose(c, k, init, out, converged, false, nullptr, nullptr);
}

void ref_optimize_intrinsics(ref_context* c, int opt_depth, int opt_color) { OptimizeIntrinsics(c, opt_depth != 0, opt_color != 0); }
void ref_get_intrinsics(ref_context* c, float depth_K[4], float color_K[4], float* a) {
  std::memcpy(depth_K, c->cfg.depth_K, sizeof(float) * 4);
  std::memcpy(color_K, c->cfg.color_K, sizeof(float) * 4);
  *a = c->a;
}
void ref_get_cfactor(ref_context* c, float* cfactor_dense) {
  cudaMemcpy2D(cfactor_dense, c->cf_w * sizeof(float), c->cfactor, c->cfactor_pitch, c->cf_w * sizeof(float), c->cf_h,
               cudaMemcpyDeviceToHost);
}
void ref_update_activation(ref_context* c) { UpdateSurfelActivation(c); cudaStreamSynchronize(c->stream); }
void ref_optimize_geometry_iteration(ref_context* c) { OptimizeGeometryIteration(c); cudaStreamSynchronize(c->stream); }

// direct_ba_alternating.cc:285-738 without the surfel lifecycle / intrinsics branches.
// `count_residuals`: run the first Gauss-Newton iteration of every keyframe with debug = true to obtain the
// reference's own residual count / cost (kernel_opt_pose.cu:224-248).
void ref_bundle_adjust(ref_context* c, const ref_ba_options* o, ref_ba_result* res, int count_residuals) {
  std::memset(res, 0, sizeof(*res));
  const int K = static_cast<int>(c->kfs.size());
  cudaStream_t s = c->stream;
  const unsigned long long launches_before = c->launches;
  const bool fixed_window = o->active_keyframe_window_start > 0 || o->active_keyframe_window_end > 0;
  const bool whole_window = !(o->active_keyframe_window_start != 0 || o->active_keyframe_window_end != K - 1);
  cudaMemsetAsync(c->active, 0, c->surfels_size, s);
  for (int iteration = 0; iteration < o->max_iterations; ++iteration) {
    ++res->iterations_done;
    if (fixed_window) {
      for (int k = 0; k < K; ++k)
        c->kfs[k].activation = (k >= o->active_keyframe_window_start && k <= o->active_keyframe_window_end) ? 0 : 2;
      DetermineCovisibleActive(c);
    }
    cudaEventRecord(c->ev[0], s);
    if (!whole_window) cudaMemsetAsync(c->active, kSurfelActiveFlag, c->surfels_size, s);
    else UpdateSurfelActivation(c);
    cudaEventRecord(c->ev[1], s);
    if (o->optimize_geometry) OptimizeGeometryIteration(c);
    cudaEventRecord(c->ev[2], s);
    int num_converged = 0;
    if (o->optimize_poses) {
      res->n_count = 0;
      res->n_depth_count = 0;
      res->cost = 0;
      for (int k = 0; k < K; ++k) {
        RefKeyframe& kf = c->kfs[k];
        if (kf.activation == 2) { ++num_converged; continue; }
        float est[7], ftg[7], diff[7], lg[6];
        int conv;
        u32 cnt = 0, dcnt = 0;
        float sum = 0;
        res->pose_iterations_total +=
            EstimateFramePose(c, k, kf.pose, est, &conv, count_residuals != 0, &cnt, &sum, count_residuals == 2 ? &dcnt : nullptr);
        res->n_count += cnt;
        res->n_depth_count += dcnt;
        res->cost += sum;
        hm_se3_inverse(kf.pose, ftg);
        hm_se3_mul(ftg, est, diff);
        hm_se3_log(diff, lg);
        const int moved = !hm_is_scale1_pose_converged(lg);
        std::memcpy(kf.pose, est, sizeof(est));
        if (moved) kf.activation = 0;
        else { kf.activation = 2; ++num_converged; }
      }
    }
    cudaEventRecord(c->ev[3], s);
    {   // direct_ba_alternating.cc:584-624 (flags gated like direct_ba.cc:427-434)
      const bool od = o->optimize_depth_intrinsics && c->cfg.use_depth_residuals;
      const bool oc = o->optimize_color_intrinsics && c->cfg.use_descriptor_residuals;
      if (od || oc) OptimizeIntrinsics(c, od, oc);
    }
    cudaEventSynchronize(c->ev[3]);
    cudaEventElapsedTime(&res->ms_surfel_activation, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&res->ms_geometry_optimization, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&res->ms_pose_optimization, c->ev[2], c->ev[3]);
    if (iteration >= o->min_iterations - 1 && (num_converged == K || !o->optimize_poses)) {
      res->converged = 1;
      break;
    }
    DetermineCovisibleActive(c);
  }
  if (o->end_tasks) res->surfels_deleted = ref_end_tasks(c);
  res->surfels_size = c->surfels_size;
  res->kernel_launches = c->launches - launches_before;
}

struct ref_pcg_options {
  int optimize_poses, optimize_geometry, optimize_depth_intrinsics, optimize_color_intrinsics;
  int min_iterations, max_iterations, max_inner_iterations, gauge_keyframe;
  int end_tasks;
};
struct ref_pcg_result {
  int iterations_done, converged, inner_iterations_total;
  float last_r_norm, ms_pcg;
  unsigned long long kernel_launches;
  unsigned int surfels_deleted, surfels_size;
};

// DirectBA::BundleAdjustmentPCG (direct_ba_pcg.cc:43-819) without the surfel lifecycle branches; the gauge keyframe is an
// argument (the reference draws rand() % K, :324).  Host-side restatement around the reference's own PCG kernels.
void ref_bundle_adjust_pcg(ref_context* c, const ref_pcg_options* o, ref_pcg_result* res) {
  std::memset(res, 0, sizeof(*res));
  cudaStream_t s = c->stream;
  const int K = static_cast<int>(c->kfs.size());
  const u32 N = c->surfels_size;
  const u32 P = static_cast<u32>(c->cf_w) * c->cf_h;
  const bool use_depth = c->cfg.use_depth_residuals != 0, use_desc = c->cfg.use_descriptor_residuals != 0;
  const bool od = o->optimize_depth_intrinsics && use_depth, oc = o->optimize_color_intrinsics && use_desc;
  const bool op = o->optimize_poses != 0, og = o->optimize_geometry != 0;
  const unsigned long long launches_before = c->launches;
  constexpr u32 kInvalid = 0xffffffffu;
  if (!c->pcg_scalars) cudaMalloc(&c->pcg_scalars, sizeof(float) * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int iteration = 0; iteration < o->max_iterations; ++iteration) {
    ++res->iterations_done;
    cudaMemsetAsync(c->active, kSurfelActiveFlag, N, s);
    if (og) UpdateSurfelNormals(c);
    u32 cur = 0;
    if (op) cur += 6 * (K - 1);
    u32 surfel_start = kInvalid, depth_start = kInvalid, a_index = kInvalid, color_start = kInvalid;
    if (og) { surfel_start = cur; cur += (use_desc ? 3 : 1) * N; }
    if (od) { depth_start = cur; cur += 5 + P; a_index = depth_start + 4; }
    if (oc) { color_start = cur; cur += 4; }
    const u32 unknown_count = cur;
    if (unknown_count > c->pcg_capacity) {
      for (float*& v : c->pcg) { cudaFree(v); cudaMalloc(&v, sizeof(float) * unknown_count); }
      c->pcg_capacity = unknown_count;
    }
    auto vec = [&](float* ptr) { return CUDABuffer_<float>(ptr, 1, static_cast<int>(c->pcg_capacity), sizeof(float) * c->pcg_capacity); };
    CUDABuffer_<float> pcg_r = vec(c->pcg[0]), pcg_M = vec(c->pcg[1]), pcg_delta = vec(c->pcg[2]), pcg_g = vec(c->pcg[3]),
                       pcg_p = vec(c->pcg[4]);
    float* sc = c->pcg_scalars;
    CUDABuffer_<float> alpha_n(sc + 0, 1, 1, sizeof(float)), alpha_d(sc + 1, 1, 1, sizeof(float)), beta_n(sc + 2, 1, 1, sizeof(float));
    cudaEventRecord(e0, s);
    cudaMemsetAsync(c->pcg[0], 0, sizeof(float) * unknown_count, s);
    cudaMemsetAsync(c->pcg[1], 0, sizeof(float) * unknown_count, s);
    const int gauge = o->gauge_keyframe;
    auto pose_index = [&](int id) -> u32 {
      if (id == gauge) return kInvalid;
      return static_cast<u32>(6 * (id < gauge ? id : id - 1));
    };
    const PixelCenterUnprojector unproj = CenterUnprojector(c->cfg.depth_K);
    for (int k = 0; k < K; ++k) {
      const RefKeyframe& kf = c->kfs[k];
      PCGInitCUDA(s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), N), DepthToColor(c->cfg), unproj, CornerProjector(c->cfg.color_K),
                  kf.tex, pose_index(k), surfel_start, (k == gauge) ? false : op, og, use_depth, use_desc, od, oc, depth_start,
                  color_start, &pcg_r, &pcg_M, N);
      ++c->launches;
    }
    PCGInit2CUDA(s, unknown_count, a_index, c->a, pcg_r, pcg_M, &pcg_delta, &pcg_g, &pcg_p, &alpha_n);
    ++c->launches;
    float prev_r_norm = INFINITY;
    int without_improvement = 0;
    for (int step = 0; step < o->max_inner_iterations; ++step) {
      cudaMemsetAsync(alpha_d.address(), 0, sizeof(float), s);
      if (step > 0) {
        std::swap(alpha_n, beta_n);
        cudaMemsetAsync(c->pcg[3], 0, sizeof(float) * unknown_count, s);
      }
      for (int k = 0; k < K; ++k) {
        const RefKeyframe& kf = c->kfs[k];
        PCGStep1CUDA(s, unknown_count, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), N), DepthToColor(c->cfg), unproj,
                     CornerProjector(c->cfg.color_K), kf.tex, pose_index(k), surfel_start, (k == gauge) ? false : op, og, use_depth,
                     use_desc, od, oc, depth_start, a_index, color_start, &pcg_p, &pcg_g, &alpha_d, N);
        c->launches += 2;
      }
      PCGStep2CUDA(s, unknown_count, a_index, pcg_r, pcg_M, &pcg_delta, &pcg_g, &pcg_p, &alpha_n, &alpha_d, &beta_n);
      ++c->launches;
      float r_norm;
      cudaMemcpyAsync(&r_norm, beta_n.address(), sizeof(float), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      r_norm = std::sqrt(r_norm);
      ++res->inner_iterations_total;
      res->last_r_norm = r_norm;
      if (r_norm < prev_r_norm - 1e-3) {
        without_improvement = 0;
      } else {
        ++without_improvement;
        if (without_improvement >= 3) break;
      }
      prev_r_norm = r_norm;
      if (step < o->max_inner_iterations - 1) {
        PCGStep3CUDA(s, unknown_count, &pcg_g, &pcg_p, &alpha_n, &beta_n);
        ++c->launches;
      }
    }
    cudaEventRecord(e1, s);
    int num_converged = 0;
    if (op) {
      std::vector<float> delta(6 * (K - 1));
      cudaMemcpyAsync(delta.data(), c->pcg[2], sizeof(float) * delta.size(), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      for (int k = 0; k < K; ++k) {
        if (k == gauge) { ++num_converged; continue; }
        float d7[7], np[7], lg[6];
        hm_se3_exp(delta.data() + pose_index(k), d7);
        hm_se3_mul(c->kfs[k].pose, d7, np);
        std::memcpy(c->kfs[k].pose, np, sizeof(np));
        hm_se3_log(d7, lg);
        if (hm_is_scale1_pose_converged(lg)) ++num_converged;
      }
    }
    if (og) {
      CUDABuffer_<float> sb = SurfelBuf(c);
      UpdateSurfelsFromPCGDeltaCUDA(s, N, &sb, use_desc, surfel_start, pcg_delta);
      ++c->launches;
    }
    if (od) {
      float b5[5];
      cudaMemcpyAsync(b5, c->pcg[2] + depth_start, sizeof(b5), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      const double old_fx_inv = 1. / c->cfg.depth_K[0], old_fy_inv = 1. / c->cfg.depth_K[1];
      const double old_cx_inv = -(c->cfg.depth_K[2] - 0.5) * old_fx_inv, old_cy_inv = -(c->cfg.depth_K[3] - 0.5) * old_fy_inv;
      const double nfx = 1. / (old_fx_inv + b5[0]), nfy = 1. / (old_fy_inv + b5[1]);
      const double ncx = -(nfx * (old_cx_inv + b5[2])) + 0.5, ncy = -(nfy * (old_cy_inv + b5[3])) + 0.5;
      c->cfg.depth_K[0] = static_cast<float>(nfx); c->cfg.depth_K[1] = static_cast<float>(nfy);
      c->cfg.depth_K[2] = static_cast<float>(ncx); c->cfg.depth_K[3] = static_cast<float>(ncy);
      c->a += b5[4];
      CUDABuffer_<float> cf(c->cfactor, c->cf_h, c->cf_w, c->cfactor_pitch);
      UpdateCFactorsFromPCGDeltaCUDA(s, &cf, depth_start + 5, pcg_delta);
      ++c->launches;
    }
    if (oc) {
      float b4[4];
      cudaMemcpyAsync(b4, c->pcg[2] + color_start, sizeof(b4), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      for (int i = 0; i < 4; ++i) c->cfg.color_K[i] = static_cast<float>(c->cfg.color_K[i] + b4[i]);
    }
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&res->ms_pcg, e0, e1);
    if (iteration >= o->min_iterations - 1 && (num_converged == K || !op)) {
      res->converged = 1;
      break;
    }
  }
  cudaStreamSynchronize(s);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (o->end_tasks) res->surfels_deleted = ref_end_tasks(c);
  res->surfels_size = c->surfels_size;
  res->kernel_launches = c->launches - launches_before;
}

// Parity hook: PCGInit (all keyframes) -> r, M; PCGInit2 -> p, alpha_n; one PCGStep1 sweep -> g, alpha_d.  No state changes.
unsigned int ref_pcg_debug(ref_context* c, const ref_pcg_options* o, float* out_r, float* out_M, float* out_p, float* out_g,
                           float* out_scalars) {
  cudaStream_t s = c->stream;
  const int K = static_cast<int>(c->kfs.size());
  const u32 N = c->surfels_size;
  const u32 P = static_cast<u32>(c->cf_w) * c->cf_h;
  const bool use_depth = c->cfg.use_depth_residuals != 0, use_desc = c->cfg.use_descriptor_residuals != 0;
  const bool od = o->optimize_depth_intrinsics && use_depth, oc = o->optimize_color_intrinsics && use_desc;
  const bool op = o->optimize_poses != 0, og = o->optimize_geometry != 0;
  constexpr u32 kInvalid = 0xffffffffu;
  if (!c->pcg_scalars) cudaMalloc(&c->pcg_scalars, sizeof(float) * 4);
  u32 cur = 0;
  if (op) cur += 6 * (K - 1);
  u32 surfel_start = kInvalid, depth_start = kInvalid, a_index = kInvalid, color_start = kInvalid;
  if (og) { surfel_start = cur; cur += (use_desc ? 3 : 1) * N; }
  if (od) { depth_start = cur; cur += 5 + P; a_index = depth_start + 4; }
  if (oc) { color_start = cur; cur += 4; }
  const u32 U = cur;
  if (!out_r) return U;
  if (U > c->pcg_capacity) {
    for (float*& v : c->pcg) { cudaFree(v); cudaMalloc(&v, sizeof(float) * U); }
    c->pcg_capacity = U;
  }
  auto vec = [&](float* ptr) { return CUDABuffer_<float>(ptr, 1, static_cast<int>(c->pcg_capacity), sizeof(float) * c->pcg_capacity); };
  CUDABuffer_<float> pcg_r = vec(c->pcg[0]), pcg_M = vec(c->pcg[1]), pcg_delta = vec(c->pcg[2]), pcg_g = vec(c->pcg[3]), pcg_p = vec(c->pcg[4]);
  float* sc = c->pcg_scalars;
  CUDABuffer_<float> alpha_n(sc + 0, 1, 1, sizeof(float)), alpha_d(sc + 1, 1, 1, sizeof(float));
  cudaMemsetAsync(c->pcg[0], 0, sizeof(float) * U, s);
  cudaMemsetAsync(c->pcg[1], 0, sizeof(float) * U, s);
  const int gauge = o->gauge_keyframe;
  auto pose_index = [&](int id) -> u32 { return id == gauge ? kInvalid : static_cast<u32>(6 * (id < gauge ? id : id - 1)); };
  const PixelCenterUnprojector unproj = CenterUnprojector(c->cfg.depth_K);
  for (int k = 0; k < K; ++k) {
    const RefKeyframe& kf = c->kfs[k];
    PCGInitCUDA(s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), N), DepthToColor(c->cfg), unproj, CornerProjector(c->cfg.color_K),
                kf.tex, pose_index(k), surfel_start, (k == gauge) ? false : op, og, use_depth, use_desc, od, oc, depth_start, color_start,
                &pcg_r, &pcg_M, N);
  }
  cudaMemcpyAsync(out_r, c->pcg[0], sizeof(float) * U, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(out_M, c->pcg[1], sizeof(float) * U, cudaMemcpyDeviceToHost, s);
  PCGInit2CUDA(s, U, a_index, c->a, pcg_r, pcg_M, &pcg_delta, &pcg_g, &pcg_p, &alpha_n);
  cudaMemcpyAsync(out_p, c->pcg[4], sizeof(float) * U, cudaMemcpyDeviceToHost, s);
  cudaMemsetAsync(alpha_d.address(), 0, sizeof(float), s);
  for (int k = 0; k < K; ++k) {
    const RefKeyframe& kf = c->kfs[k];
    PCGStep1CUDA(s, U, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), N), DepthToColor(c->cfg), unproj, CornerProjector(c->cfg.color_K),
                 kf.tex, pose_index(k), surfel_start, (k == gauge) ? false : op, og, use_depth, use_desc, od, oc, depth_start, a_index,
                 color_start, &pcg_p, &pcg_g, &alpha_d, N);
  }
  cudaMemcpyAsync(out_g, c->pcg[3], sizeof(float) * U, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(out_scalars, sc, sizeof(float) * 2, cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  return U;
}

// DirectBA::PerformBASchemeEndTasks (direct_ba.cc:566-653) with do_surfel_updates = false: DeleteSurfelsAndUpdateRadiiCUDA
// (kernel_delete_surfels.cc:40-98) + CompactSurfelsCUDA, driving the reference's own kernels.  Returns the deleted count.
unsigned int ref_end_tasks(ref_context* c) {
  if (c->surfels_size == 0) return 0;
  cudaStream_t s = c->stream;
  const size_t K = c->kfs.size();
  const int min_obs = (K < 10) ? ((K < 5) ? c->min_observation_count[0] : c->min_observation_count[1]) : c->min_observation_count[2];
  if (!c->deleted_count) cudaMalloc(&c->deleted_count, sizeof(u32));
  CallResetSurfelAccumForSurfelDeletionAndRadiusUpdateCUDAKernel(s, c->surfels_size, SurfelBuf(c), true);
  cudaMemsetAsync(c->deleted_count, 0, sizeof(u32), s);
  c->launches += 1;
  for (const RefKeyframe& kf : c->kfs) {
    CallCountObservationsAndFreeSpaceViolationsCUDAKernel(s, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size),
                                                          CUDABuffer_<u16>(kf.radius, c->cfg.depth_h, c->cfg.depth_w, kf.radius_pitch), true);
    ++c->launches;
  }
  CUDABuffer_<u32> deleted_buf(c->deleted_count, 1, 1, sizeof(u32));
  CallMarkDeletedSurfelsCUDAKernel(s, min_obs, c->surfels_size, SurfelBuf(c), &deleted_buf, true);
  ++c->launches;
  u32 deleted = 0;
  cudaMemcpyAsync(&deleted, c->deleted_count, sizeof(u32), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  u32 surfel_count = c->surfels_size - deleted;
  CUDABuffer_<float> sb = SurfelBuf(c);
  CompactSurfelsCUDA(s, &c->free_spots_temp, &c->free_spots_temp_bytes, surfel_count, &c->surfels_size, &sb, nullptr);
  cudaStreamSynchronize(s);
  return deleted;
}
unsigned int ref_surfels_size(ref_context* c) { return c->surfels_size; }
void ref_set_min_observation_counts(ref_context* c, int b1, int b2, int m) {
  c->min_observation_count[0] = b1; c->min_observation_count[1] = b2; c->min_observation_count[2] = m;
}

namespace {
void EnsureLifecycleBuffers(ref_context* c) {
  if (c->sup[0]) return;
  for (int i = 0; i < 3; ++i) cudaMallocPitch(reinterpret_cast<void**>(&c->sup[i]), &c->sup_pitch, sizeof(u32) * c->cf_w, c->cf_h);
  cudaMalloc(&c->new_flag, static_cast<size_t>(c->cfg.depth_w) * c->cfg.depth_h);
  cudaMalloc(&c->new_indices, sizeof(u32) * static_cast<size_t>(c->cfg.depth_w) * c->cfg.depth_h);
  if (!c->deleted_count) cudaMalloc(&c->deleted_count, sizeof(u32));
}
SupportingSurfelBuffers SupBuffers(ref_context* c) {
  SupportingSurfelBuffers b;
  for (int i = 0; i < 3; ++i) b.b[i] = CUDABuffer_<u32>(c->sup[i], c->cf_h, c->cf_w, c->sup_pitch);
  return b;
}
int MinObservationCount(ref_context* c) {
  const size_t K = c->kfs.size();
  return (K < 10) ? ((K < 5) ? c->min_observation_count[0] : c->min_observation_count[1]) : c->min_observation_count[2];
}
// DetermineSupportingSurfelsCUDAImpl (kernel_supporting_surfels.cc:40-118)
unsigned int DetermineSupportingSurfels(ref_context* c, const RefKeyframe& kf, bool merge) {
  cudaStream_t s = c->stream;
  for (int i = 0; i < 3; ++i) cudaMemset2DAsync(c->sup[i], c->sup_pitch, 0xff, sizeof(u32) * c->cf_w, c->cf_h, s);
  if (c->surfels_size == 0) return 0;
  CUDABuffer_<u32> deleted_buf(c->deleted_count, 1, 1, sizeof(u32));
  float cmd2 = 0, cos_thr = 0;
  if (merge) {
    cudaMemsetAsync(c->deleted_count, 0, sizeof(u32), s);
    cmd2 = static_cast<float>(c->cfg.cell) * c->cfg.cell * c->surfel_merge_dist_factor * c->surfel_merge_dist_factor;
    cos_thr = cos_normal_compatibility_threshold;
  }
  CallDetermineSupportingSurfelsCUDAKernel(s, merge, cmd2, cos_thr, MakeProjection(c, kf, MakeFrameTGlobal(kf.pose), c->surfels_size),
                                           SupBuffers(c), merge ? deleted_buf : CUDABuffer_<u32>());
  ++c->launches;
  if (!merge) return 0;
  u32 deleted = 0;
  cudaMemcpyAsync(&deleted, c->deleted_count, sizeof(u32), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  return deleted;
}
}  // namespace

// DirectBA::CreateSurfelsForKeyframe (direct_ba.cc:340-405) + CreateSurfelsForKeyframeCUDA (kernel_create_surfels.cc:40-183),
// driving the reference's own kernels.  Returns the number of surfels created.
unsigned int ref_create_surfels_for_keyframe(ref_context* c, int k, int filter_new_surfels) {
  EnsureLifecycleBuffers(c);
  cudaStream_t s = c->stream;
  const RefKeyframe& kf = c->kfs[k];
  const int w = c->cfg.depth_w, h = c->cfg.depth_h;
  DetermineSupportingSurfels(c, kf, false);
  CUDABuffer_<u16> depth(kf.depth, h, w, kf.depth_pitch), normals(kf.normals, h, w, kf.normals_pitch), radius(kf.radius, h, w, kf.radius_pitch);
  CUDABuffer_<uchar4> color(kf.color, c->cfg.color_h, c->cfg.color_w, kf.color_pitch);
  CUDABuffer_<u8> flag(c->new_flag, 1, w * h, static_cast<size_t>(w) * h);
  CUDABuffer_<u32> indices(c->new_indices, 1, w * h, sizeof(u32) * static_cast<size_t>(w) * h);
  CallCreateSurfelsForKeyframeCUDASerializingKernel(s, c->cfg.cell, depth, color, CUDABuffer_<u32>(c->sup[0], c->cf_h, c->cf_w, c->sup_pitch), flag);
  u32 new_count = CreateSurfelsForKeyframeCUDA_CountNewSurfels(s, w * h, &c->new_temp, &c->new_temp_bytes, &flag, &indices);
  c->launches += 2;
  if (new_count == 0) return 0;
  const PixelCenterUnprojector unproj = CenterUnprojector(c->cfg.depth_K);
  if (filter_new_surfels) {
    u16* obs = reinterpret_cast<u16*>(reinterpret_cast<u8*>(c->surfels) + kSurfelAccum0 * c->surfel_pitch);
    u16* viol = reinterpret_cast<u16*>(reinterpret_cast<u8*>(c->surfels) + kSurfelAccum1 * c->surfel_pitch);
    u32* list = reinterpret_cast<u32*>(reinterpret_cast<u8*>(c->surfels) + kSurfelAccum2 * c->surfel_pitch);
    CallWriteNewSurfelIndexAndInitializeObservationsCUDAKernel(s, w * h, flag, indices, obs, viol, list);
    for (int o : kf.covis) {
      const RefKeyframe& other = c->kfs[o];
      float inv[7], rel[7], M[12];
      hm_se3_inverse(other.pose, inv);
      hm_se3_mul(inv, kf.pose, rel);
      hm_se3_matrix3x4(rel, M);
      CUDAMatrix3x4 covis_T_frame;
      covis_T_frame.row0 = make_float4(M[0], M[1], M[2], M[3]);
      covis_T_frame.row1 = make_float4(M[4], M[5], M[6], M[7]);
      covis_T_frame.row2 = make_float4(M[8], M[9], M[10], M[11]);
      CallCountObservationsForNewSurfelsCUDAKernel(s, new_count, list, obs, viol, MakeDepthParams(c), unproj, depth, normals, covis_T_frame,
                                                   CornerProjector(c->cfg.depth_K), CUDABuffer_<u16>(other.depth, h, w, other.depth_pitch),
                                                   CUDABuffer_<u16>(other.normals, h, w, other.normals_pitch));
      ++c->launches;
    }
    CallFilterNewSurfelsCUDAKernel(s, static_cast<u16>(MinObservationCount(c)), new_count, list, obs, viol, flag);
    new_count = CreateSurfelsForKeyframeCUDA_CountNewSurfels(s, w * h, &c->new_temp, &c->new_temp_bytes, &flag, &indices);
    c->launches += 3;
    if (new_count == 0) return 0;
  }
  if (c->surfels_size + new_count > c->max_surfels) return 0;   // kernel_create_surfels.cc:163-166
  float gm[12];
  hm_se3_matrix3x4(kf.pose, gm);
  CUDAMatrix3x4 global_T_frame;
  global_T_frame.row0 = make_float4(gm[0], gm[1], gm[2], gm[3]);
  global_T_frame.row1 = make_float4(gm[4], gm[5], gm[6], gm[7]);
  global_T_frame.row2 = make_float4(gm[8], gm[9], gm[10], gm[11]);
  CallCreateSurfelsForKeyframeCUDACreationAppendKernel(s, unproj, DepthToColor(c->cfg), CornerProjector(c->cfg.color_K), global_T_frame,
                                                       MakeFrameTGlobal(kf.pose), MakeDepthParams(c), depth, normals, radius, kf.tex, flag, indices,
                                                       c->surfels_size, SurfelBuf(c));
  ++c->launches;
  cudaStreamSynchronize(s);
  c->surfels_size += new_count;
  return new_count;
}

// DetermineSupportingSurfelsAndMergeSurfelsCUDA for keyframe k; returns the number of merged (deleted) surfels
unsigned int ref_merge_surfels_for_keyframe(ref_context* c, int k) {
  EnsureLifecycleBuffers(c);
  return DetermineSupportingSurfels(c, c->kfs[k], true);
}

// CompactSurfelsCUDA with the active flags (direct_ba_alternating.cc:530)
unsigned int ref_compact_surfels(ref_context* c, unsigned int free_count, int with_active) {
  CUDABuffer_<float> sb = SurfelBuf(c);
  CUDABuffer_<u8> ab = ActiveBuf(c);
  CompactSurfelsCUDA(c->stream, &c->free_spots_temp, &c->free_spots_temp_bytes, c->surfels_size - free_count, &c->surfels_size, &sb,
                     with_active ? &ab : nullptr);
  cudaStreamSynchronize(c->stream);
  return c->surfels_size;
}
// BadSlam::PreprocessFrame (bad_slam.cc:692-765) followed by the ComputeMinMaxDepthCUDA of keyframe creation (bad_slam.cc:978),
// with the reference's own kernels and host launchers (cuda_depth_processing.cu, cuda_image_processing.cu).  Dense host images
// in and out; the radius buffer starts zeroed (the reference does not write the radius of pixels it drops).  Returns the
// number of kernel launches (5), or -1 on a CUDA error.
int ref_preprocess_frame(ref_context* c, float sigma_xy, float sigma_inv_depth, float radius_factor, float max_depth_m,
                         const unsigned short* raw_depth, const unsigned char* rgb, unsigned short* out_depth,
                         unsigned short* out_normals, unsigned short* out_radius, unsigned char* out_rgba, float* min_depth,
                         float* max_depth) {
  cudaStream_t s = c->stream;
  const int w = c->cfg.depth_w, h = c->cfg.depth_h, cw = c->cfg.color_w, ch = c->cfg.color_h;
  u16 *d_raw = nullptr, *d_a = nullptr, *d_b = nullptr, *d_normals = nullptr, *d_radius = nullptr;
  size_t p_raw = 0, p_a = 0, p_b = 0, p_normals = 0, p_radius = 0, p_rgb = 0, p_rgba = 0;
  uchar3* d_rgb = nullptr;
  uchar4* d_rgba = nullptr;
  float *d_init = nullptr, *d_result = nullptr;
  bool ok = true;
  auto alloc2d = [&](void** ptr, size_t* pitch, size_t row_bytes, int rows) { ok = ok && cudaMallocPitch(ptr, pitch, row_bytes, rows) == cudaSuccess; };
  alloc2d(reinterpret_cast<void**>(&d_raw), &p_raw, sizeof(u16) * w, h);
  alloc2d(reinterpret_cast<void**>(&d_a), &p_a, sizeof(u16) * w, h);
  alloc2d(reinterpret_cast<void**>(&d_b), &p_b, sizeof(u16) * w, h);
  alloc2d(reinterpret_cast<void**>(&d_normals), &p_normals, sizeof(u16) * w, h);
  alloc2d(reinterpret_cast<void**>(&d_radius), &p_radius, sizeof(u16) * w, h);
  if (rgb) {
    alloc2d(reinterpret_cast<void**>(&d_rgb), &p_rgb, sizeof(uchar3) * cw, ch);
    alloc2d(reinterpret_cast<void**>(&d_rgba), &p_rgba, sizeof(uchar4) * cw, ch);
  }
  ok = ok && cudaMalloc(&d_init, 2 * sizeof(float)) == cudaSuccess && cudaMalloc(&d_result, 2 * sizeof(float)) == cudaSuccess;
  int launches = -1;
  if (ok) {
    const float init[2] = {std::numeric_limits<float>::infinity(), 0.f};   // cuda_depth_processing.cc:41
    cudaMemcpyAsync(d_init, init, sizeof(init), cudaMemcpyHostToDevice, s);
    cudaMemcpy2DAsync(d_raw, p_raw, raw_depth, sizeof(u16) * w, sizeof(u16) * w, h, cudaMemcpyHostToDevice, s);
    cudaMemset2DAsync(d_radius, p_radius, 0, sizeof(u16) * w, h, s);
    CUDABuffer_<u16> raw_buf(d_raw, h, w, p_raw), a_buf(d_a, h, w, p_a), b_buf(d_b, h, w, p_b);
    CUDABuffer_<u16> normals_buf(d_normals, h, w, p_normals), radius_buf(d_radius, h, w, p_radius);
    launches = 0;
    if (rgb) {
      cudaMemcpy2DAsync(d_rgb, p_rgb, rgb, sizeof(uchar3) * cw, sizeof(uchar3) * cw, ch, cudaMemcpyHostToDevice, s);
      CUDABuffer_<uchar3> rgb_buf(d_rgb, ch, cw, p_rgb);
      CUDABuffer_<uchar4> rgba_buf(d_rgba, ch, cw, p_rgba);
      ComputeBrightnessCUDA(s, rgb_buf, &rgba_buf);                                                       // bad_slam.cc:692
      ++launches;
    }
    BilateralFilteringAndDepthCutoffCUDA(s, sigma_xy, sigma_inv_depth, radius_factor,
                                         max_depth_m / c->cfg.raw_to_float_depth,                         // float -> u16 (bad_slam.cc:703)
                                         c->cfg.raw_to_float_depth, raw_buf, &a_buf);                     // bad_slam.cc:698
    const PixelCenterUnprojector unproj = CenterUnprojector(c->cfg.depth_K);
    ComputeNormalsCUDA(s, unproj, MakeDepthParams(c), a_buf, &b_buf, &normals_buf);                       // bad_slam.cc:716
    ComputePointRadiiAndRemoveIsolatedPixelsCUDA(s, unproj, c->cfg.raw_to_float_depth, b_buf, &radius_buf, &a_buf);   // bad_slam.cc:754
    CUDABuffer_<float> init_buf(d_init, 1, 2, 2 * sizeof(float)), result_buf(d_result, 1, 2, 2 * sizeof(float));
    ComputeMinMaxDepthCUDA(s, a_buf, c->cfg.raw_to_float_depth, init_buf, &result_buf, min_depth, max_depth);        // bad_slam.cc:978
    launches += 4;
    c->launches += launches;
    cudaMemcpy2DAsync(out_depth, sizeof(u16) * w, d_a, p_a, sizeof(u16) * w, h, cudaMemcpyDeviceToHost, s);
    cudaMemcpy2DAsync(out_normals, sizeof(u16) * w, d_normals, p_normals, sizeof(u16) * w, h, cudaMemcpyDeviceToHost, s);
    cudaMemcpy2DAsync(out_radius, sizeof(u16) * w, d_radius, p_radius, sizeof(u16) * w, h, cudaMemcpyDeviceToHost, s);
    if (rgb) cudaMemcpy2DAsync(out_rgba, sizeof(uchar4) * cw, d_rgba, p_rgba, sizeof(uchar4) * cw, ch, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) launches = -1;
  }
  cudaFree(d_raw); cudaFree(d_a); cudaFree(d_b); cudaFree(d_normals); cudaFree(d_radius); cudaFree(d_rgb); cudaFree(d_rgba);
  cudaFree(d_init); cudaFree(d_result);
  return launches;
}

void ref_set_surfels_size(ref_context* c, unsigned int n) { c->surfels_size = n; }

// Device-side snapshot / restore of the mutable state (surfel data rows, poses, activations) for benchmarking
// repeated steps from the same starting point without host traffic.
static float* g_snapshot = nullptr;
static size_t g_snapshot_pitch = 0;
static std::vector<RefKeyframe> g_snapshot_kfs;
static unsigned int g_snapshot_size = 0;
void ref_snapshot(ref_context* c) {
  if (!g_snapshot) cudaMallocPitch(reinterpret_cast<void**>(&g_snapshot), &g_snapshot_pitch, static_cast<size_t>(c->max_surfels) * 4, 8);
  cudaMemcpy2D(g_snapshot, g_snapshot_pitch, c->surfels, c->surfel_pitch, static_cast<size_t>(c->surfels_size) * 4, 8, cudaMemcpyDeviceToDevice);
  g_snapshot_kfs = c->kfs;
  g_snapshot_size = c->surfels_size;
}
void ref_restore(ref_context* c) {
  c->surfels_size = g_snapshot_size;
  cudaMemcpy2DAsync(c->surfels, c->surfel_pitch, g_snapshot, g_snapshot_pitch, static_cast<size_t>(c->surfels_size) * 4, 8,
                    cudaMemcpyDeviceToDevice, c->stream);
  for (size_t k = 0; k < c->kfs.size(); ++k) {
    std::memcpy(c->kfs[k].pose, g_snapshot_kfs[k].pose, sizeof(float) * 7);
    c->kfs[k].activation = g_snapshot_kfs[k].activation;
  }
}
void ref_sync(ref_context* c) { cudaStreamSynchronize(c->stream); }

const char* ref_last_cuda_error(void) { return cudaGetErrorString(cudaGetLastError()); }

}  // extern "C"

// ---- image-pair odometry: BadSlam::RunOdometry (bad_slam.cc:829-950) + TrackFramePairwise (pairwise_frame_tracking.cc:153-678)
// restated on the reference's own kernels (kernel_downsample.cu, cuda_image_processing.cu, kernel_opt_pose.cu:422-1340).
namespace {

cudaTextureObject_t MakeU8Texture(u8* data, size_t pitch, int w, int h) {   // CUDABuffer::CreateTextureObject as called in pairwise_frame_tracking.cc:57-79
  cudaResourceDesc res;
  std::memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = data;
  res.res.pitch2D.desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
  res.res.pitch2D.width = w;
  res.res.pitch2D.height = h;
  res.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td;
  std::memset(&td, 0, sizeof(td));
  td.addressMode[0] = cudaAddressModeClamp;
  td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  cudaCreateTextureObject(&t, &res, &td, nullptr);
  return t;
}

void EnsureOdoBuffers(ref_context* c, int num_scales) {
  if (c->odo_scales >= num_scales) return;
  for (int f = 0; f < 2; ++f) {
    if (!c->odo_gradmag[f]) {
      cudaMallocPitch(reinterpret_cast<void**>(&c->odo_gradmag[f]), &c->odo_gradmag_pitch[f], c->cfg.color_w, c->cfg.color_h);
      c->odo_gradmag_tex[f] = MakeU8Texture(c->odo_gradmag[f], c->odo_gradmag_pitch[f], c->cfg.color_w, c->cfg.color_h);
    }
    for (int sc = c->odo_scales; sc < num_scales; ++sc) {
      ref_context::OdoImage& im = c->odo[f][sc];
      im.w = static_cast<int>(c->cfg.depth_w / pow(2, sc));
      im.h = static_cast<int>(c->cfg.depth_h / pow(2, sc));
      cudaMallocPitch(reinterpret_cast<void**>(&im.depth), &im.depth_pitch, sizeof(float) * im.w, im.h);
      if (sc >= 1) {
        cudaMallocPitch(reinterpret_cast<void**>(&im.normals), &im.normals_pitch, sizeof(u16) * im.w, im.h);
        im.owns_normals = true;
      }
      cudaMallocPitch(reinterpret_cast<void**>(&im.color), &im.color_pitch, im.w, im.h);
      im.tex = MakeU8Texture(im.color, im.color_pitch, im.w, im.h);
    }
  }
  c->odo_scales = num_scales;
}

CUDABuffer_<float> DepthBuf(const ref_context::OdoImage& im) { return CUDABuffer_<float>(im.depth, im.h, im.w, im.depth_pitch); }
CUDABuffer_<u16> NormalsBuf(const ref_context::OdoImage& im) { return CUDABuffer_<u16>(im.normals, im.h, im.w, im.normals_pitch); }
CUDABuffer_<u8> ColorBuf(const ref_context::OdoImage& im) { return CUDABuffer_<u8>(im.color, im.h, im.w, im.color_pitch); }

// PinholeCamera4f::Scaled (libvis camera.h:1086-1097,1696-1705): parameters * factor, width = factor * width + 0.5
void ScaledCamera(const float K[4], int w, int h, double factor, float out_K[4], int* out_w, int* out_h) {
  const float f = static_cast<float>(factor);
  for (int i = 0; i < 4; ++i) out_K[i] = K[i] * f;
  *out_w = static_cast<int>(factor * w + 0.5f);
  *out_h = static_cast<int>(factor * h + 0.5f);
}

struct OdoCameras {
  float dK[4], cK[4];
  int cw, ch;
};
OdoCameras ScaleCameras(ref_context* c, int scale) {   // pairwise_frame_tracking.cc:409-417
  OdoCameras r;
  const float scaling_factor = pow(2, scale);
  int dw, dh;
  ScaledCamera(c->cfg.color_K, c->cfg.color_w, c->cfg.color_h, (c->cfg.depth_w == c->cfg.color_w) ? (1.f / scaling_factor) : (2.f / scaling_factor),
               r.cK, &r.cw, &r.ch);
  ScaledCamera(c->cfg.depth_K, c->cfg.depth_w, c->cfg.depth_h, 1.f / scaling_factor, r.dK, &dw, &dh);
  return r;
}
DepthToColorPixelCorner DepthToColorScaled(const OdoCameras& k) {   // surfel_projection.h:105-124
  DepthToColorPixelCorner r;
  r.width = k.cw;
  r.height = k.ch;
  r.fx = k.cK[0] / k.dK[0];
  r.cx = -1 * k.cK[0] * k.dK[2] / k.dK[0] + k.cK[2];
  r.fy = k.cK[1] / k.dK[1];
  r.cy = -1 * k.cK[1] * k.dK[3] / k.dK[1] + k.cK[3];
  return r;
}

CUDAMatrix3x4 FrameTBase(const float base_T_frame[7]) { return MakeFrameTGlobal(base_T_frame); }   // CUDAMatrix3x4(base_T_frame.inverse().matrix3x4())

// AccumulatePoseEstimationCoeffsFromImagesCUDA, kernel_opt_pose.cc:99-191
void OdoAccumulate(ref_context* c, int scale, const float base_T_frame[7], bool use_gradmag, bool debug, u32* count, float* sum, float H[21], float b[6]) {
  cudaStream_t s = c->stream;
  const OdoCameras k = ScaleCameras(c, scale);
  const float threshold_factor = pow(2, scale);
  if (debug) {
    cudaMemsetAsync(c->residual_count, 0, sizeof(u32), s);
    cudaMemsetAsync(c->residual_sum, 0, sizeof(float), s);
    c->launches += 2;
  }
  cudaMemsetAsync(c->H, 0, sizeof(float) * 21, s);
  cudaMemsetAsync(c->b, 0, sizeof(float) * 6, s);
  c->launches += 2;
  const ref_context::OdoImage& base = c->odo[0][scale];
  const ref_context::OdoImage& trk = c->odo[1][scale];
  CUDABuffer_<u32> cnt_buf(c->residual_count, 1, 1, sizeof(u32));
  CUDABuffer_<float> sum_buf(c->residual_sum, 1, 1, sizeof(float)), H_buf(c->H, 1, 21, sizeof(float) * 21), b_buf(c->b, 1, 6, sizeof(float) * 6);
  // with debug = true the kernels also write a residual image (kernel_opt_pose.cu:578-585): it must exist
  if (debug && !c->odo_debug) cudaMallocPitch(reinterpret_cast<void**>(&c->odo_debug), &c->odo_debug_pitch, sizeof(float) * c->cfg.depth_w, c->cfg.depth_h);
  CUDABuffer_<float> debug_image(c->odo_debug, base.h, base.w, c->odo_debug_pitch);
  CUDABuffer_<float>* debug_ptr = debug ? &debug_image : nullptr;
  if (use_gradmag)
    CallAccumulatePoseEstimationCoeffsFromImagesCUDAKernel_GradMag(
        s, debug, c->cfg.use_depth_residuals != 0, c->cfg.use_descriptor_residuals != 0, CornerProjector(k.dK), CenterProjector(k.cK),
        CenterUnprojector(k.dK), c->cfg.baseline_fx, DepthToColorScaled(k), threshold_factor, FrameTBase(base_T_frame), DepthBuf(base),
        NormalsBuf(base), ColorBuf(base), DepthBuf(trk), NormalsBuf(trk), trk.tex, cnt_buf, sum_buf, H_buf, b_buf, debug_ptr);
  else
    CallAccumulatePoseEstimationCoeffsFromImagesCUDAKernel_GradientXY(
        s, debug, c->cfg.use_depth_residuals != 0, c->cfg.use_descriptor_residuals != 0, CornerProjector(k.dK), CenterProjector(k.cK),
        CenterUnprojector(k.dK), c->cfg.baseline_fx, DepthToColorScaled(k), threshold_factor, FrameTBase(base_T_frame), DepthBuf(base),
        NormalsBuf(base), ColorBuf(base), DepthBuf(trk), NormalsBuf(trk), trk.tex, cnt_buf, sum_buf, H_buf, b_buf, debug_ptr);
  ++c->launches;
  if (debug) {
    cudaMemcpyAsync(count, c->residual_count, sizeof(u32), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(sum, c->residual_sum, sizeof(float), cudaMemcpyDeviceToHost, s);
  }
  cudaMemcpyAsync(H, c->H, sizeof(float) * 21, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(b, c->b, sizeof(float) * 6, cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
}

// ComputeCostAndResidualCountFromImagesCUDA, kernel_opt_pose.cc:194-260
void OdoCost(ref_context* c, int scale, const float base_T_frame[7], bool use_gradmag, u32* count, float* cost) {
  cudaStream_t s = c->stream;
  const OdoCameras k = ScaleCameras(c, scale);
  const float threshold_factor = pow(2, scale);
  cudaMemsetAsync(c->residual_count, 0, sizeof(u32), s);
  cudaMemsetAsync(c->residual_sum, 0, sizeof(float), s);
  c->launches += 2;
  const ref_context::OdoImage& base = c->odo[0][scale];
  const ref_context::OdoImage& trk = c->odo[1][scale];
  CUDABuffer_<u32> cnt_buf(c->residual_count, 1, 1, sizeof(u32));
  CUDABuffer_<float> sum_buf(c->residual_sum, 1, 1, sizeof(float));
  if (use_gradmag)
    CallComputeCostAndResidualCountFromImagesCUDAKernel_GradMag(
        s, c->cfg.use_depth_residuals != 0, c->cfg.use_descriptor_residuals != 0, CornerProjector(k.dK), CenterUnprojector(k.dK),
        c->cfg.baseline_fx, DepthToColorScaled(k), threshold_factor, FrameTBase(base_T_frame), DepthBuf(base), NormalsBuf(base), ColorBuf(base),
        DepthBuf(trk), NormalsBuf(trk), trk.tex, cnt_buf, sum_buf);
  else
    ComputeCostAndResidualCountFromImagesCUDAKernel_GradientXY(
        s, c->cfg.use_depth_residuals != 0, c->cfg.use_descriptor_residuals != 0, CornerProjector(k.dK), CenterUnprojector(k.dK),
        c->cfg.baseline_fx, DepthToColorScaled(k), threshold_factor, FrameTBase(base_T_frame), DepthBuf(base), NormalsBuf(base), ColorBuf(base),
        DepthBuf(trk), NormalsBuf(trk), trk.tex, cnt_buf, sum_buf);
  ++c->launches;
  cudaMemcpyAsync(count, c->residual_count, sizeof(u32), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(cost, c->residual_sum, sizeof(float), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
}

}  // namespace

extern "C" {

struct ref_odometry_result {
  int iterations[8];
  int chose_initial[8];
  unsigned int residual_count;
  float residual_sum;
  unsigned long long kernel_launches;
  float ms;   // wall time of the call after the uploads (pyramids + optimisation), stream-synchronised like the reference
};

// The tracked frame comes as dense host images (depth / normals u16 [h][w], colour uchar4 [ch][cw]); base = a stored keyframe.
int ref_track_frame_pairwise(ref_context* c, int base_kf, const unsigned short* depth, const unsigned short* normals, const unsigned char* color_rgba,
                             int num_scales, int use_pyramid_level_0, int use_gradmag, int test_different_initial_estimates,
                             const float init1[7], const float init2[7], float out[7], ref_odometry_result* res) {
  const ref_config& cfg = c->cfg;
  cudaStream_t s = c->stream;
  std::memset(res, 0, sizeof(*res));
  EnsureOdoBuffers(c, num_scales);
  RefKeyframe& fr = c->odo_frame;
  if (!fr.depth) {
    cudaMallocPitch(reinterpret_cast<void**>(&fr.depth), &fr.depth_pitch, cfg.depth_w * sizeof(u16), cfg.depth_h);
    cudaMallocPitch(reinterpret_cast<void**>(&fr.normals), &fr.normals_pitch, cfg.depth_w * sizeof(u16), cfg.depth_h);
    cudaMallocPitch(reinterpret_cast<void**>(&fr.color), &fr.color_pitch, cfg.color_w * sizeof(uchar4), cfg.color_h);
    // Keyframe / frame colour texture (keyframe.cc:67-73, bad_slam.cc color_texture_): uchar4, normalised float, linear, clamp
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = fr.color;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = cfg.color_w;
    rd.res.pitch2D.height = cfg.color_h;
    rd.res.pitch2D.pitchInBytes = fr.color_pitch;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaCreateTextureObject(&fr.tex, &rd, &td, nullptr);
  }
  cudaMemcpy2D(fr.depth, fr.depth_pitch, depth, cfg.depth_w * sizeof(u16), cfg.depth_w * sizeof(u16), cfg.depth_h, cudaMemcpyHostToDevice);
  cudaMemcpy2D(fr.normals, fr.normals_pitch, normals, cfg.depth_w * sizeof(u16), cfg.depth_w * sizeof(u16), cfg.depth_h, cudaMemcpyHostToDevice);
  cudaMemcpy2D(fr.color, fr.color_pitch, color_rgba, cfg.color_w * 4, cfg.color_w * 4, cfg.color_h, cudaMemcpyHostToDevice);
  cudaStreamSynchronize(s);
  const unsigned long long launches_before = c->launches;
  cudaEventRecord(c->ev[0], s);

  const RefKeyframe& base = c->kfs[base_kf];
  CUDABuffer_<u8> base_kf_gradmag(c->odo_gradmag[0], cfg.color_h, cfg.color_w, c->odo_gradmag_pitch[0]);
  CUDABuffer_<u8> tracked_gradmag(c->odo_gradmag[1], cfg.color_h, cfg.color_w, c->odo_gradmag_pitch[1]);
  // ---- BadSlam::RunOdometry, bad_slam.cc:863-902
  if (use_gradmag) ComputeSobelGradientMagnitudeCUDA(s, base.tex, &base_kf_gradmag);
  else ComputeBrightnessCUDA(s, base.tex, &base_kf_gradmag);
  ++c->launches;
  {
    CUDABuffer_<float> d0 = DepthBuf(c->odo[0][0]);
    CUDABuffer_<u8> c0 = ColorBuf(c->odo[0][0]);
    CalibrateDepthAndTransformColorToDepthCUDA(s, DepthToColor(cfg), MakeDepthParams(c), CUDABuffer_<u16>(base.depth, cfg.depth_h, cfg.depth_w, base.depth_pitch),
                                               c->odo_gradmag_tex[0], &d0, &c0);
    ++c->launches;
  }
  if (use_gradmag) ComputeSobelGradientMagnitudeCUDA(s, fr.tex, &tracked_gradmag);
  else ComputeBrightnessCUDA(s, fr.tex, &tracked_gradmag);
  ++c->launches;

  // ---- TrackFramePairwise, pairwise_frame_tracking.cc:153-678 (kDebug = false, no convergence-sample file)
  // per-level image views: level 0 of the base uses the keyframe's normals, level 0 of the tracked frame the frame's
  ref_context::OdoImage img[2][8];
  for (int f = 0; f < 2; ++f)
    for (int l = 0; l < num_scales; ++l) img[f][l] = c->odo[f][l];
  img[0][0].normals = base.normals; img[0][0].normals_pitch = base.normals_pitch;
  img[1][0].normals = fr.normals;   img[1][0].normals_pitch = fr.normals_pitch;
  const CUDABuffer_<u16> tracked_depth_buffer(fr.depth, cfg.depth_h, cfg.depth_w, fr.depth_pitch);
  const CUDABuffer_<u16> tracked_normals_buffer(fr.normals, cfg.depth_h, cfg.depth_w, fr.normals_pitch);
  if (use_pyramid_level_0) {
    CUDABuffer_<float> d = DepthBuf(img[1][0]);
    CalibrateDepthCUDA(s, MakeDepthParams(c), tracked_depth_buffer, &d);
    ++c->launches;
    ColorBuf(img[1][0]).SetToReadModeNormalized(c->odo_gradmag_tex[1], s);
    ++c->launches;
  } else {
    if (cfg.depth_w != cfg.color_w && cfg.depth_w != 2 * cfg.color_w) return 1;   // LOG(FATAL) in the reference
    CUDABuffer_<float> d = DepthBuf(img[1][1]);
    CUDABuffer_<u16> n = NormalsBuf(img[1][1]);
    CUDABuffer_<u8> col = ColorBuf(img[1][1]);
    CalibrateAndDownsampleImagesCUDA(s, cfg.depth_w == cfg.color_w, MakeDepthParams(c), tracked_depth_buffer, tracked_normals_buffer,
                                     c->odo_gradmag_tex[1], &d, &n, &col, false);
    ++c->launches;
  }
  for (int scale = 1; scale < num_scales; ++scale) {
    if (scale >= 2 || use_pyramid_level_0) {
      CUDABuffer_<float> d = DepthBuf(img[1][scale]);
      CUDABuffer_<u16> n = NormalsBuf(img[1][scale]);
      CUDABuffer_<u8> col = ColorBuf(img[1][scale]);
      DownsampleImagesCUDA(s, DepthBuf(img[1][scale - 1]), NormalsBuf(img[1][scale - 1]), img[1][scale - 1].tex, &d, &n, &col, false);
      ++c->launches;
    }
    CUDABuffer_<float> d = DepthBuf(img[0][scale]);
    CUDABuffer_<u16> n = NormalsBuf(img[0][scale]);
    CUDABuffer_<u8> col = ColorBuf(img[0][scale]);
    DownsampleImagesCUDA(s, DepthBuf(img[0][scale - 1]), NormalsBuf(img[0][scale - 1]), img[0][scale - 1].tex, &d, &n, &col, false);
    ++c->launches;
  }
  // the views (with the level-0 normals of this call) are what the kernels and the parity hooks see
  for (int f = 0; f < 2; ++f)
    for (int l = 0; l < num_scales; ++l) { c->odo[f][l].normals = img[f][l].normals; c->odo[f][l].normals_pitch = img[f][l].normals_pitch; }

  float est[7], chosen_initial[7];
  std::memcpy(est, init1, sizeof(est));
  std::memcpy(chosen_initial, init1, sizeof(est));
  const int kMaxIterationsPerScale = 30;
  for (int scale = num_scales - 1; scale >= (use_pyramid_level_0 ? 0 : 1); --scale) {
    const float scaling_factor = pow(2, scale);
    res->chose_initial[scale] = -1;
    if (scale != num_scales - 1 || test_different_initial_estimates) {
      float last_scale[7], initial[7];
      std::memcpy(last_scale, (scale != num_scales - 1) ? est : init1, sizeof(est));
      std::memcpy(initial, (scale != num_scales - 1) ? chosen_initial : init2, sizeof(est));
      u32 count_last = 0, count_init = 0;
      float cost_last = 0, cost_init = 0;
      OdoCost(c, scale, last_scale, use_gradmag != 0, &count_last, &cost_last);
      OdoCost(c, scale, initial, use_gradmag != 0, &count_init, &cost_init);
      bool take_last;
      if (count_last > 2 * count_init) take_last = true;
      else if (count_init > 2 * count_last) take_last = false;
      else take_last = cost_last < cost_init;
      std::memcpy(est, take_last ? last_scale : initial, sizeof(est));
      res->chose_initial[scale] = take_last ? 0 : 1;
      if (scale == num_scales - 1) std::memcpy(chosen_initial, est, sizeof(est));
    }
    int iteration;
    for (iteration = 0; iteration < kMaxIterationsPerScale; ++iteration) {
      float H[21], b[6];
      double Hd[36], bd[6], xd[6];
      std::memset(Hd, 0, sizeof(Hd));
      OdoAccumulate(c, scale, est, use_gradmag != 0, false, nullptr, nullptr, H, b);
      int idx = 0;
      for (int r = 0; r < 6; ++r)
        for (int cc = r; cc < 6; ++cc) Hd[r * 6 + cc] = H[idx++];
      for (int i = 0; i < 6; ++i) bd[i] = b[i];
      hm_ldlt_solve(6, Hd, bd, xd);
      float damping = 1.f;
      if (scale == num_scales - 2) damping = 0.5f;
      else if (scale == num_scales - 1) damping = 0.25f;
      float x[6], step[6], e[7], next[7];
      for (int i = 0; i < 6; ++i) { x[i] = static_cast<float>(xd[i]); step[i] = -damping * x[i]; }
      hm_se3_exp(step, e);
      hm_se3_mul(est, e, next);
      std::memcpy(est, next, sizeof(est));
      // IsScaleNPoseEstimationConverged, convergence_analysis.h:56-63
      const float sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
      if (sq < scaling_factor * scaling_factor * 1e-08f) { ++iteration; break; }
    }
    res->iterations[scale] = iteration;
  }
  cudaEventRecord(c->ev[1], s);
  cudaEventSynchronize(c->ev[1]);
  cudaEventElapsedTime(&res->ms, c->ev[0], c->ev[1]);
  std::memcpy(out, est, sizeof(est));
  res->kernel_launches = c->launches - launches_before;
  return 0;
}

// Parity hooks on the pyramids of the last ref_track_frame_pairwise call.
int ref_odometry_get_level(ref_context* c, int which, int scale, float* depth, unsigned short* normals, unsigned char* color, int* w, int* h) {
  if (which < 0 || which > 1 || scale < 0 || scale >= c->odo_scales) return 1;
  const ref_context::OdoImage& im = c->odo[which][scale];
  if (depth) cudaMemcpy2D(depth, sizeof(float) * im.w, im.depth, im.depth_pitch, sizeof(float) * im.w, im.h, cudaMemcpyDeviceToHost);
  if (normals) cudaMemcpy2D(normals, sizeof(u16) * im.w, im.normals, im.normals_pitch, sizeof(u16) * im.w, im.h, cudaMemcpyDeviceToHost);
  if (color) cudaMemcpy2D(color, im.w, im.color, im.color_pitch, im.w, im.h, cudaMemcpyDeviceToHost);
  if (w) *w = im.w;
  if (h) *h = im.h;
  return 0;
}
void ref_odometry_coeffs(ref_context* c, int scale, int use_gradmag, const float pose_a[7], const float pose_b[7], float H[21], float b[6],
                         unsigned int* count, float* sum, unsigned int counts[2], float costs[2]) {
  OdoAccumulate(c, scale, pose_a, use_gradmag != 0, true, count, sum, H, b);
  OdoCost(c, scale, pose_a, use_gradmag != 0, &counts[0], &costs[0]);
  OdoCost(c, scale, pose_b, use_gradmag != 0, &counts[1], &costs[1]);
}

}  // extern "C"
