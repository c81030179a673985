// intrinsics.cu -- camera-intrinsics + depth-deformation step of the alternating BA (OptimizeIntrinsicsCUDA,
// applications/badslam/src/badslam/kernel_opt_intrinsics.cc:39-281).
//
// Unknowns: (fx^-1, fy^-1, cx^-1, cy^-1, a) of the depth camera + one cfactor per sparse cell (arrow-shaped normal
// equations: A 5x5, B 5xP, D diagonal P) and the four pinhole parameters of the colour camera.
//  * IntrinsicsAccumulateKernel: the reference launches AccumulateIntrinsicsCoefficientsCUDAKernel once per keyframe
//    (kernel_opt_intrinsics.cu:46-217) with 20 block-wide CUB reductions + atomics per block; here ONE persistent
//    launch walks (keyframe group, surfel tile) items group-major (images L2-resident), each lane keeps the 34 global
//    sums in registers, one transposed warp reduction + fp64 REDs per item; the per-cell terms go out as fp32 REDs.
//  * IntrinsicsSchurKernel: A -= B D^-1 B^T, b1 -= B D^-1 b2 (kernel_opt_intrinsics.cu:265-347).
//  * IntrinsicsCellUpdateKernel: cfactor -= D^-1 b2 - D^-1 B^T x1 (kernel_opt_intrinsics.cu:374-424).
// The 5x5 / 4x4 solves stay on the host in fp64 like the reference (kernel_opt_intrinsics.cc:171,272).
#include <cuda.h>

#include <algorithm>

#include "kernels.cuh"

namespace bba {

namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 16;
constexpr int kTile = 256;

__device__ __forceinline__ float WarpSum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// same butterfly as in kernels.cu (kept local: different translation unit)
__device__ __forceinline__ float TransposeReduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float lo = v[i], hi = v[i + half];
      const float send = upper ? lo : hi;
      const float keep = upper ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

struct KfLite {
  float T[12];
  const uint16_t* depth;
  const uint16_t* normals;
  cudaTextureObject_t tex;
  uint32_t depth_pitch, normals_pitch;
};

__device__ __forceinline__ void LoadKfLite(const KfDevice* __restrict__ kfs, int kf, KfLite* r) {
  const KfDevice& k = kfs[kf];
#pragma unroll
  for (int i = 0; i < 12; ++i) r->T[i] = __ldg(&k.T[i]);
  r->depth = k.depth;
  r->normals = k.normals;
  r->tex = k.tex;
  r->depth_pitch = k.depth_pitch;
  r->normals_pitch = k.normals_pitch;
}

}  // namespace

// sums layout: [0..14] A upper triangle (5x5), [15..19] b1, [20..29] colour H upper triangle (4x4), [30..33] colour b
template <bool OPT_COLOR, bool OPT_DEPTH>
__global__ void __launch_bounds__(kThreads) IntrinsicsAccumulateKernel(const __grid_constant__ IntrinsicsArgs a) {
  const uint32_t n_tiles = (a.end - a.begin + kTile - 1) / kTile;
  const uint32_t n_groups = (a.kf_count + kGroup - 1) / kGroup;
  const uint32_t n_items = n_groups * n_tiles;
  const size_t P = a.pitch;
  const int lane = threadIdx.x & 31;
  const CameraParams& cam = a.cam;
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(a.queue, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const uint32_t group = item / n_tiles, tile = item - group * n_tiles;
    const int j_begin = group * kGroup, j_end = min(a.kf_count, static_cast<int>(group + 1) * kGroup);
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    float extra0 = 0.f, extra1 = 0.f;   // sums 32, 33
    for (uint32_t sub = 0; sub < kTile / 32; ++sub) {
      const uint32_t li = a.begin + tile * kTile + sub * 32 + lane;
      const uint32_t i = SurfelShardToGlobal(li, a.shard_rank, a.shard_world);
      if (li >= a.end || i >= a.n) continue;
      const Vec3 gp = V3(a.surfels[kRowX * P + i], a.surfels[kRowY * P + i], a.surfels[kRowZ * P + i]);
      const Vec3 nrm = UnpackNormal(__float_as_uint(a.surfels[kRowNormal * P + i]));
      float radius_sq = 0.f, d1 = 0.f, d2 = 0.f;
      if (OPT_COLOR) {
        radius_sq = a.surfels[kRowRadiusSq * P + i];
        d1 = a.surfels[kRowD1 * P + i];
        d2 = a.surfels[kRowD2 * P + i];
      }
      for (int j = j_begin; j < j_end; ++j) {
        KfLite K;
        LoadKfLite(a.kfs, __ldg(a.kf_list + j), &K);
        Assoc r;
        if (!ProjectIntoImage(cam, K.T, gp, &r)) continue;
        const PixelLoads l = LoadPixel(cam, K.depth, K.depth_pitch, K.normals, K.normals_pitch, r);
        DescEval e;
        bool photo = false;
        if (OPT_COLOR) {
          float ccx, ccy;
          photo = DepthToColor(cam, r.pxf, r.pyf, &ccx, &ccy);
          float t1x, t1y, t2x, t2y;
          TangentProjections(cam, K.T, gp, nrm, radius_sq, &t1x, &t1y, &t2x, &t2y);
          EvalDescriptor(K.tex, ccx, ccy, t1x, t1y, t2x, t2y, d1, d2, &e);
        }
        if (Associate(cam, K.T, nrm, l, &r) != 3) continue;
        if (OPT_DEPTH) {
          // kernel_opt_intrinsics.cu:84-121
          const unsigned int spx = (cam.cell == 1) ? static_cast<unsigned int>(r.px) : __umulhi(static_cast<unsigned int>(r.px), cam.cell_magic);
          const unsigned int spy = (cam.cell == 1) ? static_cast<unsigned int>(r.py) : __umulhi(static_cast<unsigned int>(r.py), cam.cell_magic);
          const float raw_inv_depth = 1.0f / (cam.raw_to_float * l.measured);
          const float exp_inv_depth = expf(-cam.a * raw_inv_depth);
          const float corrected_inv_depth = l.cf * exp_inv_depth + raw_inv_depth;
          if (fabsf(corrected_inv_depth) > 1e-4f) {
            const float dot = r.nx * r.ln.x + r.ny * r.ln.y + r.ln.z;
            const float inv_stddev =
                cam.baseline_fx / (kDepthUncertaintyFactor * fabsf(r.ln.x * r.nx + r.ln.y * r.ny + r.ln.z) * (r.d * r.d));
            const float jac_base = inv_stddev * dot * exp_inv_depth / (corrected_inv_depth * corrected_inv_depth);
            float J[6];
            J[2] = inv_stddev * r.d * (nrm.x * K.T[0] + nrm.y * K.T[1] + nrm.z * K.T[2]);
            J[3] = inv_stddev * r.d * (nrm.x * K.T[4] + nrm.y * K.T[5] + nrm.z * K.T[6]);
            J[0] = r.px * J[2];
            J[1] = r.py * J[3];
            J[4] = l.cf * raw_inv_depth * jac_base;
            J[5] = -jac_base;
            const Vec3 up = V3(r.d * r.nx, r.d * r.ny, r.d);
            const float raw = inv_stddev * Dot(r.ln, up - r.lp);
            const float w = DepthWeight(raw);
            int idx = 0;
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) {
              const float wj = w * J[rr];
#pragma unroll
              for (int c = rr; c < 5; ++c) acc[idx++] += wj * J[c];
            }
            const float wr = w * raw;
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) acc[15 + rr] += wr * J[rr];
            // per-cell terms (kernel_opt_intrinsics.cu:173-190)
            const unsigned int sp = spx + spy * cam.cf_w;
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) atomicAdd(a.cell_B + static_cast<size_t>(rr) * a.cell_count + sp, w * J[rr] * J[5]);
            atomicAdd(a.cell_D + sp, w * J[5] * J[5]);
            atomicAdd(a.cell_b2 + sp, w * raw * J[5]);
            atomicAdd(a.cell_obs + sp, 1.0f);   // fp32 count: exact to 2^24, only tested against 0
          }
        }
        if (OPT_COLOR && photo) {
          // kernel_opt_intrinsics.cu:139-160,193-211: residuals that are exactly 0 are skipped
          const float J1[4] = {e.gx1 * r.nx, e.gy1 * r.ny, e.gx1, e.gy1};
          const float J2[4] = {e.gx2 * r.nx, e.gy2 * r.ny, e.gx2, e.gy2};
          const float w1 = (e.r1 != 0) ? DescWeight(e.r1) : 0.f;
          const float w2 = (e.r2 != 0) ? DescWeight(e.r2) : 0.f;
          int idx = 20;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
            for (int c = rr; c < 4; ++c) acc[idx++] += w1 * J1[rr] * J1[c] + w2 * J2[rr] * J2[c];
          }
          acc[30] += w1 * e.r1 * J1[0] + w2 * e.r2 * J2[0];
          acc[31] += w1 * e.r1 * J1[1] + w2 * e.r2 * J2[1];
          extra0 += w1 * e.r1 * J1[2] + w2 * e.r2 * J2[2];
          extra1 += w1 * e.r1 * J1[3] + w2 * e.r2 * J2[3];
        }
      }
    }
    __syncwarp();
    const float total = TransposeReduce32(acc, lane);
    if (total != 0.f) atomicAdd(a.sums + lane, static_cast<double>(total));
    if (OPT_COLOR) {
      extra0 = WarpSum(extra0);
      extra1 = WarpSum(extra1);
      if (lane == 0 && (extra0 != 0.f || extra1 != 0.f)) {
        atomicAdd(a.sums + 32, static_cast<double>(extra0));
        atomicAdd(a.sums + 33, static_cast<double>(extra1));
      }
    }
  }
}

// Schur complement over the sparse cells (kernel_opt_intrinsics.cu:265-347): A -= B D^-1 B^T, b1 -= B D^-1 b2, and
// B <- D^-1 B^T, D <- D^-1 b2 in place.  ONE block walks all cells in a fixed order and reduces in a fixed tree, so the
// result is bit-reproducible: with several ranks every replica computes exactly the same intrinsics update.
__global__ void __launch_bounds__(1024) IntrinsicsSchurKernel(uint32_t cell_count, float* __restrict__ B, float* __restrict__ D,
                                                              const float* __restrict__ b2, double* __restrict__ sums) {
  __shared__ double partial[32][20];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[20];
#pragma unroll
  for (int i = 0; i < 20; ++i) v[i] = 0.0;
  for (uint32_t p = threadIdx.x; p < cell_count; p += blockDim.x) {
    const float D_inverse = 1.0f / D[p];
    if (!(D_inverse < 1e12f)) {
      D[p] = __int_as_float(0x7fffffff);   // NaN marks cells without constraint
      continue;
    }
    const float D_inv_b2 = D_inverse * b2[p];
    D[p] = D_inv_b2;
    float Bp[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) Bp[r] = B[static_cast<size_t>(r) * cell_count + p];
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
#pragma unroll
      for (int c = r; c < 5; ++c) v[idx++] -= static_cast<double>(Bp[r] * D_inverse * Bp[c]);
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) v[15 + r] -= static_cast<double>(Bp[r] * D_inv_b2);
#pragma unroll
    for (int r = 0; r < 5; ++r) B[static_cast<size_t>(r) * cell_count + p] = D_inverse * Bp[r];
  }
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    double s = v[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) partial[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 20) {
    double s = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += partial[w][threadIdx.x];
    sums[threadIdx.x] += s;
  }
}

// kernel_opt_intrinsics.cu:374-424
__global__ void __launch_bounds__(256) IntrinsicsCellUpdateKernel(uint32_t cell_count, const float* __restrict__ obs,
                                                                  const float* __restrict__ B, const float* __restrict__ D,
                                                                  const float* __restrict__ x1, float* __restrict__ cfactor) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cell_count) return;
  float offset = D[p];
  if (isnan(offset)) {
    offset = 0;
  } else {
#pragma unroll
    for (int r = 0; r < 5; ++r) offset -= B[static_cast<size_t>(r) * cell_count + p] * x1[r];
  }
  float cf = cfactor[p] - offset;
  if (obs[p] == 0.f) cf = 0;
  cfactor[p] = cf;
}

void LaunchIntrinsicsAccumulate(const IntrinsicsArgs& a, int sm_count, bool optimize_color, bool optimize_depth, cudaStream_t stream) {
  if (a.end <= a.begin || a.kf_count <= 0) return;
  cudaMemsetAsync(a.queue, 0, sizeof(unsigned int), stream);
  const uint32_t n_tiles = (a.end - a.begin + kTile - 1) / kTile;
  const uint32_t n_groups = (a.kf_count + kGroup - 1) / kGroup;
  const uint64_t n_items = static_cast<uint64_t>(n_tiles) * n_groups;
  auto launch = [&](auto kernel) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const uint64_t ctas = std::min<uint64_t>((n_items + 7) / 8, static_cast<uint64_t>(per_sm) * sm_count);
    kernel<<<static_cast<uint32_t>(ctas), kThreads, 0, stream>>>(a);
  };
  if (optimize_color && optimize_depth) launch(IntrinsicsAccumulateKernel<true, true>);
  else if (optimize_color) launch(IntrinsicsAccumulateKernel<true, false>);
  else if (optimize_depth) launch(IntrinsicsAccumulateKernel<false, true>);
}

__global__ void IntrinsicsConvertSumsKernel(double* sums, float* head, int to_float) {
  const int i = threadIdx.x;
  if (i >= kIntrinsicsSums) return;
  if (to_float) head[i] = static_cast<float>(sums[i]);
  else sums[i] = static_cast<double>(head[i]);
}
void LaunchIntrinsicsConvertSums(double* sums, float* head, bool to_float, cudaStream_t stream) {
  IntrinsicsConvertSumsKernel<<<1, 64, 0, stream>>>(sums, head, to_float ? 1 : 0);
}

void LaunchIntrinsicsSchur(uint32_t cell_count, float* B, float* D, const float* b2, double* sums, cudaStream_t stream) {
  IntrinsicsSchurKernel<<<1, 1024, 0, stream>>>(cell_count, B, D, b2, sums);
}

void LaunchIntrinsicsCellUpdate(uint32_t cell_count, const float* obs, const float* B, const float* D, const float* x1,
                                float* cfactor, cudaStream_t stream) {
  IntrinsicsCellUpdateKernel<<<(cell_count + 255) / 256, 256, 0, stream>>>(cell_count, obs, B, D, x1, cfactor);
}

}  // namespace bba
