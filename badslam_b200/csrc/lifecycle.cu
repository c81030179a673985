// lifecycle.cu -- end-of-BA surfel maintenance (DirectBA::PerformBASchemeEndTasks, direct_ba.cc:566-653):
//  * ObservationStatsKernel: DeleteSurfelsAndUpdateRadiiCUDA (kernel_delete_surfels.cc:40-98, kernel_delete_surfels.cu:42-164).
//    Per surfel, over EVERY keyframe: observation count, free-space violations (the pixel's surface lies behind the surfel
//    by more than the association threshold) and the smallest measured radius^2; then delete (x = NaN pattern) surfels
//    with too few observations or more violations than observations, and store the new radius^2 otherwise.
//    The reference launches one kernel per keyframe that read-modify-writes three accumulator rows; here one persistent
//    launch walks (keyframe group, surfel tile) items group-major like the geometry kernels (kernels.cu).
//  * Compact*: CompactSurfelsCUDA (kernel_compact_surfels.cu:159-279): the r-th valid surfel counted from the END moves into
//    the r-th free spot counted from the front when that spot lies in front of it -- the same permutation as the reference
//    (two cub::DeviceScan passes there; one three-kernel exclusive scan here).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "kernels.cuh"

namespace bba {

namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 16;
constexpr uint32_t kDeletedPattern = 0x7fffffffu;   // CUDART_NAN_F

__device__ __forceinline__ unsigned int LoadAcquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void StoreRelease(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads) ObservationStatsKernel(const __grid_constant__ SurfelStatsArgs a) {
  const uint32_t tile_len = 1u << a.tile_shift;
  const uint32_t n_tiles = (a.n + tile_len - 1) >> a.tile_shift;
  const uint32_t n_groups = (a.kf_count + kGroup - 1) / kGroup;
  const uint32_t n_items = n_groups * n_tiles;
  const size_t P = a.pitch;
  const int lane = threadIdx.x & 31;
  const CameraParams& cam = a.cam;
  for (;;) {
    // warp-owned item (group, tile); (g, t) runs after (g - 1, t) has been retired
    unsigned int item = 0;
    if (lane == 0) {
      item = atomicAdd(a.queue, 1u);
      if (item < n_items) {
        const uint32_t g = item / n_tiles, t = item - g * n_tiles;
        while (LoadAcquire(a.tile_epoch + t) < g) __nanosleep(64);
      }
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const uint32_t group = item / n_tiles, tile = item - group * n_tiles;
    const bool first = group == 0, last = group + 1 == n_groups;
    const int j_begin = group * kGroup, j_end = min(a.kf_count, static_cast<int>(group + 1) * kGroup);
    unsigned int deleted_here = 0;
    for (uint32_t sub = 0; sub < tile_len / 32; ++sub) {
      const uint32_t i = (tile << a.tile_shift) + sub * 32 + lane;
      bool deleted = false;
      if (i < a.n) {
        float obs = 0.f, viol = 0.f, min_r2 = __int_as_float(0x7f800000);   // +inf (kernel_delete_surfels.cu:50)
        if (!first) {
          obs = __ldcg(a.surfels + (kRowAccum0 + 0) * P + i);
          viol = __ldcg(a.surfels + (kRowAccum0 + 1) * P + i);
          min_r2 = __ldcg(a.surfels + (kRowAccum0 + 2) * P + i);
        }
        const float x = a.surfels[kRowX * P + i];
        const Vec3 gp = V3(x, a.surfels[kRowY * P + i], a.surfels[kRowZ * P + i]);
        const Vec3 nrm = UnpackNormal(__float_as_uint(a.surfels[kRowNormal * P + i]));
        for (int kf = j_begin; kf < j_end; ++kf) {
          const KfDevice& K = a.kfs[kf];
          float T[12];
#pragma unroll
          for (int c = 0; c < 12; ++c) T[c] = __ldg(&K.T[c]);
          Assoc r;
          if (!ProjectIntoImage(cam, T, gp, &r)) continue;
          const PixelLoads l = LoadPixel(cam, K.depth, K.depth_pitch, K.normals, K.normals_pitch, r);
          // IsAssociatedWithPixel<return_free_space_violations = true> (surfel_projection_nvcc_only.cuh:48-127)
          if (l.measured & kInvalidDepthBit) continue;
          const float d = RawToCalibratedDepth(cam.a, l.cf, cam.raw_to_float, l.measured);
          const Vec3 ln = Rotate(T, nrm);
          const float nx = cam.fx_inv * r.px + cam.cx_inv, ny = cam.fy_inv * r.py + cam.cy_inv;
          const float thr = kDepthTukey * ((kDepthUncertaintyFactor * fabsf(ln.x * nx + ln.y * ny + ln.z) * (d * d)) / cam.baseline_fx);
          const float diff = d - r.lp.z;
          if (diff > thr) {
            viol += 1.f;
            continue;
          }
          if (diff < -thr) continue;
          if (Dot(r.lp, ln) > 0) continue;
          if (Dot(ln, U16ToImageSpaceNormal(l.kf_normal)) < kCosNormalCompat) continue;
          obs += 1.f;
          const KfRadius& R = a.radius[kf];
          const uint16_t h = __ldg(reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(R.ptr) + static_cast<size_t>(r.py) * R.pitch) + r.px);
          min_r2 = fminf(min_r2, __half2float(__ushort_as_half(h)));
        }
        if (!last) {
          __stcg(a.surfels + (kRowAccum0 + 0) * P + i, obs);
          __stcg(a.surfels + (kRowAccum0 + 1) * P + i, viol);
          __stcg(a.surfels + (kRowAccum0 + 2) * P + i, min_r2);
        } else {
          // MarkDeletedSurfelsCUDAKernel (kernel_delete_surfels.cu:129-164)
          if (obs < static_cast<float>(a.min_observation_count) || viol > obs) {
            if (__float_as_uint(x) != kDeletedPattern) {
              a.surfels[kRowX * P + i] = __uint_as_float(kDeletedPattern);
              deleted = true;
            }
          } else {
            a.surfels[kRowRadiusSq * P + i] = min_r2;
          }
        }
      }
      deleted_here += __popc(__ballot_sync(0xffffffffu, deleted));
    }
    __syncwarp();
    if (lane == 0) {
      if (deleted_here) atomicAdd(a.deleted_count, deleted_here);
      __threadfence();
      StoreRelease(a.tile_epoch + tile, group + 1);
    }
  }
}

// ---- compaction ---------------------------------------------------------------------------------------------------------
constexpr int kScanItems = 4;                               // surfels per thread
constexpr int kScanBlock = 1024 * kScanItems;               // surfels per block

// flags (1 = deleted) -> scratch row Accum2; per-block number of deleted surfels -> block_sums
__global__ void __launch_bounds__(1024) CompactFlagKernel(float* __restrict__ surfels, uint32_t pitch, uint32_t n, unsigned int* __restrict__ block_sums) {
  __shared__ unsigned int warp_sums[32];
  const size_t P = pitch;
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  unsigned int cnt = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t i = base + k;
    if (i < n) {
      const unsigned int f = __float_as_uint(surfels[kRowX * P + i]) == kDeletedPattern ? 1u : 0u;
      reinterpret_cast<unsigned int*>(surfels + (kRowAccum0 + 2) * P)[i] = f;
      cnt += f;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned int s = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
  }
}

// exclusive scan of the block sums in place (one block); block_sums[n_blocks] receives the total
__global__ void __launch_bounds__(1024) CompactScanBlocksKernel(unsigned int* __restrict__ block_sums, uint32_t n_blocks) {
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned int chunk_total;
  unsigned int carry = 0;   // identical in every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_blocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned int v = i < n_blocks ? block_sums[i] : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned int w = warp_tot[lane];
      unsigned int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;   // exclusive offset of each warp
      if (lane == 31) chunk_total = winc;
    }
    __syncthreads();
    if (i < n_blocks) block_sums[i] = carry + warp_tot[warp] + inc - v;
    carry += chunk_total;
    __syncthreads();   // warp_tot / chunk_total are rewritten by the next chunk
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}

// exclusive count of deleted surfels in front of every surfel -> scratch row Accum0; list of free spots -> row Accum3
__global__ void __launch_bounds__(1024) CompactListKernel(float* __restrict__ surfels, uint32_t pitch, uint32_t n,
                                                          const unsigned int* __restrict__ block_offsets) {
  __shared__ unsigned int warp_tot[32];
  const size_t P = pitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  const unsigned int* flags = reinterpret_cast<const unsigned int*>(surfels + (kRowAccum0 + 2) * P);
  unsigned int* before = reinterpret_cast<unsigned int*>(surfels + (kRowAccum0 + 0) * P);
  unsigned int* free_list = reinterpret_cast<unsigned int*>(surfels + (kRowAccum0 + 3) * P);
  unsigned int f[kScanItems], mine = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    f[k] = (base + k < n) ? flags[base + k] : 0u;
    mine += f[k];
  }
  unsigned int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned int w = warp_tot[lane];
    unsigned int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
  }
  __syncthreads();
  unsigned int run = block_offsets[blockIdx.x] + warp_tot[warp] + inc - mine;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t i = base + k;
    if (i < n) {
      before[i] = run;
      if (f[k]) free_list[run] = i;
      run += f[k];
    }
  }
}

// CompactSurfelsCUDAKernel (kernel_compact_surfels.cu:126-157) with adapt_active_surfels = false (direct_ba.cc:618)
__global__ void __launch_bounds__(256) CompactMoveKernel(float* __restrict__ surfels, uint32_t pitch, uint32_t n, uint32_t free_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t P = pitch;
  const unsigned int* flags = reinterpret_cast<const unsigned int*>(surfels + (kRowAccum0 + 2) * P);
  if (flags[i]) return;
  const unsigned int* before = reinterpret_cast<const unsigned int*>(surfels + (kRowAccum0 + 0) * P);
  const unsigned int* free_list = reinterpret_cast<const unsigned int*>(surfels + (kRowAccum0 + 3) * P);
  // number of valid surfels behind i = (valid total) - (valid up to and including i)
  const uint32_t valid_total = n - free_count;
  const uint32_t reverse_index = valid_total - (i + 1 - before[i]);
  if (reverse_index >= free_count) return;
  const uint32_t spot = free_list[reverse_index];
  if (spot >= i) return;
#pragma unroll
  for (int row = 0; row < kRowAccum0; ++row) surfels[row * P + spot] = surfels[row * P + i];
}

void LaunchObservationStats(SurfelStatsArgs a, int sm_count, cudaStream_t stream) {
  if (a.n == 0 || a.kf_count <= 0) return;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ObservationStatsKernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const uint64_t resident_warps = static_cast<uint64_t>(per_sm) * sm_count * (kThreads / 32);
  int shift = 8;
  while (shift > 5 && 2 * static_cast<uint64_t>((a.n + (1u << shift) - 1) >> shift) < 3 * resident_warps) --shift;
  a.tile_shift = shift;
  const uint32_t n_tiles = (a.n + (1u << shift) - 1) >> shift;
  const uint64_t n_items = static_cast<uint64_t>(n_tiles) * ((a.kf_count + kGroup - 1) / kGroup);
  cudaMemsetAsync(a.queue, 0, sizeof(unsigned int), stream);
  cudaMemsetAsync(a.tile_epoch, 0, sizeof(unsigned int) * n_tiles, stream);
  const uint64_t ctas = std::min<uint64_t>((n_items + kThreads / 32 - 1) / (kThreads / 32), static_cast<uint64_t>(per_sm) * sm_count);
  ObservationStatsKernel<<<static_cast<uint32_t>(ctas), kThreads, 0, stream>>>(a);
}

uint32_t CompactScratchWords(uint32_t n) { return (n + kScanBlock - 1) / kScanBlock + 2; }

void LaunchCompactSurfels(float* surfels, uint32_t pitch, uint32_t n, uint32_t free_count, unsigned int* block_sums, cudaStream_t stream) {
  if (n == 0 || free_count == 0) return;
  const uint32_t n_blocks = (n + kScanBlock - 1) / kScanBlock;
  CompactFlagKernel<<<n_blocks, 1024, 0, stream>>>(surfels, pitch, n, block_sums);
  CompactScanBlocksKernel<<<1, 1024, 0, stream>>>(block_sums, n_blocks);
  CompactListKernel<<<n_blocks, 1024, 0, stream>>>(surfels, pitch, n, block_sums);
  CompactMoveKernel<<<(n + 255) / 256, 256, 0, stream>>>(surfels, pitch, n, free_count);
}

}  // namespace bba
