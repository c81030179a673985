// tools/tex_probe2.cu -- 1-D staircases of the hardware bilinear weights in x and in y.
// Texture A: columns 0..3 = 0, columns 4..7 = 255  -> tex2D(x, const) = alpha_hw(x)
// Texture B: rows 0..1 = 0, rows 2..3 = 255        -> tex2D(const, y) = beta_hw(y)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
__global__ void sample(cudaTextureObject_t t, float x0, float dx, float y0, float dy, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex2D<float>(t, x0 + dx * i, y0 + dy * i);
}
static cudaTextureObject_t make(const std::vector<unsigned char>& img, int W, int H) {
  unsigned char* d;
  size_t pitch;
  cudaMallocPitch((void**)&d, &pitch, W, H);
  cudaMemcpy2D(d, pitch, img.data(), W, W, H, cudaMemcpyHostToDevice);
  cudaResourceDesc res;
  memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = d;
  res.res.pitch2D.desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
  res.res.pitch2D.width = W;
  res.res.pitch2D.height = H;
  res.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  cudaTextureObject_t tex;
  cudaCreateTextureObject(&tex, &res, &td, nullptr);
  return tex;
}
static void report(const char* name, const std::vector<float>& hw, float c0, float dc, float base) {
  const int N = (int)hw.size();
  int steps = 0;
  float prev = -1;
  double dev_round = 0, dev_trunc = 0;
  for (int i = 0; i < N; ++i) {
    float a = (c0 + dc * i) - 0.5f - base;
    if (hw[i] != prev) {
      if (steps < 4 || steps % 64 == 0) printf("  %s step %4d at frac*256=%.4f value*65535=%.3f\n", name, steps, a * 256, hw[i] * 65535.0);
      ++steps;
      prev = hw[i];
    }
    double r = floor(a * 256 + 0.5) / 256, t = floor(a * 256) / 256;
    dev_round = fmax(dev_round, fabs(hw[i] - r));
    dev_trunc = fmax(dev_trunc, fabs(hw[i] - t));
  }
  printf("%s: %d distinct values; max|hw-round8| %.3e  max|hw-trunc8| %.3e\n", name, steps, dev_round, dev_trunc);
}
int main() {
  const int W = 8, H = 4, N = 1 << 16;
  std::vector<unsigned char> A(W * H, 0), B(W * H, 0);
  for (int j = 0; j < H; ++j)
    for (int i = 4; i < W; ++i) A[j * W + i] = 255;
  for (int j = 2; j < H; ++j)
    for (int i = 0; i < W; ++i) B[j * W + i] = 255;
  cudaTextureObject_t ta = make(A, W, H), tb = make(B, W, H);
  float* dout;
  cudaMalloc(&dout, N * 4);
  std::vector<float> hw(N);
  sample<<<(N + 255) / 256, 256>>>(ta, 3.5f, 1.0f / N, 1.5f, 0.f, dout, N);
  cudaMemcpy(hw.data(), dout, N * 4, cudaMemcpyDeviceToHost);
  report("x-sweep", hw, 3.5f, 1.0f / N, 3.0f);
  sample<<<(N + 255) / 256, 256>>>(tb, 2.5f, 0.f, 1.5f, 1.0f / N, dout, N);
  cudaMemcpy(hw.data(), dout, N * 4, cudaMemcpyDeviceToHost);
  report("y-sweep", hw, 1.5f, 1.0f / N, 1.0f);
  // far from the origin (coordinate magnitude ~ 600: fewer float mantissa bits left for the fraction)
  return cudaDeviceSynchronize() != cudaSuccess;
}
