// tools/tex_probe3.cu -- 2-D weight of a single texel: texture all 0 except texel (1,1) = 255.
// For (x, y) in [1.5, 2.5)^2: tex2D = w00_hw(alpha, beta), alpha = x - 1.5, beta = y - 1.5.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
__global__ void sample(cudaTextureObject_t t, const float2* xy, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex2D<float>(t, xy[i].x, xy[i].y);
}
int main() {
  const int W = 4, H = 4, G = 512, N = G * G;
  std::vector<unsigned char> img(W * H, 0);
  img[1 * W + 1] = 255;
  unsigned char* d; size_t pitch;
  cudaMallocPitch((void**)&d, &pitch, W, H);
  cudaMemcpy2D(d, pitch, img.data(), W, W, H, cudaMemcpyHostToDevice);
  cudaResourceDesc res; memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D; res.res.pitch2D.devPtr = d;
  res.res.pitch2D.desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
  res.res.pitch2D.width = W; res.res.pitch2D.height = H; res.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &res, &td, nullptr);
  std::vector<float2> xy(N);
  for (int j = 0; j < G; ++j) for (int i = 0; i < G; ++i) { xy[j * G + i].x = 1.5f + (i + 0.37f) / G; xy[j * G + i].y = 1.5f + (j + 0.61f) / G; }
  float2* dxy; float* dout; cudaMalloc(&dxy, N * sizeof(float2)); cudaMalloc(&dout, N * 4);
  cudaMemcpy(dxy, xy.data(), N * sizeof(float2), cudaMemcpyHostToDevice);
  sample<<<(N + 255) / 256, 256>>>(tex, dxy, dout, N);
  std::vector<float> hw(N); cudaMemcpy(hw.data(), dout, N * 4, cudaMemcpyDeviceToHost);
  // hypotheses on r16 = hw * 65535
  long mism[6] = {0};
  for (int k = 0; k < N; ++k) {
    float a = xy[k].x - 0.5f - 1.0f, b = xy[k].y - 0.5f - 1.0f;
    long ai = (long)floorf(a * 256 + 0.5f), bi = (long)floorf(b * 256 + 0.5f);
    long r16 = lrint(hw[k] * 65535.0);
    long T16 = 65535;
    long h0 = ((256 - ai) * (256 - bi) * T16 + 32768) >> 16;                 // one-shot
    long top = ((256 - ai) * T16 + 128) >> 8; long h1 = ((256 - bi) * top + 128) >> 8;   // x then y
    long lef = ((256 - bi) * T16 + 128) >> 8; long h2 = ((256 - ai) * lef + 128) >> 8;   // y then x
    long w8 = ((256 - ai) * (256 - bi) + 128) >> 8; long h3 = (w8 * T16 + 128) >> 8;     // weight product rounded to 8 bits
    long w9 = ((256 - ai) * (256 - bi) + 64) >> 7; long h4 = (w9 * T16 + 256) >> 9;       // ... to 9 bits
    double we = (1.0 - a) * (1.0 - b); long h5 = lrint(we * 65535.0);                       // exact weights
    long hs[6] = {h0, h1, h2, h3, h4, h5};
    for (int q = 0; q < 6; ++q) if (hs[q] != r16) mism[q]++;
    if (k % 40009 == 0) printf("a8=%3ld b8=%3ld  a=%.5f b=%.5f hw16=%6ld  oneshot=%6ld xy=%6ld w8=%6ld w9=%6ld exact=%6ld\n", ai, bi, a, b, r16, h0, h1, h3, h4, h5);
  }
  const char* nm[6] = {"one-shot", "x-then-y", "y-then-x", "8-bit product weight", "9-bit product weight", "exact weights"};
  for (int q = 0; q < 6; ++q) printf("%-22s mismatches %ld / %d\n", nm[q], mism[q], N);
  return 0;
}
