// tools/tex_probe.cu -- which arithmetic reproduces cudaFilterModeLinear on a normalized-float u8 texture bit-exactly?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tex_probe tools/tex_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__global__ void sample(cudaTextureObject_t t, const float2* xy, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex2D<float>(t, xy[i].x, xy[i].y);
}
int main() {
  const int W = 64, H = 48, N = 1 << 20;
  std::vector<unsigned char> img(W * H);
  srand(1);
  for (auto& v : img) v = rand() & 255;
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
  std::vector<float2> xy(N);
  for (auto& p : xy) {
    p.x = -2.f + (W + 4.f) * (rand() / (float)RAND_MAX);
    p.y = -2.f + (H + 4.f) * (rand() / (float)RAND_MAX);
  }
  float2* dxy;
  float* dout;
  cudaMalloc(&dxy, N * sizeof(float2));
  cudaMalloc(&dout, N * 4);
  cudaMemcpy(dxy, xy.data(), N * sizeof(float2), cudaMemcpyHostToDevice);
  sample<<<(N + 255) / 256, 256>>>(tex, dxy, dout, N);
  std::vector<float> hw(N);
  cudaMemcpy(hw.data(), dout, N * 4, cudaMemcpyDeviceToHost);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("cuda error\n");
    return 1;
  }
  auto T = [&](long i, long j) {
    i = i < 0 ? 0 : (i > W - 1 ? W - 1 : i);
    j = j < 0 ? 0 : (j > H - 1 ? H - 1 : j);
    return (long)img[j * W + i];
  };
  const char* names[] = {"float weights exact",
                         "w=round(a*256)/256, float sum",
                         "w=trunc(a*256)/256, float sum",
                         "integer numerator / (65536*255) [round weights]",
                         "integer numerator * (1/(65536*255)) float",
                         "coords to fixed point first: rint(x*256), integer",
                         "integer [trunc weights]",
                         "lerp form, rounded weights",
                         "unorm16 one-shot round, /65535.f",
                         "unorm16 two-stage x then y (round each), /65535.f",
                         "unorm16 two-stage y then x (round each), /65535.f",
                         "unorm16 one-shot trunc, /65535.f",
                         "unorm16 one-shot round, *(1/65535.f)",
                         "unorm16 two-stage x then y, *(1/65535.f)",
                         "8-bit product weights, sum rounded once",
                         "8-bit product weights, w11 = 256 - rest",
                         "8-bit product weights, per-texel rounding",
                         "8-bit product weights, sum truncated",
                         "hierarchical: rows then columns",
                         "hierarchical: columns then rows",
                         "hierarchical: from w11",
                         "8-bit product weights, ties to even"};
  const int NC = 22;
  long mism[NC] = {0};
  double maxd[NC] = {0};
  for (int k = 0; k < N; ++k) {
    float x = xy[k].x, y = xy[k].y;
    float xb = x - 0.5f, yb = y - 0.5f;
    float fi = floorf(xb), fj = floorf(yb);
    float a = xb - fi, b = yb - fj;
    long i = (long)fi, j = (long)fj;
    float t00 = T(i, j) / 255.f, t10 = T(i + 1, j) / 255.f, t01 = T(i, j + 1) / 255.f, t11 = T(i + 1, j + 1) / 255.f;
    float c[NC];
    c[0] = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
    float ar = floorf(a * 256 + 0.5f) / 256, br = floorf(b * 256 + 0.5f) / 256;
    c[1] = (1 - ar) * (1 - br) * t00 + ar * (1 - br) * t10 + (1 - ar) * br * t01 + ar * br * t11;
    float at = floorf(a * 256) / 256, bt = floorf(b * 256) / 256;
    c[2] = (1 - at) * (1 - bt) * t00 + at * (1 - bt) * t10 + (1 - at) * bt * t01 + at * bt * t11;
    long ai = (long)floorf(a * 256 + 0.5f), bi = (long)floorf(b * 256 + 0.5f);
    long num = (256 - ai) * (256 - bi) * T(i, j) + ai * (256 - bi) * T(i + 1, j) + (256 - ai) * bi * T(i, j + 1) + ai * bi * T(i + 1, j + 1);
    c[3] = (float)((double)num / (65536.0 * 255.0));
    c[4] = (float)num * (1.0f / (65536.f * 255.f));
    {
      long x8 = lrintf(x * 256.f) - 128, y8 = lrintf(y * 256.f) - 128;
      long ii = x8 >> 8, jj = y8 >> 8;
      long a8 = x8 & 255, b8 = y8 & 255;
      long n2 = (256 - a8) * (256 - b8) * T(ii, jj) + a8 * (256 - b8) * T(ii + 1, jj) + (256 - a8) * b8 * T(ii, jj + 1) + a8 * b8 * T(ii + 1, jj + 1);
      c[5] = (float)((double)n2 / (65536.0 * 255.0));
    }
    long ati = (long)floorf(a * 256), bti = (long)floorf(b * 256);
    long n3 = (256 - ati) * (256 - bti) * T(i, j) + ati * (256 - bti) * T(i + 1, j) + (256 - ati) * bti * T(i, j + 1) + ati * bti * T(i + 1, j + 1);
    c[6] = (float)((double)n3 / (65536.0 * 255.0));
    {
      float top = t00 + ar * (t10 - t00), bot = t01 + ar * (t11 - t01);
      c[7] = top + br * (bot - top);
    }
    {
      long T00 = T(i, j) * 257, T10 = T(i + 1, j) * 257, T01 = T(i, j + 1) * 257, T11 = T(i + 1, j + 1) * 257;
      long a8 = ai, b8 = bi;
      long sum = (256 - a8) * (256 - b8) * T00 + a8 * (256 - b8) * T10 + (256 - a8) * b8 * T01 + a8 * b8 * T11;
      long one = (sum + 32768) >> 16;
      c[8] = (float)one / 65535.f;
      long top = ((256 - a8) * T00 + a8 * T10 + 128) >> 8, bot = ((256 - a8) * T01 + a8 * T11 + 128) >> 8;
      long two = ((256 - b8) * top + b8 * bot + 128) >> 8;
      c[9] = (float)two / 65535.f;
      long lef = ((256 - b8) * T00 + b8 * T01 + 128) >> 8, rig = ((256 - b8) * T10 + b8 * T11 + 128) >> 8;
      long two2 = ((256 - a8) * lef + a8 * rig + 128) >> 8;
      c[10] = (float)two2 / 65535.f;
      c[11] = (float)(sum >> 16) / 65535.f;
      c[12] = (float)one * (1.0f / 65535.f);
      c[13] = (float)two * (1.0f / 65535.f);
      // per-texel weights = 8-bit rounded products of the 1.8 fixed-point fractions (tex_probe3: exact for one texel)
      long w00 = ((256 - a8) * (256 - b8) + 128) >> 8, w10 = (a8 * (256 - b8) + 128) >> 8;
      long w01 = ((256 - a8) * b8 + 128) >> 8, w11 = (a8 * b8 + 128) >> 8;
      c[14] = (float)((w00 * T00 + w10 * T10 + w01 * T01 + w11 * T11 + 128) >> 8) / 65535.f;
      long w11b = 256 - w00 - w10 - w01;
      c[15] = (float)((w00 * T00 + w10 * T10 + w01 * T01 + w11b * T11 + 128) >> 8) / 65535.f;
      c[16] = (float)(((w00 * T00 + 128) >> 8) + ((w10 * T10 + 128) >> 8) + ((w01 * T01 + 128) >> 8) + ((w11 * T11 + 128) >> 8)) / 65535.f;
      c[17] = (float)((w00 * T00 + w10 * T10 + w01 * T01 + w11 * T11) >> 8) / 65535.f;
      {  // hierarchical splits that keep the weight sum at exactly 256
        long r0 = 256 - b8, r1 = b8;   // row weights, split between columns
        long h00 = (r0 * (256 - a8) + 128) >> 8, h10 = r0 - h00, h01 = (r1 * (256 - a8) + 128) >> 8, h11 = r1 - h01;
        c[18] = (float)((h00 * T00 + h10 * T10 + h01 * T01 + h11 * T11 + 128) >> 8) / 65535.f;
        long c0 = 256 - a8, c1 = a8;   // column weights, split between rows
        long g00 = (c0 * (256 - b8) + 128) >> 8, g01 = c0 - g00, g10 = (c1 * (256 - b8) + 128) >> 8, g11 = c1 - g10;
        c[19] = (float)((g00 * T00 + g10 * T10 + g01 * T01 + g11 * T11 + 128) >> 8) / 65535.f;
        // round the "far" weights instead
        long k11 = (a8 * b8 + 128) >> 8, k10 = c1 - k11, k01 = r1 - k11, k00 = 256 - k11 - k10 - k01;
        c[20] = (float)((k00 * T00 + k10 * T10 + k01 * T01 + k11 * T11 + 128) >> 8) / 65535.f;
        // ties to even on each product
        auto rne = [](long v) { long q = v >> 8, r = v & 255; return (r > 128 || (r == 128 && (q & 1))) ? q + 1 : q; };
        long e00 = rne((256 - a8) * (256 - b8)), e10 = rne(a8 * (256 - b8)), e01 = rne((256 - a8) * b8), e11 = rne(a8 * b8);
        c[21] = (float)((e00 * T00 + e10 * T10 + e01 * T01 + e11 * T11 + 128) >> 8) / 65535.f;
      }
    }
    for (int q = 0; q < NC; ++q) {
      if (c[q] != hw[k]) mism[q]++;
      double dd = fabs((double)c[q] - hw[k]);
      if (dd > maxd[q]) maxd[q] = dd;
    }
  }
  for (int q = 0; q < NC; ++q) printf("%-55s mismatches %8ld / %d   max|diff| %.3e\n", names[q], mism[q], N, maxd[q]);
  return 0;
}
