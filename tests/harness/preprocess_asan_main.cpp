// TEST HARNESS (CPU suite only): drives tests/harness/preprocess_host.cpp over ragged image sizes and filter radii under
// -fsanitize=address,undefined, with exactly-sized buffers, so that an out-of-bounds access of the tile program is caught.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" int harness_preprocess_frame(int w, int h, const float depth_K[4], float raw_to_float, float a, int cell, int cf_w,
                             const float* cfactor, float sigma_xy, float sigma_inv_depth, float radius_factor, float max_depth_m,
                             const uint16_t* raw_depth, uint16_t* out_depth, uint16_t* out_normals, uint16_t* out_radius,
                             int cw, int ch, const uint8_t* rgb, uint8_t* rgba, float* min_max);
int main() {
  int sizes[][2] = {{70, 45}, {33, 31}, {8, 5}, {3, 3}, {1, 1}, {64, 64}, {65, 33}, {320, 240}};
  float sig[] = {1.5f, 0.2f, 8.0f};
  for (auto& s : sizes) for (float sg : sig) {
    int w = s[0], h = s[1], cell = 4, cf_w = (w - 1) / cell + 1, cf_h = (h - 1) / cell + 1;
    std::vector<float> cf(cf_w * cf_h, 1e-3f);
    std::vector<uint16_t> raw(w * h), d(w * h), n(w * h), r(w * h);
    std::vector<uint8_t> rgb(3 * w * h), rgba(4 * w * h);
    for (int i = 0; i < w * h; ++i) { raw[i] = (rand() % 20 == 0) ? 0 : 1500 + (i % w) + rand() % 4; }
    for (auto& v : rgb) v = rand() & 255;
    float K[4] = {0.5f * h + 7, 0.5f * h + 7, 0.5f * w - 0.5f, 0.5f * h - 0.5f}, mm[2];
    int t = harness_preprocess_frame(w, h, K, 1e-3f, 0.02f, cell, cf_w, cf.data(), sg, 0.005f, 2.0f, 3.0f, raw.data(), d.data(), n.data(),
                                     r.data(), w, h, rgb.data(), rgba.data(), mm);
    printf("%dx%d sigma %.1f: %d tiles, min %.3f max %.3f\n", w, h, sg, t, mm[0], mm[1]);
  }
  return 0;
}
