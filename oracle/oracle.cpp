// oracle.cpp — CPU restatement of redwarp/kmeans-gpu's image hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it; the product library
// (kmeans-gpu_b200/csrc) never links, loads or calls anything in this directory.
//
// Parity status: PINNED by the reference's committed golden images (tests/golden/, copied
// from /root/reference/gfx) and its shader-test known answers (core/src/shader_tests.rs:169-241):
//   * find replace / find dither (3 colours and 46 colours) reproduce bit-exactly,
//   * reduce/palette goldens reproduce to +-1/255 per palette channel (the author's GPU
//     differs from IEEE f32 in the 4th digit of the centroids).
// The reference itself (Rust + wgpu) cannot be built in this image (no cargo, no Vulkan), so
// kind == "port".  Arithmetic contract: IEEE-754 binary32, every operation rounded separately
// (compile with -ffp-contract=off), sqrt and divide correctly rounded, and
//   pow_f32(x, y) := (float) pow((double) x, (double) y)          (correctly rounded f32 pow).
// Third-party arithmetic not in /root/reference: the `palette` crate 0.7.3 (Cargo.lock:846)
// sRGB<->Lab conversions used for fixed palettes and for `palette()` output are restated from
// the crate's published formulas in pal_* below.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

inline float pow_f32(float x, float y) { return (float)std::pow((double)x, (double)y); }

struct V3 {
  float x, y, z;
};

// ---------------------------------------------------------------------------------------------
// core/shaders/functions/delta_e.wgsl:1-22  — CIE94, asymmetric (SC, SH from the FIRST argument).
inline float cie94(const float* one, const float* second) {
  const float K1 = 0.045f, K2 = 0.015f;
  float dL = one[0] - second[0];
  float da = one[1] - second[1];
  float db = one[2] - second[2];
  float C1 = std::sqrt(one[1] * one[1] + one[2] * one[2]);
  float C2 = std::sqrt(second[1] * second[1] + second[2] * second[2]);
  float dCab = C1 - C2;
  float dHab = std::sqrt(std::max((da * da) + (db * db) - (dCab * dCab), 0.0f));
  float SL = 1.0f;
  float SC = 1.0f + K1 * C1;
  float SH = 1.0f + K2 * C1;
  float tL = dL / SL, tC = dCab / SC, tH = dHab / SH;
  return std::sqrt(tL * tL + tC * tC + tH * tH);
}

// core/shaders/converters/rgb_to_lab.wgsl:11-37 — sRGB decode (x100) and Lindbloom matrix.
inline float srgb_decode(float c) {
  if (c > 0.04045f) return pow_f32((c + 0.055f) / 1.055f, 2.4f);
  return c / 12.92f;
}
// core/shaders/converters/rgb_to_lab.wgsl:39-64
inline float lab_f(float t) {
  if (t > 0.008856f) return pow_f32(t, 1.0f / 3.0f);
  return (7.787f * t) + (16.0f / 116.0f);
}
inline void rgb8_to_lab(const uint8_t* px, float* out) {
  float r = srgb_decode((float)px[0] / 255.0f) * 100.0f;
  float g = srgb_decode((float)px[1] / 255.0f) * 100.0f;
  float b = srgb_decode((float)px[2] / 255.0f) * 100.0f;
  // mat3x3 * vec3 = col0*r + col1*g + col2*b, summed left to right.
  float X = (0.4124564f * r + 0.3575761f * g) + 0.1804375f * b;
  float Y = (0.2126729f * r + 0.7151522f * g) + 0.0721750f * b;
  float Z = (0.0193339f * r + 0.1191920f * g) + 0.9503041f * b;
  float x = lab_f(X / 95.0489f);
  float y = lab_f(Y / 100.0f);
  float z = lab_f(Z / 108.8840f);
  out[0] = (116.0f * y) - 16.0f;
  out[1] = 500.0f * (x - y);
  out[2] = 200.0f * (y - z);
  out[3] = 1.0f;
}

// rgba8unorm store: clamp to [0,1], scale, round to nearest (ties to even).
inline uint8_t unorm8(float v) {
  if (!(v > 0.0f)) return 0;  // also NaN
  if (v > 1.0f) v = 1.0f;
  return (uint8_t)std::nearbyintf(v * 255.0f);
}

// core/shaders/converters/lab_to_rgb.wgsl:11-66
inline float lab_finv(float t) {
  float t3 = pow_f32(t, 3.0f);
  if (t3 > 0.008856f) return t3;
  return (t - 16.0f / 116.0f) / 7.787f;
}
inline float srgb_encode(float c) {
  if (c > 0.0031308f) return 1.055f * pow_f32(c, 1.0f / 2.4f) - 0.055f;
  return 12.92f * c;
}
inline void lab_to_rgba8(const float* lab, uint8_t* out) {
  float y = (lab[0] + 16.0f) / 116.0f;
  float x = lab[1] / 500.0f + y;
  float z = y - lab[2] / 200.0f;
  x = lab_finv(x) * 95.0489f;
  y = lab_finv(y) * 100.0f;
  z = lab_finv(z) * 108.8840f;
  x = x / 100.0f;
  y = y / 100.0f;
  z = z / 100.0f;
  float r = (3.2404542f * x + -1.5371385f * y) + -0.4985314f * z;
  float g = (-0.9692660f * x + 1.8760108f * y) + 0.0415560f * z;
  float b = (0.0556434f * x + -0.2040259f * y) + 1.0572252f * z;
  out[0] = unorm8(srgb_encode(r));
  out[1] = unorm8(srgb_encode(g));
  out[2] = unorm8(srgb_encode(b));
  out[3] = 255;  // alpha = 1.0 (lab_to_rgb.wgsl:37)
}

// core/shaders/converters/rgb8u_to_rgb32f.wgsl:4-17 / rgb32f_to_rgb8u.wgsl:4-17 (ColorSpace::Rgb)
inline void rgb8_to_rgbf(const uint8_t* px, float* out) {
  out[0] = (float)px[0] / 255.0f;
  out[1] = (float)px[1] / 255.0f;
  out[2] = (float)px[2] / 255.0f;
  out[3] = (float)px[3] / 255.0f;
}
inline void rgbf_to_rgba8(const float* c, uint8_t* out) {
  out[0] = unorm8(c[0]);
  out[1] = unorm8(c[1]);
  out[2] = unorm8(c[2]);
  out[3] = unorm8(c[3]);
}

inline void to_work(const uint8_t* px, int color_space, float* out) {
  if (color_space == 0)
    rgb8_to_lab(px, out);
  else
    rgb8_to_rgbf(px, out);
}
inline void from_work(const float* c, int color_space, uint8_t* out) {
  if (color_space == 0)
    lab_to_rgba8(c, out);
  else
    rgbf_to_rgba8(c, out);
}

// ---------------------------------------------------------------------------------------------
// `palette` crate 0.7.3 (not vendored): Srgb<u8> -> Lab<D65,f32> used by
// CentroidsBuffer::fixed_centroids (core/src/structures.rs:523-553) and the palette sort key
// (core/src/lib.rs:276-284); Lab -> Srgb<u8> used by pull_values (core/src/structures.rs:581-617).
inline void pal_srgb8_to_lab(const uint8_t* px, float* out) {
  float lin[3];
  for (int i = 0; i < 3; ++i) {
    float c = (float)px[i] / 255.0f;
    lin[i] = (c <= 0.04045f) ? c / 12.92f : pow_f32((c + 0.055f) / 1.055f, 2.4f);
  }
  float X = (0.4124564f * lin[0] + 0.3575761f * lin[1]) + 0.1804375f * lin[2];
  float Y = (0.2126729f * lin[0] + 0.7151522f * lin[1]) + 0.0721750f * lin[2];
  float Z = (0.0193339f * lin[0] + 0.1191920f * lin[1]) + 0.9503041f * lin[2];
  const float eps = (float)((6.0 / 29.0) * (6.0 / 29.0) * (6.0 / 29.0));
  const float kappa = (float)(841.0 / 108.0);
  const float delta = (float)(4.0 / 29.0);
  auto f = [&](float c) { return c > eps ? std::cbrt(c) : (kappa * c) + delta; };
  float x = f(X / 0.95047f), y = f(Y / 1.0f), z = f(Z / 1.08883f);
  out[0] = (116.0f * y) - 16.0f;
  out[1] = 500.0f * (x - y);
  out[2] = 200.0f * (y - z);
}
inline void pal_lab_to_srgb8(const float* lab, uint8_t* out) {
  float y = (lab[0] + 16.0f) / 116.0f;
  float x = y + (lab[1] / 500.0f);
  float z = y - (lab[2] / 200.0f);
  const float eps = (float)(6.0 / 29.0);
  const float kappa = (float)(108.0 / 841.0);
  const float delta = (float)(4.0 / 29.0);
  auto finv = [&](float c) { return c > eps ? c * c * c : (c - delta) * kappa; };
  float X = finv(x) * 0.95047f, Y = finv(y) * 1.0f, Z = finv(z) * 1.08883f;
  float lin[3];
  lin[0] = (3.2404542f * X + -1.5371385f * Y) + -0.4985314f * Z;
  lin[1] = (-0.9692660f * X + 1.8760108f * Y) + 0.0415560f * Z;
  lin[2] = (0.0556434f * X + -0.2040259f * Y) + 1.0572252f * Z;
  for (int i = 0; i < 3; ++i) {
    float c = lin[i];
    float e = (c <= 0.0031308f) ? 12.92f * c : 1.055f * pow_f32(c, 1.0f / 2.4f) - 0.055f;
    out[i] = unorm8(e);
  }
  out[3] = 255;
}

// ---------------------------------------------------------------------------------------------
// core/src/structures.rs:67-89 — shrink size rule (note the strict `width > height`).
inline void resized_dims(uint32_t w, uint32_t h, uint32_t max_size, uint32_t* nw, uint32_t* nh) {
  if (w > h) {
    *nw = max_size;
    *nh = std::max<uint32_t>((uint32_t)((float)h * (float)max_size / (float)w), 1u);
  } else {
    *nw = std::max<uint32_t>((uint32_t)((float)w * (float)max_size / (float)h), 1u);
    *nh = max_size;
  }
}

// core/shaders/resize.wgsl:5-19 with the linear / clamp-to-edge sampler of
// core/src/structures.rs:122-133.  Sample position (gx/dw, gy/dh) — no half-texel offset.
// Bilinear filter stated in f32: unnormalised coordinate u*sw - 0.5, floor/fract, clamped taps,
// weights (1-f, f), x first then y; unorm8 store.
void resize_image(const uint8_t* src, uint32_t sw, uint32_t sh, uint8_t* dst, uint32_t dw, uint32_t dh) {
#pragma omp parallel for schedule(static)
  for (int64_t gy = 0; gy < (int64_t)dh; ++gy) {
    float v = (float)gy / (float)dh;
    float py = v * (float)sh - 0.5f;
    float fy0 = std::floor(py);
    float fy = py - fy0;
    int64_t y0 = (int64_t)fy0, y1 = y0 + 1;
    y0 = std::min<int64_t>(std::max<int64_t>(y0, 0), sh - 1);
    y1 = std::min<int64_t>(std::max<int64_t>(y1, 0), sh - 1);
    for (uint32_t gx = 0; gx < dw; ++gx) {
      float u = (float)gx / (float)dw;
      float px = u * (float)sw - 0.5f;
      float fx0 = std::floor(px);
      float fx = px - fx0;
      int64_t x0 = (int64_t)fx0, x1 = x0 + 1;
      x0 = std::min<int64_t>(std::max<int64_t>(x0, 0), sw - 1);
      x1 = std::min<int64_t>(std::max<int64_t>(x1, 0), sw - 1);
      const uint8_t* p00 = src + 4 * ((size_t)y0 * sw + x0);
      const uint8_t* p10 = src + 4 * ((size_t)y0 * sw + x1);
      const uint8_t* p01 = src + 4 * ((size_t)y1 * sw + x0);
      const uint8_t* p11 = src + 4 * ((size_t)y1 * sw + x1);
      uint8_t* o = dst + 4 * ((size_t)gy * dw + gx);
      for (int c = 0; c < 4; ++c) {
        float c00 = (float)p00[c] / 255.0f, c10 = (float)p10[c] / 255.0f;
        float c01 = (float)p01[c] / 255.0f, c11 = (float)p11[c] / 255.0f;
        float top = c00 * (1.0f - fx) + c10 * fx;
        float bot = c01 * (1.0f - fx) + c11 * fx;
        o[c] = unorm8(top * (1.0f - fy) + bot * fy);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// core/shaders/find_centroid.wgsl:27-43 — argmin, strict '<' from (100000.0, index 0).
inline uint32_t assign_one(const float* px, const float* cent, uint32_t k, float* best, float* second) {
  float min_distance = 100000.0f, second_distance = INFINITY;
  uint32_t found = 0;
  for (uint32_t i = 0; i < k; ++i) {
    float d = cie94(px, cent + 4 * i);
    if (d < min_distance) {
      second_distance = min_distance;
      min_distance = d;
      found = i;
    } else if (d < second_distance) {
      second_distance = d;
    }
  }
  if (best) *best = min_distance;
  if (second) *second = second_distance;
  return found;
}

constexpr double FIXED_SCALE = 32768.0;  // 2^15 fixed-point units per colour unit

inline int64_t to_fixed(float v) { return (int64_t)std::nearbyintf(v * 32768.0f); }

// core/shaders/choose_centroid.wgsl:73-206 — per-cluster mean + convergence flags (the scan
// machinery is an implementation detail; its f32 summation order is timing dependent, so two
// well-defined sums are offered):
//   sum_mode 0: f64 accumulation of the f32 values, divide, cast to f32;
//   sum_mode 1: exact integer accumulation of rint(v * 2^15) (order independent, what the CUDA
//               library does), mean = (double) sum / (double) count * 2^-15, cast to f32.
// Empty clusters keep their centroid and contribute 0 to the convergence count (:185-194).
uint32_t update_centroids(const float* work, const uint32_t* labels, size_t n, uint32_t k, float* cent,
                          float threshold, int sum_mode, uint64_t* counts_out) {
  std::vector<double> sum((size_t)k * 3, 0.0);
  std::vector<int64_t> isum((size_t)k * 3, 0);
  std::vector<uint64_t> cnt(k, 0);
  for (size_t p = 0; p < n; ++p) {
    uint32_t l = labels[p];
    const float* v = work + 4 * p;
    if (sum_mode == 0) {
      sum[3 * l + 0] += (double)v[0];
      sum[3 * l + 1] += (double)v[1];
      sum[3 * l + 2] += (double)v[2];
    } else {
      isum[3 * l + 0] += to_fixed(v[0]);
      isum[3 * l + 1] += to_fixed(v[1]);
      isum[3 * l + 2] += to_fixed(v[2]);
    }
    cnt[l]++;
  }
  uint32_t converged = 0;
  for (uint32_t c = 0; c < k; ++c) {
    if (counts_out) counts_out[c] = cnt[c];
    if (cnt[c] == 0) continue;
    float nc[4];
    for (int j = 0; j < 3; ++j) {
      if (sum_mode == 0)
        nc[j] = (float)(sum[3 * c + j] / (double)cnt[c]);
      else
        nc[j] = (float)(((double)isum[3 * c + j] / (double)cnt[c]) * (1.0 / FIXED_SCALE));
    }
    nc[3] = 1.0f;
    float prev[3] = {cent[4 * c], cent[4 * c + 1], cent[4 * c + 2]};
    std::memcpy(cent + 4 * c, nc, sizeof(nc));
    converged += (uint32_t)(cie94(nc, prev) < threshold);  // new first, previous second (:191)
  }
  return converged;
}

// core/shaders/plus_plus_init.wgsl:58-187 + kmeans++_calc_diff.wgsl:14-34 + host loop
// core/src/modules.rs:946-1284 — deterministic farthest-point initialisation.
// Tie rule (selectCandidate, plus_plus_init.wgsl:62-68, applied at :92,:102,:136,:142): every
// thread folds 16 consecutive pixels starting from (index 0, distance 0.0) and keeps the
// EARLIEST strict maximum; across threads and workgroups the LATER candidate wins ties.
void init_centroids(const float* work, uint32_t w, uint32_t h, uint32_t k, int32_t seed_x, int32_t seed_y,
                    float* cent, uint32_t* pick_index, float* pick_dist) {
  const size_t n = (size_t)w * h;
  size_t seed = (size_t)seed_y * w + (size_t)seed_x;
  std::memcpy(cent, work + 4 * seed, 4 * sizeof(float));  // :166-167 (copies the texel incl. w)
  if (pick_index) pick_index[0] = (uint32_t)seed;
  if (pick_dist) pick_dist[0] = 0.0f;
  std::vector<float> dmin(n, 1000000.0f);  // kmeans++_calc_diff.wgsl:27
  const size_t n_chunks = (n + 15) / 16;   // N_SEQ = 16 (plus_plus_init.wgsl:15)
  for (uint32_t j = 1; j < k; ++j) {
    const float* cprev = cent + 4 * (j - 1);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)n; ++p) dmin[p] = std::min(dmin[p], cie94(work + 4 * p, cprev));
    uint32_t best_idx = 0;
    float best_d = 0.0f;
    bool have = false;
    for (size_t c = 0; c < n_chunks; ++c) {
      uint32_t li = 0;
      float ld = 0.0f;
      size_t end = std::min(n, (c + 1) * 16);
      for (size_t p = c * 16; p < end; ++p)
        if (ld < dmin[p]) {
          ld = dmin[p];
          li = (uint32_t)p;
        }
      // fold: select(local = later, value = earlier) keeps `later` unless later.d < earlier.d
      if (!have || !(ld < best_d)) {
        best_d = ld;
        best_idx = li;
        have = true;
      }
    }
    cent[4 * j + 0] = work[4 * (size_t)best_idx + 0];
    cent[4 * j + 1] = work[4 * (size_t)best_idx + 1];
    cent[4 * j + 2] = work[4 * (size_t)best_idx + 2];
    cent[4 * j + 3] = 1.0f;  // plus_plus_init.wgsl:177
    if (pick_index) pick_index[j] = best_idx;
    if (pick_dist) pick_dist[j] = best_d;
  }
}

// core/shaders/mix_colors.wgsl:14-27
const uint32_t kBayer[16] = {0, 8, 2, 10, 12, 4, 14, 6, 3, 11, 1, 9, 15, 7, 13, 5};

// core/shaders/mix_colors.wgsl:53-68 — greedy farthest-pair scan -> dither threshold.
float dither_threshold(const float* cent, uint32_t k) {
  const float* a = cent;
  const float* b = cent + 4;
  float d_ab = cie94(a, b);
  for (uint32_t i = 2; i < k; ++i) {
    const float* ci = cent + 4 * i;
    float da = cie94(ci, a);
    float db = cie94(ci, b);
    if (da > db && da > d_ab) {
      d_ab = da;
      b = ci;
    } else if (db > d_ab) {
      d_ab = db;
      a = ci;
    }
  }
  return d_ab / std::sqrt((float)k);
}

}  // namespace

// =============================================================================================
// Exported C interface (loaded with ctypes by tests/ and bench.py's CPU-baseline legs only).

ORC_API int orc_num_threads() {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API float orc_pow_f32(float x, float y) { return pow_f32(x, y); }
ORC_API float orc_cie94(const float* one, const float* second) { return cie94(one, second); }
ORC_API float orc_srgb_decode(uint8_t v) { return srgb_decode((float)v / 255.0f); }

// K1/K3: RGBA8 -> work plane (f32x4).  color_space 0 = Lab, 1 = Rgb.
ORC_API void orc_convert(const uint8_t* rgba, size_t n, int color_space, float* work) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)n; ++p) to_work(rgba + 4 * p, color_space, work + 4 * p);
}
// K2/K4: work plane -> RGBA8.
ORC_API void orc_revert(const float* work, size_t n, int color_space, uint8_t* rgba) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)n; ++p) from_work(work + 4 * p, color_space, rgba + 4 * p);
}

// K5.  best/second may be null; they carry the winning and runner-up distances so tests can
// classify any label disagreement as a near-tie.
ORC_API void orc_assign(const float* work, size_t n, const float* cent, uint32_t k, uint32_t* labels, float* best,
                        float* second) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)n; ++p)
    labels[p] = assign_one(work + 4 * p, cent, k, best ? best + p : nullptr, second ? second + p : nullptr);
}

ORC_API uint32_t orc_update(const float* work, const uint32_t* labels, size_t n, uint32_t k, float* cent,
                            float threshold, int sum_mode, uint64_t* counts_out) {
  return update_centroids(work, labels, n, k, cent, threshold, sum_mode, counts_out);
}

// Exact fixed-point partial sums of one shard (k x 4 int64: sum0, sum1, sum2, count) — the
// quantity the multi-GPU path all-reduces.
ORC_API void orc_partial_sums(const float* work, const uint32_t* labels, size_t n, uint32_t k, int64_t* acc) {
  std::memset(acc, 0, sizeof(int64_t) * 4 * k);
  for (size_t p = 0; p < n; ++p) {
    uint32_t l = labels[p];
    acc[4 * l + 0] += to_fixed(work[4 * p + 0]);
    acc[4 * l + 1] += to_fixed(work[4 * p + 1]);
    acc[4 * l + 2] += to_fixed(work[4 * p + 2]);
    acc[4 * l + 3] += 1;
  }
}
// Finalise reduced sums -> centroids + convergence count (same arithmetic as sum_mode 1).
ORC_API uint32_t orc_finalize(const int64_t* acc, uint32_t k, float* cent, float threshold) {
  uint32_t converged = 0;
  for (uint32_t c = 0; c < k; ++c) {
    int64_t cnt = acc[4 * c + 3];
    if (cnt == 0) continue;
    float nc[4];
    for (int j = 0; j < 3; ++j) nc[j] = (float)(((double)acc[4 * c + j] / (double)cnt) * (1.0 / FIXED_SCALE));
    nc[3] = 1.0f;
    float prev[3] = {cent[4 * c], cent[4 * c + 1], cent[4 * c + 2]};
    std::memcpy(cent + 4 * c, nc, sizeof(nc));
    converged += (uint32_t)(cie94(nc, prev) < threshold);
  }
  return converged;
}

ORC_API void orc_init(const float* work, uint32_t w, uint32_t h, uint32_t k, int32_t seed_x, int32_t seed_y,
                      float* cent, uint32_t* pick_index, float* pick_dist) {
  init_centroids(work, w, h, k, seed_x, seed_y, cent, pick_index, pick_dist);
}

ORC_API void orc_resized_dims(uint32_t w, uint32_t h, uint32_t max_size, uint32_t* nw, uint32_t* nh) {
  resized_dims(w, h, max_size, nw, nh);
}
ORC_API void orc_resize(const uint8_t* src, uint32_t sw, uint32_t sh, uint8_t* dst, uint32_t dw, uint32_t dh) {
  resize_image(src, sw, sh, dst, dw, dh);
}

// core/shaders/plus_plus_init.wgsl:159-165 with correctly rounded f32 sin:
// rand(42) = 0.5625, rand(12) = 0.93359375.
ORC_API void orc_seed_pixel(uint32_t w, uint32_t h, float fx, float fy, int32_t* sx, int32_t* sy) {
  *sx = (int32_t)((float)w * fx);
  *sy = (int32_t)((float)h * fy);
}

struct OrcOpts {
  uint32_t max_dim;      // 256 (core/src/structures.rs:23); 0 = never shrink
  uint32_t max_iter;     // 128 (core/src/modules.rs:765)
  uint32_t check_every;  // 8   (core/src/modules.rs:766)
  float convergence;     // 1.0 Lab / 0.01 Rgb (core/src/lib.rs:189-194); < 0 = default
  float seed_x_frac;     // 0.5625
  float seed_y_frac;     // 0.93359375
  int32_t seed_x;        // >= 0 overrides the fraction
  int32_t seed_y;
  int32_t sum_mode;      // 0 f64, 1 fixed-point
};

// core/src/operations.rs:15-88 (extract_palette_kmeans) + core/src/modules.rs:763-840 (Lloyd
// loop / stop rule).  Returns the number of update passes executed.  `trace` (optional,
// max_iter*k*4 floats) receives the centroids after every update pass.
ORC_API uint32_t orc_kmeans(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int color_space,
                            const OrcOpts* o, float* cent, float* trace) {
  std::vector<uint8_t> shrunk;
  const uint8_t* img = rgba;
  uint32_t iw = w, ih = h;
  if (o->max_dim != 0 && (w > o->max_dim || h > o->max_dim)) {  // structures.rs:67-74
    resized_dims(w, h, o->max_dim, &iw, &ih);
    shrunk.resize((size_t)iw * ih * 4);
    resize_image(rgba, w, h, shrunk.data(), iw, ih);
    img = shrunk.data();
  }
  const size_t n = (size_t)iw * ih;
  std::vector<float> work(n * 4);
  orc_convert(img, n, color_space, work.data());
  int32_t sx = o->seed_x >= 0 ? o->seed_x : (int32_t)((float)iw * o->seed_x_frac);
  int32_t sy = o->seed_y >= 0 ? o->seed_y : (int32_t)((float)ih * o->seed_y_frac);
  std::memset(cent, 0, sizeof(float) * 4 * k);  // empty_centroids (structures.rs:501-521)
  init_centroids(work.data(), iw, ih, k, sx, sy, cent, nullptr, nullptr);
  float threshold = o->convergence >= 0.0f ? o->convergence : (color_space == 0 ? 1.0f : 0.01f);
  std::vector<uint32_t> labels(n);
  orc_assign(work.data(), n, cent, k, labels.data(), nullptr, nullptr);  // operations.rs:75-83
  uint32_t passes = 0;
  for (uint32_t it = 0; it < o->max_iter; ++it) {
    uint32_t conv = update_centroids(work.data(), labels.data(), n, k, cent, threshold, o->sum_mode, nullptr);
    passes = it + 1;
    if (trace) std::memcpy(trace + (size_t)it * k * 4, cent, sizeof(float) * 4 * k);
    bool check = it > 0 && o->check_every != 0 && it % o->check_every == 0;  // modules.rs:802
    if (check && conv >= k) break;                                           // modules.rs:827
    orc_assign(work.data(), n, cent, k, labels.data(), nullptr, nullptr);    // modules.rs:793-800
  }
  return passes;
}

// core/src/operations.rs:215-271 (find_colors: convert, assign, swap, revert).
ORC_API void orc_remap_replace(const uint8_t* rgba, size_t n, const float* cent, uint32_t k, int color_space,
                               uint8_t* out, uint32_t* labels_out, float* best_out, float* second_out) {
  std::vector<uint8_t> pal((size_t)k * 4);
  for (uint32_t c = 0; c < k; ++c) from_work(cent + 4 * c, color_space, pal.data() + 4 * c);
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)n; ++p) {
    float px[4];
    to_work(rgba + 4 * p, color_space, px);
    float b, s;
    uint32_t l = assign_one(px, cent, k, &b, &s);
    std::memcpy(out + 4 * p, pal.data() + 4 * l, 4);
    if (labels_out) labels_out[p] = l;
    if (best_out) best_out[p] = b;
    if (second_out) second_out[p] = s;
  }
}

ORC_API float orc_dither_threshold(const float* cent, uint32_t k) { return dither_threshold(cent, k); }

// core/src/operations.rs:99-155 + core/shaders/mix_colors.wgsl:50-83,92-113 (ordered dither).
ORC_API void orc_remap_dither(const uint8_t* rgba, uint32_t w, uint32_t h, const float* cent, uint32_t k,
                              int color_space, uint8_t* out, uint32_t* labels_out, float* best_out,
                              float* second_out) {
  std::vector<uint8_t> pal((size_t)k * 4);
  for (uint32_t c = 0; c < k; ++c) {
    float c4[4] = {cent[4 * c], cent[4 * c + 1], cent[4 * c + 2], 1.0f};  // vec4(closest, 1.0) :82
    from_work(c4, color_space, pal.data() + 4 * c);
  }
  const float threshold = k > 1 ? dither_threshold(cent, k) : 0.0f;
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      size_t p = (size_t)y * w + x;
      uint32_t l = 0;
      float b = 0.0f, s = INFINITY;
      if (k > 1) {  // :104-108 short-circuit for a single colour
        float px[4];
        to_work(rgba + 4 * p, color_space, px);
        float iv = (float)kBayer[(x % 4) + (y % 4) * 4] / 16.0f - 0.5f;
        float off = threshold * iv;
        float adj[3] = {px[0] + off, px[1] + off, px[2] + off};
        // :74-81 — nearest by strict '<' starting from the sentinel colour (10000,10000,10000)
        const float sentinel[3] = {10000.0f, 10000.0f, 10000.0f};
        float bd = cie94(adj, sentinel);
        bool have = false;
        for (uint32_t i = 0; i < k; ++i) {
          float d = cie94(adj, cent + 4 * i);
          if (d < bd) {
            s = have ? bd : s;
            bd = d;
            l = i;
            have = true;
          } else if (d < s) {
            s = d;
          }
        }
        b = bd;
      }
      std::memcpy(out + 4 * p, pal.data() + 4 * l, 4);
      if (labels_out) labels_out[p] = l;
      if (best_out) best_out[p] = b;
      if (second_out) second_out[p] = s;
    }
}

// core/src/operations.rs:157-213 + core/shaders/mix_colors.wgsl:29-48,85-90,115-136 (meld).
ORC_API void orc_remap_meld(const uint8_t* rgba, uint32_t w, uint32_t h, const float* cent, uint32_t k,
                            int color_space, uint8_t* out) {
  const size_t n = (size_t)w * h;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)n; ++p) {
    float res[4];
    if (k == 1) {
      std::memcpy(res, cent, sizeof(res));
    } else {
      float px[4];
      to_work(rgba + 4 * p, color_space, px);
      float closest[4] = {10000.0f, 10000.0f, 10000.0f, 10000.0f};
      float second[4] = {10000.0f, 10000.0f, 10000.0f, 10000.0f};
      for (uint32_t i = 0; i < k; ++i) {
        const float* t = cent + 4 * i;
        float td = cie94(px, t);
        if (td < cie94(px, closest)) {
          std::memcpy(second, closest, sizeof(second));
          std::memcpy(closest, t, sizeof(closest));
        } else if (td < cie94(px, second)) {
          std::memcpy(second, t, sizeof(second));
        }
      }
      float factor = cie94(px, second) / cie94(closest, second);
      for (int c = 0; c < 4; ++c) res[c] = factor * closest[c] + (1.0f - factor) * second[c];
    }
    from_work(res, color_space, out + 4 * p);
  }
}

// palette-crate helpers (R8).
ORC_API void orc_pal_srgb8_to_lab(const uint8_t* rgba, uint32_t count, float* lab4) {
  for (uint32_t i = 0; i < count; ++i) {
    pal_srgb8_to_lab(rgba + 4 * i, lab4 + 4 * i);
    lab4[4 * i + 3] = 1.0f;
  }
}
ORC_API void orc_pal_lab_to_srgb8(const float* lab4, uint32_t count, uint8_t* rgba) {
  for (uint32_t i = 0; i < count; ++i) pal_lab_to_srgb8(lab4 + 4 * i, rgba + 4 * i);
}

// Synthetic generator of SURVEY.md section 8(d) (shared definition with the CUDA library and numpy).
static inline uint32_t h32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
ORC_API void orc_synth(uint8_t* rgba, uint64_t first_pixel, uint64_t n, uint32_t frame, uint32_t seed,
                       uint32_t blobs) {
  const uint32_t fkey = h32(seed + 0x9e3779b9u * frame);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)n; ++q) {
    uint64_t i = first_pixel + (uint64_t)q;
    uint32_t lo = (uint32_t)i, hi = (uint32_t)(i >> 32);
    uint32_t folded = hi ? (lo ^ h32(hi)) : lo;
    uint32_t u = h32(folded ^ fkey);
    uint8_t* o = rgba + 4 * q;
    if (blobs == 0) {
      o[0] = (uint8_t)(u & 255);
      o[1] = (uint8_t)((u >> 8) & 255);
      o[2] = (uint8_t)((u >> 16) & 255);
    } else {
      uint32_t g = h32(u + 1u) % blobs;
      uint32_t centre = h32(0xC0FFEEu + g + seed * blobs);
      for (uint32_t c = 0; c < 3; ++c) {
        uint32_t wv = h32(u + 0x1000u * (c + 1));
        int noise = (int)(wv & 15) + (int)((wv >> 4) & 15) + (int)((wv >> 8) & 15) + (int)((wv >> 12) & 15) - 30;
        int v = (int)((centre >> (8 * c)) & 255) + noise;
        o[c] = (uint8_t)std::min(255, std::max(0, v));
      }
    }
    o[3] = 255;
  }
}
