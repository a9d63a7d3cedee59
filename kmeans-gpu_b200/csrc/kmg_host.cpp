// kmg_host.cpp — host-side colour helpers of the boundary.
//
// The reference performs these conversions on the CPU as well, through the third-party `palette`
// crate 0.7.3 (Cargo.lock:846; call sites core/src/structures.rs:534,538,603,606 and
// core/src/lib.rs:277-282).  The crate's source is not part of the reference tree; the formulas are
// restated from its published algorithm (f32, D65 white 0.95047/1.0/1.08883, epsilon (6/29)^3,
// kappa 841/108, delta 4/29, Lindbloom sRGB matrices) and pinned by the reference's golden images
// (tests/test_oracle_golden.py).  Compile with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/kmeans_gpu.h"

namespace {

inline float pow_f32(float x, float y) { return (float)std::pow((double)x, (double)y); }

inline uint8_t unorm8(float v) {
  if (!(v > 0.0f)) return 0;
  if (v > 1.0f) v = 1.0f;
  return (uint8_t)std::nearbyintf(v * 255.0f);
}

void srgb8_to_lab(const uint8_t* px, float* out) {
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
  out[3] = 1.0f;
}

void lab_to_srgb8(const float* lab, uint8_t* out) {
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

}  // namespace

// CentroidsBuffer::fixed_centroids (core/src/structures.rs:523-553)
extern "C" void kmg_fixed_centroids(const uint8_t* colors, uint32_t count, int color_space, float* out) {
  for (uint32_t i = 0; i < count; ++i) {
    if (color_space == KMG_LAB) {
      srgb8_to_lab(colors + 4 * i, out + 4 * i);
    } else {
      out[4 * i + 0] = (float)colors[4 * i + 0] / 255.0f;
      out[4 * i + 1] = (float)colors[4 * i + 1] / 255.0f;
      out[4 * i + 2] = (float)colors[4 * i + 2] / 255.0f;
      out[4 * i + 3] = 1.0f;
    }
  }
}

// CentroidsBuffer::pull_values (core/src/structures.rs:600-617)
extern "C" void kmg_centroids_to_rgba8(const float* cent, uint32_t count, int color_space, uint8_t* out) {
  for (uint32_t i = 0; i < count; ++i) {
    if (color_space == KMG_LAB) {
      lab_to_srgb8(cent + 4 * i, out + 4 * i);
    } else {
      out[4 * i + 0] = unorm8(cent[4 * i + 0]);
      out[4 * i + 1] = unorm8(cent[4 * i + 1]);
      out[4 * i + 2] = unorm8(cent[4 * i + 2]);
      out[4 * i + 3] = 255;  // Srgba::new(.., 1.0)
    }
  }
}

// kmeans_palette sort (core/src/lib.rs:276-284): by Lab L of the 8-bit colour.  The reference uses
// sort_unstable_by; ties are broken here by keeping cluster order.
extern "C" void kmg_sort_palette_by_lightness(uint8_t* colors, uint32_t count) {
  std::vector<float> key(count);
  for (uint32_t i = 0; i < count; ++i) {
    float lab[4];
    srgb8_to_lab(colors + 4 * i, lab);
    key[i] = lab[0];
  }
  std::vector<uint32_t> order(count);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  std::vector<uint8_t> tmp(colors, colors + (size_t)count * 4);
  for (uint32_t i = 0; i < count; ++i) std::memcpy(colors + 4 * i, tmp.data() + 4 * order[i], 4);
}
