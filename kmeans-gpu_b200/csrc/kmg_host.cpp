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
#include <set>
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

// ------------------------------------------------------------------------------------------------
// Octree quantiser — `Algorithm::Octree` of the reference.  It is CPU code there as well
// (core/src/octree.rs, driven by operations::extract_palette_octree, core/src/operations.rs:90-97,
// on the <= 128 px shrink that octree_palette makes, core/src/lib.rs:288-331); a Rust caller keeps
// octree.rs as it is, this entry point gives non-Rust callers the same palette.
//
// Restated behaviour: an 8-level tree keyed by the r/g/b bits from the most significant one down
// (octree.rs:12-26); a node created while walking level L carries `level = L`, so leaves have
// level 7 (octree.rs:45-57); every pixel is added to its leaf.  reduce(): all nodes that hold
// pixels, ordered by (child_count, pixel_count >> level, node id) — octree.rs:246-266 — are merged
// smallest first into their parents until at most `color_count` remain (octree.rs:66-110); a popped
// node without a parent (the root) is dropped.  The palette is the integer mean of every remaining
// node, sorted as (r,g,b,a) tuples and de-duplicated (octree.rs:104-109,128-135).
namespace {

struct OctNode {  // node ids fit 32 bits: at most 8 new nodes per pixel, < 2^28 pixels accepted
  uint32_t level = 0;
  uint32_t color_index = 0;
  int32_t parent = -1;
  int32_t children[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  uint32_t child_count = 0;
  uint64_t count = 0, r = 0, g = 0, b = 0;
};

struct OctLess {  // Node::partial_cmp, octree.rs:246-266 (a strict total order: ids are distinct)
  const std::vector<OctNode>* nodes;
  bool operator()(size_t a, size_t b) const {
    if (a == b) return false;
    const OctNode& x = (*nodes)[a];
    const OctNode& y = (*nodes)[b];
    if (x.child_count != y.child_count) return x.child_count < y.child_count;
    const uint64_t xc = x.count >> x.level, yc = y.count >> y.level;
    if (xc != yc) return xc < yc;
    return a < b;
  }
};

}  // namespace

extern "C" int kmg_octree_palette(const uint8_t* rgba, uint64_t n_pixels, uint32_t color_count, uint8_t* colors_out,
                                  uint32_t* count_out) {
  if (!count_out || (n_pixels && !rgba) || (color_count && !colors_out)) return KMG_ERR_BAD_ARG;
  *count_out = 0;
  if (color_count == 0) return KMG_OK;  // octree.rs:67-69
  if (n_pixels >= (1ull << 28)) return KMG_ERR_BAD_ARG;  // the reference hands it at most 128 x 128 pixels
  std::vector<OctNode> nodes(1);
  nodes.reserve((size_t)std::min<uint64_t>(n_pixels * 8 + 1, 1u << 20));
  for (uint64_t p = 0; p < n_pixels; ++p) {
    const uint8_t* c = rgba + 4 * p;
    size_t at = 0;
    for (uint32_t level = 0; level < 8; ++level) {
      const uint8_t mask = (uint8_t)(0x80u >> level);
      const uint32_t ci = ((c[0] & mask) ? 4u : 0u) | ((c[1] & mask) ? 2u : 0u) | ((c[2] & mask) ? 1u : 0u);
      if (nodes[at].children[ci] < 0) {
        OctNode child;
        child.level = level;
        child.color_index = ci;
        child.parent = (int32_t)at;
        nodes[at].children[ci] = (int32_t)nodes.size();
        nodes[at].child_count += 1;
        nodes.push_back(child);
      }
      at = (size_t)nodes[at].children[ci];
    }
    nodes[at].r += c[0];
    nodes[at].g += c[1];
    nodes[at].b += c[2];
    nodes[at].count += 1;
  }
  // The reference keeps a deque sorted in descending order, pops its back (the smallest node) and
  // finds / re-inserts parents by binary search — O(n) per step.  Node keys only change while the
  // node is outside the sequence, so an ordered set with the same strict total order is equivalent
  // and O(log n) per step.
  OctLess less{&nodes};
  std::set<size_t, OctLess> leaves(less);
  for (size_t i = 0; i < nodes.size(); ++i)
    if (nodes[i].count > 0) leaves.insert(i);
  while (leaves.size() > color_count) {
    const size_t id = *leaves.begin();
    leaves.erase(leaves.begin());
    OctNode& node = nodes[id];
    if (node.parent < 0) continue;
    const size_t pid = (size_t)node.parent;
    leaves.erase(pid);  // present only if the parent already holds pixels (keyed by its current state)
    OctNode& par = nodes[pid];
    par.r += node.r;
    par.g += node.g;
    par.b += node.b;
    par.count += node.count;
    par.child_count -= 1;
    par.children[node.color_index] = -1;
    node.parent = -1;
    leaves.insert(pid);
  }
  std::vector<uint32_t> pal;
  for (size_t id : leaves) {
    const OctNode& n = nodes[id];
    const uint32_t r = (uint32_t)(n.r / n.count) & 255u, g = (uint32_t)(n.g / n.count) & 255u,
                   b = (uint32_t)(n.b / n.count) & 255u;
    pal.push_back((r << 24) | (g << 16) | (b << 8) | 255u);  // big-endian tuple order for the sort
  }
  std::sort(pal.begin(), pal.end());
  pal.erase(std::unique(pal.begin(), pal.end()), pal.end());
  for (size_t i = 0; i < pal.size(); ++i) {
    colors_out[4 * i + 0] = (uint8_t)(pal[i] >> 24);
    colors_out[4 * i + 1] = (uint8_t)(pal[i] >> 16);
    colors_out[4 * i + 2] = (uint8_t)(pal[i] >> 8);
    colors_out[4 * i + 3] = 255;
  }
  *count_out = (uint32_t)pal.size();
  return KMG_OK;
}
