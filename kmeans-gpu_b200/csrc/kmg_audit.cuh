// kmg_audit.cuh — device-side audit of the near-tie certificate (test hook, KMG entry point
// kmg_dev_audit).  The production kernels trust a *certified* label without looking at it again;
// a wrong error bound would fail silently as a wrong label.  The audit runs the very device
// functions the production kernels call (argmin_small, argmin_chunked, ring_pair_search,
// remap_fast_pixel) on every pixel, also scans all k centroids with the reference's own arithmetic
// and order (find_centroid.wgsl:29-41: strict '<' from (100000.0, index 0)), and counts
//   wrong        pixels whose certified label differs from the reference scan   (must be 0)
//   uncertified  pixels that production would hand to the exact path
#pragma once
#include "kmg_kernels.cuh"
#include "kmg_lloyd_ring.cuh"

namespace kmg {

// find_centroid.wgsl:29-41 over the raw centroids; every thread scans all k for its own pixel.
__device__ __forceinline__ unsigned int reference_scan(const float4* __restrict__ cent, unsigned int k, float L, float a,
                                                       float b, float C) {
  float best = 100000.0f;
  unsigned int idx = 0;
  for (unsigned int j = 0; j < k; ++j) {
    const float4 c = cent[j];
    const float d = ex::cie94_c(L, a, b, C, c.x, c.y, c.z, ex::chroma(c.y, c.z));
    if (d < best) {
      best = d;
      idx = j;
    }
  }
  return idx;
}

// SEARCH 0: argmin_small<2, 8>   1: argmin_small<2, 16>   2: argmin_chunked<2>   3: ring_pair_search<8>
// MODE   0: Lloyd / assign (exact work plane in, no conversion error)   1: remap replace   2: remap dither
//          (RGBA8 in, fast Lab + offset; the reference label is the scan of the exact pixel)
template <int SEARCH, int MODE>
__global__ void __launch_bounds__(256) k_audit(JobPtrs J, const float4* __restrict__ work, const uint32_t* __restrict__ rgba,
                                               unsigned int w, unsigned long long n, int color_space,
                                               const float* __restrict__ lut_g, unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  __shared__ float lut[256];
  const unsigned int tid = threadIdx.x;
  const unsigned int k = J.st->k;
  const unsigned int kp = (SEARCH == 0 || SEARCH == 3) ? 8u : (SEARCH == 1 ? 16u : pad32(k));
  tab_to_smem(s_tab, J.tab, kp, tid, 256);
  lut[tid] = lut_g[tid];
  const float lmax = J.st->lmax, cmax = J.st->cmax, thr = J.st->dither_threshold;
  __syncthreads();
  float tq[8][6];
  if (SEARCH == 3) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int q = 0; q < 6; ++q) tq[j][q] = rec_at(s_tab, j)->q[q];
      tq[j][1] = __int_as_float(__float_as_int(tq[j][1]) + (9 << 23));
      tq[j][3] = __int_as_float(__float_as_int(tq[j][3]) + (9 << 23));
    }
  }
  constexpr bool CONV = MODE != 0;
  const float conv_k = color_space == 0 ? fast::LAB_ERR : fast::RGB_ERR;
  unsigned long long wrong = 0, uncert = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * 256;
  // two pixels per thread and step: p and p + stride (the searches work on pixel pairs)
  for (unsigned long long p0 = (unsigned long long)blockIdx.x * 256 + tid; p0 < n; p0 += 2 * stride) {
    Pix<2> px;
    float4 ex_px[2];
    bool valid[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned long long p = p0 + (unsigned long long)i * stride;
      valid[i] = p < n;
      if (MODE == 0) {
        const float4 v = valid[i] ? work[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        px.L[i] = v.x;
        px.a[i] = v.y;
        px.b[i] = v.z;
        px.C[i] = v.w;
        ex_px[i] = v;
      } else {
        const uint32_t v = valid[i] ? rgba[p] : 0u;
        const unsigned int xi = (unsigned int)(p % w), yi = (unsigned int)(p / w);
        float off;
        remap_fast_pixel<(MODE == 2 ? 1 : 0)>(v, lut, color_space, thr, xi, yi, px.L[i], px.a[i], px.b[i], px.C[i], off);
        ex_px[i] = remap_exact_pixel<(MODE == 2 ? 1 : 0)>(v, lut, color_space, off);
      }
    }
    float eps[2];
    unsigned int idx[2];
    bool certified[2];
    if (SEARCH == 0) {
      argmin_small<2, 8, CONV>(s_tab, px, lmax, cmax, conv_k, eps, idx, certified);
    } else if (SEARCH == 1) {
      argmin_small<2, 16, CONV>(s_tab, px, lmax, cmax, conv_k, eps, idx, certified);
    } else if (SEARCH == 2) {
      argmin_chunked<2, CONV>(s_tab, kp, px, lmax, cmax, conv_k, eps, idx, certified);
    } else {
      using RC = RingCert<8, 1024>;
      unsigned int ua, ub;
      float fla, flb;
      ring_pair_search<8, 1024>(make_float4(px.L[0], px.a[0], px.b[0], px.C[0]), make_float4(px.L[1], px.a[1], px.b[1], px.C[1]),
                                tq, lmax * 0.00138106793f, cmax * 0.001953125f, ua, ub, fla, flb);
      certified[0] = (ua & RC::CERT_MASK) == RC::CERT_ONE;
      certified[1] = (ub & RC::CERT_MASK) == RC::CERT_ONE;
      idx[0] = (ua & RC::IDX_MASK) / 1024;
      idx[1] = (ub & RC::IDX_MASK) / 1024;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!valid[i]) continue;
      if (k == 1) continue;  // production never searches a one-colour palette
      if (!certified[i]) {
        ++uncert;
        continue;
      }
      const unsigned int want = reference_scan(J.cent, k, ex_px[i].x, ex_px[i].y, ex_px[i].z, ex_px[i].w);
      wrong += idx[i] != want ? 1ull : 0ull;
    }
  }
  wrong = (unsigned long long)warp_sum_i64((long long)wrong);
  uncert = (unsigned long long)warp_sum_i64((long long)uncert);
  if ((tid & 31) == 0) {
    if (wrong) atomicAdd(counters, wrong);
    if (uncert) atomicAdd(counters + 1, uncert);
  }
}

// Margin probe: how much of the error bound is ever used?  Every pixel evaluates the fast score of
// all k centroids (the production arithmetic: pix_coef + score1, on the exact plane or on the fast
// pixel of the remap kernels) and the reference scan.  Where the fast arg-min is NOT the reference
// label the production code relies on the certificate to notice: the gap between the fast scores of
// the two labels must be below eps.  counters[0] = such pixels, counters[1] = the largest gap / eps
// seen among them, in millionths (1e6 = the bound was exactly used up; every value below leaves room).
template <int MODE>
__global__ void __launch_bounds__(256) k_audit_margin(JobPtrs J, const float4* __restrict__ work,
                                                      const uint32_t* __restrict__ rgba, unsigned int w, unsigned long long n,
                                                      int color_space, const float* __restrict__ lut_g,
                                                      unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  __shared__ float lut[256];
  const unsigned int tid = threadIdx.x;
  const unsigned int k = J.st->k;
  const unsigned int kp = pad32(k);
  tab_to_smem(s_tab, J.tab, kp, tid, 256);
  lut[tid] = lut_g[tid];
  const float lmax = J.st->lmax, cmax = J.st->cmax, thr = J.st->dither_threshold;
  __syncthreads();
  constexpr bool CONV = MODE != 0;
  const float conv_k = color_space == 0 ? fast::LAB_ERR : fast::RGB_ERR;
  unsigned long long differ = 0, worst = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * 256;
  for (unsigned long long p = (unsigned long long)blockIdx.x * 256 + tid; p < n; p += stride) {
    float L, a, b, C;
    float4 e;
    if (MODE == 0) {
      e = work[p];
      L = e.x; a = e.y; b = e.z; C = e.w;
    } else {
      const uint32_t v = rgba[p];
      float off;
      remap_fast_pixel<(MODE == 2 ? 1 : 0)>(v, lut, color_space, thr, (unsigned int)(p % w), (unsigned int)(p / w), L, a, b, C, off);
      e = remap_exact_pixel<(MODE == 2 ? 1 : 0)>(v, lut, color_space, off);
    }
    const unsigned int want = reference_scan(J.cent, k, e.x, e.y, e.z, e.w);
    const fast::PixCoef pc = fast::pix_coef(L, a, b, C);
    float m = 3.0e38f, s_want = 0.0f;
    unsigned int im = 0;
    for (unsigned int j = 0; j < k; ++j) {
      const float sc = score1(pc, rec_at(s_tab, j)->q);
      if (sc < m) {
        m = sc;
        im = j;
      }
      if (j == want) s_want = sc;
    }
    if (im != want) {
      // masked duplicates never win the fast search; the reference label is then the lowest index of
      // the duplicate group and scores the same as its twin
      const float eps = total_eps<CONV>(fast::score_eps(L, C, lmax, cmax), m, conv_k, L, C, pc.p1, pc.hs, cmax);
      const float ratio = (s_want - m) / eps;
      if (s_want < 1.0e29f) {
        ++differ;
        const unsigned long long r = (unsigned long long)(fminf(fmaxf(ratio, 0.0f), 1.0e6f) * 1.0e6f);
        worst = r > worst ? r : worst;
      }
    }
  }
  differ = (unsigned long long)warp_sum_i64((long long)differ);
  worst = warp_max_u64(worst);
  if ((tid & 31) == 0) {
    if (differ) atomicAdd(counters, differ);
    if (worst) atomicMax(counters + 1, worst);
  }
}

}  // namespace kmg
