// kmg_kernels.cuh — hand-written sm_100a kernels of the image hot path.
//
// Layout in HBM (all linear, no textures, 64-bit indexing):
//   image     : RGBA8, 4 B/px, row-major                         (reference: Rgba8Unorm texture)
//   work plane: float4 {c0,c1,c2,chroma} 16 B/px                 (reference: Rgba32Float texture; the
//               4th float, 1.0 in the reference and never read, carries sqrt(c1^2+c2^2) here)
//   dmin plane: float 4 B/px, running min distance of the init   (reference: R32Float distance map)
//   job state : one small blob per k-means problem, see JobState
//
// Kernels are HBM- or FP32-issue-bound streaming kernels (no tensor-core contraction exists on
// this path): persistent grids sized in multiples of the SM count, 128-bit coalesced loads,
// centroid tables staged in shared memory, integer (order-independent) accumulators.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kmg_math.cuh"

namespace kmg {

// One centroid as the kernels see it.  q0 = {Lc^2, Lc, C2^2, C2}, q1 = {ac, bc, 0, 0}.
// Duplicates of a lower-index centroid and padding entries carry q0.x = MASKED so they can
// never win (the reference's strict '<' scan keeps the lowest index on exact ties anyway).
struct __align__(16) CentRec {
  float4 q0;
  float4 q1;
};
constexpr float MASKED = 1.0e30f;
constexpr int MAX_K = 4096;  // table = 128 KiB of shared memory at most

struct JobState {
  unsigned int ticket;   // blocks finished in the current pass
  unsigned int conv;     // convergence[k] of the last pass (choose_centroid.wgsl:196-201)
  unsigned int passes;   // update passes done
  unsigned int done;     // stop rule fired (core/src/modules.rs:802-836)
  unsigned int k;
  unsigned int max_iter;
  unsigned int check_every;
  float conv_threshold;
  float lmax, cmax;       // max |Lc|, max C2 over live centroids (for the score error bound)
  float dither_threshold; // mix_colors.wgsl:53-68
  unsigned int pad0;
  unsigned long long slow_pixels;
  unsigned long long pad1;
};

struct JobPtrs {
  JobState* st;
  float4* cent;                // k
  CentRec* tab;                // k padded to a multiple of 32 with MASKED entries
  long long* acc;              // ACC_COPIES x k x 4  (sum0,sum1,sum2,count), fixed-point 2^-16
  long long* last;             // k x 4 — the reduced sums of the last finalised pass
  unsigned long long* keys;    // k  — arg-max keys of the init rounds
  uint32_t* pal;               // k  — centroids reverted to RGBA8
};
constexpr int ACC_COPIES = 8;

__host__ __device__ inline unsigned int pad32(unsigned int k) { return (k + 31u) & ~31u; }

// ------------------------------------------------------------------------------------------------
// Small utilities

__device__ __forceinline__ float4 ldg_stream(const float4* p) { return __ldcs(p); }

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// sRGB decode table (x100) — ex::srgb_decode100 for the 256 possible inputs.
__global__ void k_build_srgb_table(float* table) {
  unsigned int i = threadIdx.x;
  if (i < 256) table[i] = ex::srgb_decode100(i);
}

// ------------------------------------------------------------------------------------------------
// Table construction from centroids (one block).  Also fills the dither threshold and the RGBA8
// palette.  Called as a device function by the last block of a Lloyd pass and by k_prepare.
template <int THREADS>
__device__ void build_table(const JobPtrs& J, unsigned int k, int color_space, bool want_palette) {
  __shared__ float s_red[2][THREADS / 32];
  const unsigned int tid = threadIdx.x;
  const unsigned int kp = pad32(k);
  float lmax = 0.0f, cmax = 0.0f;
  for (unsigned int c = tid; c < kp; c += THREADS) {
    CentRec r;
    if (c < k) {
      float4 v = J.cent[c];
      float c2 = ex::chroma(v.y, v.z);
      bool dup = false;
      for (unsigned int i = 0; i < c; ++i) {
        float4 u = J.cent[i];
        dup |= (u.x == v.x && u.y == v.y && u.z == v.z);
      }
      r.q0 = make_float4(dup ? MASKED : v.x * v.x, v.x, c2 * c2, c2);
      r.q1 = make_float4(v.y, v.z, 0.0f, 0.0f);
      lmax = fmaxf(lmax, fabsf(v.x));
      cmax = fmaxf(cmax, c2);
      if (want_palette)
        J.pal[c] = color_space == 0 ? ex::lab_to_rgba8(v.x, v.y, v.z) : ex::rgbf_to_rgba8(v.x, v.y, v.z, v.w);
    } else {
      r.q0 = make_float4(MASKED, 0.0f, 0.0f, 0.0f);
      r.q1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    J.tab[c] = r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = lmax;
    s_red[1][tid >> 5] = cmax;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < THREADS / 32; ++w) {
      lmax = fmaxf(lmax, s_red[0][w]);
      cmax = fmaxf(cmax, s_red[1][w]);
    }
    J.st->lmax = lmax;
    J.st->cmax = cmax;
    // mix_colors.wgsl:53-68 — greedy farthest pair, asymmetric distance with centroid i first.
    float thr = 0.0f;
    if (k > 1) {
      float4 a = J.cent[0], b = J.cent[1];
      float d_ab = ex::cie94(a.x, a.y, a.z, b.x, b.y, b.z);
      for (unsigned int i = 2; i < k; ++i) {
        float4 ci = J.cent[i];
        float da = ex::cie94(ci.x, ci.y, ci.z, a.x, a.y, a.z);
        float db = ex::cie94(ci.x, ci.y, ci.z, b.x, b.y, b.z);
        if (da > db && da > d_ab) {
          d_ab = da;
          b = ci;
        } else if (db > d_ab) {
          d_ab = db;
          a = ci;
        }
      }
      thr = fdiv(d_ab, fsqrt((float)k));
    }
    J.st->dither_threshold = thr;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_prepare(JobPtrs J, int color_space, int want_palette) {
  build_table<256>(J, J.st->k, color_space, want_palette != 0);
}

// ------------------------------------------------------------------------------------------------
// Certified nearest-centroid search over a shared-memory table, P pixels per thread.
//
// Fast pass: 5 FMA per (pixel, centroid) + min / second-min tracking.  A pixel is certified when
// second - best > eps (eps bounds every rounding difference between the fast score and the
// reference's f32 distance, plus extra_eps supplied by callers whose pixel is itself approximate).
// Otherwise every centroid whose fast score is within eps of the best is re-evaluated with the
// exact reference arithmetic, scanning in index order with strict '<' from 100000.0
// (find_centroid.wgsl:29-41), which is exactly what the reference does over those candidates.
template <int P>
struct Pix {
  float L[P], a[P], b[P], C[P];
};

template <int P, int UNROLL>
__device__ __forceinline__ void argmin_fast(const CentRec* __restrict__ tab, unsigned int kp, const Pix<P>& px,
                                            float (&m1)[P], float (&m2)[P], unsigned int (&idx)[P]) {
  fast::PixCoef pc[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    pc[i] = fast::pix_coef(px.L[i], px.a[i], px.b[i], px.C[i]);
    m1[i] = 3.0e38f;
    m2[i] = 3.0e38f;
    idx[i] = 0;
  }
#pragma unroll UNROLL
  for (unsigned int j = 0; j < kp; ++j) {
    const float4 q0 = tab[j].q0;
    const float2 q1 = *reinterpret_cast<const float2*>(&tab[j].q1);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float s = fast::score(pc[i], q0, q1);
      m2[i] = fminf(m2[i], fmaxf(s, m1[i]));
      bool lt = s < m1[i];
      m1[i] = lt ? s : m1[i];
      idx[i] = lt ? j : idx[i];
    }
  }
}

// Small tables (KT <= 16): all KT scores of a pixel stay in registers.  Min + index are tracked with
// FSETP/FSEL/SEL; the certificate "no other score within eps of the best" is evaluated on the FMA
// pipe instead of the (half-rate, otherwise saturated) ALU pipe:
//   S = sum_j sat((s_j - m1) / eps)   is  >= KT - 1  iff every other score is >= eps away.
template <int P, int KT>
__device__ __forceinline__ void argmin_saved(const CentRec* __restrict__ tab, const Pix<P>& px, float lmax,
                                             float cmax, float (&m1)[P], float (&eps)[P], unsigned int (&idx)[P],
                                             bool (&certified)[P]) {
  fast::PixCoef pc[P];
  float s[KT][P];
#pragma unroll
  for (int i = 0; i < P; ++i) pc[i] = fast::pix_coef(px.L[i], px.a[i], px.b[i], px.C[i]);
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float4 q0 = tab[j].q0;
    const float2 q1 = *reinterpret_cast<const float2*>(&tab[j].q1);
#pragma unroll
    for (int i = 0; i < P; ++i) s[j][i] = fast::score(pc[i], q0, q1);
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    float m = s[0][i];
    unsigned int ix = 0;
#pragma unroll
    for (int j = 1; j < KT; ++j) {
      bool lt = s[j][i] < m;
      m = lt ? s[j][i] : m;
      ix = lt ? (unsigned int)j : ix;
    }
    m1[i] = m;
    idx[i] = ix;
    eps[i] = fast::score_eps(px.L[i], px.C[i], lmax, cmax);
    const float g = fast::rcp(eps[i]);
    const float base = -m * g;
    float S = 0.0f;
#pragma unroll
    for (int j = 0; j < KT; ++j) S += __saturatef(fmaf(s[j][i], g, base));
    certified[i] = S > (float)(KT - 1) - 1.0e-3f;
  }
}

// Packed variant of argmin_saved for P even: pixels are processed two at a time with FFMA2/FADD2,
// the table is read as duplicated pairs {q,q} (tab2: KT x 6 float2), the minimum is a 3-input-min
// tournament and the winning index is recovered on the FMA pipe from the certificate terms:
// with t_j = sat((s_j - m1)/eps) in {~0 (winner), 1 (everyone else)} for a certified pixel,
//   R_m = sum_{j>=m} t_j (suffix sums),  S = R_0,  sum_j j*t_j = sum_{m>=1} R_m,
// so idx = KT(KT-1)/2 - sum_{m>=1} R_m.  Uncertified pixels take the exact path, which recomputes
// the index itself.
template <int P, int KT>
__device__ __forceinline__ void argmin_saved_x2(const float2* __restrict__ tab2, const Pix<P>& px, float lmax,
                                                float cmax, float (&m1)[P], float (&eps)[P],
                                                unsigned int (&idx)[P], bool (&certified)[P]) {
  static_assert(P % 2 == 0, "pairs of pixels");
  static_assert((KT & (KT - 1)) == 0, "table length must be a power of two");
  constexpr int H = P / 2;
  fast::f32x2 pp[H][5];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    fast::PixCoef c0 = fast::pix_coef(px.L[2 * h], px.a[2 * h], px.b[2 * h], px.C[2 * h]);
    fast::PixCoef c1 = fast::pix_coef(px.L[2 * h + 1], px.a[2 * h + 1], px.b[2 * h + 1], px.C[2 * h + 1]);
    pp[h][0] = fast::pack2(c0.p0, c1.p0);
    pp[h][1] = fast::pack2(c0.p1, c1.p1);
    pp[h][2] = fast::pack2(c0.p2, c1.p2);
    pp[h][3] = fast::pack2(c0.p3, c1.p3);
    pp[h][4] = fast::pack2(c0.p4, c1.p4);
  }
  fast::f32x2 s2[KT][H];
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float4 qa = *reinterpret_cast<const float4*>(tab2 + j * 6);      // {Lc^2,Lc^2, Lc,Lc}
    const float4 qb = *reinterpret_cast<const float4*>(tab2 + j * 6 + 2);  // {C2^2,C2^2, C2,C2}
    const float4 qc = *reinterpret_cast<const float4*>(tab2 + j * 6 + 4);  // {ac,ac, bc,bc}
    const fast::f32x2 q0x = fast::pack2(qa.x, qa.y), q0y = fast::pack2(qa.z, qa.w);
    const fast::f32x2 q0z = fast::pack2(qb.x, qb.y), q0w = fast::pack2(qb.z, qb.w);
    const fast::f32x2 q1x = fast::pack2(qc.x, qc.y), q1y = fast::pack2(qc.z, qc.w);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      fast::f32x2 s = fast::fma2(pp[h][0], q0y, q0x);
      s = fast::fma2(pp[h][1], q0z, s);
      s = fast::fma2(pp[h][2], q0w, s);
      s = fast::fma2(pp[h][3], q1x, s);
      s = fast::fma2(pp[h][4], q1y, s);
      s2[j][h] = s;
    }
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float sa[KT], sb[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) fast::unpack2(s2[j][h], sa[j], sb[j]);
    float ma = sa[0], mb = sb[0];
#pragma unroll
    for (int j = 1; j + 1 < KT; j += 2) {
      ma = fast::min3(ma, sa[j], sa[j + 1]);
      mb = fast::min3(mb, sb[j], sb[j + 1]);
    }
    if ((KT & 1) == 0) {
      ma = fminf(ma, sa[KT - 1]);
      mb = fminf(mb, sb[KT - 1]);
    }
    const float ea = fast::score_eps(px.L[2 * h], px.C[2 * h], lmax, cmax);
    const float eb = fast::score_eps(px.L[2 * h + 1], px.C[2 * h + 1], lmax, cmax);
    const float ga = fast::rcp(ea), gb = fast::rcp(eb);
    const float ba = -ma * ga, bb = -mb * gb;
    // t_j pairs, then a pairwise tree: at every level the sum of the odd-position entries is the
    // count of indices with that bit set, so sum_j j*t_j = sum_b 2^b * B_b with depth log2(KT).
    fast::f32x2 v[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j)
      v[j] = fast::pack2(__saturatef(fmaf(sa[j], ga, ba)), __saturatef(fmaf(sb[j], gb, bb)));
    fast::f32x2 I = fast::pack2(0.0f, 0.0f);
    float wgt = 1.0f;
#pragma unroll
    for (int n = KT; n > 1; n >>= 1) {
      fast::f32x2 odd = v[1];
#pragma unroll
      for (int m = 1; m < n / 2; ++m) odd = fast::add2(odd, v[2 * m + 1]);
      I = fast::fma2(odd, fast::pack2(wgt, wgt), I);
      wgt *= 2.0f;
#pragma unroll
      for (int m = 0; m < n / 2; ++m) v[m] = fast::add2(v[2 * m], v[2 * m + 1]);
    }
    const fast::f32x2 R = v[0];
    float Sa, Sb, Ia, Ib;
    fast::unpack2(R, Sa, Sb);
    fast::unpack2(I, Ia, Ib);
    constexpr float TRI = (float)(KT * (KT - 1) / 2);
    m1[2 * h] = ma;
    m1[2 * h + 1] = mb;
    eps[2 * h] = ea;
    eps[2 * h + 1] = eb;
    certified[2 * h] = Sa > (float)(KT - 1) - 1.0e-3f;
    certified[2 * h + 1] = Sb > (float)(KT - 1) - 1.0e-3f;
    idx[2 * h] = (unsigned int)__float2int_rn(TRI - Ia) & (unsigned int)(KT - 1);
    idx[2 * h + 1] = (unsigned int)__float2int_rn(TRI - Ib) & (unsigned int)(KT - 1);
  }
}

// Exact re-evaluation for one pixel (exact components + exact chroma).
__device__ __noinline__ unsigned int argmin_exact(const CentRec* __restrict__ tab, unsigned int k, float L, float a,
                                                  float b, float C, float bound) {
  fast::PixCoef pc = fast::pix_coef(L, a, b, C);
  float best = 100000.0f;
  unsigned int found = 0;
  for (unsigned int j = 0; j < k; ++j) {
    const float4 q0 = tab[j].q0;
    const float4 q1 = tab[j].q1;
    float s = fast::score(pc, q0, make_float2(q1.x, q1.y));
    if (s <= bound) {
      float d = ex::cie94_c(L, a, b, C, q0.y, q1.x, q1.y, q0.w);
      if (d < best) {
        best = d;
        found = j;
      }
    }
  }
  return found;
}

// ------------------------------------------------------------------------------------------------
// K1/K3: RGBA8 -> work plane, exact.  4 pixels (one 128-bit load) per thread per step.
__global__ void __launch_bounds__(256) k_convert(const uint32_t* __restrict__ rgba, unsigned long long n,
                                                 int color_space, const float* __restrict__ lut_g,
                                                 float4* __restrict__ work) {
  __shared__ float lut[256];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    uint32_t v = __ldcs(rgba + p);
    float4 o;
    if (color_space == 0)
      o = ex::lin100_to_lab(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u]);
    else
      o = ex::rgb8_to_rgbf(v);
    work[p] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// K15: bilinear shrink, exact restatement (see oracle resize_image).
__global__ void __launch_bounds__(256) k_resize(const uint32_t* __restrict__ src, unsigned int sw, unsigned int sh,
                                                uint32_t* __restrict__ dst, unsigned int dw, unsigned int dh) {
  const unsigned long long n = (unsigned long long)dw * dh;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    unsigned int gx = (unsigned int)(p % dw), gy = (unsigned int)(p / dw);
    float py = fsub(fmul(fdiv((float)gy, (float)dh), (float)sh), 0.5f);
    float px = fsub(fmul(fdiv((float)gx, (float)dw), (float)sw), 0.5f);
    float fy0 = floorf(py), fx0 = floorf(px);
    float fy = fsub(py, fy0), fx = fsub(px, fx0);
    long long y0 = (long long)fy0, x0 = (long long)fx0;
    long long y1 = y0 + 1, x1 = x0 + 1;
    y0 = min(max(y0, 0ll), (long long)sh - 1);
    y1 = min(max(y1, 0ll), (long long)sh - 1);
    x0 = min(max(x0, 0ll), (long long)sw - 1);
    x1 = min(max(x1, 0ll), (long long)sw - 1);
    uint32_t p00 = __ldg(src + (size_t)y0 * sw + x0), p10 = __ldg(src + (size_t)y0 * sw + x1);
    uint32_t p01 = __ldg(src + (size_t)y1 * sw + x0), p11 = __ldg(src + (size_t)y1 * sw + x1);
    float wx0 = fsub(1.0f, fx), wy0 = fsub(1.0f, fy);
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float c00 = fdiv((float)((p00 >> (8 * c)) & 255u), 255.0f);
      float c10 = fdiv((float)((p10 >> (8 * c)) & 255u), 255.0f);
      float c01 = fdiv((float)((p01 >> (8 * c)) & 255u), 255.0f);
      float c11 = fdiv((float)((p11 >> (8 * c)) & 255u), 255.0f);
      float top = fadd(fmul(c00, wx0), fmul(c10, fx));
      float bot = fadd(fmul(c01, wx0), fmul(c11, fx));
      o |= ex::unorm8(fadd(fmul(top, wy0), fmul(bot, fy))) << (8 * c);
    }
    dst[p] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Farthest-point initialisation (K8-K11).
//
// Round j (1 <= j < k) folds centroid j-1 into the running min-distance plane and finds the
// arg-max for centroid j in the same pass: 16 B + 4 B read, 4 B written per pixel.  The arg-max
// is an order-independent 64-bit atomicMax of  (distance bits << 32) | (pixel ^ 15):
// among equal maxima the highest 16-pixel chunk wins and, inside it, the lowest pixel — the tie
// rule of selectCandidate (plus_plus_init.wgsl:62-68,:92,:102,:136,:142).  A zero maximum selects
// pixel 0 (every thread starts from Candidate(0, 0.0)).
__device__ __forceinline__ unsigned long long key_to_pixel(unsigned long long key) {
  return (key >> 32) == 0ull ? 0ull : ((key & 0xffffffffull) ^ 15ull);
}

__global__ void k_init_seed(JobPtrs J, const float4* __restrict__ work, unsigned long long seed_local,
                            int seed_is_local) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (seed_is_local) {
      float4 v = work[seed_local];
      J.cent[0] = make_float4(v.x, v.y, v.z, 1.0f);
    }
    for (unsigned int i = 0; i < J.st->k; ++i) J.keys[i] = 0ull;
  }
}

// pixel_offset: global index of this shard's first pixel (0 on a single GPU).
template <bool FIRST>
__global__ void __launch_bounds__(256) k_init_round(JobPtrs J, const float4* __restrict__ work,
                                                    float* __restrict__ dmin, unsigned long long n,
                                                    unsigned long long pixel_offset, unsigned int j) {
  __shared__ unsigned long long s_key[8];
  // centroid j-1 was resolved into J.cent[j-1] by k_init_pick (or k_init_seed for j == 1).
  const float4 c = J.cent[j - 1];
  const float cc = ex::chroma(c.y, c.z);
  unsigned long long best = 0ull;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    float4 v = ldg_stream(work + p);
    float d = ex::cie94_c(v.x, v.y, v.z, v.w, c.x, c.y, c.z, cc);
    float dm = FIRST ? fminf(1000000.0f, d) : fminf(dmin[p], d);  // kmeans++_calc_diff.wgsl:27-31
    dmin[p] = dm;
    unsigned long long key =
        ((unsigned long long)__float_as_uint(dm) << 32) | (((pixel_offset + p) & 0xffffffffull) ^ 15ull);
    best = key > best ? key : best;
  }
  best = warp_max_u64(best);
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = s_key[w] > best ? s_key[w] : best;
    atomicMax(J.keys + j, best);
  }
}

// Resolve the winner of round j into centroid j (plus_plus_init.wgsl:170-179).
__global__ void k_init_pick(JobPtrs J, const float4* __restrict__ work, unsigned long long n,
                            unsigned long long pixel_offset, unsigned int j) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long p = key_to_pixel(J.keys[j]);
    if (p >= pixel_offset && p - pixel_offset < n) {
      float4 v = work[p - pixel_offset];
      J.cent[j] = make_float4(v.x, v.y, v.z, 1.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lloyd pass = assignment (K5) + centroid update (K6/K7 for every cluster) in ONE sweep over the
// cached work plane: 16 B/px read, nothing written but k x 4 integer sums.
//
// Sums are exact integers, rint(v * 2^16) accumulated in int32 thread-private shared-memory slots
// (ACC_PRIVATE: conflict-free 128-bit read-modify-write, flushed before they can overflow) or sent
// straight to L2 with 64-bit reductions (ACC_GLOBAL, large k).  Integer addition commutes, so the
// result is independent of block scheduling, grid size and of how many GPUs share the image.
// The last block to finish turns the sums into the new centroids, convergence flags and the next
// table, so a pass is exactly one launch and needs no host round trip.

template <int THREADS>
__device__ void finalize_pass(const JobPtrs& J, int color_space, bool distributed_partial) {
  JobState* st = J.st;
  const unsigned int k = st->k;
  const unsigned int tid = threadIdx.x;
  __shared__ unsigned int s_conv;
  if (tid == 0) s_conv = 0;
  __syncthreads();
  unsigned int conv = 0;
  for (unsigned int c = tid; c < k; c += THREADS) {
    long long s[4] = {0, 0, 0, 0};
    for (int copy = 0; copy < ACC_COPIES; ++copy) {
      long long* a = J.acc + ((size_t)copy * k + c) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        s[q] += __ldcg(a + q);
        if (!distributed_partial) a[q] = 0;
      }
    }
    if (distributed_partial) {
      // multi-GPU: leave the folded partial in copy 0 for the all-reduce; finalised later
      long long* a0 = J.acc + (size_t)c * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) a0[q] = s[q];
      for (int copy = 1; copy < ACC_COPIES; ++copy) {
        long long* a = J.acc + ((size_t)copy * k + c) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = 0;
      }
      continue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) J.last[(size_t)c * 4 + q] = s[q];
    if (s[3] > 0) {  // choose_centroid.wgsl:185-194
      const double cnt = (double)s[3];
      float4 prev = J.cent[c];
      float4 nc;
      nc.x = (float)(((double)s[0] / cnt) * (1.0 / 65536.0));
      nc.y = (float)(((double)s[1] / cnt) * (1.0 / 65536.0));
      nc.z = (float)(((double)s[2] / cnt) * (1.0 / 65536.0));
      nc.w = 1.0f;
      J.cent[c] = nc;
      conv += (ex::cie94(nc.x, nc.y, nc.z, prev.x, prev.y, prev.z) < st->conv_threshold) ? 1u : 0u;
    }
  }
  if (distributed_partial) {
    __syncthreads();
    if (tid == 0) st->ticket = 0;
    return;
  }
  if (conv) atomicAdd(&s_conv, conv);
  __threadfence();
  __syncthreads();
  build_table<THREADS>(J, k, color_space, false);
  if (tid == 0) {
    const unsigned int it = st->passes;  // 0-based iteration index of this pass
    st->conv = s_conv;
    st->passes = it + 1;
    // core/src/modules.rs:802,827 — tested only when it > 0 && it % 8 == 0; also the hard cap.
    bool check = it > 0 && st->check_every != 0 && (it % st->check_every) == 0;
    if ((check && s_conv >= k) || it + 1 >= st->max_iter) st->done = 1;
    st->ticket = 0;
  }
}

// Finalise after an external all-reduce of acc copy 0 (multi-GPU).
__global__ void __launch_bounds__(256) k_finalize(JobPtrs J, int color_space) {
  if (J.st->done) return;
  finalize_pass<256>(J, color_space, false);
}

template <int THREADS, int P, bool CHECK>
__device__ __forceinline__ void lloyd_load(const float4* __restrict__ work, unsigned long long base,
                                           unsigned long long n, float4 (&v)[P]) {
#pragma unroll
  for (int i = 0; i < P; ++i) {
    unsigned long long p = base + (unsigned long long)i * THREADS;
    v[i] = (!CHECK || p < n) ? ldg_stream(work + p) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int KT, int THREADS, int P, bool SAVED, bool CHECK>
__device__ __forceinline__ void lloyd_tile(const CentRec* __restrict__ s_tab, const float2* __restrict__ s_tab2,
                                           int4* __restrict__ s_acc,
                                           const float4 (&v)[P], unsigned long long base, unsigned long long n,
                                           unsigned int k, float lmax, float cmax, unsigned int tid,
                                           unsigned int& slow) {
  Pix<P> px;
  bool valid[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    valid[i] = CHECK ? (base + (unsigned long long)i * THREADS) < n : true;
    px.L[i] = v[i].x;
    px.a[i] = v[i].y;
    px.b[i] = v[i].z;
    px.C[i] = v[i].w;
  }
  float m1[P], eps[P];
  unsigned int idx[P];
  bool certified[P];
  if (SAVED) {
    argmin_saved_x2<P, KT>(s_tab2, px, lmax, cmax, m1, eps, idx, certified);
  } else {
    float m2[P];
    argmin_fast<P, KT>(s_tab, KT, px, m1, m2, idx);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      eps[i] = fast::score_eps(px.L[i], px.C[i], lmax, cmax);
      certified[i] = m2[i] - m1[i] > eps[i];
    }
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (!certified[i] && valid[i]) {
      idx[i] = argmin_exact(s_tab, k, px.L[i], px.a[i], px.b[i], px.C[i], m1[i] + eps[i]);
      ++slow;
    }
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (valid[i]) {
      int4* slot = s_acc + idx[i] * THREADS + tid;
      int4 a = *slot;
      a.x += ex::to_fixed(px.L[i]);
      a.y += ex::to_fixed(px.a[i]);
      a.z += ex::to_fixed(px.b[i]);
      a.w += 1;
      *slot = a;
    }
  }
}

template <int KT, int THREADS, int P, bool SAVED, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_lloyd_private(JobPtrs J, const float4* __restrict__ work,
                                                                 unsigned long long n, int color_space,
                                                                 int distributed_partial) {
  // KT: compile-time table length (k padded with MASKED entries), fully unrolled.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4* s_acc = reinterpret_cast<int4*>(smem_raw);  // [KT][THREADS]
  __shared__ CentRec s_tab[KT];
  __shared__ __align__(16) float2 s_tab2[KT * 6];  // the same table as duplicated pairs for FFMA2
  __shared__ bool s_last;
  JobState* st = J.st;
  if (st->done) return;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  for (unsigned int c = tid; c < KT; c += THREADS) {
    CentRec r = J.tab[c];
    s_tab[c] = r;
    s_tab2[c * 6 + 0] = make_float2(r.q0.x, r.q0.x);
    s_tab2[c * 6 + 1] = make_float2(r.q0.y, r.q0.y);
    s_tab2[c * 6 + 2] = make_float2(r.q0.z, r.q0.z);
    s_tab2[c * 6 + 3] = make_float2(r.q0.w, r.q0.w);
    s_tab2[c * 6 + 4] = make_float2(r.q1.x, r.q1.x);
    s_tab2[c * 6 + 5] = make_float2(r.q1.y, r.q1.y);
  }
#pragma unroll
  for (int c = 0; c < KT; ++c) s_acc[c * THREADS + tid] = make_int4(0, 0, 0, 0);
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();

  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long full_tiles = n / TILE;
  unsigned int since_flush = 0;
  unsigned int slow = 0;

  auto flush = [&]() {
    // Every warp folds the private slots of its own 32 threads for all clusters (no block sync
    // needed: a warp only reads what it wrote) and sends 4 reductions per cluster to L2.
    const unsigned int lane = tid & 31;
    long long* dst = J.acc + (size_t)(blockIdx.x % ACC_COPIES) * k * 4;
    for (unsigned int c = 0; c < k; ++c) {
      int4 v = s_acc[c * THREADS + tid];
      s_acc[c * THREADS + tid] = make_int4(0, 0, 0, 0);
      long long s0 = warp_sum_i64(v.x), s1 = warp_sum_i64(v.y), s2 = warp_sum_i64(v.z), s3 = warp_sum_i64(v.w);
      if (lane == 0 && s3 != 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(dst + c * 4 + 0), (unsigned long long)s0);
        atomicAdd(reinterpret_cast<unsigned long long*>(dst + c * 4 + 1), (unsigned long long)s1);
        atomicAdd(reinterpret_cast<unsigned long long*>(dst + c * 4 + 2), (unsigned long long)s2);
        atomicAdd(reinterpret_cast<unsigned long long*>(dst + c * 4 + 3), (unsigned long long)s3);
      }
    }
    since_flush = 0;
  };

  // full tiles: no bounds checks in the hot loop; the next tile is in flight (registers) while the
  // current one is processed, so HBM latency is hidden even at 2 blocks per SM.
  {
    float4 cur[P], nxt[P];
    unsigned long long tile = blockIdx.x;
    if (tile < full_tiles) lloyd_load<THREADS, P, false>(work, tile * TILE + tid, n, cur);
    for (; tile < full_tiles; tile += gridDim.x) {
      const unsigned long long next = tile + gridDim.x;
      if (next < full_tiles) lloyd_load<THREADS, P, false>(work, next * TILE + tid, n, nxt);
      lloyd_tile<KT, THREADS, P, SAVED, false>(s_tab, s_tab2, s_acc, cur, tile * TILE + tid, n, k, lmax, cmax, tid, slow);
      since_flush += P;
      // |v| < 2^7 colour units -> |fixed| < 2^23; 240 pixels stay below 2^31.
      if (since_flush + P > 240) flush();
#pragma unroll
      for (int i = 0; i < P; ++i) cur[i] = nxt[i];
    }
  }
  // ragged tail (< TILE pixels), taken by the block whose turn it would be
  if (full_tiles * TILE < n && blockIdx.x == (unsigned int)(full_tiles % gridDim.x)) {
    if (since_flush + P > 240) flush();
    float4 tail[P];
    lloyd_load<THREADS, P, true>(work, full_tiles * TILE + tid, n, tail);
    lloyd_tile<KT, THREADS, P, SAVED, true>(s_tab, s_tab2, s_acc, tail, full_tiles * TILE + tid, n, k, lmax, cmax, tid, slow);
  }
  flush();
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);

  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    finalize_pass<THREADS>(J, color_space, distributed_partial != 0);
  }
}

// Large k: table in dynamic shared memory, runtime loop, sums reduced straight into L2.
template <int THREADS, int P>
__global__ void __launch_bounds__(THREADS) k_lloyd_global(JobPtrs J, const float4* __restrict__ work,
                                                          unsigned long long n, int color_space,
                                                          int distributed_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  __shared__ bool s_last;
  JobState* st = J.st;
  if (st->done) return;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  const unsigned int kp = pad32(k);
  for (unsigned int c = tid; c < kp; c += THREADS) s_tab[c] = J.tab[c];
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();

  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long tiles = (n + TILE - 1) / TILE;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(J.acc + (size_t)(blockIdx.x % ACC_COPIES) * k * 4);
  unsigned int slow = 0;
  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned long long base = tile * TILE + tid;
    Pix<P> px;
    bool valid[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      unsigned long long p = base + (unsigned long long)i * THREADS;
      valid[i] = p < n;
      float4 v = valid[i] ? ldg_stream(work + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      px.L[i] = v.x;
      px.a[i] = v.y;
      px.b[i] = v.z;
      px.C[i] = v.w;
    }
    float m1[P], m2[P];
    unsigned int idx[P];
    argmin_fast<P, 4>(s_tab, kp, px, m1, m2, idx);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float eps = fast::score_eps(px.L[i], px.C[i], lmax, cmax);
      if (m2[i] - m1[i] <= eps && valid[i]) {
        idx[i] = argmin_exact(s_tab, k, px.L[i], px.a[i], px.b[i], px.C[i], m1[i] + eps);
        ++slow;
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (valid[i]) {
        unsigned long long* a = acc + (size_t)idx[i] * 4;
        atomicAdd(a + 0, (unsigned long long)(long long)ex::to_fixed(px.L[i]));
        atomicAdd(a + 1, (unsigned long long)(long long)ex::to_fixed(px.a[i]));
        atomicAdd(a + 2, (unsigned long long)(long long)ex::to_fixed(px.b[i]));
        atomicAdd(a + 3, 1ull);
      }
    }
  }
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);

  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    finalize_pass<THREADS>(J, color_space, distributed_partial != 0);
  }
}

// ------------------------------------------------------------------------------------------------
// K5 alone: labels for a work plane (stage-level parity test hook, also used by meld-free tools).
template <int THREADS, int P>
__global__ void __launch_bounds__(THREADS) k_assign(JobPtrs J, const float4* __restrict__ work,
                                                    unsigned long long n, uint32_t* __restrict__ labels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  JobState* st = J.st;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  const unsigned int kp = pad32(k);
  for (unsigned int c = tid; c < kp; c += THREADS) s_tab[c] = J.tab[c];
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();
  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long tiles = (n + TILE - 1) / TILE;
  unsigned int slow = 0;
  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned long long base = tile * TILE + tid;
    Pix<P> px;
    bool valid[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      unsigned long long p = base + (unsigned long long)i * THREADS;
      valid[i] = p < n;
      float4 v = valid[i] ? ldg_stream(work + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      px.L[i] = v.x;
      px.a[i] = v.y;
      px.b[i] = v.z;
      px.C[i] = v.w;
    }
    float m1[P], m2[P];
    unsigned int idx[P];
    argmin_fast<P, 4>(s_tab, kp, px, m1, m2, idx);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float eps = fast::score_eps(px.L[i], px.C[i], lmax, cmax);
      if (m2[i] - m1[i] <= eps && valid[i]) {
        idx[i] = argmin_exact(s_tab, k, px.L[i], px.a[i], px.b[i], px.C[i], m1[i] + eps);
        ++slow;
      }
      if (valid[i]) labels[base + (unsigned long long)i * THREADS] = idx[i];
    }
  }
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);
}

// ------------------------------------------------------------------------------------------------
// Fused remap (K1 + K5 + K12 + K2, or K1 + K13 + K2): RGBA8 in, RGBA8 out, 8 B/px of HBM traffic
// instead of the reference's 72-80 B/px over 3-4 dispatches.  Lab is computed approximately in
// registers; only pixels whose certificate fails recompute it exactly (FP64 pow) and re-evaluate
// the near-tied candidates with reference arithmetic.  The output colour of cluster c is the
// pre-reverted palette entry pal[c] (swap.wgsl:22-24 + lab_to_rgb.wgsl of a constant).
// mix_colors.wgsl:14-17
__device__ __constant__ float c_bayer[16] = {0.f, 8.f, 2.f, 10.f, 12.f, 4.f, 14.f, 6.f,
                                             3.f, 11.f, 1.f, 9.f, 15.f, 7.f, 13.f, 5.f};

template <int MODE, int KT>
__device__ __noinline__ unsigned int remap_exact(const CentRec* __restrict__ tab, unsigned int k, uint32_t v,
                                                 const float* __restrict__ lut, int color_space, float off,
                                                 float slack) {
  float4 e = color_space == 0 ? ex::lin100_to_lab(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u])
                              : ex::rgb8_to_rgbf(v);
  float L = e.x, a = e.y, b = e.z, C = e.w;
  if (MODE == 1) {  // mix_colors.wgsl:70-72
    L = fadd(L, off);
    a = fadd(a, off);
    b = fadd(b, off);
    C = ex::chroma(a, b);
  }
  // candidate bound around this pixel's own best fast score
  fast::PixCoef pc = fast::pix_coef(L, a, b, C);
  float m1 = 3.0e38f;
  for (unsigned int j = 0; j < k; ++j) {
    const float4 q0 = tab[j].q0;
    const float4 q1 = tab[j].q1;
    m1 = fminf(m1, fast::score(pc, q0, make_float2(q1.x, q1.y)));
  }
  return argmin_exact(tab, k, L, a, b, C, m1 + slack);
}

template <int MODE, int KT, int THREADS>
__global__ void __launch_bounds__(THREADS) k_remap(JobPtrs J, const uint32_t* __restrict__ rgba, unsigned int w,
                                                   unsigned long long n, int color_space,
                                                   const float* __restrict__ lut_g, uint32_t* __restrict__ out) {
  // KT > 0: compile-time table length; KT == 0: runtime length in dynamic shared memory.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  __shared__ float lut[256];
  JobState* st = J.st;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  const unsigned int kp = KT > 0 ? (unsigned int)KT : pad32(k);
  uint32_t* s_pal = reinterpret_cast<uint32_t*>(s_tab + kp);
  for (unsigned int c = tid; c < kp; c += THREADS) {
    s_tab[c] = J.tab[c];
    s_pal[c] = c < k ? J.pal[c] : 0u;
  }
  for (unsigned int c = tid; c < 256; c += THREADS) lut[c] = lut_g[c];
  const float lmax = st->lmax, cmax = st->cmax;
  const float thr = st->dither_threshold;
  __syncthreads();

  constexpr int P = 4;
  const unsigned long long groups = (n + P - 1) / P;
  const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
  unsigned int slow = 0;
  for (unsigned long long g = (unsigned long long)blockIdx.x * THREADS + tid; g < groups; g += stride) {
    const unsigned long long p0 = g * P;
    uint32_t v[P];
    const bool full = p0 + P <= n;
    if (full && (reinterpret_cast<uintptr_t>(rgba) & 15) == 0) {
      uint4 t = __ldcs(reinterpret_cast<const uint4*>(rgba) + g);
      v[0] = t.x;
      v[1] = t.y;
      v[2] = t.z;
      v[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = p0 + i < n ? rgba[p0 + i] : 0u;
    }
    if (MODE != 1 && k == 1) {
      // single colour: the scan trivially returns index 0
    }
    Pix<P> px;
    float off[P];
    unsigned int x = 0, y = 0;
    if (MODE == 1) {
      x = (unsigned int)(p0 % w);
      y = (unsigned int)(p0 / w);
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float L, a, b;
      if (color_space == 0) {
        float3 lab = fast::lin100_to_lab(lut[v[i] & 255u], lut[(v[i] >> 8) & 255u], lut[(v[i] >> 16) & 255u]);
        L = lab.x;
        a = lab.y;
        b = lab.z;
      } else {
        L = (float)(v[i] & 255u) * (1.0f / 255.0f);
        a = (float)((v[i] >> 8) & 255u) * (1.0f / 255.0f);
        b = (float)((v[i] >> 16) & 255u) * (1.0f / 255.0f);
      }
      off[i] = 0.0f;
      if (MODE == 1) {
        unsigned int xi = x + i, yi = y;
        if (xi >= w) {  // group straddles a row end (w % 4 != 0)
          unsigned long long p = p0 + i;
          xi = (unsigned int)(p % w);
          yi = (unsigned int)(p / w);
        }
        float iv = c_bayer[(xi & 3u) + ((yi & 3u) << 2)] * 0.0625f - 0.5f;  // mix_colors.wgsl:21-27,70
        off[i] = fmul(thr, iv);
        L += off[i];
        a += off[i];
        b += off[i];
      }
      px.L[i] = L;
      px.a[i] = a;
      px.b[i] = b;
      px.C[i] = sqrtf(fmaf(a, a, b * b));
    }
    float m1[P], m2[P];
    unsigned int idx[P];
    if (KT > 0)
      argmin_fast<P, (KT > 0 ? KT : 4)>(s_tab, kp, px, m1, m2, idx);
    else
      argmin_fast<P, 4>(s_tab, kp, px, m1, m2, idx);
    uint32_t o[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      // error bound: score rounding + sensitivity of the score gap to the approximate Lab:
      // |grad d^2| <= 2.5 * D_E, D_E <= SC * d, two candidates -> 5 * SC * d * LAB_ERR.
      float eps = fast::score_eps(px.L[i], px.C[i], lmax, cmax);
      float SC = fmaf(0.045f, px.C[i], 1.0f);
      float pconst = fmaf(px.L[i], px.L[i], (px.C[i] * px.C[i]) / (SC * SC));
      float d2 = fmaxf(m1[i] + pconst, 0.0f) + eps;
      float eps_conv = color_space == 0 ? 5.0f * fast::LAB_ERR * SC * sqrtf(d2) + 3.0f * fast::LAB_ERR * fast::LAB_ERR
                                        : 5.0f * 2.4e-7f * SC * sqrtf(d2);
      float slack = eps + eps_conv;
      if (m2[i] - m1[i] <= slack && p0 + i < n && k > 1) {
        idx[i] = remap_exact<MODE, KT>(s_tab, k, v[i], lut, color_space, off[i], 2.0f * slack);
        ++slow;
      }
      o[i] = s_pal[idx[i]];
    }
    if (full && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i)
        if (p0 + i < n) out[p0 + i] = o[i];
    }
  }
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);
}

// Meld (K14, mix_colors.wgsl:29-48,85-90,115-136): continuous output, evaluated exactly per pixel.
__global__ void __launch_bounds__(256) k_remap_meld(JobPtrs J, const uint32_t* __restrict__ rgba,
                                                    unsigned long long n, int color_space,
                                                    const float* __restrict__ lut_g, uint32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_cent = reinterpret_cast<float4*>(smem_raw);
  __shared__ float lut[256];
  const unsigned int k = J.st->k;
  for (unsigned int c = threadIdx.x; c < k; c += blockDim.x) s_cent[c] = J.cent[c];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    uint32_t v = __ldcs(rgba + p);
    float4 res;
    if (k == 1) {
      res = s_cent[0];
    } else {
      float4 e = color_space == 0 ? ex::lin100_to_lab(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u])
                                  : ex::rgb8_to_rgbf(v);
      float4 closest = make_float4(10000.f, 10000.f, 10000.f, 10000.f), second = closest;
      float dc = ex::cie94_c(e.x, e.y, e.z, e.w, closest.x, closest.y, closest.z, ex::chroma(closest.y, closest.z));
      float ds = dc;
      for (unsigned int i = 0; i < k; ++i) {
        float4 t = s_cent[i];
        float td = ex::cie94_c(e.x, e.y, e.z, e.w, t.x, t.y, t.z, ex::chroma(t.y, t.z));
        if (td < dc) {
          second = closest;
          ds = dc;
          closest = t;
          dc = td;
        } else if (td < ds) {
          second = t;
          ds = td;
        }
      }
      float factor = fdiv(ds, ex::cie94(closest.x, closest.y, closest.z, second.x, second.y, second.z));
      float g = fsub(1.0f, factor);
      res.x = fadd(fmul(factor, closest.x), fmul(g, second.x));
      res.y = fadd(fmul(factor, closest.y), fmul(g, second.y));
      res.z = fadd(fmul(factor, closest.z), fmul(g, second.z));
      res.w = fadd(fmul(factor, closest.w), fmul(g, second.w));
    }
    out[p] = color_space == 0 ? ex::lab_to_rgba8(res.x, res.y, res.z) : ex::rgbf_to_rgba8(res.x, res.y, res.z, res.w);
  }
}

// ------------------------------------------------------------------------------------------------
// Synthetic images, identical to oracle orc_synth.
__device__ __forceinline__ uint32_t h32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256) k_synth(uint32_t* __restrict__ rgba, unsigned long long first,
                                               unsigned long long n, uint32_t frame, uint32_t seed, uint32_t blobs) {
  const uint32_t fkey = h32(seed + 0x9e3779b9u * frame);
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    unsigned long long i = first + q;
    uint32_t lo = (uint32_t)i, hi = (uint32_t)(i >> 32);
    uint32_t folded = hi ? (lo ^ h32(hi)) : lo;
    uint32_t u = h32(folded ^ fkey);
    uint32_t o;
    if (blobs == 0) {
      o = (u & 0x00ffffffu) | 0xff000000u;
    } else {
      uint32_t g = h32(u + 1u) % blobs;
      uint32_t centre = h32(0xC0FFEEu + g + seed * blobs);
      o = 0xff000000u;
#pragma unroll
      for (uint32_t c = 0; c < 3; ++c) {
        uint32_t wv = h32(u + 0x1000u * (c + 1));
        int noise = (int)(wv & 15) + (int)((wv >> 4) & 15) + (int)((wv >> 8) & 15) + (int)((wv >> 12) & 15) - 30;
        int val = (int)((centre >> (8 * c)) & 255) + noise;
        val = min(255, max(0, val));
        o |= (uint32_t)val << (8 * c);
      }
    }
    rgba[q] = o;
  }
}

// Max |fast Lab - exact Lab| over all 2^24 colours (test hook for the LAB_ERR bound).
__global__ void __launch_bounds__(256) k_fast_lab_error(const float* __restrict__ lut_g, float* __restrict__ out_max) {
  __shared__ float lut[256];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  float worst = 0.0f;
  const unsigned int stride = gridDim.x * blockDim.x;
  for (unsigned int v = blockIdx.x * blockDim.x + threadIdx.x; v < (1u << 24); v += stride) {
    float r = lut[v & 255u], g = lut[(v >> 8) & 255u], b = lut[(v >> 16) & 255u];
    float4 e = ex::lin100_to_lab(r, g, b);
    float3 f = fast::lin100_to_lab(r, g, b);
    float dx = e.x - f.x, dy = e.y - f.y, dz = e.z - f.z;
    worst = fmaxf(worst, sqrtf(dx * dx + dy * dy + dz * dz));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(worst));
}

}  // namespace kmg
