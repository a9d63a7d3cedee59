// kmg_math.cuh — device arithmetic for the image hot path.
//
// Two families:
//   ex::   "exact" restatement of the reference shaders in IEEE binary32, every operation rounded
//          separately (intrinsics __f*_rn are never contracted into FMAs), sqrt/div correctly
//          rounded, pow_f32(x,y) := (float)pow((double)x,(double)y).  Bit-identical to
//          oracle/oracle.cpp.  Used where values are stored or compared exactly (work plane,
//          farthest-point distances, centroid finalisation, palette reversion, near-tie re-checks).
//   fast:: algebraically reduced CIE94 score and approximate Lab, only ever used together with a
//          certificate (score gap > error bound); when the certificate fails the caller re-evaluates
//          the candidates with ex:: so the result is the reference's, bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kmg {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

namespace ex {

__device__ __forceinline__ float pow_f32(float x, float y) { return (float)pow((double)x, (double)y); }

// core/shaders/functions/delta_e.wgsl:1-22.  `one` supplies SC/SH (asymmetric).
// c1 = sqrt(one.a^2 + one.b^2), c2 likewise for `second` (both evaluated exactly by the caller).
__device__ __forceinline__ float chroma(float a, float b) { return fsqrt(fadd(fmul(a, a), fmul(b, b))); }

__device__ __forceinline__ float cie94_c(float l1, float a1, float b1, float c1, float l2, float a2, float b2,
                                         float c2) {
  float dL = fsub(l1, l2);
  float da = fsub(a1, a2);
  float db = fsub(b1, b2);
  float dC = fsub(c1, c2);
  float dH = fsqrt(fmaxf(fsub(fadd(fmul(da, da), fmul(db, db)), fmul(dC, dC)), 0.0f));
  float SC = fadd(1.0f, fmul(0.045f, c1));
  float SH = fadd(1.0f, fmul(0.015f, c1));
  float tL = dL;  // dL / 1.0 (kL * SL = 1, delta_e.wgsl:17): x / 1 == x for every x, no division needed
  float tC = fdiv(dC, SC);
  float tH = fdiv(dH, SH);
  return fsqrt(fadd(fadd(fmul(tL, tL), fmul(tC, tC)), fmul(tH, tH)));
}
__device__ __forceinline__ float cie94(float l1, float a1, float b1, float l2, float a2, float b2) {
  return cie94_c(l1, a1, b1, chroma(a1, b1), l2, a2, b2, chroma(a2, b2));
}

// core/shaders/converters/rgb_to_lab.wgsl:16-33 for one 8-bit channel, including the x100.
__device__ __forceinline__ float srgb_decode100(uint32_t v) {
  float c = fdiv((float)v, 255.0f);
  float r = (c > 0.04045f) ? pow_f32(fdiv(fadd(c, 0.055f), 1.055f), 2.4f) : fdiv(c, 12.92f);
  return fmul(r, 100.0f);
}
// pow_f32(t, 1.0f / 3.0f) for 2^-10 < t < 2^4 without the ~200 FP64 instructions of pow() and
// without any f32<->f64 conversion (those run on the quarter-rate XU pipe): one Newton step of
// the cube root from a MUFU seed y0, carried as an unevaluated sum y0 + lo of two floats.
//   y0^3 - t  is formed exactly with FMA error-free products  (y0^2 = p + e1, p*y0 = q + e2,
//             q - t exact by Sterbenz), its f32 rounding is ~6e-14 t
//   lo        = -(y0^3 - t) * r + y0 * c,  r ~ 1/(3 y0^2),  c = (e - 1/3) ln t for the f32 exponent
//             e = 0.3333333432674408;  y0 + lo = t^e (1 + ~4e-13)
//   f         = RN(y0 + lo) — ONE rounding of the exact sum of two floats; rem = lo - (f - y0) is
//             its exact residual (Fast2Sum)
// Ziv's test: when y0 + lo lies within 2e-12 f of an f32 rounding boundary (|rem| close to half an
// ulp), or f is a power of two, the full pow() decides — so the value returned is always the one
// pow_f32 returns (checked over all 2^24 colours by test_convert_lab_all_16m_colours).
__device__ __noinline__ float pow_third_slow(float t) { return pow_f32(t, 1.0f / 3.0f); }
__device__ __forceinline__ float pow_third(float t) {
  float lg, y0, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__fmul_rn(lg, 0.33333334f)));
  const float p = __fmul_rn(y0, y0);
  const float e1 = __fmaf_rn(y0, y0, -p);
  const float q = __fmul_rn(p, y0);
  const float e2 = __fmaf_rn(p, y0, -q);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(3.0f, p)));
  const float res = __fadd_rn(__fsub_rn(q, t), __fmaf_rn(e1, y0, e2));
  const float c = __fmul_rn(6.8858033e-9f, lg);  // (e - 1/3) * ln 2 * log2 t
  const float lo = __fmaf_rn(y0, c, -__fmul_rn(res, r));
  const float f = __fadd_rn(y0, lo);
  const float rem = __fsub_rn(lo, __fsub_rn(f, y0));
  const float h = __int_as_float((__float_as_int(f) & 0x7f800000) - (24 << 23));  // half an ulp of f
  const float gap = fabsf(__fsub_rn(fabsf(rem), h));
  if (gap < __fmul_rn(2.0e-12f, f) || (__float_as_int(f) & 0x007fffff) == 0) return pow_third_slow(t);
  return f;
}

// x / c for one of the three white-point constants, correctly rounded like __fdiv_rn but in three
// instructions instead of ~10 plus a slow-path call (Markstein): q = RN(x * rc) with rc = RN(1 / c),
// r = x - q * c exactly (one FMA), q' = RN(q + r * rc).  rc is the correctly rounded reciprocal for all
// three constants (checked against exact rationals); the quotient is the IEEE one for every X, Y, Z that
// an sRGB8 colour can produce — test_convert_lab_all_16m_colours compares all 2^24 colours with the
// oracle, which divides.
__device__ __forceinline__ float div_white(float x, float c, float rc) {
  const float q = __fmul_rn(x, rc);
  const float r = __fmaf_rn(-q, c, x);
  return __fmaf_rn(r, rc, q);
}
// core/shaders/converters/rgb_to_lab.wgsl:39-64
__device__ __forceinline__ float lab_f(float t) {
  if (t > 0.008856f) return pow_third(t);
  return fadd(fmul(7.787f, t), 16.0f / 116.0f);
}
// r,g,b already decoded and scaled by 100 (table lookup).
__device__ __forceinline__ float4 lin100_to_lab(float r, float g, float b) {
  float X = fadd(fadd(fmul(0.4124564f, r), fmul(0.3575761f, g)), fmul(0.1804375f, b));
  float Y = fadd(fadd(fmul(0.2126729f, r), fmul(0.7151522f, g)), fmul(0.0721750f, b));
  float Z = fadd(fadd(fmul(0.0193339f, r), fmul(0.1191920f, g)), fmul(0.9503041f, b));
  float x = lab_f(div_white(X, 95.0489f, 0x1.58bfb6p-7f));
  float y = lab_f(div_white(Y, 100.0f, 0x1.47ae14p-7f));
  float z = lab_f(div_white(Z, 108.8840f, 0x1.2cf1b2p-7f));
  float4 o;
  o.x = fsub(fmul(116.0f, y), 16.0f);
  o.y = fmul(500.0f, fsub(x, y));
  o.z = fmul(200.0f, fsub(y, z));
  o.w = chroma(o.y, o.z);
  return o;
}
// core/shaders/converters/rgb8u_to_rgb32f.wgsl:16-17 (ColorSpace::Rgb): unorm8 load.
__device__ __forceinline__ float4 rgb8_to_rgbf(uint32_t px) {
  float4 o;
  o.x = fdiv((float)(px & 255u), 255.0f);
  o.y = fdiv((float)((px >> 8) & 255u), 255.0f);
  o.z = fdiv((float)((px >> 16) & 255u), 255.0f);
  o.w = chroma(o.y, o.z);
  return o;
}

// rgba8unorm store: clamp, scale, round to nearest even.
__device__ __forceinline__ uint32_t unorm8(float v) {
  if (!(v > 0.0f)) return 0u;
  if (v > 1.0f) v = 1.0f;
  return (uint32_t)__float2int_rn(fmul(v, 255.0f));
}
// pow_f32(t, 3.0f): the double product t*t*t is within 2.3e-16 of t^3; Ziv's test as in pow_third.
__device__ __noinline__ float pow_cube_slow(float t) { return pow_f32(t, 3.0f); }
__device__ __forceinline__ float pow_cube(float t) {
  const double td = (double)t;
  const double yd = td * td * td;
  const float f = __double2float_rn(yd);
  if (!(fabsf(f) > 1.0e-30f) || !(fabsf(f) < 1.0e30f)) return pow_cube_slow(t);  // zero, tiny, huge, NaN
  const double h = (double)__int_as_float((__float_as_int(f) & 0x7f800000) - (24 << 23));
  const double gap = fabs(fabs(yd - (double)f) - h);
  if (gap < 1.0e-14 * fabs(yd) || (__float_as_int(f) & 0x007fffff) == 0) return pow_cube_slow(t);
  return f;
}
// core/shaders/converters/lab_to_rgb.wgsl:40-66
__device__ __forceinline__ float lab_finv(float t) {
  float t3 = pow_cube(t);
  if (t3 > 0.008856f) return t3;
  return fdiv(fsub(t, 16.0f / 116.0f), 7.787f);
}
// core/shaders/converters/lab_to_rgb.wgsl:11-38
__device__ __forceinline__ float srgb_encode(float c) {
  if (c > 0.0031308f) return fsub(fmul(1.055f, pow_f32(c, 1.0f / 2.4f)), 0.055f);
  return fmul(12.92f, c);
}
// Lab -> linear sRGB (lab_to_rgb.wgsl:40-82 up to the transfer function).
__device__ __forceinline__ float3 lab_to_linear_rgb(float L, float A, float B) {
  float y = fdiv(fadd(L, 16.0f), 116.0f);
  float x = fadd(fdiv(A, 500.0f), y);
  float z = fsub(y, fdiv(B, 200.0f));
  x = fmul(lab_finv(x), 95.0489f);
  y = fmul(lab_finv(y), 100.0f);
  z = fmul(lab_finv(z), 108.8840f);
  x = fdiv(x, 100.0f);
  y = fdiv(y, 100.0f);
  z = fdiv(z, 100.0f);
  float3 o;
  o.x = fadd(fadd(fmul(3.2404542f, x), fmul(-1.5371385f, y)), fmul(-0.4985314f, z));
  o.y = fadd(fadd(fmul(-0.9692660f, x), fmul(1.8760108f, y)), fmul(0.0415560f, z));
  o.z = fadd(fadd(fmul(0.0556434f, x), fmul(-0.2040259f, y)), fmul(1.0572252f, z));
  return o;
}
__device__ __forceinline__ uint32_t lab_to_rgba8(float L, float A, float B) {
  const float3 c = lab_to_linear_rgb(L, A, B);
  return unorm8(srgb_encode(c.x)) | (unorm8(srgb_encode(c.y)) << 8) | (unorm8(srgb_encode(c.z)) << 16) | 0xFF000000u;
}
// core/shaders/converters/rgb32f_to_rgb8u.wgsl:16-17 (alpha component w).
__device__ __forceinline__ uint32_t rgbf_to_rgba8(float r, float g, float b, float w) {
  return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(w) << 24);
}

// Fixed-point unit of the centroid sums: rint(v * 2^15), ties to even.  For |v| < 128 — every Lab
// or RGB component of an sRGB8 colour — the IEEE sum v + 384.0f lies in [256, 512), whose ulp is
// 2^-15: the one rounding of the addition is the rounding wanted, and the integer sits in the low
// mantissa bits.  One FADD and one integer subtraction, no conversion instruction (F2I runs on the
// quarter-rate XU pipe); kernels that accumulate many pixels add the raw bits and take
// count * FIXED_MAGIC_BITS out once (kmg_lloyd_ring.cuh).
constexpr float FIXED_MAGIC = 384.0f;
constexpr unsigned int FIXED_MAGIC_BITS = 0x43C00000u;
constexpr double FIXED_UNIT = 1.0 / 32768.0;
__device__ __forceinline__ int to_fixed(float v) {
  return __float_as_int(__fadd_rn(v, FIXED_MAGIC)) - (int)FIXED_MAGIC_BITS;
}

}  // namespace ex

namespace fast {

// MUFU.RCP, 1 ulp, one instruction (the IEEE __frcp_rn costs ~12 with a slow-path call).
__device__ __forceinline__ float rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Blackwell packed-f32x2 arithmetic (SASS FFMA2 / FADD2): one issue slot, two lanes of work.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Per-pixel coefficients of the reduced CIE94 score (see DESIGN.md "assignment").  With
//   da^2 + db^2 - dC^2 = 2*(C1*C2 - a*ac - b*bc)
// half the squared distance is
//   d^2(p,c)/2 = [L^2/2 + C1^2/(2 SC^2)] + Lc^2/2 + L*(-Lc) + p1*(C2^2/2) + p2*C2 + p3*(-ac) + p4*(-bc)
// with p1 = 1/SC^2, p2 = C1*(1/SH^2 - 1/SC^2), p3 = a/SH^2, p4 = b/SH^2.  The bracket does not
// depend on c; the table record of a centroid is {Lc^2/2, -Lc, C2^2/2, C2, -ac, -bc}, so L itself
// is the first coefficient and no sign or factor 2 is ever applied per pixel (10 operations).
struct PixCoef {
  float p0, p1, p2, p3, p4;
  float hs;  // 1 / SH^2
};
__device__ __forceinline__ PixCoef pix_coef(float L, float a, float b, float C1) {
  float SC = fmaf(0.045f, C1, 1.0f);
  float SH = fmaf(0.015f, C1, 1.0f);
  float rSC = rcp(SC);
  float rSH = rcp(SH);
  PixCoef p;
  p.p1 = rSC * rSC;
  float hs = rSH * rSH;
  p.hs = hs;
  p.p0 = L;
  p.p2 = C1 * (hs - p.p1);
  p.p3 = hs * a;
  p.p4 = hs * b;
  return p;
}
// Absolute error bound of a (half-distance) score difference: rounding of the 5-term FMA chain,
// of the coefficients, and the reference's own f32 rounding, generous by > 4x:
//   eps = 2^-19 * [ (|L| + Lmax)^2 + 2 * (C1 + Cmax)^2 ]   (the scale is folded into the two terms).
__device__ __forceinline__ float score_eps(float L, float C1, float lmax, float cmax) {
  constexpr float KU = 0.00138106793f;  // 2^-9.5
  constexpr float KV = 0.001953125f;    // 2^-9
  float u = fmaf(fabsf(L), KU, lmax * KU);
  float v = fmaf(C1, KV, cmax * KV);
  return fmaf(u, u, v * v);
}

// Approximate f(t) of xyz_to_lab: relative error <= ~2^-21.
__device__ __forceinline__ float lab_f(float t) {
  float lg = lg2(t);
  float y0 = ex2(lg * 0.33333334f);
  // one Newton step for the cube root, then the (1/3)_f32 vs 1/3 exponent correction
  float r = rcp(y0 * y0);
  float y1 = fmaf(y0, 0.6666667f, 0.33333334f * t * r);
  y1 = y1 * fmaf(6.8862e-9f, lg, 1.0f);  // t^(0.3333333432674408 - 1/3) = 2^(9.934e-9 * log2 t)
  float lin = fmaf(7.787f, t, 16.0f / 116.0f);
  return t > 0.008856f ? y1 : lin;
}
// r,g,b decoded and scaled by 100.  Absolute error of the result <= LAB_ERR per component.
__device__ __forceinline__ float3 lin100_to_lab(float r, float g, float b) {
  float X = fmaf(0.1804375f, b, fmaf(0.3575761f, g, 0.4124564f * r));
  float Y = fmaf(0.0721750f, b, fmaf(0.7151522f, g, 0.2126729f * r));
  float Z = fmaf(0.9503041f, b, fmaf(0.1191920f, g, 0.0193339f * r));
  float x = lab_f(X * (1.0f / 95.0489f));
  float y = lab_f(Y * (1.0f / 100.0f));
  float z = lab_f(Z * (1.0f / 108.8840f));
  float3 o;
  o.x = fmaf(116.0f, y, -16.0f);
  o.y = 500.0f * (x - y);
  o.z = 200.0f * (y - z);
  return o;
}
// Bound on |approximate pixel - exact pixel| (Euclidean norm) in the remap kernels:
//   9.92e-5  largest |fast Lab - exact Lab| over ALL 2^24 sRGB colours (exhaustive on B200, the MUFU
//            results being deterministic; tests/test_gpu_parity.py::test_fast_lab_error_bound)
// + 1.4e-5   the dither offset is added to the approximate value and to the exact one: two
//            roundings of at most half an ulp(128) = 2^-18 per component
// + 2.4e-5   the chroma handed to the score is sqrt.approx (2^-23 relative) of a rounded a^2 + b^2,
//            C <= 134
// = 1.37e-4, rounded up.
constexpr float LAB_ERR = 1.5e-4f;
// RGB colour space: v * (1/255) against the exact v / 255 (one ulp of 1.0 per component), plus the
// same two terms for values <= 1.
constexpr float RGB_ERR = 6.0e-7f;

}  // namespace fast
}  // namespace kmg
