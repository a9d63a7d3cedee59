// kmg_lloyd_ring.cuh — the small-k Lloyd pass (k <= 8) as a warp-specialised TMA pipeline.
//
// Same work as k_lloyd<8,...> (kmg_kernels.cuh): assignment (find_centroid.wgsl:15-44) + per-cluster
// sums (choose_centroid.wgsl:73-178) in one sweep over the cached work plane, 16 B/px read.  What
// differs is how the pixels reach the registers and what the hot loop issues:
//
//  * a block is 8 consumer warps + 1 producer warp around a ring of D stages in shared memory; the
//    producer re-arms a stage with one bulk copy (cp.async.bulk -> SASS UBLKCP, the TMA engine) of
//    the block's next 256 x P pixels as soon as the eight consumers have released it, so D - 1 tiles
//    per block are in flight whatever the occupancy; no consumer issues a global load or an address
//    computation for one, no "next tile" register buffer exists and nothing is copied out of it;
//  * the consumer loop contains no call: a pixel pair with an uncertified pixel (1e-4 of the pixels)
//    leaves the loop, the exact path runs outside and re-enters.  With no call inside, the 48 table
//    values stay resident (40 in uniform registers, the 8 addends in vector registers) for the whole
//    sweep instead of being re-loaded from the constant bank every tile;
//  * the fixed-point unit of the sums is produced by one FADD (v + 384.0f puts rint(v * 2^15) into
//    the low mantissa bits) and the bias is taken out again when the slots are flushed;
//  * the certificate's weighted flag sum carries the cluster index already multiplied by the slot
//    stride, so one LOP3 turns it into the shared-memory address of the pixel's accumulator slot.
#pragma once
#include "kmg_kernels.cuh"

namespace kmg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

using ex::FIXED_MAGIC;
using ex::FIXED_MAGIC_BITS;

// Layout of the dynamic shared memory of k_lloyd_ring<KT, CWARPS, P, D>:
//   [pad to a multiple of SLOT_SPAN]  slots  int32 [4 components][KT clusters][256 consumer threads]
//   ring   float4 [D][CWARPS][P][32]          one stage = one contiguous run of the plane
//   bars   uint64 full[D], empty[D]
template <int KT, int CWARPS, int P, int D>
struct RingLayout {
  static constexpr unsigned int CTHREADS = CWARPS * 32;
  static constexpr unsigned int THREADS = CTHREADS + 32;      // + the producer warp
  static constexpr unsigned int SLOT_SPAN = KT * CTHREADS * 4;  // bytes of one component; a power of two
  static constexpr unsigned int WARP_BYTES = 32 * P * 16;
  static constexpr unsigned int STAGE_BYTES = CWARPS * WARP_BYTES;
  static constexpr unsigned int RING_BYTES = D * STAGE_BYTES;
  static constexpr unsigned int BAR_BYTES = 2 * D * 8;
  static constexpr unsigned int BYTES = SLOT_SPAN /* alignment slack */ + 4 * SLOT_SPAN + RING_BYTES + BAR_BYTES;
  static_assert((SLOT_SPAN & (SLOT_SPAN - 1)) == 0, "slot span must be a power of two");
  static_assert((D & (D - 1)) == 0, "ring depth must be a power of two");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// One pixel into (SIGN = +1) or out of (SIGN = -1) the private slots of a cluster: three biased
// fixed-point values and the count, four fire-and-forget shared-memory reductions.
// L_BIASED: the first value already is L + FIXED_MAGIC.
template <unsigned int SPAN, int SIGN = 1, bool L_BIASED = false>
__device__ __forceinline__ void slot_add(uint32_t addr, float L, float a, float b) {
  const unsigned int f0 = __float_as_uint(L_BIASED ? L : __fadd_rn(L, FIXED_MAGIC)), f1 = __float_as_uint(__fadd_rn(a, FIXED_MAGIC)),
                     f2 = __float_as_uint(__fadd_rn(b, FIXED_MAGIC));
  asm volatile(
      "red.shared.add.u32 [%0], %1;\n"
      "red.shared.add.u32 [%0+%5], %2;\n"
      "red.shared.add.u32 [%0+%6], %3;\n"
      "red.shared.add.u32 [%0+%7], %4;" ::"r"(addr),
      "r"(SIGN > 0 ? f0 : 0u - f0), "r"(SIGN > 0 ? f1 : 0u - f1), "r"(SIGN > 0 ? f2 : 0u - f2), "r"(SIGN > 0 ? 1u : 0xffffffffu),
      "n"(SPAN), "n"(2 * SPAN), "n"(3 * SPAN)
      : "memory");
}

// The certified search of one pixel pair against the resident table (tq[j][1] and tq[j][3] carry
// -Lc * 2^9 and C2 * 2^9).
// Returns, per pixel, the bits of 2^23 + V with V = sum_j (KT + j) * STRIDE * [s_j <= min + eps]:
// exactly one score within eps of the minimum  <=>  (V & CERT_MASK) == CERT_ONE, and then
// V & IDX_MASK = idx * STRIDE, the byte offset of cluster idx's slot.
template <int KT, unsigned int STRIDE>
struct RingCert {
  static constexpr unsigned int IDX_MASK = (KT - 1) * STRIDE;
  static constexpr unsigned int CERT_MASK = 0x7fffffu & ~(unsigned int)(KT * STRIDE - 1);
  static constexpr unsigned int CERT_ONE = KT * STRIDE;
};
template <int KT, unsigned int STRIDE>
__device__ __forceinline__ void ring_pair_search(const float4& va, const float4& vb, const float (&tq)[KT][6], float lmax_u,
                                                 float cmax_v, unsigned int& ua, unsigned int& ub, float& fixed_la,
                                                 float& fixed_lb) {
  constexpr float TOTAL = (float)(KT * KT + KT * (KT - 1) / 2);  // sum of all weights (KT + j)
  constexpr float WSCALE = (float)STRIDE;
  // The per-pixel coefficients of the reduced score (fast::pix_coef) for both pixels at once, in
  // packed arithmetic wherever the operands already sit in register pairs.  L and the pixel's
  // chroma travel as L * 2^-9 and C * 2^-9 against table entries scaled by 2^9 (all exact): the pairs
  // then are results of multiplications — a pair merely put together from two loads is cloned by
  // ptxas before almost every use (ten moves per pixel pair).
  fast::f32x2 pp[5];
  pp[0] = fast::pack2(va.x * 0.001953125f, vb.x * 0.001953125f);
  const fast::f32x2 c2 = fast::pack2(va.w * 0.001953125f, vb.w * 0.001953125f);
  const fast::f32x2 one2 = fast::pack2(1.0f, 1.0f);
  float sca, scb, sha, shb;
  fast::unpack2(fast::fma2(c2, fast::pack2(0.045f * 512.0f, 0.045f * 512.0f), one2), sca, scb);  // SC = 1 + 0.045 C
  fast::unpack2(fast::fma2(c2, fast::pack2(0.015f * 512.0f, 0.015f * 512.0f), one2), sha, shb);  // SH = 1 + 0.015 C
  const fast::f32x2 rsc = fast::pack2(fast::rcp(sca), fast::rcp(scb)), rsh = fast::pack2(fast::rcp(sha), fast::rcp(shb));
  pp[1] = fast::mul2(rsc, rsc);                          // 1 / SC^2
  const fast::f32x2 hs2 = fast::mul2(rsh, rsh);          // 1 / SH^2
  pp[2] = fast::mul2(c2, fast::sub2(hs2, pp[1]));        // C (1/SH^2 - 1/SC^2) * 2^-9
  float hsa, hsb;
  fast::unpack2(hs2, hsa, hsb);
  pp[3] = fast::pack2(hsa * va.y, hsb * vb.y);           // a / SH^2
  pp[4] = fast::pack2(hsa * va.z, hsb * vb.z);           // b / SH^2
  float sa[KT], sb[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) fast::unpack2(score2(pp, tq[j]), sa[j], sb[j]);
  float ma = fast::min3(sa[0], sa[1], sa[2]), mb = fast::min3(sb[0], sb[1], sb[2]);
  ma = fast::min3(ma, sa[3], sa[4]);
  mb = fast::min3(mb, sb[3], sb[4]);
  ma = fast::min3(ma, sa[5], sa[6]);
  mb = fast::min3(mb, sb[5], sb[6]);
  ma = fminf(ma, sa[7]);
  mb = fminf(mb, sb[7]);
  // threshold = minimum + fast::score_eps, the addition folded into the last FMA, both pixels at once:
  // u = L * 2^-9.5 + lmax * 2^-9.5 from the pair that already holds L * 2^-9 (L >= 0 in both colour
  // spaces; a rounding-sized negative L would shrink the bound by parts in 10^9), v = C * 2^-9 + cmax * 2^-9
  const fast::f32x2 u2 = fast::fma2(pp[0], fast::pack2(0.70710678f, 0.70710678f), fast::pack2(lmax_u, lmax_u));
  const fast::f32x2 v2 = fast::add2(c2, fast::pack2(cmax_v, cmax_v));
  float ta, tb;
  fast::unpack2(fast::fma2(u2, u2, fast::fma2(v2, v2, fast::pack2(ma, mb))), ta, tb);
  // V = sum_j (KT + j) * stride * [s_j <= t]: exactly one score within eps of the minimum  <=>
  // V == (KT + idx) * stride, and then V & IDX_MASK is the byte offset of cluster idx's slot
  fast::f32x2 acc0 = fast::pack2(8388608.0f + TOTAL * WSCALE, 8388608.0f + TOTAL * WSCALE);
  fast::f32x2 acc1 = fast::pack2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float fa = sa[j] > ta ? 1.0f : 0.0f;
    const float fb = sb[j] > tb ? 1.0f : 0.0f;
    const float w = -(float)(KT + j) * WSCALE;
    if (j & 1)
      acc1 = fast::fma2(fast::pack2(fa, fb), fast::pack2(w, w), acc1);
    else
      acc0 = fast::fma2(fast::pack2(fa, fb), fast::pack2(w, w), acc0);
  }
  float Va, Vb;
  fast::unpack2(fast::add2(acc0, acc1), Va, Vb);
  ua = __float_as_uint(Va);
  ub = __float_as_uint(Vb);
  // rint(L * 2^15) of both pixels from the pair L * 2^-9: (L * 2^-9) * 512 + 384 = L + 384, one rounding
  fast::unpack2(fast::fma2(pp[0], fast::pack2(512.0f, 512.0f), fast::pack2(FIXED_MAGIC, FIXED_MAGIC)), fixed_la, fixed_lb);
}

// The exact path for one pixel of the stage that left the hot loop (cold; every lane of the warp
// calls it).  The hot loop has already added every pixel to the slot its flag sum pointed at (`hit`,
// the byte offset of a cluster): an uncertified pixel is taken out of that slot again and put into
// the cluster the exact search finds.
template <unsigned int SPAN>
__device__ __noinline__ void ring_slow_pixel(const CentRec* __restrict__ g_tab, unsigned int k, float4 v, bool nd,
                                             uint32_t hit, float lmax, float cmax, uint32_t slot_tid_u32,
                                             unsigned int cluster_stride, unsigned int& slow) {
  if (!__any_sync(0xffffffffu, nd)) return;
  const float eps = fast::score_eps(v.x, v.w, lmax, cmax);
  const unsigned int idx = warp_exact_argmin<true>(g_tab, k, nd, v.x, v.y, v.z, v.w, eps, 0u);
  if (nd) {
    slot_add<SPAN, -1>(slot_tid_u32 | hit, v.x, v.y, v.z);
    slot_add<SPAN, 1>(slot_tid_u32 + idx * cluster_stride, v.x, v.y, v.z);
    ++slow;
  }
}

// FLAGS: 8 = timing experiment, the memory side alone (consumers wait, touch the stage and release
// it; the sums are dropped and the centroids kept)
template <int KT, int CWARPS, int P, int D, int MINB, int FLAGS = 0>
__global__ void __launch_bounds__((CWARPS + 1) * 32, MINB)
    k_lloyd_ring(JobPtrs J, const float4* __restrict__ work, unsigned long long n, int color_space, int distributed_mode,
                 PeerXchg X, int cslot, unsigned int k_arg) {
  static_assert(KT == 8, "the uniform-register table holds 8 centroids");
  static_assert(P == 2 || P == 4 || P == 8, "whole packed pixel pairs per lane and stage");
  using L = RingLayout<KT, CWARPS, P, D>;
  constexpr unsigned int THREADS = L::THREADS, CTHREADS = L::CTHREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ bool s_last;
  JobState* st = J.st;
  if (st->done) return;
  const unsigned int tid = threadIdx.x, lane = tid & 31u;
  const unsigned int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const unsigned int k = k_arg;

  // carve the dynamic shared memory (slot region aligned to its own span so that OR can replace ADD)
  const uint32_t raw = smem_u32(smem_raw);
  uint32_t slots_u32 = (raw + L::SLOT_SPAN - 1) & ~(L::SLOT_SPAN - 1);
  asm volatile("mov.u32 %0, %0;" : "+r"(slots_u32));  // opaque: keeps ptxas from re-deriving it inside the loop
  int* slots = reinterpret_cast<int*>(smem_raw + (slots_u32 - raw));
  const uint32_t ring_u32 = slots_u32 + 4 * L::SLOT_SPAN;
  const uint32_t full_u32 = ring_u32 + L::RING_BYTES, empty_u32 = full_u32 + D * 8;

  if (tid < CTHREADS) {
#pragma unroll 4
    for (int c = 0; c < 4 * KT; ++c) slots[c * CTHREADS + tid] = 0;
  }
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      mbar_init(full_u32 + d * 8, 1);
      mbar_init(empty_u32 + d * 8, CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();

  // block tiles (= stages) of CTHREADS * P pixels, dealt round-robin to the blocks of the grid
  constexpr unsigned long long BT = (unsigned long long)CTHREADS * P;
  const unsigned long long full_tiles = n / BT;
  const unsigned int my_tiles =
      blockIdx.x < full_tiles ? (unsigned int)((full_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  unsigned int slow = 0;

  if (warp == CWARPS) {
    // ---- producer warp: one bulk copy per stage ----------------------------------------------------
    const uint64_t policy = l2_evict_first_policy();
    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(work) + (unsigned long long)blockIdx.x * (BT * 16);
    const unsigned long long gstep = (unsigned long long)gridDim.x * (BT * 16);
    for (unsigned int t = 0; t < my_tiles; ++t) {
      const unsigned int s = t & (D - 1);
      if (t >= D) mbar_wait(empty_u32 + s * 8, ((t / D) - 1u) & 1u);
      if (lane == 0) {
        mbar_expect_tx(full_u32 + s * 8, L::STAGE_BYTES);
        bulk_g2s(ring_u32 + s * L::STAGE_BYTES, gsrc, L::STAGE_BYTES, full_u32 + s * 8, policy);
      }
      gsrc += gstep;
    }
  } else {
    // ---- consumer warps ------------------------------------------------------------------------------
    const uint32_t ring_lane_u32 = ring_u32 + warp * L::WARP_BYTES + lane * 16;
    const uint32_t slot_tid_u32 = slots_u32 + tid * 4;  // component 0, cluster 0 of this thread
    int* slot_tid = slots + tid;
    unsigned long long* g_acc = reinterpret_cast<unsigned long long*>(J.acc + (size_t)(blockIdx.x % J.acc_copies) * k * 4);
    unsigned int since_flush = 0;

    // The table, loaded once: q[1..5] of the 8 records through warp-uniform constant loads (they land
    // in uniform registers, the scalar operand of the packed FFMA2), q[0] — the addend of the first
    // FMA, whose other scalar already is uniform — through a per-thread address into vector registers.
    // `volatile` keeps ptxas from re-materialising the loads inside the loop.
    float tq[KT][6];
    {
      const size_t ct = __cvta_generic_to_constant(c_tab[cslot]);  // ld.const takes a constant-space address
      const unsigned int zero = threadIdx.y;                       // always 0, but not provably uniform
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        asm volatile("ld.const.f32 %0, [%1];" : "=f"(tq[j][0]) : "l"(ct + 4 * (6 * j + zero)));
#pragma unroll
        for (int q = 1; q < 6; ++q) asm volatile("ld.const.f32 %0, [%1];" : "=f"(tq[j][q]) : "l"(ct + 4 * (6 * j + q)));
        // -Lc * 2^9 by an integer addition to the exponent field (stays in the uniform datapath; a zero
        // becomes 2^-118, which still multiplies to nothing)
        tq[j][1] = __int_as_float(__float_as_int(tq[j][1]) + (9 << 23));
        tq[j][3] = __int_as_float(__float_as_int(tq[j][3]) + (9 << 23));  // C2 * 2^9 likewise
      }
    }
    const float lmax_u = lmax * 0.00138106793f, cmax_v = cmax * 0.001953125f;  // see fast::score_eps

    // every warp folds the private slots of its own 32 threads (no block sync: a warp only reads what
    // it wrote) and sends 4 reductions per non-empty cluster to L2; the magic-number bias leaves here
    auto flush = [&]() {
      for (unsigned int c = 0; c < k; ++c) {
        int* s = slot_tid + c * CTHREADS;
        const int cnt = s[3 * KT * CTHREADS];
        const int v0 = s[0] - cnt * (int)FIXED_MAGIC_BITS;
        const int v1 = s[KT * CTHREADS] - cnt * (int)FIXED_MAGIC_BITS;
        const int v2 = s[2 * KT * CTHREADS] - cnt * (int)FIXED_MAGIC_BITS;
        s[0] = 0;
        s[KT * CTHREADS] = 0;
        s[2 * KT * CTHREADS] = 0;
        s[3 * KT * CTHREADS] = 0;
        const long long s3 = (long long)__reduce_add_sync(0xffffffffu, cnt);
        if (s3 == 0) continue;  // warp-uniform
        const long long s0 = warp_sum_i32(v0), s1 = warp_sum_i32(v1), s2 = warp_sum_i32(v2);
        if (lane == 0) {
          atomicAdd(g_acc + c * 4 + 0, (unsigned long long)s0);
          atomicAdd(g_acc + c * 4 + 1, (unsigned long long)s1);
          atomicAdd(g_acc + c * 4 + 2, (unsigned long long)s2);
          atomicAdd(g_acc + c * 4 + 3, (unsigned long long)s3);
        }
      }
      since_flush = 0;
    };

    // |v| < 2^7 colour units -> |fixed| < 2^22: 480 pixels per slot stay below 2^31
    constexpr unsigned int FLUSH_PX = 480;
    using RC = RingCert<KT, CTHREADS * 4>;  // slot stride of a cluster in bytes
    constexpr unsigned int IDX_MASK = RC::IDX_MASK, CERT_MASK = RC::CERT_MASK, CERT_ONE = RC::CERT_ONE;

    unsigned int t = 0;  // stages (tiles) of this block done so far by this warp
    while (t < my_tiles) {
      const unsigned int t_stop = min(my_tiles, t + (FLUSH_PX - since_flush) / P);
      const unsigned int t_begin = t;
      unsigned int need = 0, hit[P];
      float4 vk[P];  // the pixels of the stage that left the loop (its ring slot has been handed back)
      // ---- hot loop: one stage (P / 2 pixel pairs per lane) per iteration, no calls, table resident ----
      for (; t < t_stop; ++t) {
        const unsigned int s = t & (D - 1);
        mbar_wait(full_u32 + s * 8, (t / D) & 1u);
        const uint32_t stage_addr = ring_lane_u32 + s * L::STAGE_BYTES;
        float4 v[P];
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = lds128(stage_addr + i * 512);
        // The pixels are in registers: hand the stage back at once (not after the arithmetic), so the
        // refill is in flight while this warp computes.  The arrive must not overtake the loads: its
        // address is made to depend on the last value of every one of them (chroma >= 0: the sign bits
        // are 0), which makes the hardware wait for the data first.
        {
          unsigned int dep = 0;
#pragma unroll
          for (int i = 0; i < P; ++i) dep |= __float_as_uint(v[i].w);
          if (lane == 0) mbar_arrive(empty_u32 + s * 8 + (dep >> 31));
        }
        if (FLAGS & 8) {  // timing experiment: the memory side alone
          if (v[0].x + v[P - 1].x == 12345.678f) slow++;
          continue;
        }
        bool all_cert = true;
#pragma unroll
        for (int h = 0; h < P / 2; ++h) {
          unsigned int ua, ub;
          float fla, flb;
          ring_pair_search<KT, CTHREADS * 4>(v[2 * h], v[2 * h + 1], tq, lmax_u, cmax_v, ua, ub, fla, flb);
          // every pixel goes to the slot its flag sum points at; the rare uncertified one is moved by
          // the exact path below (no predicate, no branch around the reductions)
          slot_add<L::SLOT_SPAN, 1, true>((ua & IDX_MASK) | slot_tid_u32, fla, v[2 * h].y, v[2 * h].z);
          slot_add<L::SLOT_SPAN, 1, true>((ub & IDX_MASK) | slot_tid_u32, flb, v[2 * h + 1].y, v[2 * h + 1].z);
          all_cert &= (ua & CERT_MASK) == CERT_ONE && (ub & CERT_MASK) == CERT_ONE;
          hit[2 * h] = ua;
          hit[2 * h + 1] = ub;
        }
        if (__any_sync(0xffffffffu, !all_cert)) {
#pragma unroll
          for (int i = 0; i < P; ++i) {
            need |= ((hit[i] & CERT_MASK) == CERT_ONE ? 0u : 1u) << i;
            vk[i] = v[i];
          }
          break;
        }
      }
      // ---- cold: exact path of the stage that left the loop, slot flush ------------------------------
      if (t < t_stop) {
#pragma unroll
        for (int i = 0; i < P; ++i)
          ring_slow_pixel<L::SLOT_SPAN>(J.tab, k, vk[i], (need >> i) & 1u, hit[i] & IDX_MASK, lmax, cmax, slot_tid_u32,
                                        CTHREADS * 4, slow);
        __syncwarp();
        ++t;
      }
      since_flush += P * (t - t_begin);
      if (since_flush + P > FLUSH_PX) flush();
    }

    // ragged tail (< CTHREADS * P pixels): the block whose turn it would be, straight from global
    // memory; every tail pixel takes the exact path with an unbounded slack (the in-order scan)
    if (full_tiles * BT < n && blockIdx.x == (unsigned int)(full_tiles % gridDim.x)) {
      for (unsigned long long p0 = full_tiles * BT + warp * 32; p0 < n; p0 += CTHREADS) {
        if (since_flush + 1 > FLUSH_PX) flush();
        const unsigned long long p = p0 + lane;
        const bool valid = p < n;
        const float4 v = valid ? work[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        const unsigned int idx = warp_exact_argmin<true>(J.tab, k, valid, v.x, v.y, v.z, v.w, 3.0e38f, 0u);
        if (valid) slot_add<L::SLOT_SPAN>(slot_tid_u32 + idx * CTHREADS * 4, v.x, v.y, v.z);
        ++since_flush;
      }
    }
    flush();
    if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);
  }

  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (FLAGS & 8) {
      // timing experiments produce wrong sums: drop them, keep the centroids, count the pass
      for (unsigned int c = tid; c < J.acc_copies * k * 4; c += THREADS) J.acc[c] = 0;
      if (tid == 0) {
        st->passes += 1;
        st->ticket = 0;
      }
    } else {
      finalize_pass<THREADS>(J, color_space, distributed_mode, X);
    }
  }
}

}  // namespace kmg
