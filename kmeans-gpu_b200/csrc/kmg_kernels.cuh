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

// One centroid as the kernels see it: the six constants of the reduced (half-distance) score,
//   q = {Lc^2/2, -Lc, C2^2/2, C2, -ac, -bc}   (24 bytes, records are dense: 8 of them = twelve float4).
// Packed FFMA2 takes them as scalar-broadcast operands, so nothing is duplicated.  Duplicates of
// a lower-index centroid and padding entries carry q[0] = MASKED so they can never win (the
// reference's strict '<' scan keeps the lowest index on exact ties anyway).
struct __align__(8) CentRec {
  float q[6];
};
constexpr float MASKED = 1.0e30f;
constexpr int MAX_K = 4096;  // table = 104 KiB of shared memory at most

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
  // lazy farthest-point initialisation (kmg_init_lazy.cuh)
  unsigned int init_tau16;       // candidates of the next sweep: pixels whose 16-bit upper bound is >= this
  unsigned int init_ncmax;       // 1 + the largest 16-bit bound among the pixels the sweep skipped (0: none skipped)
  unsigned int init_done_round;  // last round resolved
  unsigned int init_attempts;    // sweeps so far (statistics)
  unsigned long long init_refreshed;  // pixel refreshes so far (statistics)
  unsigned long long init_folds;      // (pixel, centroid) pairs folded so far (statistics)
  unsigned long long init_exact;      // of those, pairs that needed the exact distance (statistics)
};

struct JobPtrs {
  JobState* st;
  float4* cent;                // k
  CentRec* tab;                // k padded to a multiple of 32 with MASKED entries
  long long* acc;              // acc_copies x k x 4  (sum0,sum1,sum2,count), fixed-point 2^-15
  long long* last;             // k x 4 — the reduced sums of the last finalised pass
  unsigned long long* keys;    // k  — arg-max keys of the init rounds
  uint32_t* pal;               // k  — centroids reverted to RGBA8
  unsigned int acc_copies;     // privatised accumulator copies (block b adds into copy b % acc_copies)
  float* cbig;                 // c_big addressed as global memory when this job holds it (NULL otherwise)
  float* ctab;                 // the job's slot of the constant-bank table c_tab, addressed as global memory
                               // (NULL: none): build_table refreshes it, so that no copy has to be enqueued
                               // between two passes (constant caches are invalidated at kernel boundaries)
};

__host__ __device__ inline unsigned int pad32(unsigned int k) { return (k + 31u) & ~31u; }

// Multi-GPU pixel sharding without a separate collective: the last block of a Lloyd pass on every
// GPU stores its k x 4 partial sums straight into every peer's mailbox (peer-mapped memory, the
// stores travel over NVLink / NVSwitch), waits for the peers' sums in its own mailbox and finalises
// — reduction and pass are one kernel.  Every 64-bit value travels as two 8-byte words
// {32 payload bits, sequence number}: an aligned 8-byte store arrives whole, so a word that shows the
// expected sequence number is valid by itself — no fence, no separate flag, one NVLink flight.
// Mailbox of a GPU:
//   mbox  [2 parities][MAX_PEERS ranks][xcap] words     flags [2 parities][MAX_PEERS ranks] u32 (poison only)
constexpr unsigned int MAX_PEERS = 8;
struct PeerXchg {
  unsigned int n_ranks;  // 0: not in use
  unsigned int rank;
  unsigned int xcap;     // 8-byte words per (parity, rank), >= 8 k
  unsigned int seq_base; // words carry seq_base + pass number + 1 (distinct per job)
  long long* mbox[MAX_PEERS];      // rank r's mailbox as mapped on this GPU
  unsigned int* flags[MAX_PEERS];
};
constexpr unsigned int PASS_FAULT = 0xffffffffu;  // JobState::conv when a peer did not answer
constexpr unsigned int XCHG_POISON = 0xdeadbeefu; // flag value a rank that gave up leaves in every mailbox

// Word w of this rank's stretch (pass parity par) in peer r's mailbox / of rank r's stretch in this GPU's own.
__device__ __forceinline__ volatile unsigned long long* xchg_out(const PeerXchg& X, unsigned int par, unsigned int r,
                                                                 size_t w) {
  return reinterpret_cast<volatile unsigned long long*>(X.mbox[r]) + ((size_t)par * MAX_PEERS + X.rank) * X.xcap + w;
}
__device__ __forceinline__ const volatile unsigned long long* xchg_in(const PeerXchg& X, unsigned int par, unsigned int r,
                                                                       size_t w) {
  return reinterpret_cast<const volatile unsigned long long*>(X.mbox[X.rank]) + ((size_t)par * MAX_PEERS + r) * X.xcap + w;
}
// 64-bit value number i of this rank's stretch, to peer r: words 2 i and 2 i + 1
__device__ __forceinline__ void xchg_post64(const PeerXchg& X, unsigned int par, unsigned int seq, unsigned int r, size_t i,
                                            unsigned long long v) {
  volatile unsigned long long* d = xchg_out(X, par, r, 2 * i);
  d[0] = ((unsigned long long)seq << 32) | (v & 0xffffffffull);
  d[1] = ((unsigned long long)seq << 32) | (v >> 32);
}
// A rank that waits longer than 4 s for a word (a peer never launched the matching kernel) gives up
// — and poisons its flag in EVERY mailbox, both parities, so that the peers fault in their current or
// next exchange as well instead of finishing the pass and running on with centroids this rank never
// got: all ranks of a job fail together (KMG_ERR_NCCL on each), and the communicator stays poisoned
// until kmg_comm_destroy / kmg_comm_init.
__device__ __forceinline__ bool xchg_gave_up(const PeerXchg& X, unsigned int par, unsigned int r, unsigned long long t0,
                                          unsigned int* s_fault) {
  bool fault = *(volatile unsigned int*)(X.flags[X.rank] + par * MAX_PEERS + r) == XCHG_POISON;
  if (!fault) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    fault = t1 - t0 > 4000000000ull;
  }
  if (fault) {
    *s_fault = 1;
    for (unsigned int q = 0; q < X.n_ranks; ++q) {
      *(volatile unsigned int*)(X.flags[q] + X.rank) = XCHG_POISON;
      *(volatile unsigned int*)(X.flags[q] + MAX_PEERS + X.rank) = XCHG_POISON;
    }
  }
  return fault;
}
// 64-bit value number i of every rank's stretch, summed in rank order (ADD) or handed back one by one.
// All 2 n words are requested before the first one is looked at; stragglers are polled again.
template <typename F>
__device__ __forceinline__ void xchg_gather64(const PeerXchg& X, unsigned int par, unsigned int seq, size_t i,
                                              unsigned int* s_fault, F&& use) {
  unsigned long long t0 = 0;
#pragma unroll 1
  for (unsigned int r0 = 0; r0 < X.n_ranks; r0 += 4) {
    unsigned long long lo[4], hi[4];
#pragma unroll
    for (unsigned int q = 0; q < 4; ++q) {
      if (r0 + q < X.n_ranks) {
        const volatile unsigned long long* a = xchg_in(X, par, r0 + q, 2 * i);
        lo[q] = a[0];
        hi[q] = a[1];
      }
    }
#pragma unroll
    for (unsigned int q = 0; q < 4; ++q) {
      if (r0 + q < X.n_ranks) {
        const volatile unsigned long long* a = xchg_in(X, par, r0 + q, 2 * i);
        unsigned int spins = 0;
        while ((unsigned int)(lo[q] >> 32) != seq || (unsigned int)(hi[q] >> 32) != seq) {
          if (t0 == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          if ((spins++ & 255u) == 0 && xchg_gave_up(X, par, r0 + q, t0, s_fault)) break;
          lo[q] = a[0];
          hi[q] = a[1];
        }
        use(r0 + q, (lo[q] & 0xffffffffull) | (hi[q] << 32));
      }
    }
  }
}

// The job blob of frame f in a batch: every pointer shifted by f * blob_stride bytes.
__device__ __forceinline__ JobPtrs job_at(JobPtrs J, size_t off) {
  JobPtrs R;
  R.st = reinterpret_cast<JobState*>(reinterpret_cast<unsigned char*>(J.st) + off);
  R.cent = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(J.cent) + off);
  R.tab = reinterpret_cast<CentRec*>(reinterpret_cast<unsigned char*>(J.tab) + off);
  R.acc = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(J.acc) + off);
  R.last = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(J.last) + off);
  R.keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(J.keys) + off);
  R.pal = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(J.pal) + off);
  R.acc_copies = J.acc_copies;
  R.ctab = nullptr;
  R.cbig = nullptr;
  return R;
}

// Shared-memory copy of the table: every chunk of 8 records (192 B) is followed by 16 B of padding,
// so chunk bases advance by an odd number of 16-byte bank groups and lanes that read *different*
// chunks (winning-chunk rescan, cooperative exact path) do not all collide on the same banks.
constexpr unsigned int CHUNK_BYTES = 8 * 24 + 16;
__host__ __device__ inline size_t tab_smem_bytes(unsigned int kp) { return (size_t)(kp / 8) * CHUNK_BYTES; }
__device__ __forceinline__ const CentRec* rec_at(const CentRec* tab, unsigned int j) {
  return reinterpret_cast<const CentRec*>(reinterpret_cast<const unsigned char*>(tab) + (size_t)(j >> 3) * CHUNK_BYTES +
                                          (j & 7u) * 24u);
}
__device__ __forceinline__ void tab_to_smem(CentRec* s_tab, const CentRec* __restrict__ g_tab, unsigned int kp,
                                            unsigned int tid, unsigned int nthreads) {
  for (unsigned int c = tid; c < kp; c += nthreads) *const_cast<CentRec*>(rec_at(s_tab, c)) = g_tab[c];
}

// Constant-bank copy of a small table (k <= 16), dense records, one slot per job in flight.  Values
// read from the constant bank with a warp-uniform address live in *uniform* registers, and packed
// FFMA2 takes a scalar-broadcast operand straight from a uniform register: the hot loop of a
// small-k Lloyd pass then keeps no table in vector registers and issues no shared-memory loads for
// it, which buys a third resident block per SM.  The host copies J.tab into the job's slot
// (device-to-device, 24 k bytes) before each pass.
constexpr int CTAB_SLOTS = 64;
constexpr int CTAB_RECORDS = 16;
constexpr int CTAB_FLOATS = CTAB_RECORDS * 6;
__constant__ float c_tab[CTAB_SLOTS][CTAB_FLOATS];
// One large table (k <= CBIG_MAX_K) for the chunked search, laid out per chunk of 8 centroids as
// [q0 x 8 | q1 x 8 | ... | q5 x 8]: the five multiplier rows of a chunk arrive with ten 128-bit
// warp-uniform loads in uniform registers, where a packed FFMA2 reads its scalar operand without
// touching the register-file banks of its two vector pairs (2.0 instead of 2.25-2.5 cycles per FFMA2
// with a vector-register scalar).  One job per device holds it at a time.
constexpr unsigned int CBIG_MAX_K = 1024;
__constant__ float c_big[CBIG_MAX_K * 6];
// (Feeding the chunk loop of the k > 32 search from the constant bank the same way was measured
// slower — 48 uniform registers per chunk leave no room to prefetch the next chunk — and removed.)

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
// Sum of one int32 per lane (any values) with two REDUX: the signed high 16 bits and the unsigned
// low 16 bits are reduced separately, neither partial sum can overflow 32 bits.
__device__ __forceinline__ long long warp_sum_i32(int v) {
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  const unsigned int lo = __reduce_add_sync(0xffffffffu, (unsigned int)v & 0xffffu);
  return (long long)hi * 65536 + (long long)lo;
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
      r.q[0] = dup ? MASKED : 0.5f * (v.x * v.x);
      r.q[1] = -v.x;
      r.q[2] = 0.5f * (c2 * c2);
      r.q[3] = c2;
      r.q[4] = -v.y;
      r.q[5] = -v.z;
      lmax = fmaxf(lmax, fabsf(v.x));
      cmax = fmaxf(cmax, c2);
      if (want_palette)
        J.pal[c] = color_space == 0 ? ex::lab_to_rgba8(v.x, v.y, v.z) : ex::rgbf_to_rgba8(v.x, v.y, v.z, v.w);
    } else {
      r.q[0] = MASKED;
      r.q[1] = r.q[2] = r.q[3] = r.q[4] = r.q[5] = 0.0f;
    }
    J.tab[c] = r;
    if (J.ctab && c < CTAB_RECORDS) {
#pragma unroll
      for (int q = 0; q < 6; ++q) J.ctab[6 * c + q] = r.q[q];
    }
    if (J.cbig && c < CBIG_MAX_K) {
#pragma unroll
      for (int q = 0; q < 6; ++q) J.cbig[(c >> 3) * 48 + q * 8 + (c & 7u)] = r.q[q];
    }
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
  }
  // Only the remap needs the dither threshold (k_prepare passes want_palette): 2 (k - 2) exact
  // distances walked by one thread have no place in the tail of every Lloyd pass.
  if (tid == 0 && want_palette) {
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

// The current table into c_big (chunk-major rows, see c_big) when a job acquires it between two passes.
__global__ void __launch_bounds__(256) k_big_table(JobPtrs J, unsigned int kp) {
  const unsigned int c = blockIdx.x * 256 + threadIdx.x;
  if (c < kp && c < CBIG_MAX_K) {
    const CentRec r = J.tab[c];
#pragma unroll
    for (int q = 0; q < 6; ++q) J.cbig[(c >> 3) * 48 + q * 8 + (c & 7u)] = r.q[q];
  }
}

__global__ void __launch_bounds__(256) k_prepare(JobPtrs J, int color_space, int want_palette) {
  build_table<256>(J, J.st->k, color_space, want_palette != 0);
}

// ------------------------------------------------------------------------------------------------
// Certified nearest-centroid search over a shared-memory table.
//
// Fast pass: the reduced CIE94 score (5 FMA per pixel x centroid, see kmg_math.cuh) evaluated for
// two pixels at a time with Blackwell's packed FFMA2, the table holding every value as a
// duplicated pair.  A pixel is *certified* when no other centroid's score lies within eps of the
// best one, eps bounding every rounding difference between the fast score and the reference's f32
// distance (plus, in the remap kernels, the error of the approximate Lab).  The certificate
// "S = sum_j sat((s_j - m1)/eps) >= K - 1" runs on the FMA pipe, and the same terms give the index:
// sum_j j*t_j = sum_b 2^b * (sum of t_j over j with bit b set), evaluated as a pairwise tree.
// Uncertified pixels are re-evaluated with exact reference arithmetic (warp_exact_argmin), so
// labels are the reference's bit for bit, not "up to near ties".

template <int P>
struct Pix {
  float L[P], a[P], b[P], C[P];
};

__device__ __forceinline__ float tournament8(const float (&s)[8]) {
  float m = fast::min3(s[0], s[1], s[2]);
  m = fast::min3(m, s[3], s[4]);
  m = fast::min3(m, s[5], s[6]);
  return fminf(m, s[7]);
}

// Certificate + index over KT saved scores of one pixel pair (sa: first pixel, sb: second).
// Flag t_j = (s_j > m + eps) is an exact 0/1 (FSET on the ALU pipe); the weighted sum
//   V = sum_j (KT + j) * (1 - t_j) = KT * z + (sum of the indices of the z unflagged scores)
// is accumulated exactly in the integer range of an f32 sitting on 2^23, two pixels per FFMA2, so
// the low mantissa bits of the result are V itself: z == 1  <=>  (V & ~(KT-1)) == KT, and then
// V & (KT-1) is the index of the only score within eps of the minimum, i.e. the certified arg-min.
template <int KT>
__device__ __forceinline__ void certify_pair(const float (&sa)[KT], const float (&sb)[KT], float ma, float mb,
                                             float ea, float eb, bool& ca, bool& cb, unsigned int& ia,
                                             unsigned int& ib) {
  static_assert((KT & (KT - 1)) == 0 && KT <= 64, "table length must be a small power of two");
  const float ta = ma + ea, tb = mb + eb;
  constexpr float TOTAL = (float)(KT * KT + KT * (KT - 1) / 2);
  fast::f32x2 acc0 = fast::pack2(8388608.0f + TOTAL, 8388608.0f + TOTAL);
  fast::f32x2 acc1 = fast::pack2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float fa = sa[j] > ta ? 1.0f : 0.0f;
    const float fb = sb[j] > tb ? 1.0f : 0.0f;
    const float w = -(float)(KT + j);
    if (j & 1)
      acc1 = fast::fma2(fast::pack2(fa, fb), fast::pack2(w, w), acc1);
    else
      acc0 = fast::fma2(fast::pack2(fa, fb), fast::pack2(w, w), acc0);
  }
  float Va, Vb;
  fast::unpack2(fast::add2(acc0, acc1), Va, Vb);
  const unsigned int ua = __float_as_uint(Va), ub = __float_as_uint(Vb);
  constexpr unsigned int HI = 0xffffu & ~(unsigned int)(KT - 1);
  ca = (ua & HI) == (unsigned int)KT;
  cb = (ub & HI) == (unsigned int)KT;
  ia = ua & (unsigned int)(KT - 1);
  ib = ub & (unsigned int)(KT - 1);
}

__device__ __forceinline__ void pack_coefs(const fast::PixCoef& c0, const fast::PixCoef& c1, fast::f32x2 (&pp)[5]) {
  pp[0] = fast::pack2(c0.p0, c1.p0);
  pp[1] = fast::pack2(c0.p1, c1.p1);
  pp[2] = fast::pack2(c0.p2, c1.p2);
  pp[3] = fast::pack2(c0.p3, c1.p3);
  pp[4] = fast::pack2(c0.p4, c1.p4);
}
__device__ __forceinline__ fast::f32x2 score2(const fast::f32x2 (&pp)[5], const float* q) {
  fast::f32x2 s = fast::fma2(pp[0], fast::pack2(q[1], q[1]), fast::pack2(q[0], q[0]));
  s = fast::fma2(pp[1], fast::pack2(q[2], q[2]), s);
  s = fast::fma2(pp[2], fast::pack2(q[3], q[3]), s);
  s = fast::fma2(pp[3], fast::pack2(q[4], q[4]), s);
  s = fast::fma2(pp[4], fast::pack2(q[5], q[5]), s);
  return s;
}
__device__ __forceinline__ float score1(const fast::PixCoef& p, const float* q) {
  float s = fmaf(p.p0, q[1], q[0]);
  s = fmaf(p.p1, q[2], s);
  s = fmaf(p.p2, q[3], s);
  s = fmaf(p.p3, q[4], s);
  s = fmaf(p.p4, q[5], s);
  return s;
}
// One chunk of 8 records = twelve 128-bit shared-memory loads into registers.
__device__ __forceinline__ void load_chunk(const CentRec* chunk, float (&f)[48]) {
  const float4* c4 = reinterpret_cast<const float4*>(chunk);
#pragma unroll
  for (int u = 0; u < 12; ++u) {
    const float4 t = c4[u];
    f[4 * u + 0] = t.x;
    f[4 * u + 1] = t.y;
    f[4 * u + 2] = t.z;
    f[4 * u + 3] = t.w;
  }
}

// Error bound of a score gap.  CONV (remap kernels): the pixel itself is approximate (fast Lab,
// approximate chroma, dither offset added to the approximate value): |p~ - p| <= lab_err.  A score
// gap between two candidates then moves by at most lab_err * (|grad s_c| + |grad s_c'|), s = d^2/2.
// With tC = dC/SC, tH = dH/SH (both <= d = sqrt(d^2)), u = (a,b)/C1:
//   d(dL^2)/dL           = 2 dL                                          <= 2 d
//   d(tC^2)/d(a,b)       = 2 tC (1 - 0.045 tC) u / SC                    <= 2 d (1 + 0.045 d)
//   d(dH^2/SH^2)/d(a,b)  = 2 (C2 u - (a2,b2)) / SH^2 - 0.03 tH^2 u / SH  <= 4 Cmax / SH^2 + 0.03 d^2
// so |grad d^2| <= G(d) = 4 d + 0.12 d^2 + 4 Cmax / SH^2 for either candidate (their distances
// differ by less than the bound being computed; d^2 is inflated by 2 eps0 and G by 2 %).
// d^2 = 2 * best score + the pixel-only part of the squared distance, L^2 + C^2 / SC^2.
template <bool CONV>
__device__ __forceinline__ float total_eps(float eps0, float m, float lab_err, float L, float C, float inv_sc2, float hs,
                                           float cmax) {
  if (!CONV) return eps0;
  const float pconst = fmaf(L, L, C * C * inv_sc2);
  const float d2 = fmaxf(fmaf(2.0f, m, pconst), 0.0f) + 2.0f * eps0;
  const float d = fast::sqrt_approx(d2);
  const float G = fmaf(d, 4.0f, fmaf(0.12f, d2, 4.0f * cmax * hs));
  return fmaf(1.02f * lab_err, G, eps0) + 1.5e-6f;
}

// Small tables (KT = 8 or 16, compile time): all KT scores of a pixel stay in registers.
template <int P, int KT, bool CONV, bool CT = false>
__device__ __forceinline__ void argmin_small(const CentRec* __restrict__ tab, const Pix<P>& px, float lmax,
                                             float cmax, float conv_k, float (&eps)[P], unsigned int (&idx)[P],
                                             bool (&certified)[P], float* thr_out = nullptr) {
  static_assert(P % 2 == 0, "pairs of pixels");
  constexpr int H = P / 2;
  fast::f32x2 pp[H][5];
  float inv_sc2[P], inv_sh2[P];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const fast::PixCoef c0 = fast::pix_coef(px.L[2 * h], px.a[2 * h], px.b[2 * h], px.C[2 * h]);
    const fast::PixCoef c1 = fast::pix_coef(px.L[2 * h + 1], px.a[2 * h + 1], px.b[2 * h + 1], px.C[2 * h + 1]);
    inv_sc2[2 * h] = c0.p1;
    inv_sc2[2 * h + 1] = c1.p1;
    inv_sh2[2 * h] = c0.hs;
    inv_sh2[2 * h + 1] = c1.hs;
    pack_coefs(c0, c1, pp[h]);
  }
  fast::f32x2 s2[KT][H];
#pragma unroll
  for (int c = 0; c < KT; c += 8) {
    float f[48];
    if (CT) {
      // tab points into the constant bank (dense records): uniform-register operands
      const float* ct = reinterpret_cast<const float*>(tab) + 6 * c;
#pragma unroll
      for (int u = 0; u < 48; ++u) f[u] = ct[u];
    } else {
      load_chunk(rec_at(tab, c), f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int h = 0; h < H; ++h) s2[c + j][h] = score2(pp[h], f + 6 * j);
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
    ma = fminf(ma, sa[KT - 1]);
    mb = fminf(mb, sb[KT - 1]);
    eps[2 * h] = total_eps<CONV>(fast::score_eps(px.L[2 * h], px.C[2 * h], lmax, cmax), ma, conv_k, px.L[2 * h],
                                 px.C[2 * h], inv_sc2[2 * h], inv_sh2[2 * h], cmax);
    eps[2 * h + 1] = total_eps<CONV>(fast::score_eps(px.L[2 * h + 1], px.C[2 * h + 1], lmax, cmax), mb, conv_k,
                                     px.L[2 * h + 1], px.C[2 * h + 1], inv_sc2[2 * h + 1], inv_sh2[2 * h + 1], cmax);
    certify_pair<KT>(sa, sb, ma, mb, eps[2 * h], eps[2 * h + 1], certified[2 * h], certified[2 * h + 1], idx[2 * h],
                     idx[2 * h + 1]);
    if (thr_out) {  // smallest fast score + eps: what the exact path admits candidates up to
      thr_out[2 * h] = ma + eps[2 * h];
      thr_out[2 * h + 1] = mb + eps[2 * h + 1];
    }
  }
}

// Any table length (kp = multiple of 8, padded with MASKED entries): centroids are visited in
// chunks of 8; only each chunk's minimum is kept (3-input-min tournament) together with the best
// and second-best chunk minima.  The precise certificate then runs on the winning chunk alone, so
// the half-rate ALU pipe sees ~1.1 min/select operations per (pixel, centroid) instead of 5 and
// the loop is bound by the 5 FMAs of the score.
template <int P, bool CONV, bool BIG = false>
__device__ __forceinline__ void argmin_chunked(const CentRec* __restrict__ tab, unsigned int kp, const Pix<P>& px,
                                               float lmax, float cmax, float conv_k, float (&eps)[P],
                                               unsigned int (&idx)[P], bool (&certified)[P], float* thr_out = nullptr) {
  static_assert(P % 2 == 0, "pairs of pixels");
  constexpr int H = P / 2;
  fast::PixCoef pc[P];
  fast::f32x2 pp[H][5];
#pragma unroll
  for (int i = 0; i < P; ++i) pc[i] = fast::pix_coef(px.L[i], px.a[i], px.b[i], px.C[i]);
#pragma unroll
  for (int h = 0; h < H; ++h) pack_coefs(pc[2 * h], pc[2 * h + 1], pp[h]);
  float m1c[P], m2c[P];
  unsigned int ic[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    m1c[i] = 3.0e38f;
    m2c[i] = 3.0e38f;
    ic[i] = 0;
  }
  // BIG: multipliers from c_big through warp-uniform 128-bit constant loads (uniform registers), the
  // addends through a per-thread address (vector registers: an FFMA2 takes one uniform operand).  Rows
  // 1-2 of the next chunk are requested while rows 3-5 of this one are being used, rows 3-5 of this one
  // while its rows 1-2 are being used, so no FFMA2 waits for the constant cache.
  const size_t cb = BIG ? __cvta_generic_to_constant(c_big) : 0;
  const unsigned int zero = BIG ? threadIdx.y : 0u;  // always 0, but not provably uniform
  float q0n[8], q1n[8], q2n[8];
  auto ldc4u = [&](float* d, size_t addr) {
    asm volatile("ld.const.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(addr));
  };
  if (BIG) {
    ldc4u(q0n, cb + 4 * zero);
    ldc4u(q0n + 4, cb + 16 + 4 * zero);
    ldc4u(q1n, cb + 32);
    ldc4u(q1n + 4, cb + 48);
    ldc4u(q2n, cb + 64);
    ldc4u(q2n + 4, cb + 80);
  }
  for (unsigned int c = 0; c < kp; c += 8) {
    fast::f32x2 s2[8][H];
    if (BIG) {
      const size_t base = cb + (size_t)(c >> 3) * 192;
      float q3[8], q4[8], q5[8];
      ldc4u(q3, base + 96);
      ldc4u(q3 + 4, base + 112);
      ldc4u(q4, base + 128);
      ldc4u(q4 + 4, base + 144);
      ldc4u(q5, base + 160);
      ldc4u(q5 + 4, base + 176);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          s2[j][h] = fast::fma2(pp[h][0], fast::pack2(q1n[j], q1n[j]), fast::pack2(q0n[j], q0n[j]));
          s2[j][h] = fast::fma2(pp[h][1], fast::pack2(q2n[j], q2n[j]), s2[j][h]);
        }
      }
      if (c + 8 < kp) {  // rows 0-2 of the next chunk
        ldc4u(q0n, base + 192 + 4 * zero);
        ldc4u(q0n + 4, base + 208 + 4 * zero);
        ldc4u(q1n, base + 224);
        ldc4u(q1n + 4, base + 240);
        ldc4u(q2n, base + 256);
        ldc4u(q2n + 4, base + 272);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          s2[j][h] = fast::fma2(pp[h][2], fast::pack2(q3[j], q3[j]), s2[j][h]);
          s2[j][h] = fast::fma2(pp[h][3], fast::pack2(q4[j], q4[j]), s2[j][h]);
          s2[j][h] = fast::fma2(pp[h][4], fast::pack2(q5[j], q5[j]), s2[j][h]);
        }
      }
    } else {
      float f[48];
      load_chunk(rec_at(tab, c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int h = 0; h < H; ++h) s2[j][h] = score2(pp[h], f + 6 * j);
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float sa[8], sb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) fast::unpack2(s2[j][h], sa[j], sb[j]);
      const float cm[2] = {tournament8(sa), tournament8(sb)};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = 2 * h + q;
        m2c[i] = fminf(m2c[i], fmaxf(cm[q], m1c[i]));
        const bool lt = cm[q] < m1c[i];
        m1c[i] = lt ? cm[q] : m1c[i];
        ic[i] = lt ? c : ic[i];
      }
    }
  }
  // precise certificate + index inside the winning chunk (per-lane table addresses)
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float sa[8], sb[8];
    {
      float f[48];
      load_chunk(rec_at(tab, ic[2 * h]), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sa[j] = score1(pc[2 * h], f + 6 * j);
      load_chunk(rec_at(tab, ic[2 * h + 1]), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sb[j] = score1(pc[2 * h + 1], f + 6 * j);
    }
    const float ma = tournament8(sa), mb = tournament8(sb);
    const float ea = total_eps<CONV>(fast::score_eps(px.L[2 * h], px.C[2 * h], lmax, cmax), ma, conv_k, px.L[2 * h],
                                     px.C[2 * h], pc[2 * h].p1, pc[2 * h].hs, cmax);
    const float eb = total_eps<CONV>(fast::score_eps(px.L[2 * h + 1], px.C[2 * h + 1], lmax, cmax), mb, conv_k,
                                     px.L[2 * h + 1], px.C[2 * h + 1], pc[2 * h + 1].p1, pc[2 * h + 1].hs, cmax);
    bool ca, cb;
    unsigned int ia, ib;
    certify_pair<8>(sa, sb, ma, mb, ea, eb, ca, cb, ia, ib);
    eps[2 * h] = ea;
    eps[2 * h + 1] = eb;
    certified[2 * h] = ca && (m2c[2 * h] - ma > ea);
    certified[2 * h + 1] = cb && (m2c[2 * h + 1] - mb > eb);
    idx[2 * h] = ic[2 * h] + ia;
    idx[2 * h + 1] = ic[2 * h + 1] + ib;
    if (thr_out) {  // the winning chunk's minimum is the smallest of all fast scores
      thr_out[2 * h] = ma + ea;
      thr_out[2 * h + 1] = mb + eb;
    }
  }
}

// Exact re-evaluation, warp-cooperative.  Every lane of the warp must call this together (the
// callers' loops are warp-uniform).  For each lane whose `need` flag is set, the 32 lanes split the
// centroid range (j = lane, lane+32, ...): they find the pixel's smallest fast score, then evaluate
// the exact reference distance (delta_e.wgsl:1-22 order, IEEE) for every centroid whose fast score
// lies within `slack` of it, and reduce the 64-bit key (distance bits << 32 | index) with a min.
// The minimum of that key is the lowest index among the smallest distances, i.e. exactly the
// result of the reference's in-order scan with strict '<' starting from (100000.0, index 0)
// (find_centroid.wgsl:29-41).  (L,a,b,C) must be the exact pixel (C = exact chroma).
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// HAVE_BOUND: `slack` already is the absolute admission threshold (smallest fast score + eps, as the
// search that failed to certify computed it with the same arithmetic): the first sweep over the table,
// which only looks for that smallest score, is skipped.
template <bool DENSE = false, bool HAVE_BOUND = false>
__device__ __noinline__ unsigned int warp_exact_argmin(const CentRec* __restrict__ tab, unsigned int k, bool need,
                                                       float L, float a, float b, float C, float slack,
                                                       unsigned int idx_in) {
  unsigned int mask = __ballot_sync(0xffffffffu, need);
  const unsigned int lane = threadIdx.x & 31u;
  unsigned int result = idx_in;
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const float pL = __shfl_sync(0xffffffffu, L, src), pa = __shfl_sync(0xffffffffu, a, src);
    const float pb = __shfl_sync(0xffffffffu, b, src), pC = __shfl_sync(0xffffffffu, C, src);
    const float pslack = __shfl_sync(0xffffffffu, slack, src);
    const fast::PixCoef pc = fast::pix_coef(pL, pa, pb, pC);
    float bound = pslack;
    if (!HAVE_BOUND) {
      float m = 3.0e38f;
      for (unsigned int j = lane; j < k; j += 32) m = fminf(m, score1(pc, (DENSE ? tab + j : rec_at(tab, j))->q));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
      bound = m + pslack;
    }
    unsigned long long best = ~0ull;
    for (unsigned int j = lane; j < k; j += 32) {
      const CentRec r = DENSE ? tab[j] : *rec_at(tab, j);
      if (score1(pc, r.q) <= bound) {
        const float d = ex::cie94_c(pL, pa, pb, pC, -r.q[1], -r.q[4], -r.q[5], r.q[3]);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | j;
        best = key < best ? key : best;
      }
    }
    best = warp_min_u64(best);
    if ((int)lane == src) {
      const float d = __uint_as_float((unsigned int)(best >> 32));
      result = (best != ~0ull && d < 100000.0f) ? (unsigned int)(best & 0xffffffffull) : 0u;
    }
  }
  return result;
}

// ------------------------------------------------------------------------------------------------
// K1/K3: RGBA8 -> work plane, exact (rgb_to_lab.wgsl:11-80, rgb8u_to_rgb32f.wgsl:4-17).
// Four pixels per thread and step, lane-consecutive within each of the four (so every load
// instruction of a warp reads 128 contiguous bytes and every store instruction writes 512), all four
// loads issued before the first conversion starts and the four conversions independent of each other:
// the kernel was a latency-bound stream with one dependent load -> ~150 instructions -> store per
// thread (0.44 of the HBM peak); what remains is the arithmetic of the exact cube root.
__global__ void __launch_bounds__(256) k_convert(const uint32_t* __restrict__ rgba, unsigned long long n,
                                                 int color_space, const float* __restrict__ lut_g,
                                                 float4* __restrict__ work) {
  __shared__ float lut[256];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  constexpr int U = 4;
  const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x * U;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < n; base += step) {
    uint32_t v[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned long long p = base + (unsigned long long)i * blockDim.x;
      v[i] = p < n ? __ldcs(rgba + p) : 0u;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned long long p = base + (unsigned long long)i * blockDim.x;
      float4 o;
      if (color_space == 0)
        o = ex::lin100_to_lab(lut[v[i] & 255u], lut[(v[i] >> 8) & 255u], lut[(v[i] >> 16) & 255u]);
      else
        o = ex::rgb8_to_rgbf(v[i]);
      if (p < n) __stcs(work + p, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K15: bilinear shrink, exact restatement (see oracle resize_image).  One destination pixel p of
// the dw x dh image sampled from the sw x sh source: resize.wgsl:5-19 with a linear / clamp-to-edge
// sampler at (gx/dw, gy/dh), unorm8 store.  Split into tap selection and blend so that callers can
// keep the loads of several pixels in flight.
struct ResizeTaps {
  size_t i00, i10, i01, i11;  // source pixel indices
  float fx, fy;               // blend weights of the second column / row
};
__device__ __forceinline__ ResizeTaps resize_taps(unsigned int sw, unsigned int sh, unsigned int dw, unsigned int dh,
                                                  unsigned long long p) {
  unsigned int gx = (unsigned int)(p % dw), gy = (unsigned int)(p / dw);
  float py = fsub(fmul(fdiv((float)gy, (float)dh), (float)sh), 0.5f);
  float px = fsub(fmul(fdiv((float)gx, (float)dw), (float)sw), 0.5f);
  float fy0 = floorf(py), fx0 = floorf(px);
  ResizeTaps t;
  t.fy = fsub(py, fy0);
  t.fx = fsub(px, fx0);
  long long y0 = (long long)fy0, x0 = (long long)fx0;
  long long y1 = y0 + 1, x1 = x0 + 1;
  y0 = min(max(y0, 0ll), (long long)sh - 1);
  y1 = min(max(y1, 0ll), (long long)sh - 1);
  x0 = min(max(x0, 0ll), (long long)sw - 1);
  x1 = min(max(x1, 0ll), (long long)sw - 1);
  t.i00 = (size_t)y0 * sw + x0;
  t.i10 = (size_t)y0 * sw + x1;
  t.i01 = (size_t)y1 * sw + x0;
  t.i11 = (size_t)y1 * sw + x1;
  return t;
}
// u8f(v) must return fdiv((float)v, 255.0f) (directly, or from a table of those 256 quotients).
template <class U8F>
__device__ __forceinline__ uint32_t resize_blend(uint32_t p00, uint32_t p10, uint32_t p01, uint32_t p11, float fx, float fy,
                                                 U8F u8f) {
  float wx0 = fsub(1.0f, fx), wy0 = fsub(1.0f, fy);
  uint32_t o = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float c00 = u8f((p00 >> (8 * c)) & 255u);
    float c10 = u8f((p10 >> (8 * c)) & 255u);
    float c01 = u8f((p01 >> (8 * c)) & 255u);
    float c11 = u8f((p11 >> (8 * c)) & 255u);
    float top = fadd(fmul(c00, wx0), fmul(c10, fx));
    float bot = fadd(fmul(c01, wx0), fmul(c11, fx));
    o |= ex::unorm8(fadd(fmul(top, wy0), fmul(bot, fy))) << (8 * c);
  }
  return o;
}
__device__ __forceinline__ uint32_t resize_pixel(const uint32_t* __restrict__ src, unsigned int sw, unsigned int sh,
                                                 unsigned int dw, unsigned int dh, unsigned long long p) {
  const ResizeTaps t = resize_taps(sw, sh, dw, dh, p);
  return resize_blend(__ldg(src + t.i00), __ldg(src + t.i10), __ldg(src + t.i01), __ldg(src + t.i11), t.fx, t.fy,
                      [](uint32_t v) { return fdiv((float)v, 255.0f); });
}

__global__ void __launch_bounds__(256) k_resize(const uint32_t* __restrict__ src, unsigned int sw, unsigned int sh,
                                                uint32_t* __restrict__ dst, unsigned int dw, unsigned int dh) {
  const unsigned long long n = (unsigned long long)dw * dh;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    dst[p] = resize_pixel(src, sw, sh, dw, dh, p);
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

// The mailbox exchange of one init round between the GPUs that share an image: thread 0 holds this
// rank's (key, global pixel or ~0, colour) and gets the winner's.  All threads of the block call it
// (at least 32 of them).  Four 64-bit values per rank: thread t posts value t & 3 to rank t >> 2,
// threads 0-3 gather one value of every rank each.
__device__ __forceinline__ bool init_exchange(const PeerXchg& X, unsigned int k, unsigned int j, unsigned long long& key,
                                              unsigned long long pix, float4& col, unsigned int* s_fault) {
  __shared__ unsigned long long s_mine[4];
  __shared__ unsigned long long s_all[MAX_PEERS][4];
  const unsigned int par = (k - j) & 1u;
  const unsigned int seq = X.seq_base - j;
  const unsigned int t = threadIdx.x;
  if (t == 0) {
    *s_fault = 0;
    s_mine[0] = key;
    s_mine[1] = pix;
    s_mine[2] = ((unsigned long long)__float_as_uint(col.y) << 32) | __float_as_uint(col.x);
    s_mine[3] = __float_as_uint(col.z);
  }
  __syncthreads();
  if (t < 4 * X.n_ranks) xchg_post64(X, par, seq, t >> 2, t & 3u, s_mine[t & 3u]);
  if (t < 4) xchg_gather64(X, par, seq, t, s_fault, [&](unsigned int r, unsigned long long v) { s_all[r][t] = v; });
  __syncthreads();
  if (*s_fault) return false;
  if (t == 0) {
    // largest key over the ranks (keys of different shards never tie: they embed the global pixel
    // index; all-zero maxima resolve to global pixel 0), colour from the rank that holds it
    unsigned long long kmax = 0ull;
    for (unsigned int r = 0; r < X.n_ranks; ++r) kmax = s_all[r][0] > kmax ? s_all[r][0] : kmax;
    const unsigned long long want = key_to_pixel(kmax);
    for (unsigned int r = 0; r < X.n_ranks; ++r) {
      if (s_all[r][1] == want) {
        const unsigned long long la = s_all[r][2];
        col = make_float4(__uint_as_float((unsigned int)la), __uint_as_float((unsigned int)(la >> 32)),
                          __uint_as_float((unsigned int)s_all[r][3]), 1.0f);
      }
    }
    key = kmax;
  }
  return true;
}

// rgba != NULL: the work plane does not exist yet (the first init round will write it, fused with the
// conversion): the seed pixel is converted here.
__global__ void k_init_seed(JobPtrs J, const float4* __restrict__ work, unsigned long long seed_local,
                            int seed_is_local, const uint32_t* __restrict__ rgba = nullptr,
                            const float* __restrict__ lut_g = nullptr, int color_space = 0) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (seed_is_local) {
      float4 v;
      if (rgba) {
        const uint32_t px = rgba[seed_local];
        v = color_space == 0 ? ex::lin100_to_lab(lut_g[px & 255u], lut_g[(px >> 8) & 255u], lut_g[(px >> 16) & 255u])
                             : ex::rgb8_to_rgbf(px);
      } else {
        v = work[seed_local];
      }
      J.cent[0] = make_float4(v.x, v.y, v.z, 1.0f);
    }
    for (unsigned int i = 0; i < J.st->k; ++i) J.keys[i] = 0ull;
    J.st->init_tau16 = 0;
    J.st->init_ncmax = 0;
    J.st->init_done_round = 0;
    J.st->init_attempts = 0;
    J.st->init_refreshed = 0ull;
    J.st->init_folds = 0ull;
    J.st->init_exact = 0ull;
  }
}

// pixel_offset: global index of this shard's first pixel (0 on a single GPU).
// PICK 0: only the arg-max key is produced (k_init_pick resolves it; NCCL path of a sharded image).
// PICK 1: the last block to finish resolves the winner into centroid j itself (single GPU).
// PICK 2: sharded image with peer mailboxes — the last block on every GPU posts its local winner
//   (key, global pixel index, colour) to every rank's mailbox over NVLink, waits for the peers'
//   flags, and every rank takes the colour of the largest key: the arg-max all-reduce and the
//   colour broadcast of the round happen inside the round's own launch (no NCCL, no extra launch).
//   Parity (k - j) & 1 alternates down to the first Lloyd pass (parity 0); flags carry
//   seq_base - j, disjoint from the passes' seq_base + pass + 1.
// ub (may be NULL): the 16-bit upper bounds of the running minima, kept up to date for the lazy
// rounds that follow (kmg_init_lazy.cuh).
// CONVERT (with FIRST): the first consumer of the work plane is this round, so the conversion is
// fused into it (operations.rs:63-83 runs convert -> init -> assign as separate dispatches): RGBA8 in
// (4 B/px), work plane + running minimum out (20 B/px) instead of a convert sweep (20 B/px) followed by a
// round that reads the plane again (16 + 4 B/px).
template <bool FIRST, int PICK, bool CONVERT = false>
__global__ void __launch_bounds__(256) k_init_round(JobPtrs J, const float4* __restrict__ work,
                                                    float* __restrict__ dmin, unsigned long long n,
                                                    unsigned long long pixel_offset, unsigned int j, PeerXchg X,
                                                    unsigned short* __restrict__ ub = nullptr,
                                                    const uint32_t* __restrict__ rgba = nullptr,
                                                    float4* __restrict__ work_out = nullptr,
                                                    const float* __restrict__ lut_g = nullptr, int color_space = 0) {
  static_assert(!CONVERT || FIRST, "the conversion is fused into the first round only");
  __shared__ unsigned long long s_key[8];
  __shared__ bool s_last;
  __shared__ float lut[CONVERT ? 256 : 1];
  if (CONVERT) {
    lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
  }
  // a peer is gone (an earlier exchange of this job gave up): every later round would wait for it again
  if (PICK == 2 && J.st->conv == PASS_FAULT) return;
  // centroid j-1 was resolved into J.cent[j-1] by the previous round (or k_init_seed for j == 1).
  const float4 c = J.cent[j - 1];
  const float cc = ex::chroma(c.y, c.z);
  unsigned long long best = 0ull;
  // four pixels per thread and step, all eight loads in flight before the first distance is
  // computed: the round is a latency-bound stream otherwise (one dependent load -> ~100
  // instructions -> store per pixel)
  constexpr int U = 4;
  const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x * U;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < n; base += step) {
    float4 v[U];
    float old[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned long long p = base + (unsigned long long)i * blockDim.x;
      const bool ok = p < n;
      if (CONVERT) {
        const uint32_t px = ok ? __ldcs(rgba + p) : 0u;
        v[i] = color_space == 0 ? ex::lin100_to_lab(lut[px & 255u], lut[(px >> 8) & 255u], lut[(px >> 16) & 255u])
                                : ex::rgb8_to_rgbf(px);
        if (ok) work_out[p] = v[i];
      } else {
        v[i] = ok ? ldg_stream(work + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      old[i] = (!FIRST && ok) ? dmin[p] : 1000000.0f;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned long long p = base + (unsigned long long)i * blockDim.x;
      if (p < n) {
        const float d = ex::cie94_c(v[i].x, v[i].y, v[i].z, v[i].w, c.x, c.y, c.z, cc);
        // kmeans++_calc_diff.wgsl:27-31 (the first round starts from 1000000.0).  The plane is only
        // written where the minimum decreases: in round j about 1/j of the pixels, so most 32-byte
        // sectors of the plane stay clean in HBM.
        const float dm = fminf(old[i], d);
        if (FIRST || d < old[i]) dmin[p] = dm;
        if (ub && (FIRST || d < old[i])) ub[p] = (unsigned short)((__float_as_uint(dm) + 0xffffu) >> 16);
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(dm) << 32) | (((pixel_offset + p) & 0xffffffffull) ^ 15ull);
        best = key > best ? key : best;
      }
    }
  }
  best = warp_max_u64(best);
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = s_key[w] > best ? s_key[w] : best;
    atomicMax(J.keys + j, best);
  }
  if (PICK == 0) return;
  // the last block resolves the round
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&J.st->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __shared__ unsigned int s_fault;
  if (threadIdx.x == 0) s_fault = 0;
  __syncthreads();
  unsigned long long key = 0ull, pix = 0ull;
  float4 col = make_float4(0.f, 0.f, 0.f, 1.0f);
  if (threadIdx.x == 0) {
    __threadfence();
    key = __ldcg(J.keys + j);
    pix = key_to_pixel(key);
    if (pix >= pixel_offset && pix - pixel_offset < n) {
      const float4 v = CONVERT ? __ldcg(work_out + (pix - pixel_offset)) : work[pix - pixel_offset];
      col = make_float4(v.x, v.y, v.z, 1.0f);
    } else {
      pix = ~0ull;  // zero maximum on a shard that does not hold pixel 0: no candidate
    }
  }
  if (PICK == 2) {
    if (!init_exchange(X, J.st->k, j, key, pix, col, &s_fault) && threadIdx.x == 0) {
      J.st->conv = PASS_FAULT;
      J.st->done = 1;
    }
  }
  if (threadIdx.x == 0) {
    J.cent[j] = col;
    J.keys[j] = key;
    J.st->ticket = 0;
    if (ub) {  // threshold of the first lazy round: farthest-point distances never grow
      J.st->init_tau16 = __float_as_uint(__uint_as_float((unsigned int)(key >> 32)) * 0.98f) >> 16;
      J.st->init_done_round = j;
    }
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
// Sums are exact integers, rint(v * 2^15) accumulated in int32 thread-private shared-memory slots
// (PRIVATE: conflict-free 128-bit read-modify-write, flushed before they can overflow) or sent
// straight to L2 with 64-bit reductions into a per-block copy (large k).  Integer addition
// commutes, so the result is independent of block scheduling, grid size and of how many GPUs share
// the image.  The last block to finish turns the sums into the new centroids, convergence flags
// and the next table, so a pass is exactly one launch and needs no host round trip.

// mode 0: single GPU.  mode 1: leave the folded partial sums in accumulator copy 0 for an external
// all-reduce (k_finalize completes the pass).  mode 2: exchange them with the peers right here.
template <int THREADS>
__device__ void finalize_pass(const JobPtrs& J, int color_space, int mode, const PeerXchg& X) {
  JobState* st = J.st;
  const unsigned int k = st->k;
  const unsigned int tid = threadIdx.x;
  __shared__ unsigned int s_conv;
  __shared__ unsigned int s_fault;
  if (tid == 0) {
    s_conv = 0;
    s_fault = 0;
  }
  __syncthreads();
  const unsigned int par = st->passes & 1u;
  const unsigned int seq = X.seq_base + st->passes + 1u;
  if (mode == 2) {
    // fold the local copies and post them to every rank (own mailbox included)
    for (unsigned int c = tid; c < k; c += THREADS) {
      long long s[4] = {0, 0, 0, 0};
      for (unsigned int copy = 0; copy < J.acc_copies; ++copy) {
        long long* a = J.acc + ((size_t)copy * k + c) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          s[q] += __ldcg(a + q);
          a[q] = 0;
        }
      }
      for (unsigned int r = 0; r < X.n_ranks; ++r) {
#pragma unroll
        for (int q = 0; q < 4; ++q) xchg_post64(X, par, seq, r, (size_t)c * 4 + q, (unsigned long long)s[q]);
      }
    }
    // gather: one thread per (centroid, component), the ranks' values added in rank order on every GPU
    for (unsigned int i = tid; i < 4 * k; i += THREADS) {
      long long sum = 0;
      xchg_gather64(X, par, seq, i, &s_fault, [&](unsigned int, unsigned long long v) { sum += (long long)v; });
      J.last[i] = sum;
    }
    __syncthreads();
    if (s_fault) {
      if (tid == 0) {
        st->conv = PASS_FAULT;
        st->done = 1;
        st->ticket = 0;
      }
      return;
    }
  }
  unsigned int conv = 0;
  for (unsigned int c = tid; c < k; c += THREADS) {
    long long s[4] = {0, 0, 0, 0};
    if (mode == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] = __ldcg(J.last + (size_t)c * 4 + q);
    } else {
      for (unsigned int copy = 0; copy < J.acc_copies; ++copy) {
        long long* a = J.acc + ((size_t)copy * k + c) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          s[q] += __ldcg(a + q);
          if (mode == 0 || copy > 0) a[q] = 0;
        }
      }
    }
    if (mode == 1) {
      // leave the folded partial in copy 0 for the all-reduce; k_finalize completes it
      long long* a0 = J.acc + (size_t)c * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) a0[q] = s[q];
      continue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) J.last[(size_t)c * 4 + q] = s[q];
    if (s[3] > 0) {  // choose_centroid.wgsl:185-194
      const double cnt = (double)s[3];
      float4 prev = J.cent[c];
      float4 nc;
      nc.x = (float)(((double)s[0] / cnt) * ex::FIXED_UNIT);
      nc.y = (float)(((double)s[1] / cnt) * ex::FIXED_UNIT);
      nc.z = (float)(((double)s[2] / cnt) * ex::FIXED_UNIT);
      nc.w = 1.0f;
      J.cent[c] = nc;
      conv += (ex::cie94(nc.x, nc.y, nc.z, prev.x, prev.y, prev.z) < st->conv_threshold) ? 1u : 0u;
    }
  }
  if (mode == 1) {
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
  PeerXchg none;
  none.n_ranks = 0;
  none.rank = 0;
  none.xcap = 0;
  none.seq_base = 0;
  finalize_pass<256>(J, color_space, 0, none);
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

// One tile of THREADS x P pixels.  KT > 0: compile-time table length (saved-score search);
// KT == 0: runtime length kp (chunked search).  PRIVATE: thread-private int4 slots in s_acc.
// CT: s_tab is the job's slot in the constant bank (argmin) and x_tab the dense table in global
// memory for the rare exact path; otherwise both are the shared-memory copy.
// ATOM (with PRIVATE): the private slots are laid out [component][cluster][thread] and updated with
// four fire-and-forget shared-memory atomic adds per pixel instead of a 128-bit read-modify-write,
// whose load -> add -> store chains serialise when two pixels of a thread share a cluster (the
// compiler has to assume they do); cstride = distance between components in ints.
template <int KT, int THREADS, int P, bool PRIVATE, bool CHECK, bool CT = false, bool ATOM = false>
__device__ __forceinline__ void lloyd_tile(const CentRec* __restrict__ s_tab, const CentRec* __restrict__ x_tab,
                                           unsigned int kp, int4* __restrict__ s_acc, unsigned int cstride,
                                           unsigned long long* __restrict__ g_acc, const float4 (&v)[P],
                                           unsigned long long base, unsigned long long n, unsigned int k,
                                           float lmax, float cmax, unsigned int tid, unsigned int& slow) {
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
  float eps[P], thr[P];
  unsigned int idx[P];
  bool certified[P];
  if (KT > 0)
    argmin_small<P, (KT > 0 ? KT : 8), false, CT>(s_tab, px, lmax, cmax, 0.0f, eps, idx, certified, thr);
  else
    argmin_chunked<P, false, (CT && KT == 0)>(s_tab, kp, px, lmax, cmax, 0.0f, eps, idx, certified, thr);
  // one vote per tile: the exact path is rare (1e-4 .. 1e-2 of the pixels)
  bool need[P], any_need = false;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    need[i] = !certified[i] && valid[i];
    any_need |= need[i];
  }
  if (__any_sync(0xffffffffu, any_need)) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (__any_sync(0xffffffffu, need[i])) {
        idx[i] = warp_exact_argmin<(CT && KT > 0), true>(x_tab, k, need[i], px.L[i], px.a[i], px.b[i], px.C[i], thr[i],
                                                         idx[i]);
        slow += need[i] ? 1u : 0u;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (valid[i]) {
      if (!PRIVATE && ATOM) {
        // block accumulators [7][kp]: every fixed-point value goes in as its low 12 bits (unsigned)
        // and its high part (signed), so 32-bit sums cannot overflow for 2^19 pixels per block and
        // seven fire-and-forget shared-memory atomics replace four 64-bit reductions through L2
        int* a = reinterpret_cast<int*>(s_acc) + idx[i];
        const int f0 = ex::to_fixed(px.L[i]), f1 = ex::to_fixed(px.a[i]), f2 = ex::to_fixed(px.b[i]);
        atomicAdd(a, f0 & 0xfff);
        atomicAdd(a + cstride, f0 >> 12);
        atomicAdd(a + 2 * cstride, f1 & 0xfff);
        atomicAdd(a + 3 * cstride, f1 >> 12);
        atomicAdd(a + 4 * cstride, f2 & 0xfff);
        atomicAdd(a + 5 * cstride, f2 >> 12);
        atomicAdd(a + 6 * cstride, 1);
      } else if (PRIVATE && ATOM) {
        int* slot = reinterpret_cast<int*>(s_acc) + idx[i] * THREADS + tid;
        atomicAdd(slot, ex::to_fixed(px.L[i]));
        atomicAdd(slot + cstride, ex::to_fixed(px.a[i]));
        atomicAdd(slot + 2 * cstride, ex::to_fixed(px.b[i]));
        atomicAdd(slot + 3 * cstride, 1);
      } else if (PRIVATE) {
        int4* slot = s_acc + idx[i] * THREADS + tid;
        int4 a = *slot;
        a.x += ex::to_fixed(px.L[i]);
        a.y += ex::to_fixed(px.a[i]);
        a.z += ex::to_fixed(px.b[i]);
        a.w += 1;
        *slot = a;
      } else {
        unsigned long long* a = g_acc + (size_t)idx[i] * 4;
        atomicAdd(a + 0, (unsigned long long)(long long)ex::to_fixed(px.L[i]));
        atomicAdd(a + 1, (unsigned long long)(long long)ex::to_fixed(px.a[i]));
        atomicAdd(a + 2, (unsigned long long)(long long)ex::to_fixed(px.b[i]));
        atomicAdd(a + 3, 1ull);
      }
    }
  }
}

// KT > 0: table of KT entries in static shared memory.  KT == 0: table of pad32(k) entries at the
// start of dynamic shared memory (followed, if PRIVATE, by the accumulator slots for KCAP clusters).
// CT (KT > 0, PRIVATE): the table is read from slot `cslot` of the constant bank instead.
template <int KT, int KCAP, int THREADS, int P, bool PRIVATE, int MINB, bool CT = false, bool ATOM = false>
__global__ void __launch_bounds__(THREADS, MINB) k_lloyd(JobPtrs J, const float4* __restrict__ work,
                                                         unsigned long long n, int color_space,
                                                         int distributed_mode, PeerXchg X, int cslot,
                                                         unsigned int k_arg) {
  static_assert(!CT || KT == 0 || (KT * 6 <= CTAB_FLOATS && PRIVATE), "constant-bank tables: small compile-time k, or c_big");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) unsigned char s_tab_static[(KT > 0 && !CT) ? (KT / 8) * CHUNK_BYTES : 16];
  __shared__ bool s_last;
  JobState* st = J.st;
  if (st->done) return;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = k_arg;  // == st->k; a kernel parameter is provably warp-uniform (loop counters over
                                 // the table then live in uniform registers and can index the constant bank)
  const unsigned int kp = KT > 0 ? (unsigned int)KT : pad32(k);
  // CT: no shared-memory table at all
  constexpr bool CT_SMALL = CT && KT > 0;
  const CentRec* s_tab = CT_SMALL ? reinterpret_cast<const CentRec*>(c_tab[cslot])
                                  : reinterpret_cast<const CentRec*>(KT > 0 ? s_tab_static : smem_raw);
  const CentRec* x_tab = CT_SMALL ? J.tab : s_tab;
  // PRIVATE: [KCAP][THREADS] slots after the (compile-time sized) table; !PRIVATE && ATOM: block
  // accumulators [7][kp] ints after the runtime-sized table
  constexpr bool BLOCK_ACC = !PRIVATE && ATOM;
  int4* s_acc = reinterpret_cast<int4*>(smem_raw + (KT > 0 ? 0 : tab_smem_bytes(BLOCK_ACC ? kp : pad32(KCAP))));
  if (!CT_SMALL) tab_to_smem(const_cast<CentRec*>(s_tab), J.tab, kp, tid, THREADS);
  const unsigned int CSTRIDE = BLOCK_ACC ? kp : (unsigned int)(KCAP > 0 ? KCAP : 1) * THREADS;
  if (PRIVATE) {
#pragma unroll 4
    for (int c = 0; c < KCAP; ++c) s_acc[c * THREADS + tid] = make_int4(0, 0, 0, 0);
  }
  if (BLOCK_ACC)
    for (unsigned int c = tid; c < 7 * kp; c += THREADS) reinterpret_cast<int*>(s_acc)[c] = 0;
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();

  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long full_tiles = n / TILE;
  unsigned long long* g_acc = reinterpret_cast<unsigned long long*>(J.acc + (size_t)(blockIdx.x % J.acc_copies) * k * 4);
  unsigned int since_flush = 0;
  unsigned int slow = 0;
  // block accumulators without a constant-bank table: `cslot` carries log2 of the drain interval in
  // pixels per block (19 in production: 4095 * 2^19 < 2^31; tests lower it to exercise the drain)
  const unsigned int block_flush_px = (BLOCK_ACC && !CT) ? (1u << cslot) : (1u << 19);

  // block accumulators: all threads of the block call this together
  auto flush_block = [&]() {
    if (!BLOCK_ACC) return;
    __syncthreads();
    int* a = reinterpret_cast<int*>(s_acc);
    for (unsigned int c = tid; c < k; c += THREADS) {
      const int cnt = a[c + 6 * CSTRIDE];
      if (cnt != 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const long long v = (long long)a[c + (2 * q + 1) * CSTRIDE] * 4096 + (long long)(unsigned int)a[c + 2 * q * CSTRIDE];
          atomicAdd(g_acc + c * 4 + q, (unsigned long long)v);
        }
        atomicAdd(g_acc + c * 4 + 3, (unsigned long long)(unsigned int)cnt);
      }
#pragma unroll
      for (int q = 0; q < 7; ++q) a[c + q * CSTRIDE] = 0;
    }
    __syncthreads();
    since_flush = 0;
  };
  auto flush = [&]() {
    if (!PRIVATE) return;
    // Every warp folds the private slots of its own 32 threads for all clusters (no block sync
    // needed: a warp only reads what it wrote) and sends 4 reductions per cluster to L2.
    const unsigned int lane = tid & 31;
    for (unsigned int c = 0; c < k; ++c) {
      int4 v;
      if (ATOM) {
        int* slot = reinterpret_cast<int*>(s_acc) + c * THREADS + tid;
        v = make_int4(slot[0], slot[CSTRIDE], slot[2 * CSTRIDE], slot[3 * CSTRIDE]);
        slot[0] = 0;
        slot[CSTRIDE] = 0;
        slot[2 * CSTRIDE] = 0;
        slot[3 * CSTRIDE] = 0;
      } else {
        v = s_acc[c * THREADS + tid];
        s_acc[c * THREADS + tid] = make_int4(0, 0, 0, 0);
      }
      const long long s3 = (long long)__reduce_add_sync(0xffffffffu, v.w);  // <= 32 x 240
      if (s3 == 0) continue;                                                  // warp-uniform
      const long long s0 = warp_sum_i32(v.x), s1 = warp_sum_i32(v.y), s2 = warp_sum_i32(v.z);
      if (lane == 0) {
        atomicAdd(g_acc + c * 4 + 0, (unsigned long long)s0);
        atomicAdd(g_acc + c * 4 + 1, (unsigned long long)s1);
        atomicAdd(g_acc + c * 4 + 2, (unsigned long long)s2);
        atomicAdd(g_acc + c * 4 + 3, (unsigned long long)s3);
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
      lloyd_tile<KT, THREADS, P, PRIVATE, false, CT, ATOM>(s_tab, x_tab, kp, s_acc, CSTRIDE, g_acc, cur,
                                                           tile * TILE + tid, n, k, lmax, cmax, tid, slow);
      since_flush += P;
      // |v| < 2^7 colour units -> |fixed| < 2^22; 240 pixels stay far below 2^31.
      if (PRIVATE && since_flush + P > 240) flush();
      // block accumulators: 2^19 pixels of the block (since_flush counts pixels per thread)
      if (BLOCK_ACC && (since_flush + P) * THREADS > block_flush_px) flush_block();
#pragma unroll
      for (int i = 0; i < P; ++i) cur[i] = nxt[i];
    }
  }
  // ragged tail (< TILE pixels), taken by the block whose turn it would be
  if (full_tiles * TILE < n && blockIdx.x == (unsigned int)(full_tiles % gridDim.x)) {
    if (PRIVATE && since_flush + P > 240) flush();
    float4 tail[P];
    lloyd_load<THREADS, P, true>(work, full_tiles * TILE + tid, n, tail);
    lloyd_tile<KT, THREADS, P, PRIVATE, true, CT, ATOM>(s_tab, x_tab, kp, s_acc, CSTRIDE, g_acc, tail,
                                                        full_tiles * TILE + tid, n, k, lmax, cmax, tid, slow);
  }
  flush();
  flush_block();
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);

  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    finalize_pass<THREADS>(J, color_space, distributed_mode, X);
  }
}

// ------------------------------------------------------------------------------------------------
// K5 alone: labels for a work plane (stage-level parity test hook).
template <int THREADS, int P>
__global__ void __launch_bounds__(THREADS) k_assign(JobPtrs J, const float4* __restrict__ work,
                                                    unsigned long long n, uint32_t* __restrict__ labels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  JobState* st = J.st;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  const unsigned int kp = pad32(k);
  tab_to_smem(s_tab, J.tab, kp, tid, THREADS);
  const float lmax = st->lmax, cmax = st->cmax;
  __syncthreads();
  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long tiles = (n + TILE - 1) / TILE;
  unsigned int slow = 0;
  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned long long base = tile * TILE + tid;
    float4 v[P];
    lloyd_load<THREADS, P, true>(work, base, n, v);
    Pix<P> px;
  #pragma unroll
    for (int i = 0; i < P; ++i) {
      px.L[i] = v[i].x;
      px.a[i] = v[i].y;
      px.b[i] = v[i].z;
      px.C[i] = v[i].w;
      }
    float eps[P];
    unsigned int idx[P];
    bool certified[P];
    argmin_chunked<P, false>(s_tab, kp, px, lmax, cmax, 0.0f, eps, idx, certified);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const bool valid = base + (unsigned long long)i * THREADS < n;
      const bool need = !certified[i] && valid;
      if (__any_sync(0xffffffffu, need)) {
        idx[i] = warp_exact_argmin(s_tab, k, need, px.L[i], px.a[i], px.b[i], px.C[i], eps[i], idx[i]);
        slow += need ? 1u : 0u;
      }
      if (valid) labels[base + (unsigned long long)i * THREADS] = idx[i];
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
constexpr unsigned long long bayer_nibbles() {
  const unsigned long long m[16] = {0, 8, 2, 10, 12, 4, 14, 6, 3, 11, 1, 9, 15, 7, 13, 5};
  unsigned long long v = 0;
  for (int j = 0; j < 16; ++j) v |= m[j] << (4 * j);
  return v;
}
constexpr unsigned long long BAYER_NIBBLES = bayer_nibbles();

// Exact pixel of the remap kernels: Lab through the FP64 pow path (+ the dither offset).
template <int MODE>
__device__ __noinline__ float4 remap_exact_pixel(uint32_t v, const float* __restrict__ lut, int color_space, float off) {
  float4 e = color_space == 0 ? ex::lin100_to_lab(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u])
                              : ex::rgb8_to_rgbf(v);
  if (MODE == 1) {  // mix_colors.wgsl:70-72
    e.x = fadd(e.x, off);
    e.y = fadd(e.y, off);
    e.z = fadd(e.z, off);
    e.w = ex::chroma(e.y, e.z);
  }
  return e;
}

// The approximate pixel of the remap kernels: fast Lab (or v / 255), plus the ordered-dither offset
// of cell (x % 4, y % 4) in MODE 1, and the approximate chroma the score takes.
template <int MODE>
__device__ __forceinline__ void remap_fast_pixel(uint32_t v, const float* __restrict__ lut, int color_space, float thr,
                                                 unsigned int xi, unsigned int yi, float& L, float& a, float& b, float& C,
                                                 float& off) {
  if (color_space == 0) {
    float3 lab = fast::lin100_to_lab(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u]);
    L = lab.x;
    a = lab.y;
    b = lab.z;
  } else {
    L = (float)(v & 255u) * (1.0f / 255.0f);
    a = (float)((v >> 8) & 255u) * (1.0f / 255.0f);
    b = (float)((v >> 16) & 255u) * (1.0f / 255.0f);
  }
  off = 0.0f;
  if (MODE == 1) {
    // index_matrix[x % 4 + 4 * (y % 4)] / 16 - 0.5 (mix_colors.wgsl:14-17,21-27,70); the sixteen
    // 4-bit entries sit in one 64-bit constant (no divergent constant-bank load)
    const unsigned int cell = (xi & 3u) + ((yi & 3u) << 2);
    float iv = (float)((unsigned int)(BAYER_NIBBLES >> (4u * cell)) & 15u) * 0.0625f - 0.5f;
    off = fmul(thr, iv);
    L += off;
    a += off;
    b += off;
  }
  C = fast::sqrt_approx(fmaf(a, a, b * b));
}

template <int MODE, int KT, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_remap(JobPtrs J0, const uint32_t* __restrict__ rgba, unsigned int w,
                                                   unsigned long long n, int color_space,
                                                   const float* __restrict__ lut_g, uint32_t* __restrict__ out,
                                                   size_t blob_stride) {
  // KT > 0: compile-time table length; KT == 0: runtime length.  Table + palette in dynamic smem.
  // blockIdx.y = frame of a batch (frames of n pixels back to back, job blobs blob_stride apart).
  const JobPtrs J = job_at(J0, (size_t)blockIdx.y * blob_stride);
  rgba += (size_t)blockIdx.y * n;
  out += (size_t)blockIdx.y * n;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CentRec* s_tab = reinterpret_cast<CentRec*>(smem_raw);
  __shared__ float lut[256];
  JobState* st = J.st;
  const unsigned int tid = threadIdx.x;
  const unsigned int k = st->k;
  const unsigned int kp = KT > 0 ? (unsigned int)KT : pad32(k);
  uint32_t* s_pal = reinterpret_cast<uint32_t*>(smem_raw + tab_smem_bytes(kp));
  tab_to_smem(s_tab, J.tab, kp, tid, THREADS);
  for (unsigned int c = tid; c < kp; c += THREADS) s_pal[c] = c < k ? J.pal[c] : 0u;
  for (unsigned int c = tid; c < 256; c += THREADS) lut[c] = lut_g[c];
  const float lmax = st->lmax, cmax = st->cmax;
  const float thr = st->dither_threshold;
  const unsigned long long w_magic = ~0ull / w + 1ull;
  __syncthreads();

  constexpr int P = 4;
  const unsigned long long groups = (n + P - 1) / P;
  const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
  unsigned int slow = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(rgba) & 15) == 0;
  auto load_group = [&](unsigned long long g, uint32_t (&v)[P]) {
    const unsigned long long p0 = g * P;
    if (g < groups && p0 + P <= n && aligned) {
      uint4 t = __ldcs(reinterpret_cast<const uint4*>(rgba) + g);
      v[0] = t.x;
      v[1] = t.y;
      v[2] = t.z;
      v[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = (g < groups && p0 + i < n) ? rgba[p0 + i] : 0u;
    }
  };
  // warp-uniform trip count (the exact path is warp-cooperative); lanes past the end idle.
  // The next group's pixels are loaded before the current group is processed.
  uint32_t v[P], vn[P];
  load_group((unsigned long long)blockIdx.x * THREADS + tid, v);
  for (unsigned long long g0 = (unsigned long long)blockIdx.x * THREADS; g0 < groups; g0 += stride) {
    const unsigned long long g = g0 + tid;
    const unsigned long long p0 = g * P;
    const bool live = g < groups;
    const bool full = live && p0 + P <= n;
    load_group(g + stride, vn);
    Pix<P> px;
    float off[P];
    unsigned int x = 0, y = 0;
    if (MODE == 1) {  // p0 / w by multiplication with ceil(2^64 / w): exact for p0 < 2^33
      y = w == 1 ? (unsigned int)p0 : (unsigned int)__umul64hi(p0, w_magic);
      x = (unsigned int)(p0 - (unsigned long long)y * w);
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      unsigned int xi = x + i, yi = y;
      if (MODE == 1) {
        while (xi >= w) {  // group straddles a row end (w % 4 != 0)
          xi -= w;
          ++yi;
        }
      }
      remap_fast_pixel<MODE>(v[i], lut, color_space, thr, xi, yi, px.L[i], px.a[i], px.b[i], px.C[i], off[i]);
    }
    const float conv_k = color_space == 0 ? fast::LAB_ERR : fast::RGB_ERR;  // |approximate pixel - exact pixel|
    float eps[P];
    unsigned int idx[P];
    bool certified[P];
    if (KT > 0)
      argmin_small<P, (KT > 0 ? KT : 8), true>(s_tab, px, lmax, cmax, conv_k, eps, idx, certified);
    else
      argmin_chunked<P, true>(s_tab, kp, px, lmax, cmax, conv_k, eps, idx, certified);
    uint32_t o[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const bool need = !certified[i] && live && p0 + i < n && k > 1;
      if (__any_sync(0xffffffffu, need)) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need) e = remap_exact_pixel<MODE>(v[i], lut, color_space, off[i]);
        // the exact pixel is used here, so only the score-rounding part of the bound is needed
        idx[i] = warp_exact_argmin(s_tab, k, need, e.x, e.y, e.z, e.w, fast::score_eps(e.x, e.w, lmax, cmax), idx[i]);
        slow += need ? 1u : 0u;
      }
      o[i] = s_pal[k > 1 ? idx[i] : 0u];
    }
    if (full && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i)
        if (live && p0 + i < n) out[p0 + i] = o[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = vn[i];
  }
  if (slow) atomicAdd(&st->slow_pixels, (unsigned long long)slow);
}

// Meld (K14, mix_colors.wgsl:29-48,85-90,115-136): continuous output, evaluated exactly per pixel.
__global__ void __launch_bounds__(256) k_remap_meld(JobPtrs J0, const uint32_t* __restrict__ rgba,
                                                    unsigned long long n, int color_space,
                                                    const float* __restrict__ lut_g, uint32_t* __restrict__ out,
                                                    size_t blob_stride) {
  const JobPtrs J = job_at(J0, (size_t)blockIdx.y * blob_stride);
  rgba += (size_t)blockIdx.y * n;
  out += (size_t)blockIdx.y * n;
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
// FP32 (non-tensor) issue peak of the device, for the FP32 side of the roofline: 16 independent
// FFMA chains per thread with a constant-bank multiplier (the form that issues every cycle: two
// register reads in different banks), 8 resident blocks per SM.
__constant__ float c_peak_mul[4] = {1.0000001f, 0.9999999f, 1.0000002f, 0.9999998f};
__global__ void __launch_bounds__(256) k_fp32_peak(float* __restrict__ out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + (float)(threadIdx.x + i) * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], c_peak_mul[r], 1e-7f);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

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
