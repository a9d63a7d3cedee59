// kmg_small.cuh — the whole k-means of one small image in ONE launch, one thread-block cluster per
// image: shrink (K15) -> convert (K1/K3) -> farthest-point init (K8-K11) -> Lloyd loop with the
// reference stop rule (K5 + K6/K7, core/src/modules.rs:763-840) -> table / palette for the remap.
//
// The reference never clusters more than 256 x 256 pixels (core/src/structures.rs:23,67-74), and
// spends ~300 dispatches and >= 20 blocking host syncs on them.  Here the
// shrunk image lives in the distributed shared memory of a cluster of 8 or 16 CTAs (16 B/px work
// plane + 4 B/px running min distance), so after the one read of the source image nothing touches
// HBM until the final centroids are written:
//   * init round: every CTA scans its slice, the per-CTA arg-max keys are exchanged with remote
//     shared-memory stores + one cluster barrier, and the winning pixel's colour is read straight
//     out of its owner's shared memory;
//   * Lloyd pass: certified nearest-centroid search (same code as k_lloyd) into thread-private
//     integer accumulators, a block fold, an all-to-all of the k x 4 int64 partial sums through
//     DSMEM, one cluster barrier, and a redundant, fixed-order finalisation in every CTA — so all
//     CTAs hold identical centroids and take the same stop decision without further traffic.
// A batch of frames is one launch with one cluster per frame (BASELINE config 5).
//
// Results are bit-identical to the multi-launch path (k_resize, k_convert, k_init_round, k_lloyd):
// every value that is stored or compared is computed by the same ex:: arithmetic, and the centroid
// sums are integer.
#pragma once
#include <cooperative_groups.h>

#include "kmg_kernels.cuh"

namespace kmg {
namespace cg = cooperative_groups;

constexpr unsigned int SMALL_MAX_CLUSTER = 16;

// Development aid (make TRACE=1): warp 0 / lane 0 of every CTA of frame 0 records clock64() at the
// phase boundaries.  Not compiled into the product library.
#ifdef KMG_TRACE
__device__ unsigned long long g_small_trace[SMALL_MAX_CLUSTER][512];
#define KMG_TRACE_DECL unsigned int trace_n = 0
#define KMG_TRACE_MARK()                                                              \
  do {                                                                                \
    if (tid == 0 && frame == 0 && trace_n < 512) g_small_trace[rank][trace_n++] = clock64(); \
  } while (0)
#else
#define KMG_TRACE_DECL
#define KMG_TRACE_MARK() do {} while (0)
#endif

struct SmallParams {
  const uint32_t* src;           // frame 0, full size
  unsigned long long frame_px;   // pixels per source frame (stride between frames)
  unsigned int sw, sh;           // source size
  unsigned int dw, dh;           // clustered size (== source size when no shrink is needed)
  int shrink;
  unsigned int ppc;              // pixels per CTA (multiple of 4), ppc * cluster size >= dw * dh
  unsigned int seed;             // seed pixel index on the clustered image
  unsigned int k;
  unsigned int max_iter, check_every;
  float conv_threshold;
  int color_space;
  int want_palette;
  size_t blob_stride;            // bytes between the job blobs of consecutive frames
  const float* lut;
};

// exchange slots per rank: k x 4 sums + {exact-path pixel count, pad}
__host__ __device__ inline unsigned int small_xslots(unsigned int kcap) { return kcap * 4 + 2; }
__host__ __device__ inline size_t small_smem_bytes(unsigned int ppc, unsigned int kcap, unsigned int threads,
                                                   unsigned int csize) {
  return (size_t)ppc * 20 + (size_t)kcap * threads * 16 + (size_t)2 * csize * small_xslots(kcap) * 8;
}

// KCAP: table / accumulator capacity (8, 16: saved-score search; 32: chunked search).
template <int KCAP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_kmeans_small(SmallParams prm, JobPtrs J0) {
  static_assert(KCAP == 8 || KCAP == 16 || KCAP == 32, "table capacity");
  static_assert(THREADS % KCAP == 0, "fold groups");
  constexpr int P = KCAP == 8 ? 4 : 2;
  constexpr int G = THREADS / KCAP < 32 ? THREADS / KCAP : 32;  // lanes that fold one cluster's slots
  constexpr int NW = THREADS / 32;
  constexpr int TAB_BYTES = (KCAP / 8) * CHUNK_BYTES;

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int csize = cluster.num_blocks();
  const unsigned int rank = cluster.block_rank();
  const unsigned int frame = blockIdx.x / csize;
  const unsigned int tid = threadIdx.x;
  const unsigned int lane = tid & 31u, warp = tid >> 5;
  const unsigned int k = prm.k;
  const unsigned int ppc = prm.ppc;
  const unsigned int XS = small_xslots(KCAP);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_work = reinterpret_cast<float4*>(smem_raw);
  float* s_dmin = reinterpret_cast<float*>(smem_raw + (size_t)ppc * 16);
  int4* s_acc = reinterpret_cast<int4*>(smem_raw + (size_t)ppc * 20);
  long long* s_x = reinterpret_cast<long long*>(smem_raw + (size_t)ppc * 20 + (size_t)KCAP * THREADS * 16);
  // every warp keeps its own copy of the centroids and of the search table, so the finalisation
  // of a pass needs no block-wide barrier (all warps compute identical values)
  __shared__ __align__(16) unsigned char s_tab_raw[NW][TAB_BYTES];
  __shared__ float4 s_cent[NW][KCAP];
  __shared__ float s_lut[256];
  __shared__ unsigned long long s_keys[2][SMALL_MAX_CLUSTER];
  __shared__ unsigned long long s_red[NW];
  __shared__ unsigned int s_slow;
  CentRec* s_tab = reinterpret_cast<CentRec*>(s_tab_raw[warp]);
  float4* my_cent = s_cent[warp];

  const JobPtrs J = job_at(J0, (size_t)frame * prm.blob_stride);
  const uint32_t* src = prm.src + (size_t)frame * prm.frame_px;
  const unsigned int N = prm.dw * prm.dh;
  const unsigned int first = rank * ppc;
  const unsigned int n_local = first < N ? min(ppc, N - first) : 0u;

  KMG_TRACE_DECL;
  KMG_TRACE_MARK();  // 0: start
  // ---- shrink + convert into the local slice of the work plane --------------------------------
  for (unsigned int c = tid; c < 256; c += THREADS) s_lut[c] = prm.lut[c];
  if (tid == 0) s_slow = 0;
  {
    // pass 1: RGBA8 of the clustered image (the loads of several pixels in flight), parked in the
    // distance plane; pass 2: exact conversion (FP64 pow) out of shared memory
    uint32_t* s_px = reinterpret_cast<uint32_t*>(s_dmin);
#pragma unroll 4
    for (unsigned int i = tid; i < n_local; i += THREADS)
      s_px[i] = prm.shrink ? resize_pixel(src, prm.sw, prm.sh, prm.dw, prm.dh, first + i) : __ldg(src + first + i);
    __syncthreads();
    KMG_TRACE_MARK();  // 1: resized
    for (unsigned int i = tid; i < n_local; i += THREADS) {
      const uint32_t v = s_px[i];
      s_work[i] = prm.color_space == 0
                      ? ex::lin100_to_lab(s_lut[v & 255u], s_lut[(v >> 8) & 255u], s_lut[(v >> 16) & 255u])
                      : ex::rgb8_to_rgbf(v);
    }
  }
#pragma unroll 4
  for (int q = 0; q < KCAP; ++q) s_acc[q * THREADS + tid] = make_int4(0, 0, 0, 0);
  KMG_TRACE_MARK();  // 2: converted
  cluster.sync();
  KMG_TRACE_MARK();  // 3: cluster sync

  // ---- farthest-point init (plus_plus_init.wgsl, kmeans++_calc_diff.wgsl) ----------------------
  auto pixel_colour = [&](unsigned int g) -> float4 {  // any pixel of the image, through DSMEM
    const float4* owner = cluster.map_shared_rank(s_work, g / ppc);
    float4 v = owner[g % ppc];
    v.w = 1.0f;
    return v;
  };
  float4 c = pixel_colour(prm.seed);
  if (lane == 0) my_cent[0] = c;
  if (tid == 0 && rank == 0) J.keys[0] = 0ull;
  for (unsigned int j = 1; j < k; ++j) {
    const float cc = ex::chroma(c.y, c.z);
    unsigned long long best = 0ull;
    for (unsigned int i = tid; i < n_local; i += THREADS) {
      const float4 v = s_work[i];
      const float d = ex::cie94_c(v.x, v.y, v.z, v.w, c.x, c.y, c.z, cc);
      const float dm = j == 1 ? fminf(1000000.0f, d) : fminf(s_dmin[i], d);  // kmeans++_calc_diff.wgsl:27-31
      s_dmin[i] = dm;
      const unsigned long long key = ((unsigned long long)__float_as_uint(dm) << 32) | (unsigned long long)((first + i) ^ 15u);
      best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    if (lane == 0) s_red[warp] = best;
    KMG_TRACE_MARK();  // init a: scanned
    __syncthreads();
    KMG_TRACE_MARK();  // init b: block barrier
    if (tid < csize) {  // thread r hands this CTA's maximum to rank r
      unsigned long long b = s_red[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) b = s_red[w] > b ? s_red[w] : b;
      *cluster.map_shared_rank(&s_keys[j & 1u][rank], tid) = b;
    }
    KMG_TRACE_MARK();  // init c: sent
    cluster.sync();
    KMG_TRACE_MARK();  // init d: cluster barrier
    unsigned long long gk = 0ull;
    for (unsigned int r = 0; r < csize; ++r) gk = s_keys[j & 1u][r] > gk ? s_keys[j & 1u][r] : gk;
    c = pixel_colour((unsigned int)key_to_pixel(gk));
    if (lane == 0) my_cent[j] = c;
    if (tid == 0 && rank == 0) J.keys[j] = gk;
    KMG_TRACE_MARK();  // init e: colour fetched
  }
  __syncwarp();

  // ---- Lloyd loop --------------------------------------------------------------------------------
  unsigned int it = 0, conv = 0;
  unsigned long long slow_total = 0;
  long long last[4] = {0, 0, 0, 0};  // lane c: reduced sums of cluster c in the last pass
  bool done = false;
  const unsigned int tiles = (ppc + THREADS * P - 1) / (THREADS * P);
  while (true) {
    // this warp's table of the current centroids (KCAP <= 32 entries, one lane each)
    float lmax = 0.0f, cmax = 0.0f;
    if (lane < KCAP) {
      CentRec r;
      if (lane < k) {
        const float4 v = my_cent[lane];
        const float c2 = ex::chroma(v.y, v.z);
        bool dup = false;
        for (unsigned int i = 0; i < lane; ++i) {
          const float4 u = my_cent[i];
          dup |= (u.x == v.x && u.y == v.y && u.z == v.z);
        }
        r.q[0] = dup ? MASKED : 0.5f * (v.x * v.x);
        r.q[1] = -v.x;
        r.q[2] = 0.5f * (c2 * c2);
        r.q[3] = c2;
        r.q[4] = -v.y;
        r.q[5] = -v.z;
        lmax = fabsf(v.x);
        cmax = c2;
      } else {
        r.q[0] = MASKED;
        r.q[1] = r.q[2] = r.q[3] = r.q[4] = r.q[5] = 0.0f;
      }
      *const_cast<CentRec*>(rec_at(s_tab, lane)) = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    __syncwarp();
    KMG_TRACE_MARK();  // pass a: table
    if (done) break;  // the table of the final centroids is not needed (build_table below redoes it in HBM)

    // assignment + thread-private accumulation over the local slice
    unsigned int slow = 0;
    for (unsigned int t = 0; t < tiles; ++t) {
      Pix<P> px;
      bool valid[P];
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const unsigned int p = (t * P + i) * THREADS + tid;
        valid[i] = p < n_local;
        const float4 v = valid[i] ? s_work[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        px.L[i] = v.x;
        px.a[i] = v.y;
        px.b[i] = v.z;
        px.C[i] = v.w;
      }
      float eps[P];
      unsigned int idx[P];
      bool certified[P];
      if (KCAP <= 16)
        argmin_small<P, (KCAP <= 16 ? KCAP : 8), false>(s_tab, px, lmax, cmax, 0.0f, eps, idx, certified);
      else
        argmin_chunked<P, false>(s_tab, KCAP, px, lmax, cmax, 0.0f, eps, idx, certified);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const bool need = !certified[i] && valid[i];
        if (__any_sync(0xffffffffu, need)) {
          idx[i] = warp_exact_argmin(s_tab, k, need, px.L[i], px.a[i], px.b[i], px.C[i], eps[i], idx[i]);
          slow += need ? 1u : 0u;
        }
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
    if (__any_sync(0xffffffffu, slow != 0)) {
      slow = __reduce_add_sync(0xffffffffu, slow);
      if (lane == 0) atomicAdd(&s_slow, slow);
    }
    KMG_TRACE_MARK();  // pass b: assigned
    __syncthreads();
    KMG_TRACE_MARK();  // pass c: block barrier

    // block fold: G consecutive lanes own one cluster's THREADS slots (and clear them for the next
    // pass), then hand the four sums to every rank of the cluster (all-to-all through DSMEM)
    const unsigned int par = it & 1u;
    if (tid < G * KCAP) {  // whole warps
      const unsigned int cl = tid / G, sub = tid % G;
      long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 4
      for (unsigned int u = sub; u < THREADS; u += G) {
        const int4 a = s_acc[cl * THREADS + u];
        s_acc[cl * THREADS + u] = make_int4(0, 0, 0, 0);
        s0 += a.x;
        s1 += a.y;
        s2 += a.z;
        s3 += a.w;
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      }
      long long* mine = s_x + ((size_t)par * csize + rank) * XS + cl * 4;
      for (unsigned int r = sub; r < csize; r += G) {
        longlong2* dst = reinterpret_cast<longlong2*>(cluster.map_shared_rank(mine, r));
        dst[0] = make_longlong2(s0, s1);
        dst[1] = make_longlong2(s2, s3);
      }
      if (warp == 0) {
        const unsigned int sl = s_slow;
        if (lane < csize) *cluster.map_shared_rank(s_x + ((size_t)par * csize + rank) * XS + KCAP * 4, lane) = (long long)sl;
        __syncwarp();
        if (lane == 0) s_slow = 0;
      }
    }
    KMG_TRACE_MARK();  // pass d: folded + sent
    cluster.sync();
    KMG_TRACE_MARK();  // pass e: cluster barrier

    // finalisation, redundantly and in the same fixed order in every warp of every CTA
    // (choose_centroid.wgsl:180-206; see finalize_pass)
    bool flag = false;
    if (lane < k) {
      long long s[4] = {0, 0, 0, 0};
      for (unsigned int r = 0; r < csize; ++r) {
        const longlong2* a = reinterpret_cast<const longlong2*>(s_x + ((size_t)par * csize + r) * XS + lane * 4);
        const longlong2 a0 = a[0], a1 = a[1];
        s[0] += a0.x;
        s[1] += a0.y;
        s[2] += a1.x;
        s[3] += a1.y;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) last[q] = s[q];
      if (s[3] > 0) {
        const double cnt = (double)s[3];
        const float4 prev = my_cent[lane];
        float4 nc;
        nc.x = (float)(((double)s[0] / cnt) * (1.0 / 65536.0));
        nc.y = (float)(((double)s[1] / cnt) * (1.0 / 65536.0));
        nc.z = (float)(((double)s[2] / cnt) * (1.0 / 65536.0));
        nc.w = 1.0f;
        my_cent[lane] = nc;
        flag = ex::cie94(nc.x, nc.y, nc.z, prev.x, prev.y, prev.z) < prm.conv_threshold;
      }
    }
    {
      long long sl = lane < csize ? s_x[((size_t)par * csize + lane) * XS + KCAP * 4] : 0ll;
      slow_total += (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)sl);
    }
    conv = __popc(__ballot_sync(0xffffffffu, flag));
    // core/src/modules.rs:802,827 — tested only when it > 0 && it % 8 == 0; also the hard cap.
    const bool check = it > 0 && prm.check_every != 0 && (it % prm.check_every) == 0;
    done = (check && conv >= k) || it + 1 >= prm.max_iter;
    ++it;
    __syncwarp();
    KMG_TRACE_MARK();  // pass f: finalised
  }

  // ---- results: centroids, state, table / dither threshold / RGBA8 palette for the remap --------
  if (rank == 0) {
    if (warp == 0 && lane < k) {
      J.cent[lane] = my_cent[lane];
#pragma unroll
      for (int q = 0; q < 4; ++q) J.last[lane * 4 + q] = last[q];
    }
    if (tid == 0) {
      JobState* st = J.st;
      st->ticket = 0;
      st->conv = conv;
      st->passes = it;
      st->done = 1;
      st->k = k;
      st->max_iter = prm.max_iter;
      st->check_every = prm.check_every;
      st->conv_threshold = prm.conv_threshold;
      st->slow_pixels = slow_total;
    }
    __syncthreads();
    build_table<THREADS>(J, k, prm.color_space, prm.want_palette != 0);
  }
  KMG_TRACE_MARK();  // end
}

}  // namespace kmg
