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
// A batch of frames is one launch with one cluster per frame (BASELINE config 5).  Large batches
// switch to a throughput mode of the same code: cluster size 1, one persistent CTA per SM looping
// over frames, the two planes in an L2-resident scratch slice — a cluster finishes one image sooner
// but idles its SMs during barriers and the serial finalisation, a lone CTA keeps its SM busy.
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
  int tail;                      // 0: centroids only, 1: + search table and RGBA8 palette, 2: + dither threshold
  size_t blob_stride;            // bytes between the job blobs of consecutive frames
  const float* lut;
  // Throughput mode for large batches: one persistent CTA per SM (no cluster) walks the frames
  // blockIdx.x, blockIdx.x + gridDim.x, ...; its work / distance planes live in a per-CTA slice of
  // this L2-resident scratch (ppc * 20 bytes each) instead of shared memory.  NULL: cluster mode.
  unsigned char* gscratch;
  unsigned int n_frames;
};

// exchange slots per rank: k x 4 sums + {exact-path pixel count, pad}
__host__ __device__ inline unsigned int small_xslots(unsigned int kcap) { return kcap * 4 + 2; }
__host__ __device__ inline size_t small_smem_bytes(unsigned int ppc, unsigned int kcap, unsigned int threads,
                                                   unsigned int csize) {
  return (size_t)ppc * 20 + (size_t)kcap * threads * 16 + (size_t)2 * csize * small_xslots(kcap) * 8;
}

// Warp-wide maximum of a 64-bit key in two 32-bit hardware reductions (REDUX).
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long v) {
  const unsigned int hi = (unsigned int)(v >> 32);
  const unsigned int mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned int ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned int)v : 0u);
  return ((unsigned long long)mh << 32) | ml;
}
// Warp-wide sum of 64-bit integers whose absolute values stay below 2^43: two 32-bit reductions.
__device__ __forceinline__ long long warp_sum_split(long long v) {
  const int hi = (int)(v >> 20);
  const int lo = (int)(v & 0xfffff);
  return ((long long)__reduce_add_sync(0xffffffffu, hi) << 20) + (long long)__reduce_add_sync(0xffffffffu, lo);
}

// KCAP: table / accumulator capacity (8, 16: saved-score search; 32: chunked search).
template <int KCAP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_kmeans_small(SmallParams prm, JobPtrs J0) {
  static_assert(KCAP == 8 || KCAP == 16 || KCAP == 32, "table capacity");
  static_assert(THREADS % KCAP == 0, "fold groups");
  constexpr int P = 2;
  constexpr int G = THREADS / KCAP < 32 ? THREADS / KCAP : 32;  // lanes that fold one cluster's slots
  constexpr int NW = THREADS / 32;
  constexpr int TAB_BYTES = (KCAP / 8) * CHUNK_BYTES;

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int csize = cluster.num_blocks();
  const unsigned int rank = cluster.block_rank();
  const bool gmode = prm.gscratch != nullptr;
  const unsigned int tid = threadIdx.x;
  const unsigned int lane = tid & 31u, warp = tid >> 5;
  const unsigned int k = prm.k;
  const unsigned int ppc = prm.ppc;
  const unsigned int XS = small_xslots(KCAP);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* plane_base = gmode ? prm.gscratch + (size_t)blockIdx.x * ((size_t)ppc * 20) : smem_raw;
  unsigned char* rest = gmode ? smem_raw : smem_raw + (size_t)ppc * 20;
  float4* s_work = reinterpret_cast<float4*>(plane_base);
  float* s_dmin = reinterpret_cast<float*>(plane_base + (size_t)ppc * 16);
  int4* s_acc = reinterpret_cast<int4*>(rest);
  long long* s_x = reinterpret_cast<long long*>(rest + (size_t)KCAP * THREADS * 16);
  // every warp searches its own copy of the table (built redundantly from the shared centroids)
  __shared__ __align__(16) unsigned char s_tab_raw[NW][TAB_BYTES];
  __shared__ float4 s_cent[KCAP];
  __shared__ long long s_last[KCAP * 4];
  __shared__ unsigned int s_flag[KCAP];
  __shared__ unsigned int s_pal[KCAP];
  __shared__ float s_lut[256];
  __shared__ float s_u8f[256];
  __shared__ unsigned long long s_keys[2][SMALL_MAX_CLUSTER];
  __shared__ unsigned long long s_red[NW];
  __shared__ unsigned int s_slow;
  CentRec* s_tab = reinterpret_cast<CentRec*>(s_tab_raw[warp]);

  const unsigned int N = prm.dw * prm.dh;
  const unsigned int first = rank * ppc;
  const unsigned int n_local = first < N ? min(ppc, N - first) : 0u;
  for (unsigned int c = tid; c < 256; c += THREADS) {
    s_lut[c] = prm.lut[c];
    s_u8f[c] = fdiv((float)c, 255.0f);
  }

  // cluster mode: exactly one frame per cluster; throughput mode: this CTA's share of the batch
  const unsigned int frame_step = gmode ? gridDim.x : 0xffffffffu;
  for (unsigned int frame = gmode ? blockIdx.x : blockIdx.x / csize; frame < prm.n_frames;
       frame = frame_step > prm.n_frames - frame ? prm.n_frames : frame + frame_step) {
  const JobPtrs J = job_at(J0, (size_t)frame * prm.blob_stride);
  const uint32_t* src = prm.src + (size_t)frame * prm.frame_px;

  KMG_TRACE_DECL;
  KMG_TRACE_MARK();  // 0: start
  // ---- shrink + convert into the local slice of the work plane --------------------------------
  if (tid < KCAP) s_flag[tid] = 0;
  if (tid == 0) s_slow = 0;
  __syncthreads();
  {
    // RGBA8 of the clustered image: taps of four pixels in flight per thread
    auto to_work = [&](uint32_t v) -> float4 {
      return prm.color_space == 0 ? ex::lin100_to_lab(s_lut[v & 255u], s_lut[(v >> 8) & 255u], s_lut[(v >> 16) & 255u])
                                  : ex::rgb8_to_rgbf(v);
    };
    auto u8f = [&](uint32_t v) { return s_u8f[v]; };
    constexpr int U = 4;
    for (unsigned int i0 = tid; i0 < n_local; i0 += U * THREADS) {
      uint32_t v[U];
      if (prm.shrink) {
        ResizeTaps t[U];
        uint32_t tap[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned int i = min(i0 + u * THREADS, n_local - 1);
          t[u] = resize_taps(prm.sw, prm.sh, prm.dw, prm.dh, first + i);
          tap[u][0] = __ldcs(src + t[u].i00);  // streaming: the source is read once
          tap[u][1] = __ldcs(src + t[u].i10);
          tap[u][2] = __ldcs(src + t[u].i01);
          tap[u][3] = __ldcs(src + t[u].i11);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = resize_blend(tap[u][0], tap[u][1], tap[u][2], tap[u][3], t[u].fx, t[u].fy, u8f);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(src + first + min(i0 + u * THREADS, n_local - 1));
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * THREADS < n_local) s_work[i0 + u * THREADS] = to_work(v[u]);
    }
  }
#pragma unroll 4
  for (int q = 0; q < KCAP; ++q) s_acc[q * THREADS + tid] = make_int4(0, 0, 0, 0);
  KMG_TRACE_MARK();  // 1: converted
  cluster.sync();
  KMG_TRACE_MARK();  // 2: cluster barrier

  // ---- farthest-point init (plus_plus_init.wgsl, kmeans++_calc_diff.wgsl) ----------------------
  auto pixel_colour = [&](unsigned int g) -> float4 {  // any pixel of the image, through DSMEM
    const float4* owner = gmode ? s_work : cluster.map_shared_rank(s_work, g / ppc);
    float4 v = owner[g % ppc];
    v.w = 1.0f;
    return v;
  };
  float4 c = pixel_colour(prm.seed);
  if (tid == 0) {
    s_cent[0] = c;
    if (rank == 0) J.keys[0] = 0ull;
  }
  for (unsigned int j = 1; j < k; ++j) {
    const float cc = ex::chroma(c.y, c.z);
    unsigned long long best = 0ull;
    constexpr int UI = 4;  // pixels in flight per thread (the planes may sit in L2)
    for (unsigned int i0 = tid; i0 < n_local; i0 += UI * THREADS) {
      float4 v[UI];
      float prev[UI];
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const unsigned int i = i0 + u * THREADS;
        const bool ok = i < n_local;
        v[u] = ok ? s_work[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        prev[u] = (ok && j > 1) ? s_dmin[i] : 1000000.0f;  // kmeans++_calc_diff.wgsl:27-31
      }
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const unsigned int i = i0 + u * THREADS;
        if (i < n_local) {
          const float d = ex::cie94_c(v[u].x, v[u].y, v[u].z, v[u].w, c.x, c.y, c.z, cc);
          const float dm = fminf(prev[u], d);
          s_dmin[i] = dm;
          const unsigned long long key = ((unsigned long long)__float_as_uint(dm) << 32) | (unsigned long long)((first + i) ^ 15u);
          best = key > best ? key : best;
        }
      }
    }
    best = warp_max_key(best);
    if (lane == 0) s_red[warp] = best;
    KMG_TRACE_MARK();  // init a: scanned
    __syncthreads();
    if (warp == 0) {  // lane r hands this CTA's maximum to rank r
      const unsigned long long b = warp_max_key(lane < NW ? s_red[lane] : 0ull);
      if (lane < csize) *cluster.map_shared_rank(&s_keys[j & 1u][rank], lane) = b;
    }
    KMG_TRACE_MARK();  // init b: sent
    cluster.sync();
    KMG_TRACE_MARK();  // init c: cluster barrier
    const unsigned long long gk = warp_max_key(lane < csize ? s_keys[j & 1u][lane] : 0ull);
    c = pixel_colour((unsigned int)key_to_pixel(gk));
    if (tid == 0) {
      s_cent[j] = c;
      if (rank == 0) J.keys[j] = gk;
    }
    KMG_TRACE_MARK();  // init d: colour fetched
  }
  __syncthreads();

  // ---- Lloyd loop --------------------------------------------------------------------------------
  unsigned int it = 0, conv = 0;
  unsigned long long slow_total = 0;
  bool done = false;
  float lmax = 0.0f, cmax = 0.0f;
  const unsigned int tiles = (ppc + THREADS * P - 1) / (THREADS * P);
  while (true) {
    // this warp's table of the current centroids (KCAP <= 32 entries, one lane each)
    {
      float lm = 0.0f, cm = 0.0f;
      if (lane < KCAP) {
        CentRec r;
        if (lane < k) {
          const float4 v = s_cent[lane];
          const float c2 = ex::chroma(v.y, v.z);
          bool dup = false;
#pragma unroll
          for (unsigned int i = 0; i < KCAP - 1; ++i) {
            const float4 u = s_cent[i];
            dup |= (i < lane) && (u.x == v.x && u.y == v.y && u.z == v.z);
          }
          r.q[0] = dup ? MASKED : 0.5f * (v.x * v.x);
          r.q[1] = -v.x;
          r.q[2] = 0.5f * (c2 * c2);
          r.q[3] = c2;
          r.q[4] = -v.y;
          r.q[5] = -v.z;
          lm = fabsf(v.x);
          cm = c2;
        } else {
          r.q[0] = MASKED;
          r.q[1] = r.q[2] = r.q[3] = r.q[4] = r.q[5] = 0.0f;
        }
        *const_cast<CentRec*>(rec_at(s_tab, lane)) = r;
      }
      // non-negative floats order like their bit patterns
      lmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lm)));
      cmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(cm)));
      __syncwarp();
    }
    KMG_TRACE_MARK();  // pass a: table
    if (done) break;

    // assignment + thread-private accumulation over the local slice
    unsigned int slow = 0;
    float4 nxt[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const unsigned int p = i * THREADS + tid;
      nxt[i] = p < n_local ? s_work[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (unsigned int t = 0; t < tiles; ++t) {
      Pix<P> px;
      bool valid[P];
#pragma unroll
      for (int i = 0; i < P; ++i) {
        valid[i] = (t * P + i) * THREADS + tid < n_local;
        px.L[i] = nxt[i].x;
        px.a[i] = nxt[i].y;
        px.b[i] = nxt[i].z;
        px.C[i] = nxt[i].w;
      }
      // the next tile's pixels are in flight while this one is searched (L2 latency in throughput mode)
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const unsigned int p = ((t + 1) * P + i) * THREADS + tid;
        nxt[i] = p < n_local ? s_work[p] : make_float4(0.f, 0.f, 0.f, 0.f);
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
          // private int32 slots, no flush inside a pass: a thread sees at most 150,000 / (8 CTAs x 256
          // threads) < 80 pixels (<= 256 in throughput mode, 65,536 px on one CTA), each |fixed| < 2^22
          // (|v| < 128 at 2^-15) — far below 2^31
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
      if (G == 32) {
        s0 = warp_sum_split(s0);
        s1 = warp_sum_split(s1);
        s2 = warp_sum_split(s2);
        s3 = warp_sum_split(s3);
      } else {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        }
      }
      long long* mine = s_x + ((size_t)par * csize + rank) * XS + cl * 4;
      for (unsigned int r = sub; r < csize; r += G) {
        longlong2* dst = reinterpret_cast<longlong2*>(cluster.map_shared_rank(mine, r));
        dst[0] = make_longlong2(s0, s1);
        dst[1] = make_longlong2(s2, s3);
      }
    }
    if (warp == NW - 1) {
      const unsigned int sl = s_slow;
      if (lane < csize) *cluster.map_shared_rank(s_x + ((size_t)par * csize + rank) * XS + KCAP * 4, lane) = (long long)sl;
      __syncwarp();
      if (lane == 0) s_slow = 0;
    }
    KMG_TRACE_MARK();  // pass c: folded + sent
    cluster.sync();
    KMG_TRACE_MARK();  // pass d: cluster barrier

    // finalisation (choose_centroid.wgsl:180-206; see finalize_pass): warp w reduces the partial
    // sums of clusters w, w + NW, ... over the ranks in parallel lanes and publishes the new
    // centroid; every CTA does the same work on the same numbers, so all CTAs stay identical
    for (unsigned int cl = warp; cl < k; cl += NW) {
      long long s[4] = {0, 0, 0, 0};
      if (lane < csize) {
        const longlong2* a = reinterpret_cast<const longlong2*>(s_x + ((size_t)par * csize + lane) * XS + cl * 4);
        const longlong2 a0 = a[0], a1 = a[1];
        s[0] = a0.x;
        s[1] = a0.y;
        s[2] = a1.x;
        s[3] = a1.y;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] = warp_sum_split(s[q]);
      bool flag = false;
      const float4 prev = s_cent[cl];
      __syncwarp();
      if (s[3] > 0) {  // warp-uniform
        const double cnt = (double)s[3];
        const long long mine = lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2]);
        const float comp = (float)(((double)mine / cnt) * ex::FIXED_UNIT);
        float4 nc;
        nc.x = __shfl_sync(0xffffffffu, comp, 0);
        nc.y = __shfl_sync(0xffffffffu, comp, 1);
        nc.z = __shfl_sync(0xffffffffu, comp, 2);
        nc.w = 1.0f;
        flag = ex::cie94(nc.x, nc.y, nc.z, prev.x, prev.y, prev.z) < prm.conv_threshold;
        if (lane == 0) s_cent[cl] = nc;
      }
      if (lane < 4) s_last[cl * 4 + lane] = s[lane];
      if (lane == 0) s_flag[cl] = flag ? 1u : 0u;
    }
    {
      const long long sl = lane < csize ? s_x[((size_t)par * csize + lane) * XS + KCAP * 4] : 0ll;
      slow_total += (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)sl);
    }
    __syncthreads();
    conv = __popc(__ballot_sync(0xffffffffu, lane < k && s_flag[lane] != 0));
    // core/src/modules.rs:802,827 — tested only when it > 0 && it % 8 == 0; also the hard cap.
    const bool check = it > 0 && prm.check_every != 0 && (it % prm.check_every) == 0;
    done = (check && conv >= k) || it + 1 >= prm.max_iter;
    ++it;
    KMG_TRACE_MARK();  // pass e: finalised
  }

  // ---- results: centroids and state; for the remap also the table, RGBA8 palette and dither
  //      threshold (k_prepare / build_table, spread over the threads of rank 0) -------------------
  if (rank == 0) {
    JobState* st = J.st;
    if (tid < k) {
      J.cent[tid] = s_cent[tid];
#pragma unroll
      for (int q = 0; q < 4; ++q) J.last[tid * 4 + q] = s_last[tid * 4 + q];
    }
    if (tid == 0) {
      st->ticket = 0;
      st->conv = conv;
      st->passes = it;
      st->done = 1;
      st->k = k;
      st->max_iter = prm.max_iter;
      st->check_every = prm.check_every;
      st->conv_threshold = prm.conv_threshold;
      st->slow_pixels = slow_total;
      st->lmax = lmax;
      st->cmax = cmax;
      st->dither_threshold = 0.0f;
    }
    if (prm.tail >= 1) {
      // warp 0's table holds the final centroids (the loop leaves after building it)
      if (tid < pad32(k)) {
        CentRec r;
        if (tid < KCAP) {
          r = *rec_at(reinterpret_cast<const CentRec*>(s_tab_raw[0]), tid);
        } else {
          r.q[0] = MASKED;
          r.q[1] = r.q[2] = r.q[3] = r.q[4] = r.q[5] = 0.0f;
        }
        J.tab[tid] = r;
      }
      // palette: one thread per (centroid, channel) — the transfer function is the expensive part
      if (tid < k) s_pal[tid] = 0xFF000000u;
      __syncthreads();
      if (tid < 3 * k) {
        const unsigned int cl = tid / 3, ch = tid % 3;
        const float4 v = s_cent[cl];
        unsigned int byte;
        if (prm.color_space == 0) {
          const float3 lin = ex::lab_to_linear_rgb(v.x, v.y, v.z);
          byte = ex::unorm8(ex::srgb_encode(ch == 0 ? lin.x : (ch == 1 ? lin.y : lin.z)));
        } else {
          byte = ex::unorm8(ch == 0 ? v.x : (ch == 1 ? v.y : v.z));
        }
        atomicOr(&s_pal[cl], byte << (8 * ch));
      }
      // dither threshold (mix_colors.wgsl:53-68): all pair distances in parallel, then the greedy
      // farthest-pair scan is a walk over a k x k table
      float* s_pair = reinterpret_cast<float*>(s_acc);  // KCAP * KCAP floats
      if (prm.tail >= 2 && k > 1) {
        for (unsigned int e = tid; e < k * k; e += THREADS) {
          const float4 ci = s_cent[e / k], cj = s_cent[e % k];
          s_pair[e] = ex::cie94(ci.x, ci.y, ci.z, cj.x, cj.y, cj.z);  // centroid i first
        }
      }
      __syncthreads();
      if (tid < k) {
        unsigned int pv = s_pal[tid];
        if (prm.color_space != 0) pv = (pv & 0x00ffffffu) | (ex::unorm8(s_cent[tid].w) << 24);
        J.pal[tid] = pv;
      }
      if (tid == 0 && prm.tail >= 2 && k > 1) {
        unsigned int a = 0, b = 1;
        float d_ab = s_pair[0 * k + 1];
        for (unsigned int i = 2; i < k; ++i) {
          const float da = s_pair[i * k + a], db = s_pair[i * k + b];
          if (da > db && da > d_ab) {
            d_ab = da;
            b = i;
          } else if (db > d_ab) {
            d_ab = db;
            a = i;
          }
        }
        st->dither_threshold = fdiv(d_ab, fsqrt((float)k));
      }
    }
  }
  KMG_TRACE_MARK();  // end
  __syncthreads();  // throughput mode: the next frame reuses the shared state
  }
}

}  // namespace kmg
