// kmg_init_lazy.cuh — farthest-point initialisation (plus_plus_init.wgsl:70-187,
// kmeans++_calc_diff.wgsl:14-34; host loop core/src/modules.rs:946-1284) as ONE persistent launch
// that only touches the pixels that can still win.
//
// The reference recomputes, for every centroid j, the distance of EVERY pixel to all j centroids
// chosen so far (O(k^2 n) pair evaluations in 3 k dispatches); k_init_round (kmg_kernels.cuh) keeps
// a running minimum and sweeps 24 B/px per round (O(k n)).  Here a round is lazy and still exact:
//
//   * the running minimum of a pixel only ever decreases, so a stale value is an upper bound.
//     Every pixel carries  dmin (f32, exact w.r.t. the first `fold` centroids),  fold (u16)  and
//     ub (u16) = the top 16 bits of dmin, rounded up — a 2 B/px array that one sweep reads at
//     stream speed;
//   * round j sweeps ub only: pixels with ub >= tau are candidates.  Their indices are compacted
//     per block, then each candidate is refreshed by one thread — the centroids fold .. j-1 are
//     folded into its minimum with the exact IEEE distance — and takes part in the 64-bit arg-max
//     key (distance bits << 32 | pixel ^ 15), the tie rule of selectCandidate
//     (plus_plus_init.wgsl:62-68).  The skipped pixels contribute the largest bound among them;
//   * the round is resolved when the best exact distance is strictly above every skipped pixel's
//     bound: no skipped pixel can reach it, not even tie with it, so the winner — and its tie-break
//     — is the one a full sweep finds.  Otherwise tau drops (to the best exact distance found, or
//     below the largest skipped bound when nothing was found) and the sweep repeats;
//   * tau of the next round is RHO x the distance just found (farthest-point distances never grow).
//
// The first rounds are full sweeps (k_init_round, which also maintains the bounds): while only a few
// centroids exist almost every pixel is a candidate, and a streaming sweep beats scattered
// refreshes.  The remaining rounds run inside one cooperative launch (grid-wide barriers between sweep and
// resolution): no launch gaps, the centroid list stays in shared memory, and on a sharded image
// block 0 does the NVLink mailbox exchange of the round right there (PICK 2, as k_init_round<.,2>).
#pragma once
#include <cooperative_groups.h>

#include "kmg_kernels.cuh"

namespace kmg {

namespace cg = cooperative_groups;

constexpr float LAZY_RHO = 0.98f;

// smallest 16-bit value u with float(u << 16) >= d (d >= 0, finite)
__device__ __forceinline__ unsigned int up16(float d) { return (__float_as_uint(d) + 0xffffu) >> 16; }

constexpr unsigned int LAZY_QCAP = 1280;  // candidate indices a warp can hold (a step adds at most 256, four steps per check)

// PICK 1: single GPU.  PICK 2: sharded image, peer mailboxes.
// Dynamic shared memory: float4 cent[k] (x, y, z, chroma) + uint32 queue[8 warps][LAZY_QCAP].
// Every warp sweeps its own 256-pixel steps (one 128-bit load of eight 16-bit bounds per lane, four
// steps in flight), collects the candidates in its own queue and refreshes them 32 at a time when the
// queue fills up or the sweep ends — no block barrier inside a sweep.
// (the sharded variant carries the exchange code: three blocks per SM keep its sweep free of spills,
// measured faster on 2 GPUs; the single-GPU variant is faster with four)
template <int PICK>
__global__ void __launch_bounds__(256, PICK == 2 ? 3 : 4) k_init_lazy(JobPtrs J, const float4* __restrict__ work, float* __restrict__ dmin,
                                                   unsigned short* __restrict__ ub, unsigned short* __restrict__ fold,
                                                   unsigned long long n, unsigned long long pixel_offset, PeerXchg X,
                                                   unsigned int j0) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  JobState* st = J.st;
  const unsigned int k = st->k;
  float4* s_cent = reinterpret_cast<float4*>(smem_raw);
  const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  unsigned int* s_queue = reinterpret_cast<unsigned int*>(smem_raw + (size_t)k * 16) + warp * LAZY_QCAP;
  __shared__ unsigned int s_nc, s_fault, s_resolved;
  __shared__ unsigned int s_qn[8];  // entries reserved in each warp's queue
  __shared__ unsigned long long s_key[8];
  if (tid == 0) s_nc = 0;
  if (tid < 8) s_qn[tid] = 0;
  // n < 2^32 (validate_dims): step numbers and the per-thread statistics fit 32 bits
  const unsigned int steps = (unsigned int)((n + 255) / 256);  // warp steps of 32 lanes x 8 pixels
  const unsigned int gwarp = blockIdx.x * 8 + warp, gwarps = gridDim.x * 8;
  unsigned int refreshed = 0, folds = 0, exact = 0;
  unsigned int fails = 0;  // block 0, thread 0: unresolved sweeps of the current round

  // rounds 1 .. j0-1 were full sweeps (k_init_round with bounds): centroids 0 .. j0-1 exist, every
  // minimum is exact w.r.t. centroids 0 .. j0-2
  for (unsigned int i = tid; i + 1 < j0; i += 256) {
    const float4 c = __ldcg(J.cent + i);
    s_cent[i] = make_float4(c.x, c.y, c.z, ex::chroma(c.y, c.z));
  }
  for (unsigned int j = j0; j < k; ++j) {
    // centroid j-1: resolved by block 0 before the last grid barrier (by round 1 for j == 2)
    if (tid == 0) {
      const float4 c = __ldcg(J.cent + (j - 1));
      s_cent[j - 1] = make_float4(c.x, c.y, c.z, ex::chroma(c.y, c.z));
    }
    __syncthreads();
    for (;;) {
      const unsigned int tau16 = __ldcg(&st->init_tau16);
      unsigned long long best = 0ull;
      unsigned int ncmax = 0;  // 1 + largest skipped bound
      unsigned int qn = 0;     // entries in this warp's queue (warp-uniform)

      // refresh the queued candidates, 32 at a time: fold centroids fold .. j-1 into the minimum
      auto drain = [&]() {
        __syncwarp();
        for (unsigned int q = lane; q < qn; q += 32) {
          const unsigned int p = s_queue[q];
          const float4 v = work[p];
          float d = __ldcg(dmin + p);
          const unsigned int f = max((unsigned int)__ldcg(fold + p), j0 - 1u);  // rounds < j0 were full sweeps: centroids 0 .. j0-2 are in
          // A centroid can only lower the minimum if the lower bound
          //   dL^2 + (da^2 + db^2) / SC^2  <=  dL^2 + (dC / SC)^2 + (dH / SH)^2 = d^2   (SC >= SH, dC^2 + dH^2 = da^2 + db^2)
          // lies below it; the exact distance (2 sqrt, 2 div) is evaluated only then.  The margin covers
          // the rounding of the bound and of the reference's own f32 evaluation many times over.
          const float sc = fmaf(0.045f, v.w, 1.0f);
          const float inv_sc2 = fast::rcp(sc * sc);
          float d2 = d * d * 1.00001f;
          for (unsigned int i = f; i < j; ++i) {
            const float4 c = s_cent[i];
            const float dl = v.x - c.x, da = v.y - c.y, db = v.z - c.z;
            const float lb2 = fmaf(dl, dl, fmaf(da, da, db * db) * inv_sc2);
            if (lb2 < d2) {
              const float dd = ex::cie94_c(v.x, v.y, v.z, v.w, c.x, c.y, c.z, c.w);
              ++exact;
              if (dd < d) {
                d = dd;
                d2 = d * d * 1.00001f;
              }
            }
          }
          if (f < j) {
            dmin[p] = d;
            fold[p] = (unsigned short)j;
            ub[p] = (unsigned short)up16(d);
            folds += j - f;
          }
          ++refreshed;
          const unsigned long long key =
              ((unsigned long long)__float_as_uint(d) << 32) | (((pixel_offset + p) & 0xffffffffull) ^ 15ull);
          best = key > best ? key : best;
        }
        if (lane == 0) s_qn[warp] = 0;
        __syncwarp();
        qn = 0;
      };

      {
        // sweep of the bounds: lane l of a step holds pixels step * 256 + 8 l .. + 7, two 16-bit bounds
        // per word.  The upper bound of a word is tested in place (word >= tau << 16 <=> upper half >= tau,
        // and the largest skipped word carries the largest skipped upper half), the lower one shifted up.
        const unsigned int tau_hi = tau16 << 16;
        unsigned int nc_hi = 0, nc_lo = 0, skipped = 0;
        // candidates of one step into the warp's queue: the lanes that hold some reserve their places
        // with one shared-memory add (the order inside the queue does not matter)
        auto push = [&](unsigned int step, unsigned int mine) {
          qn += __reduce_add_sync(0xffffffffu, __popc(mine));
          if (mine) {
            unsigned int at = atomicAdd(&s_qn[warp], (unsigned int)__popc(mine));
            const unsigned int p0 = step * 256 + lane * 8;
            do {
              s_queue[at++] = p0 + (unsigned int)__ffs((int)mine) - 1u;
              mine &= mine - 1u;
            } while (mine);
          }
        };
        auto take = [&](unsigned int step, const uint4& q) {
          const unsigned int w[4] = {q.x, q.y, q.z, q.w};
          unsigned int mine = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const unsigned int lo = w[e] << 16;
            if (w[e] >= tau_hi)
              mine |= 2u << (2 * e);
            else
              nc_hi = max(nc_hi, w[e]);
            if (lo >= tau_hi)
              mine |= 1u << (2 * e);
            else
              nc_lo = max(nc_lo, lo);
          }
          skipped |= mine ^ 0xffu;
          push(step, mine);
        };
        const unsigned int full = (unsigned int)(n / 256);  // steps with all 256 pixels
        auto load = [&](unsigned int step) {
          return step < full ? __ldcg(reinterpret_cast<const uint4*>(ub + (size_t)step * 256 + lane * 8))
                             : make_uint4(0u, 0u, 0u, 0u);
        };
        unsigned int step = gwarp;
        uint4 q0 = load(step), q1 = load(step + gwarps), q2 = load(step + 2 * gwarps), q3 = load(step + 3 * gwarps);
        for (; step < full; step += 4 * gwarps) {  // four loads in flight per lane
          take(step, q0);
          q0 = load(step + 4 * gwarps);
          if (step + gwarps < full) take(step + gwarps, q1);
          q1 = load(step + 5 * gwarps);
          if (step + 2 * gwarps < full) take(step + 2 * gwarps, q2);
          q2 = load(step + 6 * gwarps);
          if (step + 3 * gwarps < full) take(step + 3 * gwarps, q3);
          q3 = load(step + 7 * gwarps);
          if (qn > LAZY_QCAP - 1024) drain();
        }
        if (full < steps && full % gwarps == gwarp) {  // the ragged last step, pixel by pixel
          const unsigned long long p0 = (unsigned long long)full * 256 + lane * 8;
          unsigned int mine = 0;
          for (unsigned int e = 0; e < 8 && p0 + e < n; ++e) {
            const unsigned int u = __ldcg(ub + p0 + e);
            if (u >= tau16) {
              mine |= 1u << e;
            } else {
              nc_lo = max(nc_lo, u << 16);
              skipped = 1;
            }
          }
          push(full, mine);
        }
        drain();
        ncmax = skipped ? (max(nc_hi, nc_lo) >> 16) + 1u : 0u;
      }
      // warp -> block -> grid
      best = warp_max_u64(best);
      ncmax = __reduce_max_sync(0xffffffffu, ncmax);
      if (lane == 0) {
        s_key[warp] = best;
        if (ncmax) atomicMax(&s_nc, ncmax);
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < 8; ++w) best = s_key[w] > best ? s_key[w] : best;
        if (best) atomicMax(J.keys + j, best);
        if (s_nc) atomicMax(&st->init_ncmax, s_nc);
        s_nc = 0;
      }
      __threadfence();
      grid.sync();
      // resolution by block 0
      if (blockIdx.x == 0) {
        unsigned long long key = 0ull, pix = ~0ull;
        float4 col = make_float4(0.f, 0.f, 0.f, 1.0f);
        if (tid == 0) {
          key = __ldcg(J.keys + j);
          const unsigned int nc = __ldcg(&st->init_ncmax);
          const unsigned int best_bits = (unsigned int)(key >> 32);
          // strictly above every skipped bound: no skipped pixel can reach or tie the best distance
          // (nothing skipped: the sweep was a full one)
          const bool resolved = nc == 0 || (key != 0ull && best_bits > ((nc - 1u) << 16));
          s_resolved = resolved ? 1u : 0u;
          st->init_attempts += 1;
          if (!resolved) {
            // candidates found: next, everything that can reach the best of them (resolves for sure).
            // None found: drop below the largest skipped bound, in growing steps (16-bit float prefixes:
            // 128 units = a factor of two), down to 0 = every pixel
            unsigned int t16;
            if (key != 0ull) {
              t16 = best_bits >> 16;
            } else {
              const unsigned int drop = 8u << min(fails, 12u);
              t16 = nc - 1u > drop ? nc - 1u - drop : 0u;
              ++fails;
            }
            if (t16 >= tau16) t16 = tau16 - 1;  // always make progress (tau16 > 0 here: with 0 nothing is skipped)
            st->init_tau16 = t16;
            st->init_ncmax = 0;
          } else {
            fails = 0;
            pix = key_to_pixel(key);
            if (pix >= pixel_offset && pix - pixel_offset < n) {
              const float4 v = work[pix - pixel_offset];
              col = make_float4(v.x, v.y, v.z, 1.0f);
            } else {
              pix = ~0ull;  // zero maximum on a shard that does not hold pixel 0: no candidate
            }
          }
        }
        __syncthreads();
        if (s_resolved) {
          bool ok = true;
          if (PICK == 2) ok = init_exchange(X, k, j, key, pix, col, &s_fault);
          if (tid == 0) {
            if (!ok) {
              st->conv = PASS_FAULT;
              st->done = 1;
            }
            J.cent[j] = col;
            J.keys[j] = key;
            // farthest-point distances never grow: the next winner lies at or below this one
            const float bj = __uint_as_float((unsigned int)(key >> 32));
            st->init_tau16 = __float_as_uint(bj * LAZY_RHO) >> 16;
            st->init_ncmax = 0;
            st->init_done_round = j;
          }
        }
      }
      __threadfence();
      grid.sync();
      // a peer is gone: every later round would wait 4 s for it again — all blocks leave together
      if (PICK == 2 && __ldcg(&st->conv) == PASS_FAULT) return;
      if (__ldcg(&st->init_done_round) == j) break;
    }
  }
  const unsigned long long refreshed_w = (unsigned long long)warp_sum_i64((long long)refreshed);
  const unsigned long long folds_w = (unsigned long long)warp_sum_i64((long long)folds);
  const unsigned long long exact_w = (unsigned long long)warp_sum_i64((long long)exact);
  if (lane == 0 && refreshed_w) {
    atomicAdd(&st->init_refreshed, refreshed_w);
    atomicAdd(&st->init_folds, folds_w);
  }
  if (lane == 0 && exact_w) atomicAdd(&st->init_exact, exact_w);
}

}  // namespace kmg
