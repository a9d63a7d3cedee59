/* kmeans_gpu.h — C ABI of the B200-native image hot path of redwarp/kmeans-gpu.
 *
 * This is the drop-in boundary: it sits where core/src/lib.rs (kept) calls into
 * core/src/operations.rs / modules.rs / structures.rs (replaced).  Every entry point cites the
 * reference interface it replaces (paths relative to the reference repository root).  Plain
 * pointers and sizes only; the library owns all device memory, streams and scratch.
 *
 * Conventions
 *  - images: tightly packed row-major RGBA8, w*h*4 bytes, borrowed for the duration of the call,
 *    never mutated (core/src/image.rs:20-48).  Output alpha is always 255.
 *  - centroids: k x 4 floats in colour-space units (Lab: L,a,b,1.0 — Rgb: r,g,b,1.0 in [0,1]),
 *    the data[] part of the reference's CentroidsBuffer (core/src/structures.rs:495-553).
 *  - return value: KMG_OK (0) or a kmg_status; kmg_last_error() gives the message for the calling
 *    thread.  Mirrors anyhow::Result at the Rust boundary.
 *  - threading: a kmg_ctx may be used from many host threads at once (the reference shares one
 *    ImageProcessor through Arc across 14 threads, core/examples/parallel.rs:23,36-51).  Calls block
 *    until the result is in the caller's buffer (pollster::block_on, cli/src/main.rs:27-40).
 *  - there is no CPU fallback: every compute entry point fails with KMG_ERR_CUDA without a GPU.
 */
#ifndef KMEANS_GPU_H_
#define KMEANS_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMG_ABI_VERSION 1

typedef struct kmg_ctx kmg_ctx;

typedef enum kmg_status {
  KMG_OK = 0,
  KMG_ERR_BAD_ARG = 1,
  KMG_ERR_CUDA = 2,
  KMG_ERR_OOM = 3,
  KMG_ERR_NCCL = 4,
  KMG_ERR_UNSUPPORTED = 5
} kmg_status;

/* core/src/lib.rs:168-171 (ColorSpace) */
typedef enum kmg_color_space { KMG_LAB = 0, KMG_RGB = 1 } kmg_color_space;
/* core/src/lib.rs:235-239 (ReduceMode) */
typedef enum kmg_reduce_mode { KMG_REPLACE = 0, KMG_DITHER = 1, KMG_MELD = 2 } kmg_reduce_mode;

/* The reference's hard-coded constants, made explicit.  kmg_default_opts() fills in:
 *   max_dim 256 (core/src/structures.rs:23; 0 = never shrink, needed for BASELINE config 4),
 *   max_iter 128 and check_every 8 (core/src/modules.rs:765-766),
 *   convergence < 0 -> 1.0 for Lab / 0.01 for Rgb (core/src/lib.rs:189-194),
 *   seed fractions 0.5625 / 0.93359375 = rand(42.0), rand(12.0) of
 *   core/shaders/plus_plus_init.wgsl:58-60,161-165 under correctly rounded f32 sin,
 *   seed_x / seed_y = -1 (>= 0 selects an explicit seed pixel on the clustered image),
 *   flags 0. */
typedef struct kmg_opts {
  uint32_t struct_size; /* sizeof(kmg_opts), for forward compatibility */
  uint32_t max_dim;
  uint32_t max_iter;
  uint32_t check_every;
  float convergence;
  float seed_x_frac;
  float seed_y_frac;
  int32_t seed_x;
  int32_t seed_y;
  uint32_t flags; /* KMG_OPT_* bits, default 0 */
} kmg_opts;

/* Run the k-means of small (<= 256 x 256 after the shrink) images stage by stage (resize, convert,
 * one launch per init round and per Lloyd pass) instead of the single fused thread-block-cluster
 * launch.  Results are bit-identical either way; the flag exists for the parity tests. */
#define KMG_OPT_NO_FUSED_KMEANS 1u

/* ---- lifetime ------------------------------------------------------------------------------- */

/* ImageProcessor::new (core/src/lib.rs:38-65): bind to CUDA device `device`. */
int kmg_create(int device, kmg_ctx** out);
/* Drop of ImageProcessor. */
void kmg_destroy(kmg_ctx* ctx);
/* Message of the last failure on the calling thread (never NULL). */
const char* kmg_last_error(void);
int kmg_abi_version(void);
void kmg_default_opts(kmg_opts* opts);

/* ---- the reference's three operations, host buffers in and out ------------------------------ */

/* operations::extract_palette_kmeans (core/src/operations.rs:15-88) + the raw read-back half of
 * CentroidsBuffer::pull_values (core/src/structures.rs:581-598): shrink, convert, farthest-point
 * init, Lloyd loop with the reference stop rule.  centroids_out: k*4 floats, cluster order.
 * passes_out (optional): number of centroid-update passes executed (1..max_iter). */
int kmg_kmeans_palette(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int color_space,
                       const kmg_opts* opts, float* centroids_out, uint32_t* passes_out);

/* operations::find_colors / dither_colors / meld_colors (core/src/operations.rs:99-271) followed by
 * OutputTexture::pull_image (core/src/structures.rs:441-470): remap every pixel of the full-size
 * image onto the given centroids.  out_rgba: w*h*4 bytes. */
int kmg_remap(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, const float* centroids, uint32_t k,
              int color_space, int mode, uint8_t* out_rgba);

/* ImageProcessor::reduce with Algorithm::Kmeans (core/src/lib.rs:116-164): the two calls above
 * fused, the image uploaded once and the centroids never leaving the device.
 * centroids_out / passes_out are optional. */
int kmg_reduce(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int color_space, int mode,
               const kmg_opts* opts, uint8_t* out_rgba, float* centroids_out, uint32_t* passes_out);

/* InputTexture::resized (core/src/structures.rs:76-182) + pull_image — the bilinear shrink the
 * octree path needs (core/src/lib.rs:288-310).  out must hold out_w*out_h*4 bytes as given by
 * kmg_resized_dims. */
void kmg_resized_dims(uint32_t w, uint32_t h, uint32_t max_size, uint32_t* out_w, uint32_t* out_h);
int kmg_resize(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t max_size, uint8_t* out);

/* Batch of n_frames equally sized frames (BASELINE config 5): frame f occupies
 * rgba[f*w*h*4 ...].  Same result as calling kmg_reduce on every frame.  centroids_out
 * (n_frames*k*4) and passes_out (n_frames) are optional. */
int kmg_reduce_batch(kmg_ctx* ctx, const uint8_t* rgba, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t k,
                     int color_space, int mode, const kmg_opts* opts, uint8_t* out_rgba, float* centroids_out,
                     uint32_t* passes_out);

/* Page-locked host buffers.  The reference hands wgpu pageable slices (queue.write_texture,
 * core/src/structures.rs:47-63); with CUDA the host<->device copies of a call only run at full PCIe
 * speed, and the chunks of kmg_reduce_batch only overlap, when the caller's buffers are page-locked.
 * Optional: every entry point accepts pageable memory as well. */
void* kmg_alloc_pinned(size_t bytes);
void kmg_free_pinned(void* p);

/* ---- host-side colour helpers (CPU in the reference too: `palette` crate 0.7.3) -------------- */

/* CentroidsBuffer::fixed_centroids (core/src/structures.rs:523-553): sRGB8 -> k x 4 floats. */
void kmg_fixed_centroids(const uint8_t* colors_rgba8, uint32_t count, int color_space, float* centroids_out);
/* CentroidsBuffer::pull_values (core/src/structures.rs:600-617): k x 4 floats -> sRGB8 (alpha 255). */
void kmg_centroids_to_rgba8(const float* centroids, uint32_t count, int color_space, uint8_t* colors_out);
/* Sort key of kmeans_palette (core/src/lib.rs:276-284): stable sort of RGBA8 colours by Lab L. */
void kmg_sort_palette_by_lightness(uint8_t* colors_rgba8, uint32_t count);
/* operations::extract_palette_octree (core/src/operations.rs:90-97) over ColorTree::{add_color,reduce}
 * (core/src/octree.rs:41-110): the reference's CPU octree quantiser, CPU here too.  rgba = the
 * pixels octree_palette hands it (the image, or its <= 128 px shrink from kmg_resize,
 * core/src/lib.rs:288-316).  colors_out must hold color_count x 4 bytes; *count_out <= color_count
 * colours are written, sorted as (r,g,b,a) tuples and de-duplicated; the caller then sorts them
 * with kmg_sort_palette_by_lightness (lib.rs:318-329).  n_pixels < 2^28 (the reference passes at
 * most 128 x 128).  Host-only: needs no device and no kmg_ctx. */
int kmg_octree_palette(const uint8_t* rgba, uint64_t n_pixels, uint32_t color_count, uint8_t* colors_out,
                       uint32_t* count_out);

/* ---- device-resident entry points (inputs already in HBM; used by batch callers, the bench and
 *      the stage-level parity tests).  Pointers are device pointers on ctx's device; `stream` is a
 *      cudaStream_t (NULL = the CUDA default stream).  Calls are asynchronous on `stream` unless
 *      they return values to host memory, in which case they synchronise that stream. ---------- */

typedef struct kmg_job kmg_job; /* one k-means problem resident on the device */

/* K1/K3 (core/shaders/converters/rgb_to_lab.wgsl, rgb8u_to_rgb32f.wgsl): RGBA8 -> work plane,
 * n x float4 {c0,c1,c2, sqrt(c1^2+c2^2)}; bit-exact with the oracle. */
int kmg_dev_convert(kmg_ctx* ctx, const uint8_t* d_rgba, uint64_t n, int color_space, float* d_work, void* stream);
/* K15 (core/shaders/resize.wgsl). */
int kmg_dev_resize(kmg_ctx* ctx, const uint8_t* d_src, uint32_t sw, uint32_t sh, uint8_t* d_dst, uint32_t dw,
                   uint32_t dh, void* stream);
/* K5 (core/shaders/find_centroid.wgsl): labels for a work plane. d_labels: n x uint32. */
int kmg_dev_assign(kmg_ctx* ctx, const float* d_work, uint64_t n, const float* centroids_host, uint32_t k,
                   uint32_t* d_labels, void* stream);

/* A resident k-means job over a work plane of n pixels (w*h).  The plane is borrowed and must
 * outlive the job. */
int kmg_job_create(kmg_ctx* ctx, const float* d_work, uint32_t w, uint32_t h, uint32_t k, int color_space,
                   const kmg_opts* opts, kmg_job** out);
void kmg_job_destroy(kmg_job* job);
/* K8-K11 (core/shaders/plus_plus_init.wgsl, kmeans++_calc_diff.wgsl; host loop
 * core/src/modules.rs:946-1284).  pick_index/pick_dist (k each, host, optional) return the chosen
 * pixel and its max-min distance per round. */
int kmg_job_init(kmg_job* job, uint32_t* pick_index, float* pick_dist, void* stream);
int kmg_job_set_centroids(kmg_job* job, const float* centroids_host, void* stream);
int kmg_job_get_centroids(kmg_job* job, float* centroids_host, void* stream);
/* One fused assign+update pass = K5 + K6/K7 for all clusters (core/src/modules.rs:773-800).
 * Asynchronous.  `count` passes are enqueued back to back. */
int kmg_job_step(kmg_job* job, uint32_t count, void* stream);
/* The whole Lloyd loop with the reference stop rule (core/src/modules.rs:763-840). */
int kmg_job_run(kmg_job* job, uint32_t* passes_out, void* stream);
/* Convergence count of the last pass (convergence[k], core/shaders/choose_centroid.wgsl:196-201),
 * passes so far, and pixels that needed the exact re-evaluation path since job creation. */
int kmg_job_stats(kmg_job* job, uint32_t* converged_out, uint32_t* passes_out, uint64_t* slow_pixels_out,
                  void* stream);

/* Work done by the lazy farthest-point rounds of the last kmg_job_init (kmg_init_lazy.cuh): sweeps over
 * the 2 B/px bound array (one per lazy round if every round resolved at once), pixels refreshed,
 * (pixel, centroid) pairs folded into a running minimum, and how many of those pairs needed the exact
 * distance — full sweeps would fold (k - 1) * w * h pairs, all of them exactly. */
int kmg_job_init_stats(kmg_job* job, uint32_t* sweeps_out, uint64_t* refreshed_out, uint64_t* pairs_out,
                       uint64_t* exact_out, void* stream);

/* Reduced integer sums of the last pass: k x {sum0, sum1, sum2, count}, sums in units of 2^-15
 * (the k x 4 accumulators the multi-GPU path all-reduces; sums of shards add up exactly). */
int kmg_job_get_sums(kmg_job* job, int64_t* sums_out, void* stream);

/* Multi-GPU pixel sharding (BASELINE config 4): each rank owns a row block of one image as its
 * job's work plane; the per-pass k x 4 integer sums are all-reduced over NCCL so every rank holds
 * identical centroids.  kmg_comm_unique_id fills a 128-byte NCCL id on rank 0; the caller ships it
 * to the other ranks (torch.distributed / MPI / files) and every rank calls kmg_comm_init. */
int kmg_comm_unique_id(kmg_ctx* ctx, uint8_t id_out[128]);
int kmg_comm_init(kmg_ctx* ctx, const uint8_t id[128], int n_ranks, int rank);
int kmg_comm_destroy(kmg_ctx* ctx);
/* How the per-pass sums of a sharded job travel: 0 = no communicator, 1 = NCCL all-reduce after the
 * pass kernel, 2 = inside the pass kernel, through peer-mapped mailboxes over NVLink (chosen by
 * kmg_comm_init when CUDA IPC peer mapping works between all ranks; at most 8 ranks). */
int kmg_comm_mode(kmg_ctx* ctx);
/* Marks the job as one shard of a distributed problem: global_w/global_h describe the whole image
 * and row_offset the first row of this shard (used by the init seed and tie rule).
 * One live sharded job per context (KMG_ERR_BAD_ARG otherwise): the peer mailboxes of mode 2 are per
 * context.  kmg_job_step is asynchronous, but the ranks of a sharded job must launch matching passes
 * within 4 s of each other: a rank that waits longer gives up and poisons the mailboxes, every rank of
 * the job then fails with KMG_ERR_NCCL at its next state read (kmg_job_run / kmg_job_stats / ...), and
 * the communicator has to be re-created (kmg_comm_destroy + kmg_comm_init). */
int kmg_job_set_shard(kmg_job* job, uint32_t global_w, uint32_t global_h, uint32_t row_offset);

/* Fused remap on device buffers (mode as kmg_reduce_mode).  centroids_host: k x 4 floats. */
int kmg_dev_remap(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t w, uint32_t h, const float* centroids_host,
                  uint32_t k, int color_space, int mode, uint8_t* d_out, void* stream);
/* Same, taking the centroids of a resident job (no host round trip). */
int kmg_dev_remap_job(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t w, uint32_t h, kmg_job* job, int mode,
                      uint8_t* d_out, void* stream);
/* Batched reduce on device-resident frames. */
int kmg_dev_reduce_batch(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t k,
                         int color_space, int mode, const kmg_opts* opts, uint8_t* d_out, float* centroids_out_host,
                         uint32_t* passes_out_host, void* stream);

/* Synthetic image generator of SURVEY.md section 8(d) (uniform when blobs == 0). */
int kmg_dev_synth(kmg_ctx* ctx, uint8_t* d_rgba, uint64_t first_pixel, uint64_t n, uint32_t frame, uint32_t seed,
                  uint32_t blobs, void* stream);
/* sRGB decode table (256 floats, x100) as built on the device — exposed for the parity test. */
int kmg_dev_srgb_table(kmg_ctx* ctx, float table_out[256]);
/* Largest |approximate Lab - exact Lab| (Euclidean) over all 2^24 sRGB colours, as used by the
 * remap kernels' near-tie certificate — exposed for the parity test of that bound. */
int kmg_dev_fast_lab_error(kmg_ctx* ctx, float* max_err_out);
/* Measured FP32 (non-tensor) peak of the device in fused multiply-adds per second (x 2 = FLOP/s):
 * the FP32 side of the roofline bench.py reports next to the HBM side. */
int kmg_dev_fp32_peak(kmg_ctx* ctx, double* fma_per_second_out);
/* Audit of the near-tie certificate (test hook; tests/test_gpu_parity.py::test_certificate_audit_*).
 * The production searches trust a certified label without evaluating the reference distance; this
 * runs the same device functions on every pixel, also scans all k centroids with the reference's
 * arithmetic and order (core/shaders/find_centroid.wgsl:29-41), and returns in *wrong_out the number
 * of pixels whose certified label differs from that scan (any value but 0 is a bug in an error
 * bound) and in *uncertified_out the pixels production would hand to its exact path.
 *   search  0: 8-entry table, all scores kept   1: 16-entry table   2: chunked search, any k
 *           3: the resident-table search of the k <= 8 Lloyd pass (kmg_lloyd_ring.cuh)
 *           4: margin probe — *wrong_out = pixels whose fast arg-min differs from the reference label (the
 *              certificate has to catch every one of them), *uncertified_out = the largest
 *              (fast score gap between the two labels) / eps among them, in millionths: how much of the
 *              error bound is ever used
 *   mode    0: Lloyd pass / assignment — d_work is the exact work plane of the w*h pixels
 *           1: remap, replace   2: remap, ordered dither — d_rgba is the RGBA8 image (fast Lab in the
 *              kernel; the reference label is the scan of the exact pixel + dither offset) */
int kmg_dev_audit(kmg_ctx* ctx, const float* d_work, const uint8_t* d_rgba, uint32_t w, uint32_t h,
                  const float* centroids_host, uint32_t k, int color_space, int search, int mode, uint64_t* wrong_out,
                  uint64_t* uncertified_out, void* stream);
/* Number of kernels this library launched on behalf of ctx since creation. */
uint64_t kmg_launch_count(kmg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KMEANS_GPU_H_ */
