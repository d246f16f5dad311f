/*
 * rs_cuda.h -- the thin C-ABI between the C/C++ host (imageSynth/engine, host
 * prep) and the sm_100a kernels.  Plain pointers and sizes only; no C++ types,
 * no exceptions cross it; every function returns 0 or a non-zero code with the
 * text available from rs_cuda_last_error().
 *
 * What each entry point replaces in the reference:
 *   rs_job_run            the pass loop            lib/refiner.h:75-121, lib/refinerThreaded.h:318-358
 *     (kernel k_synth_pass)  per-target loop       lib/synthesize.h:426-642
 *                            neighbour gathering   lib/synthesize.h:189-241
 *                            candidates + probes   lib/synthesize.h:537-604, lib/engine.c:434-443
 *                            patch distance        lib/synthesize.h:266-400
 *                            commit                lib/synthesize.h:620-639
 *   rs_bestfit_batch      computeBestFit alone on caller-given patches and candidate lists
 *   rs_job_create/upload  state preparation        lib/engine.c:207-224,314-327,338-431
 *   rs_job_download       result write-back        (engine() mutates targetMap in place)
 */
#ifndef RS_CUDA_H
#define RS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RsJob RsJob; /* opaque */

typedef struct {
  int32_t tw, th, cw, ch;        /* target / corpus image dimensions */
  int32_t bpp;                   /* bytes per pixel of the raw internal pixmaps (TFormatIndices.total_bpp) */
  int32_t n_color;               /* img_match_bpp: 1 or 3; colours start at byte 1 */
  int32_t n_map;                 /* map_match_bpp: 0..3 */
  int32_t map_bip;               /* map_start_bip */
  int32_t alpha_bip;             /* byte index of alpha, or -1 */
  int32_t alpha_target;          /* isAlphaTarget */
  int32_t alpha_source;          /* isAlphaSource */
  int32_t htile, vtile;          /* wrap neighbour coordinates in x / y */
  int32_t use_context;           /* matchContextType != 0 */
  uint32_t patch_size;           /* as passed by the caller (<= 64) */
  uint32_t max_probes;
  uint32_t seed;                 /* key of the per-probe counter hash */
  uint32_t pass_end[6];          /* prefix length of each pass (lib/passes.h:67-93) */
  uint32_t n_passes;             /* <= 6 */
  double terminate_fraction;     /* stop after a pass when (float)betters/n, widened to double, is below this (0.1) */
  int32_t ordered_visits;        /* hint: the visit order is spatially sorted (matchContextType 2..8), not a shuffle:
                                    pass 0 is then a narrow dependency front and runs in latency mode throughout */
  int32_t reserved;
} RsJobDesc;

typedef struct {
  uint64_t visits, evals, evals_issued, compares, offset_scans, heur_evals, heur_skips, perfect;
  uint64_t betters[6], pass_visits[6], sum_best[6];
  uint32_t passes_run;
  uint32_t n_corpus;             /* corpus points (counted on the device when not given by the caller) */
  float ms_passes;               /* CUDA-event time of all pass kernels of the last run */
  float ms_pass[6];              /* device-clock duration of each pass that ran (first claim to last CTA out) */
  float ms_synth;                /* CUDA-event time of the pass kernels alone (ms_passes minus the pass-0 patch gather) */
  uint32_t kernel_launches;      /* kernels this job launched: upload/init, pass-0 gather, passes, write-back */
  uint32_t synth_launches_run;   /* pass-kernel launches that did work (launches of passes after the stop rule exit at once) */
} RsJobCounters;

/* The target selection (mask byte != 0) as the device sees it: count, first/last row, 128-bit digest. */
typedef struct {
  uint64_t h1, h2;
  uint32_t n, ymin, ymax, xmin, xmax, pad; /* number of target points, the rows and the columns that hold them */
} RsTargetDigest;
/* What a visit order is a function of (lib/orderTarget.h): the selection, the image size, matchContextType and the
 * seed of the ordering PRNG stream. */
typedef struct {
  uint64_t h1, h2;
  uint32_t n;
  int32_t tw, th, mode;
  uint32_t seed;
} RsOrderKey;

/* Called on the host, in order, for every (pass, index) with (index & 4095) == 0 that the device
 * has started (lib/synthesize.h:493-497).  Return non-zero to cancel the job. */
typedef int (*RsTickFn)(void *ctx, uint32_t pass, uint32_t index);

const char *rs_cuda_last_error(void);
int rs_cuda_set_device(int ordinal);
int rs_cuda_device_count(void);
/* Host cores this process may use for helper threads: hardware cores / LOCAL_WORLD_SIZE (one process per GPU under
 * torchrun), or RS_HOST_THREADS. */
unsigned rs_host_cores(void);
/* Number of jobs the caller intends to keep in flight per device (default 1).  With k > 1 every job's persistent
 * kernels take ceil(1/k) of the SMs so that k jobs run side by side instead of queueing behind each other. */
void rs_cuda_set_job_slots(int slots);

int rs_job_create(const RsJobDesc *desc, RsJob **out);
/* Jobs of one batch (any non-zero id, unique per batch call) that pass the SAME corpus pixmap to rs_job_stage share one
 * device-resident corpus per device -- canonical pixels, point list, bitmap, count -- built by the first job that needs
 * it; a second device copies it from the first (peer copy) instead of staging it from the host again.  Call before
 * rs_job_stage.  rs_cuda_drop_shared_corpora releases a batch's entries; the stats count builds / reuses / peer copies
 * of this process. */
void rs_job_share_corpus(RsJob *job, unsigned long long batch);
void rs_cuda_drop_shared_corpora(unsigned long long batch);
void rs_cuda_shared_corpus_stats(unsigned long long *builds, unsigned long long *hits, unsigned long long *peer_copies);
/* Host -> device.  target_raw/corpus_raw: tw*th*bpp and cw*ch*bpp bytes, pixel = [mask][colours][alpha?][maps].
 * targets: n points packed x | y<<16 in visit order.  corpus_points: C points packed likewise.
 * offsets: n_offsets neighbour offsets packed (int16 x | int16 y << 16), ascending distance, entry 0 = (0,0);
 *          NULL = build the reference's full table for these image sizes on the device (cached per workspace).
 * color_lut[256], map_lut[256]: metric by absolute difference; map_lut_max = mapsMetric[0]. */
int rs_job_upload(RsJob *job, const uint8_t *target_raw, const uint8_t *corpus_raw,
                  const uint32_t *targets, uint32_t n_targets,
                  const uint32_t *corpus_points, uint32_t n_corpus,
                  const uint32_t *offsets, uint32_t n_offsets,
                  const uint32_t *color_lut256, const uint32_t *map_lut256, uint32_t map_lut_max);
/* The same upload in two asynchronous phases, so that the host can compute the visit order while the images are
 * copied and prepared: (1) everything but the order -- y_min..y_max = rows of the target image that contain target
 * points; corpus_points may be NULL: the device then builds the list (row-major, mask 0xFF and not transparent);
 * (2) the visit order (n_targets points). */
int rs_job_upload_images(RsJob *job, const uint8_t *target_raw, const uint8_t *corpus_raw, uint32_t n_targets,
                         uint32_t y_min, uint32_t y_max, const uint32_t *corpus_points, uint32_t n_corpus,
                         const uint32_t *offsets, uint32_t n_offsets,
                         const uint32_t *color_lut256, const uint32_t *map_lut256, uint32_t map_lut_max);
int rs_job_upload_order(RsJob *job, const uint32_t *targets);
/* Runs all passes (early termination decided on the device), calling tick from the waiting host thread. */
/* The same upload with the target points found on the device (no per-pixel work on the host):
 *   rs_job_stage       host -> device, state init, corpus points, offsets table, selection digest; asynchronous
 *   rs_job_digest      waits for the digest only (the first thing the stream computes)
 *   rs_job_bind_order  1: a visit order for this key is cached on the device and now bound to the job; 0: miss
 *   rs_job_set_order   miss path: the points as ordered by the host (with a key they stay on the device for later
 *                      jobs; NULL: not cached)
 * rs_cuda_order_cache(0) drops the cached orders and disables the cache. */
int rs_job_stage(RsJob *job, const uint8_t *target_raw, const uint8_t *corpus_raw, const uint32_t *color_lut256,
                 const uint32_t *map_lut256, uint32_t map_lut_max);
int rs_job_digest(RsJob *job, RsTargetDigest *out);
/* rs_job_stage for the simple API (lib/imageSynth.h:31-52): ONE image of tw x th pixels, bpp - 1 channels, rows
 * img_row_bytes apart, its selection mask and optionally an explicit corpus mask (imageSynth2; NULL: the corpus is what
 * is not selected).  The internal [mask][channels] pixmaps are built on the device.  rs_job_download_simple copies
 * the rows that hold target points back into the caller's image. */
int rs_job_stage_simple(RsJob *job, const uint8_t *img, size_t img_row_bytes, const uint8_t *mask, size_t mask_row_bytes,
                        const uint8_t *mask2, size_t mask2_row_bytes, const uint32_t *color_lut256,
                        const uint32_t *map_lut256, uint32_t map_lut_max);
int rs_job_download_simple(RsJob *job, uint8_t *img, size_t img_row_bytes);
int rs_job_bind_order(RsJob *job, const RsTargetDigest *digest, const RsOrderKey *key);
int rs_job_set_order(RsJob *job, const uint32_t *ordered_points, const RsOrderKey *key);
void rs_cuda_order_cache(int enabled);
/* The shuffling orders (matchContextType 0, 1) built on the device, exactly: `draws` = the n = digest.n draws j_i of the
 * reference's loop  for i: swap(a[i], a[j_i])  (lib/orderTarget.h:38-53), made by the host's PRNG stream.  The device
 * compacts the target points, and finds what ends at each position by walking the chain of swaps backwards.  Replaces
 * rs_job_set_order on a cache miss; ordered_out (or NULL) receives the order. */
int rs_job_shuffle_order(RsJob *job, const uint32_t *draws, const RsOrderKey *key, uint32_t *ordered_out);
/* The same from the RAW 32-bit words of the PRNG stream: rs_job_raw_buffer returns a pinned buffer for n_words words
 * that the caller's producer fills (it can start before the number of target points is known);
 * rs_job_shuffle_order_raw sends the first n_raw of them up and the device applies g_rand_int_range's rejection rule
 * and modulo itself.  n_raw must leave room for the rejections (n / 2^32 per word); if fewer than digest.n words
 * survive, the job is flagged faulty and rs_job_run returns an error. */
uint32_t *rs_job_raw_buffer(RsJob *job, size_t n_words);
int rs_job_shuffle_order_raw(RsJob *job, uint32_t n_raw, const RsOrderKey *key, uint32_t *ordered_out);
/* The same with the PRNG stream made on the device as well: the raw words of GLib's GRand (MT19937, g_rand_new_with_seed /
 * g_rand_int; the reference seeds it in lib/engine.c and draws in lib/orderTarget.h:38-53) come from one CTA beside the
 * image upload; nothing of the order is computed on or copied from the host.  The whole pipeline of the three
 * rs_job_shuffle_order* calls runs on the job's side stream from the moment the selection is on the device.
 * rs_cuda_mt19937_raw (test entry): the first n_words words of the stream of `seed`, made on the device, to the host. */
int rs_job_shuffle_order_seed(RsJob *job, uint32_t seed, const RsOrderKey *key, uint32_t *ordered_out);
int rs_cuda_mt19937_raw(uint32_t seed, uint32_t n_words, uint32_t *out_host);
/* Stable ascending radix sort of n (key, value) pairs on the low key_bits bits of the keys, host buffers in place, on a
 * side stream of the job (it runs beside the staging).  The sort step of the target orderings 2-8. */
int rs_job_sort_pairs(RsJob *job, uint32_t *keys, uint32_t *vals, uint32_t n, int key_bits);
/* Pass schedule (prefix length of each pass), when it was not known at rs_job_create (target points counted on the
 * device). */
void rs_job_set_passes(RsJob *job, const uint32_t *pass_end, uint32_t n_passes);

int rs_job_run(RsJob *job, RsTickFn tick, void *tick_ctx);
/* target_raw_out: the caller's tw*th*bpp pixmap; the rows containing target points are overwritten with the
 * device's copy, which differs from the uploaded one only in the colour bytes of target points.
 * sources_out (may be NULL; needs rs_job_want_sources(job,1) before rs_job_run): packed best corpus point of
 * each target point in visit order, 0xFFFFFFFF if none. */
int rs_job_download(RsJob *job, uint8_t *target_raw_out, uint32_t *sources_out);
void rs_job_want_sources(RsJob *job, int yes);
/* Page-locked caller buffers (cudaHostAlloc / cudaHostRegister; a pinned torch tensor) are copied from and to directly,
 * without the staging copy through the workspace's pinned memory (the reference's callers pass malloc'ed ImageBuffers,
 * lib/imageBuffer.h: those are staged).  rs_cuda_host_is_pinned: 1 if `p` is such memory.  rs_job_result_direct(job, 1)
 * before rs_job_run: the result rows are not staged; rs_job_download / rs_job_download_simple then copy them from the
 * device into the (page-locked) destination they are given. */
int rs_cuda_host_is_pinned(const void *p);
void rs_job_result_direct(RsJob *job, int yes);
int rs_job_counters(RsJob *job, RsJobCounters *out);
void rs_job_destroy(RsJob *job);
/* Throughput profile of a pass of the last run: out_ns[i] = ns from the start of the pass to the claim of visit
 * 4096 * i.  Returns the number of entries written (0 if the pass did not run). */
uint32_t rs_job_timeline(RsJob *job, uint32_t pass, uint64_t *out_ns, uint32_t cap);
/* Copies the job's neighbour-offset table back (parity tests of the device-side build). */
int rs_job_read_offsets(RsJob *job, uint32_t *out, uint32_t cap);
/* Frees the pooled workspaces (device buffers, pinned staging, streams) that jobs leave behind for reuse. */
void rs_cuda_release_cached(void);

/* computeBestFit over explicit inputs, for bit-exact kernel tests (lib/synthesize.h:266-400).
 * corpus_raw: cw*ch*bpp bytes.  For visit v in [0,n_visits): patch entries [nb_begin[v], nb_begin[v+1]) of
 * nb_offsets (packed int16 pairs) / nb_pixels (8 raw pixel bytes each); candidates
 * [cand_begin[v], cand_begin[v+1]) of cands (packed x|y<<16).  Outputs per visit: best sum (0xFFFFFFFF if no
 * candidate), index of the winning candidate within the visit's list (-1 if none). */
/* Micro-benchmark: sustained rate of independent, uniformly random aligned loads of elem_bytes (4 or 8) from a
 * device buffer of buffer_bytes, in loads per second (best of `repeats`, CUDA events).  The distance loop does one
 * such load per neighbour compare, so this is its practical ceiling on this GPU. */
int rs_cuda_gather_rate(size_t buffer_bytes, int elem_bytes, int repeats, double *loads_per_s);

/* The launch plan of one pass (host logic only, no device needed): a pass runs as up to four persistent launches over
 * consecutive segments of the visit order; widths4[k] = warps per visit of segment k (8, 4, 2: latency kernel
 * k_synth_pass_team; 1: throughput kernel k_synth_pass), ends4[k] = its end (exclusive; the last one is pass_end).
 * ordered_visits as in RsJobDesc.  Returns the number of segments. */
int rs_cuda_plan_pass(uint32_t n_targets, uint32_t pass_end, int ordered_visits, int patch_size, uint32_t pass,
                      uint32_t *ends4, uint32_t *widths4);

int rs_bestfit_batch(const RsJobDesc *desc, const uint8_t *corpus_raw,
                     const uint32_t *color_lut256, const uint32_t *map_lut256, uint32_t map_lut_max,
                     uint32_t n_visits, const uint32_t *nb_begin, const uint32_t *nb_offsets,
                     const uint8_t *nb_pixels, const uint32_t *cand_begin, const uint32_t *cands,
                     uint32_t *best_sum_out, int32_t *best_index_out);

#ifdef __cplusplus
}
#endif
#endif
