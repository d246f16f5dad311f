/*
 * resynthesizer.h -- the drop-in boundary of libresynthesizer_b200.so.
 *
 * These are the entry points and ABI types a caller of the reference library
 * binds.  Layouts are binary-compatible with the
 * reference so that its callers (src/testSynth.c:77,168; the GIMP engine
 * plug-in src/resynthesizer/resynthesizer.c:505) relink unchanged.  Behind
 * them the synthesis passes run as sm_100a CUDA kernels (include/rs_cuda.h);
 * there is no CPU fallback: without a usable CUDA device every synthesis call
 * returns RS_ERROR_CUDA.
 */
#ifndef RESYNTHESIZER_B200_H
#define RESYNTHESIZER_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* replaces lib/imageBuffer.h:12-19 -- interleaved pixels, rows padded to rowBytes */
typedef struct _ImageBuffer {
  unsigned char *data;
  unsigned int width;
  unsigned int height;
  size_t rowBytes;
} ImageBuffer;

/* replaces lib/imageFormat.h:36-42 */
typedef enum ImageFormat { T_RGB, T_RGBA, T_Gray, T_GrayA } TImageFormat;

/* replaces lib/engineParams.h:13-28 */
typedef enum ImageSynthError {
  IMAGE_SYNTH_SUCCESS,
  IMAGE_SYNTH_ERROR_INVALID_IMAGE_FORMAT,
  IMAGE_SYNTH_ERROR_IMAGE_MASK_MISMATCH,
  IMAGE_SYNTH_ERROR_PATCH_SIZE_EXCEEDED,
  IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE,
  IMAGE_SYNTH_ERROR_EMPTY_TARGET,
  IMAGE_SYNTH_ERROR_EMPTY_CORPUS
} TImageSynthError;

/* Outside the reference's enum range: the CUDA layer failed (no device, out of
 * memory, launch error).  rs_last_error() holds the text.  (SURVEY.md section 5:
 * "map CUDA failure to a non-zero code outside the enum range".) */
#define RS_ERROR_CUDA 100

/* replaces lib/engineParams.h:31-86 (field order and types fixed by the ABI) */
typedef struct ImageSynthParametersStruct {
  int isMakeSeamlesslyTileableHorizontally;
  int isMakeSeamlesslyTileableVertically;
  int matchContextType;          /* 0..8, selects orderTarget mode (lib/orderTarget.h:268-343) */
  double mapWeight;
  double sensitivityToOutliers;
  unsigned int patchSize;        /* <= 64 */
  unsigned int maxProbeCount;
} TImageSynthParameters;

/* replaces lib/imageFormatIndicies.h:46-58 -- byte positions inside the internal pixel
 * [mask][colour x1..3][alpha?][map x0..3] */
typedef unsigned char TPixelelIndex;
typedef struct indicesStruct {
  TPixelelIndex colorEndBip;
  TPixelelIndex alpha_bip;
  TPixelelIndex map_start_bip;
  TPixelelIndex map_end_bip;
  TPixelelIndex img_match_bpp;
  TPixelelIndex map_match_bpp;
  TPixelelIndex total_bpp;
  int isAlphaTarget;
  int isAlphaSource;
} TFormatIndices;

/* replaces lib/glibProxy.h:86-92 / GLib's public GArray head; lib/map.h:28-46 */
#ifndef RS_NO_GLIB_NAMES
typedef struct _GArray { char *data; unsigned int len; } GArray;
#endif
typedef struct {
  unsigned int width;
  unsigned int height;
  unsigned int depth;
  GArray *data;
} Map;
typedef struct { int x; int y; } Coordinates;
typedef unsigned char Pixelel;

/* ---- simple API: replaces lib/imageSynth.h:31-52 (impl lib/imageSynth.c:61-216) ---- */
int imageSynth(ImageBuffer *imageBuffer, ImageBuffer *mask, TImageFormat imageFormat,
               TImageSynthParameters *parameters, /* NULL -> defaults */
               void (*progressCallback)(int, void *), void *contextInfo, int *cancelFlag);
int imageSynth2(ImageBuffer *imageBuffer, ImageBuffer *mask, ImageBuffer *mask2, TImageFormat imageFormat,
                TImageSynthParameters *parameters,
                void (*progressCallback)(int, void *), void *contextInfo, int *cancelFlag);

/* ---- full API: replaces lib/engine.h:3-12 (impl lib/engine.c:539-690) ---- */
int engine(TImageSynthParameters parameters, TFormatIndices *indices, Map *targetMap, Map *corpusMap,
           void (*progressCallback)(int, void *), void *contextInfo, int *cancelFlag);

/* replaces lib/engineParams.h:91-94 */
void setDefaultParams(TImageSynthParameters *param);

/* replaces lib/imageFormatIndicies.h:60-84 */
unsigned int countPixelelsPerPixelForFormat(TImageFormat format);
int prepareImageFormatIndicesFromFormatType(TFormatIndices *indices, TImageFormat format);
void prepareImageFormatIndices(TFormatIndices *indices, unsigned int count_color_channels_target,
                               unsigned int count_color_channels_map, int is_alpha_target,
                               int is_alpha_source, int isMap);
void prepareDefaultFormatIndices(TFormatIndices *formatIndices);

/* replaces lib/map.h:49-99 (used by the GIMP adapter, src/resynthesizer/adaptGimp.h:190-276) */
void free_map(Map *map);
void new_pixmap(Map *map, unsigned int width, unsigned int height, unsigned int depth);
void new_bytemap(Map *map, unsigned int width, unsigned int height);
void new_intmap(Map *map, unsigned int width, unsigned int height);
void new_coordmap(Map *map, unsigned int width, unsigned int height);
void set_bytemap(Map *map, unsigned char value);
void invert_bytemap(Map *map);
void interleave_mask(Map *pixmap, Map *mask);

/* ---- additions (not in the reference) ---- */
/* Text of the last CUDA-layer failure on this thread ("" if none). */
const char *rs_last_error(void);
/* Select the CUDA device subsequent calls on this thread use (default: RESYNTH_CUDA_DEVICE or 0). */
int rs_set_device(int ordinal);

/* Counters of the most recent engine()/imageSynth() call on this thread. */
typedef struct {
  unsigned long long visits, evals, evals_issued, compares, offset_scans, heur_evals, heur_skips, perfect;
  unsigned long long betters[6], pass_visits[6], sum_best[6];
  unsigned int passes_run, n_targets, n_corpus;
  float ms_prep, ms_h2d, ms_kernels, ms_d2h, ms_total; /* host prep / copies / device passes (CUDA events) */
  float ms_pass[6];                                    /* device-clock duration of each pass that ran */
  float ms_synth;                                      /* CUDA-event time of the pass kernels alone */
  unsigned int kernel_launches, synth_launches_run;    /* kernels launched by the job; pass launches that did work */
  unsigned int order_cache_hit;                        /* 1: the visit order came from the device-side cache */
} RsStats;
/* The visit order of the target points is a pure function of (selection, image size, matchContextType, seed) and is
 * kept on the device for later jobs with the same key (16 entries / 1 GiB, least recently used out first).
 * rs_order_cache(0) drops the entries and disables the cache, rs_order_cache(1) enables it (default). */
void rs_order_cache(int enabled);
/* Orderings 2-8 (matchContextType) sort the target points by a geometric key: the keys are computed on the host (the
 * brushfire ray index needs libm's atan2 bit for bit), the pairs are sorted on the device when there are at least
 * n_points of them (default 65536; below, the host's radix sort is quicker than the copies).  Same result either way. */
void rs_set_device_sort_min(unsigned int n_points);
/* Orderings 0 and 1 are the reference's loop  for i: swap(a[i], a[rand(0,n)])  (lib/orderTarget.h:38-53).  The host makes
 * the draws with the reference's PRNG stream; from n_points points on (default 32768) the device resolves the chain of
 * swaps itself -- exactly, by walking it backwards for every position -- instead of the host swapping and uploading. */
void rs_set_device_shuffle_min(unsigned int n_points);
/* Throughput profile of the last engine() call on this thread (needs rs_keep_result(1)): ns from the start of
 * `pass` to the claim of its visit 4096 * i.  Returns the number of entries written. */
unsigned int rs_get_timeline(unsigned int pass, unsigned long long *out_ns, unsigned int cap);
void rs_get_stats(RsStats *out);
/* CUDA kernels launched by all engine() calls of this process so far (any thread). */
unsigned long long rs_total_kernel_launches(void);
/* Batch of independent jobs (the reference has no such call; its users loop over engine()).  Runs `n_jobs`
 * engine() calls on `slots` host threads that share the current CUDA device; each job's kernels take 1/slots of
 * the SMs.  Small jobs are latency-bound on their dependency chains, so running several side by side multiplies
 * throughput; `slots` is an upper bound: jobs of 16 k+ target points run at most 4 at a time, 200 k+ at most 2 (they
 * fill the GPU on their own; the rest only overlaps host work and copies with kernels).  errors_out[i] receives engine()'s return value for job i.  Progress/cancel are not forwarded.
 * Returns 0 if every job returned 0, else the first non-zero code. */
int rs_engine_batch(int n_jobs, const TImageSynthParameters *params, TFormatIndices *const *indices,
                    Map *const *targetMaps, Map *const *corpusMaps, int slots, int *errors_out);
/* The same batch dealt over several GPUs of the box from ONE process (a job never shards, a batch does; no collective):
 * `devices[0..n_devices)` are CUDA ordinals (n_devices 0 / devices NULL: the calling thread's device).  Every device
 * gets `slots` host threads; all of them pull the next job from one queue, longest estimated job first, so a device
 * that finishes early takes more.  What the reference's callers do with a loop over engine()
 * (PluginScripts/plugin-heal-selection.py:148 per image) becomes one call.  rs_last_error() holds the text of the first
 * failed job. */
int rs_engine_batch_multi(int n_jobs, const TImageSynthParameters *params, TFormatIndices *const *indices,
                          Map *const *targetMaps, Map *const *corpusMaps, int n_devices, const int *devices, int slots,
                          int *errors_out);
/* Jobs of one rs_engine_batch(_multi) call that pass the same corpusMap share it on the device: the corpus is staged,
 * canonicalised and indexed once per device (and copied device to device for a second GPU) -- one texture or style
 * source, many targets.  Counters of this process: corpora built, reuses, peer copies. */
void rs_shared_corpus_stats(unsigned long long *builds, unsigned long long *hits, unsigned long long *peer_copies);
/* The batch call for simple-API jobs: imageSynth(images[i], masks[i], format, params, ...) for every i, or imageSynth2
 * where masks2 (may be NULL) holds an explicit corpus mask for that job; params NULL = defaults.  Images change in place
 * exactly as imageSynth() changes them.  Same dealing as rs_engine_batch_multi. */
int rs_image_synth_batch(int n_jobs, ImageBuffer *const *images, ImageBuffer *const *masks, ImageBuffer *const *masks2,
                         TImageFormat format, const TImageSynthParameters *params, int n_devices, const int *devices,
                         int slots, int *errors_out);
/* rs_keep_result(1): engine() calls on this thread also fetch what rs_get_last_result() returns (off by default). */
void rs_keep_result(int yes);
/* Visit order and final source (best corpus point) of each target point of the last engine() call on this
 * thread, packed x | y << 16 (0xFFFFFFFF = no source).  Returns the number of target points. */
unsigned int rs_get_last_result(unsigned int *targets_out, unsigned int *sources_out, unsigned int cap);
/* Seed of the per-probe counter hash (default 1198472, the reference's PRNG seed, lib/engine.c:643). */
void rs_set_seed(unsigned int seed);

#ifdef __cplusplus
}
#endif
#endif
