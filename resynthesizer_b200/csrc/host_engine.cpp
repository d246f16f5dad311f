// The reference-facing host layer: imageSynth()/imageSynth2()/engine() with the reference's signatures,
// error codes, progress and cancel behaviour (include/resynthesizer.h), driving the CUDA passes through
// include/rs_cuda.h.  There is no CPU synthesis path in this library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/resynthesizer.h"
#include "../../include/rs_cuda.h"
#include "host_prep.h"

namespace {

thread_local std::string t_err;
thread_local RsStats t_stats;
thread_local uint32_t t_seed = 1198472u;  // lib/engine.c:643
thread_local bool t_device_chosen = false;
thread_local int t_device = -1;  // ordinal chosen by rs_set_device / ensure_device on this thread
thread_local bool t_quiet = false;  // batch workers: no progress callback to serve, rs_job_run may sleep instead of polling
thread_local unsigned long long t_batch = 0;  // != 0 inside a batch call: jobs naming the same corpus pixmap share it on the device
thread_local bool t_keep_result = false;  // rs_keep_result(): also fetch per-target sources (tests, quality metrics)
thread_local std::vector<uint32_t> t_last_sources, t_last_targets;  // of the last engine() call, visit order
thread_local std::vector<uint64_t> t_timeline[6];                   // of the last engine() call (rs_keep_result)

// point lists shorter than this are sorted on the host (orderings 2-8): a device sort costs two small copies and a sync
std::atomic<size_t> g_device_sort_min{std::getenv("RS_DEVICE_SORT_MIN") ? (size_t)std::atoll(std::getenv("RS_DEVICE_SORT_MIN"))
                                                                        : ((size_t)1 << 16)};
// shuffling orders (matchContextType 0, 1) of at least this many points are resolved on the device
std::atomic<size_t> g_device_shuffle_min{(size_t)1 << 15};
std::atomic<unsigned long long> g_kernel_launches{0};  // every kernel any engine() call of this process launched

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

extern "C" const char *rs_cuda_peek_error(void);
void dbg(const char *where) {
  static const bool on = std::getenv("RS_DEBUG") != nullptr;
  if (!on) return;
  if (const char *e = rs_cuda_peek_error()) std::fprintf(stderr, "[rs debug] pending CUDA error at %s: %s\n", where, e);
}

int ensure_device() {
  if (t_device_chosen) return 0;
  int ordinal = 0;
  if (const char *e = std::getenv("RESYNTH_CUDA_DEVICE")) ordinal = std::atoi(e);
  if (rs_cuda_device_count() <= 0) { t_err = "no CUDA device available (this library has no CPU path)"; return RS_ERROR_CUDA; }
  if (rs_cuda_set_device(ordinal)) { t_err = rs_cuda_last_error(); return RS_ERROR_CUDA; }
  t_device_chosen = true;
  t_device = ordinal;
  return 0;
}

// Progress bookkeeping of lib/progress.c:53-65 driven by device ticks.
struct TickState {
  void (*cb)(int, void *);
  void *ctx;
  int *cancel;
  uint32_t completed, estimated, prior_percent;
};

int on_tick(void *p, uint32_t /*pass*/, uint32_t /*index*/) {
  TickState *t = static_cast<TickState *>(p);
  t->completed += 4095u;  // IMAGE_SYNTH_CALLBACK_COUNT
  const uint32_t percent = (uint32_t)(((float)t->completed / t->estimated) * 100);
  if (percent > t->prior_percent) {
    t->cb((int)percent, t->ctx);
    t->prior_percent = percent;
  }
  return *t->cancel ? 1 : 0;  // polled right after the callback (lib/synthesize.h:493-497)
}

}  // namespace

extern "C" const char *rs_last_error(void) { return t_err.c_str(); }
extern "C" void rs_get_stats(RsStats *out) { *out = t_stats; }
extern "C" void rs_set_seed(unsigned int seed) { t_seed = seed; }
extern "C" void rs_order_cache(int enabled) { rs_cuda_order_cache(enabled); }
extern "C" void rs_set_device_sort_min(unsigned int n_points) { g_device_sort_min.store(n_points); }
extern "C" void rs_set_device_shuffle_min(unsigned int n_points) { g_device_shuffle_min.store(n_points); }
extern "C" unsigned long long rs_total_kernel_launches(void) { return g_kernel_launches.load(); }
extern "C" void rs_keep_result(int yes) { t_keep_result = yes != 0; }
extern "C" int rs_set_device(int ordinal) {
  if (rs_cuda_set_device(ordinal)) { t_err = rs_cuda_last_error(); return RS_ERROR_CUDA; }
  t_device_chosen = true;
  t_device = ordinal;
  return 0;
}

// replaces lib/engineParams.c:8-19
extern "C" void setDefaultParams(TImageSynthParameters *p) {
  p->isMakeSeamlesslyTileableHorizontally = 0;
  p->isMakeSeamlesslyTileableVertically = 0;
  p->matchContextType = 1;
  p->mapWeight = 0.5;
  p->sensitivityToOutliers = 0.117;
  p->patchSize = 30;
  p->maxProbeCount = 200;
}

// replaces lib/imageFormat.c:39-56
extern "C" unsigned int countPixelelsPerPixelForFormat(TImageFormat f) {
  switch (f) {
    case T_RGB: return 3;
    case T_RGBA: return 4;
    case T_Gray: return 1;
    case T_GrayA: return 2;
    default: return 0;
  }
}

// replaces lib/imageFormat.c:116-207: [mask][colours][alpha if either image has one][map channels]
extern "C" void prepareImageFormatIndices(TFormatIndices *o, unsigned int n_color, unsigned int n_map, int alpha_target,
                                          int alpha_source, int is_map) {
  o->img_match_bpp = (TPixelelIndex)n_color;
  o->colorEndBip = (TPixelelIndex)(1 + n_color);
  if (alpha_target || alpha_source) {
    o->alpha_bip = o->colorEndBip;
    o->map_start_bip = (TPixelelIndex)(o->colorEndBip + 1);
  } else {
    o->map_start_bip = o->colorEndBip;  // alpha_bip stays undefined, as in the reference
  }
  o->map_match_bpp = (TPixelelIndex)(is_map ? n_map : 0);
  o->map_end_bip = (TPixelelIndex)(o->map_start_bip + o->map_match_bpp);
  o->total_bpp = o->map_end_bip;
  o->isAlphaTarget = alpha_target;
  o->isAlphaSource = alpha_source;
}

// replaces lib/imageFormat.c:65-113
extern "C" int prepareImageFormatIndicesFromFormatType(TFormatIndices *o, TImageFormat f) {
  switch (f) {
    case T_RGB: prepareImageFormatIndices(o, 3, 0, 0, 0, 0); return 0;
    case T_RGBA: prepareImageFormatIndices(o, 3, 0, 1, 1, 0); return 0;
    case T_Gray: prepareImageFormatIndices(o, 1, 0, 0, 0, 0); return 0;
    case T_GrayA: prepareImageFormatIndices(o, 1, 0, 1, 1, 0); return 0;
    default: return IMAGE_SYNTH_ERROR_INVALID_IMAGE_FORMAT;
  }
}

// replaces lib/imageFormat.c:210-241 (test default: M R G B A)
extern "C" void prepareDefaultFormatIndices(TFormatIndices *o) {
  prepareImageFormatIndices(o, 3, 0, 1, 1, 0);
}

// ------------------------------------------------------------------------------------------ the job, host side
namespace {

// Where the pixels of a job are.  Full API: two caller-owned internal pixmaps [mask][colours][alpha?][maps].
// Simple API: ONE image + selection mask (+ explicit corpus mask); nothing is repacked on the host -- the planes go to
// the device as they are and the pixmaps are built there (lib/imageSynth.c:61-137 does it with per-pixel loops).
struct PixelSource {
  int tw = 0, th = 0, cw = 0, ch = 0, bpp = 0;
  uint8_t *tpix = nullptr;         // full API
  const uint8_t *cpix = nullptr;
  ImageBuffer *img = nullptr, *mask = nullptr, *mask2 = nullptr;  // simple API
  bool simple() const { return img != nullptr; }
};

// Any non-zero byte in p[0..n): 64 bytes per step (a byte loop with an early exit costs a millisecond on the two
// megabytes of mask above a centred hole in a 2048x2048 image -- more than a pass of the synthesis itself).
bool any_nonzero(const uint8_t *p, size_t n) {
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    uint64_t a[8];
    std::memcpy(a, p + i, 64);
    if (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) return true;
  }
  for (; i < n; i++) if (p[i]) return true;
  return false;
}
bool any_target(const PixelSource &s) {
  if (!s.simple()) return rs::has_target_point(s.tpix, s.tw, s.th, s.bpp);
  for (int y = 0; y < s.th; y++)
    if (any_nonzero(s.mask->data + (size_t)y * s.mask->rowBytes, (size_t)s.tw)) return true;
  return false;
}
bool any_corpus(const PixelSource &s, const TFormatIndices &fi) {
  if (!s.simple()) return rs::has_corpus_point(s.cpix, s.cw, s.ch, s.bpp, fi);
  const int nc = s.bpp - 1;
  const bool alpha = fi.isAlphaSource != 0;
  for (int y = 0; y < s.th; y++) {
    const uint8_t *mrow = s.mask2 ? s.mask2->data + (size_t)y * s.mask2->rowBytes : s.mask->data + (size_t)y * s.mask->rowBytes;
    const uint8_t *irow = s.img->data + (size_t)y * s.img->rowBytes;
    for (int x = 0; x < s.tw; x++) {
      const bool selected = s.mask2 ? mrow[x] == 0xFF : mrow[x] == 0;  // inverted selection mask, or the explicit one
      if (selected && (!alpha || irow[(size_t)x * nc + (fi.alpha_bip - 1)] != 0)) return true;
    }
  }
  return false;
}

// Everything engine() and imageSynth() share.  write_back: results go to the caller's buffers (imageSynth skips that
// when cancelled, lib/imageSynth.c:108).
int synth_core(TImageSynthParameters prm, TFormatIndices *fi, const PixelSource &src,
               void (*progressCallback)(int, void *), void *contextInfo, int *cancelFlag, bool skip_write_back_if_cancelled) {
  const double t0 = now_ms();
  std::memset(&t_stats, 0, sizeof t_stats);
  t_err.clear();
  if (prm.patchSize > 64) return IMAGE_SYNTH_ERROR_PATCH_SIZE_EXCEEDED;  // lib/engine.c:591
  const int tw = src.tw, th = src.th, cw = src.cw, ch = src.ch, bpp = src.bpp;

  // Empty target / corpus and the context-type range, in the reference's order (lib/engine.c:605-610, 620-627, 645-647).
  // Without a device they are detected on the host before CUDA is touched.  With one, the empty-target scan -- megabytes of
  // mask above a centred hole, 0.3 ms of a 50 ms call -- is left to the device: the selection digest that every job waits
  // for counts the target points anyway (an empty target then costs a staged job instead of a scan: the error path).
  // The host scans only where the answer decides which error comes first.
  const bool ctx_ok = prm.matchContextType >= 0 && prm.matchContextType <= 8;
  const bool count_on_device = ctx_ok && !std::getenv("RS_HOST_TARGET_SCAN") && (t_device_chosen || rs_cuda_device_count() > 0);
  if (!count_on_device && !any_target(src)) return IMAGE_SYNTH_ERROR_EMPTY_TARGET;
  if (!any_corpus(src, *fi)) return (count_on_device && !any_target(src)) ? IMAGE_SYNTH_ERROR_EMPTY_TARGET : IMAGE_SYNTH_ERROR_EMPTY_CORPUS;
  if (!ctx_ok) return IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE;  // orderTargetPoints' default case

  if (tw > 32767 || th > 32767 || cw > 32767 || ch > 32767) {
    t_err = "image dimensions above 32767 are not supported by the packed device layout";
    return RS_ERROR_CUDA;
  }
  if (int e = ensure_device()) return e;

  uint16_t c512[512];
  uint32_t m512[512];
  rs::build_metric_tables(prm.sensitivityToOutliers, prm.mapWeight, c512, m512);
  // metric by |difference| (both functions are even: matchWeighting.h:56-58,176)
  uint32_t c256[256], m256[256];
  for (int d = 0; d < 256; d++) { c256[d] = c512[256 + d]; m256[d] = m512[256 + d]; }

  RsJobDesc desc;
  std::memset(&desc, 0, sizeof desc);
  desc.tw = tw; desc.th = th; desc.cw = cw; desc.ch = ch; desc.bpp = bpp;
  desc.n_color = fi->img_match_bpp; desc.n_map = fi->map_match_bpp; desc.map_bip = fi->map_start_bip;
  desc.alpha_bip = (fi->isAlphaTarget || fi->isAlphaSource) ? fi->alpha_bip : -1;
  desc.alpha_target = fi->isAlphaTarget ? 1 : 0;
  desc.alpha_source = fi->isAlphaSource ? 1 : 0;
  desc.htile = prm.isMakeSeamlesslyTileableHorizontally ? 1 : 0;
  desc.vtile = prm.isMakeSeamlesslyTileableVertically ? 1 : 0;
  desc.use_context = prm.matchContextType != 0;
  desc.patch_size = prm.patchSize; desc.max_probes = prm.maxProbeCount; desc.seed = t_seed;
  desc.n_passes = 6;
  desc.ordered_visits = prm.matchContextType >= 2 ? 1 : 0;
  desc.terminate_fraction = 0.1;  // IMAGE_SYNTH_TERMINATE_FRACTION, a double (lib/refiner.h:111)
  const double t1 = now_ms();

  // Stage the images (asynchronous: host -> device, state init, corpus points, offsets table) and get the digest of
  // the target selection back: number of target points, their rows, and the key of the visit-order cache.
  // The shuffling orders need one PRNG draw per target point.  By default the device makes the stream itself
  // (rs_job_shuffle_order_seed).  RS_HOST_PRNG=1 keeps the host's producer: the raw words of the stream depend on the seed
  // only, so a producer thread starts making them now, while the images are staged and counted (rejection rate < n / 2^32).
  RsJob *job = nullptr;
  dbg("before create");
  if (rs_job_create(&desc, &job)) { t_err = rs_cuda_last_error(); return RS_ERROR_CUDA; }
  if (t_batch && !src.simple()) rs_job_share_corpus(job, t_batch);
  dbg("after create");
  const size_t npx = (size_t)tw * th, raw_cap = npx + npx / 32 + 65536;
  std::unique_ptr<rs::RawStream> raw;
  bool raw_pinned = false;
  const bool host_prng = std::getenv("RS_HOST_PRNG") != nullptr || std::getenv("RS_NO_RAW_STREAM") != nullptr;
  if (host_prng && prm.matchContextType <= 1 && npx >= g_device_shuffle_min.load() * 4 && npx <= ((size_t)1 << 26) &&
      !std::getenv("RS_NO_RAW_STREAM")) {  // (the switch lets the tests take the host-reduced path below)
    uint32_t *pinned = rs_job_raw_buffer(job, raw_cap);  // the producer writes where the H2D copy will read
    raw_pinned = pinned != nullptr;
    raw.reset(new rs::RawStream(t_seed, raw_cap, pinned));
  }
  RsTargetDigest dg;
  const char *host_fault = nullptr;  // a failure detected by this layer (not by the CUDA layer)
  int rc = src.simple()
               ? rs_job_stage_simple(job, src.img->data, src.img->rowBytes, src.mask->data, src.mask->rowBytes,
                                     src.mask2 ? src.mask2->data : nullptr, src.mask2 ? src.mask2->rowBytes : 0, c256, m256, m512[0])
               : rs_job_stage(job, src.tpix, src.cpix, c256, m256, m512[0]);
  if (!rc) rc = rs_job_digest(job, &dg);
  if (rc) {
    t_err = rs_cuda_last_error();
    raw.reset();  // the producer writes into the job's pinned buffer: stop it before the workspace goes back to the pool
    rs_job_destroy(job);
    return RS_ERROR_CUDA;
  }
  dbg("after stage+digest");
  if (dg.n == 0) {  // (count_on_device: no host scan was made)
    raw.reset();
    rs_job_destroy(job);
    return IMAGE_SYNTH_ERROR_EMPTY_TARGET;
  }
  const uint32_t n = dg.n;
  uint32_t pass_end[6];
  const uint32_t estimated = rs::pass_schedule(n, pass_end);
  rs_job_set_passes(job, pass_end, 6);
  const double t2 = now_ms();
  // The visit order (lib/engine.c:643-645) is a pure function of the selection, the image size, the context type and
  // the seed: jobs that share them (a batch, the frames of a clip) order their points once; the list stays on the device.
  RsOrderKey key;
  std::memset(&key, 0, sizeof key);
  key.h1 = dg.h1; key.h2 = dg.h2; key.n = n; key.tw = tw; key.th = th; key.mode = prm.matchContextType; key.seed = t_seed;
  static thread_local std::vector<uint32_t> targets;  // reused across calls: no page faults per job
  int hit = t_keep_result ? 0 : rs_job_bind_order(job, &dg, &key);
  dbg("after bind");
  if (hit == 0) {
    if (t_keep_result) hit = rs_job_bind_order(job, &dg, nullptr) == 100 ? 100 : 0;  // sizes only; the order is wanted on the host
    if (hit == 0 && prm.matchContextType <= 1 && n >= g_device_shuffle_min.load()) {
      // miss, shuffling order: the device compacts the target points and resolves the chain of swaps from the draws of
      // the reference's PRNG stream (rs_job_shuffle_order*)
      static thread_local std::vector<uint32_t> draws;
      uint32_t *ordered = nullptr;
      if (t_keep_result) { targets.resize(n); ordered = targets.data(); }
      const RsOrderKey *ckey = t_keep_result ? nullptr : &key;
      if (!host_prng) {
        // the PRNG stream, the rejection rule, the modulo and the chain of swaps all on the device
        rc = rs_job_shuffle_order_seed(job, t_seed, ckey, ordered);
      } else if (raw && raw_pinned) {
        // raw words straight from the producer's pinned buffer; rejection rule and modulo applied on the device
        const size_t n_raw = std::min(raw_cap, (size_t)n + n / 32 + 65536);
        raw->wait_ready(n_raw);
        rc = rs_job_shuffle_order_raw(job, (uint32_t)n_raw, ckey, ordered);
      } else {
        draws.resize(n);
        if (raw) {
          raw->reduce(n, draws.data(), n);
        } else {
          rs::GRandMT prng(t_seed);
          prng.fill_int_range(n, draws.data(), n);
        }
        rc = rs_job_shuffle_order(job, draws.data(), ckey, ordered);
      }
      hit = 2;
    }
    if (hit == 0) {  // miss: collect and order the points on the host (the reference's PRNG stream) while the device stages
      const uint8_t *mask0 = src.simple() ? src.mask->data : src.tpix;
      const size_t pstride = src.simple() ? 1 : (size_t)bpp, rstride = src.simple() ? src.mask->rowBytes : (size_t)tw * bpp;
      // orderings 2-8 sort the points by a geometric key: the host computes the keys, the device sorts the pairs
      const size_t device_sort_min = g_device_sort_min.load();
      const rs::PairSorter sorter = [job, device_sort_min](uint32_t *keys, uint32_t *vals, size_t cnt) {
        if (cnt < device_sort_min) return false;
        dbg("before sort");
        if (rs_job_sort_pairs(job, keys, vals, (uint32_t)cnt, 32) == 0) { dbg("after sort"); return true; }
        if (std::getenv("RS_DEBUG")) std::fprintf(stderr, "rs_job_sort_pairs failed: %s\n", rs_cuda_last_error());
        return false;
      };
      if (rs::collect_and_order(prm.matchContextType, mask0, tw, th, pstride, rstride, n, t_seed, targets, &sorter) != 0 || targets.size() != n) {
        host_fault = "target point count differs between host and device";
        rc = 100;
      }
      if (!rc) rc = rs_job_set_order(job, targets.data(), t_keep_result ? nullptr : &key);
    }
  }
  raw.reset();
  if (hit == 100) rc = 100;
  dbg("after order");
  const double t2b = now_ms();
  TickState ts{progressCallback, contextInfo, cancelFlag, 0u, estimated, 0u};
  // a page-locked destination gets its rows straight from the device (decided before the run: the staged copy is skipped)
  const bool direct = !t_keep_result && rs_cuda_host_is_pinned(src.simple() ? (const void *)src.img->data : (const void *)src.tpix) != 0;
  if (!rc) { rs_job_want_sources(job, t_keep_result ? 1 : 0); rs_job_result_direct(job, direct ? 1 : 0); rc = rs_job_run(job, t_quiet ? nullptr : on_tick, &ts); }
  const double t3 = now_ms();
  if (!rc) {
    // engine() mutates the colour bytes of targetMap in place (lib/synthesize.h:403-419); alpha and maps untouched
    const bool write_back = !(skip_write_back_if_cancelled && *cancelFlag);
    if (t_keep_result) {
      t_last_sources.assign(n, 0xFFFFFFFFu);
      t_last_targets = targets;
      rc = rs_job_download(job, src.simple() || !write_back ? nullptr : src.tpix, t_last_sources.data());
    } else if (!src.simple() && write_back) {
      rc = rs_job_download(job, src.tpix, nullptr);
    }
    if (!rc && src.simple() && write_back) rc = rs_job_download_simple(job, src.img->data, src.img->rowBytes);
  }
  if (rc) {
    t_err = host_fault ? host_fault : rs_cuda_last_error();  // the CUDA layer's text only when the failure was its own
    rs_job_destroy(job);
    return RS_ERROR_CUDA;
  }
  dbg("after run+download");
  RsJobCounters jc;
  rs_job_counters(job, &jc);
  if (t_keep_result)
    for (uint32_t p = 0; p < 6; p++) {
      t_timeline[p].assign(512, 0);
      t_timeline[p].resize(rs_job_timeline(job, p, t_timeline[p].data(), 512));
    }
  rs_job_destroy(job);
  dbg("after destroy");
  const double t4 = now_ms();
  t_stats.visits = jc.visits; t_stats.evals = jc.evals; t_stats.evals_issued = jc.evals_issued;
  t_stats.compares = jc.compares; t_stats.offset_scans = jc.offset_scans; t_stats.heur_evals = jc.heur_evals;
  t_stats.heur_skips = jc.heur_skips; t_stats.perfect = jc.perfect;
  for (int p = 0; p < 6; p++) { t_stats.betters[p] = jc.betters[p]; t_stats.pass_visits[p] = jc.pass_visits[p]; t_stats.sum_best[p] = jc.sum_best[p]; }
  t_stats.passes_run = jc.passes_run; t_stats.n_targets = n; t_stats.n_corpus = jc.n_corpus;
  t_stats.ms_prep = (float)((t1 - t0) + (t2b - t2)); t_stats.ms_h2d = (float)(t2 - t1); t_stats.ms_kernels = jc.ms_passes;
  t_stats.order_cache_hit = hit == 1 ? 1u : 0u;
  t_stats.ms_d2h = (float)(t4 - t3); t_stats.ms_total = (float)(t4 - t0);
  for (int p = 0; p < 6; p++) t_stats.ms_pass[p] = jc.ms_pass[p];
  g_kernel_launches.fetch_add(jc.kernel_launches);
  if (std::getenv("RS_DEBUG"))  // where the call went, host clock: the run phase includes whatever of the upload was still queued
    std::fprintf(stderr, "[rs debug] call %.2f ms: prep %.2f | stage+digest %.2f | order %.2f | run %.2f (kernels %.2f) | read-back %.2f\n",
                 t4 - t0, t1 - t0, t2 - t1, t2b - t2, t3 - t2b, (double)jc.ms_passes, t4 - t3);
  t_stats.ms_synth = jc.ms_synth; t_stats.kernel_launches = jc.kernel_launches; t_stats.synth_launches_run = jc.synth_launches_run;
  return 0;  // success, also when cancelled (lib/engine.c:689)
}

}  // namespace

// ------------------------------------------------------------------------------------------ engine()
extern "C" int engine(TImageSynthParameters prm, TFormatIndices *fi, Map *targetMap, Map *corpusMap,
                      void (*progressCallback)(int, void *), void *contextInfo, int *cancelFlag) {
  PixelSource src;
  src.tw = (int)targetMap->width; src.th = (int)targetMap->height;
  src.cw = (int)corpusMap->width; src.ch = (int)corpusMap->height;
  src.bpp = fi->total_bpp;
  src.tpix = reinterpret_cast<uint8_t *>(targetMap->data->data);
  src.cpix = reinterpret_cast<const uint8_t *>(corpusMap->data->data);
  return synth_core(prm, fi, src, progressCallback, contextInfo, cancelFlag, false);
}

// ------------------------------------------------------------------------------- batch of independent jobs
// The reference's users loop over engine() / imageSynth() (PluginScripts/plugin-heal-selection.py:148 once per image);
// a job never shards (a visit reads pixels written by arbitrary earlier visits), a batch does.  The dealer below runs
// ONE queue of jobs over `n_devices` GPUs of the box from one process: every device gets `slots` host threads (each
// with its own workspace and stream on that device), every thread pulls the next job from the queue -- longest jobs
// first when their estimated costs differ -- so a GPU that finishes early takes more.  No collective, no peer copies:
// jobs are independent.
extern "C" void rs_cuda_set_job_slots(int slots);
extern "C" void rs_cuda_shared_corpus_stats(unsigned long long *builds, unsigned long long *hits, unsigned long long *peer_copies);
extern "C" void rs_shared_corpus_stats(unsigned long long *builds, unsigned long long *hits, unsigned long long *peer_copies) {
  rs_cuda_shared_corpus_stats(builds, hits, peer_copies);
}
namespace {
struct BatchPlan {
  int n_jobs = 0, n_devices = 1, slots = 1;
  bool share_corpora = false;  // some jobs name the same corpus pixmap: it is staged and prepared once per device
  bool share_sms = true;     // side-by-side jobs each take 1/slots of the SMs; false: every job launches full-width
                             // grids and the jobs in flight only overlap their copies and host work with kernels
  const int *devices = nullptr;
  std::vector<double> cost;  // per job: ~ target points x (patch + probes); empty = all equal
};
// Estimated target points of a selection from a sparse sample of its mask bytes (stride `px_stride` bytes).
size_t estimate_targets(const uint8_t *mask0, size_t npx, size_t px_stride) {
  const size_t step = 61;
  size_t hits = 0;
  for (size_t i = 0; i < npx; i += step) hits += mask0[i * px_stride] != 0;
  return hits * step;
}
// Jobs in flight per device.  Side-by-side jobs (each on 1/slots of the SMs) pay off while a job is latency-bound (a
// few thousand target points: B200 sweeps in profiles/); from ~16 k points on a job fills the GPU by itself and more
// than a few in flight only overlap host work and copies with kernels.
int cap_slots(int slots, size_t n_est) {
  int cap = n_est >= 200000 ? 2 : (n_est >= 16384 ? 4 : 8);
  if (const char *e = getenv("RS_SLOTS_CAP")) { const int c = atoi(e); if (c > 0) cap = c; }  // sweeps
  return slots > cap ? cap : (slots < 1 ? 1 : slots);
}
// Jobs in flight on a device each take 1/slots of the SMs.  The alternative -- full-width grids, the jobs in flight only
// overlapping copies and host work with kernels -- measured worse on B200 even for 65 k-point jobs (64 heal jobs
// 2048x2048: 3.83 ms per job shared at 4 slots, 4.36 full-width; one call at a time 5.99): RS_BATCH_SHARE=0 selects it.
bool share_sms_for(size_t /*n_est*/) {
  if (const char *e = std::getenv("RS_BATCH_SHARE")) return std::atoi(e) != 0;
  return true;
}
int run_batch(const BatchPlan &plan, const std::function<int(int)> &run_one, int *errors_out) {
  const int n_jobs = plan.n_jobs;
  int n_dev = rs_cuda_device_count();
  if (n_dev <= 0) { t_err = "no CUDA device available (this library has no CPU path)"; return RS_ERROR_CUDA; }
  std::vector<int> devices;
  if (plan.devices && plan.n_devices > 0) {
    for (int d = 0; d < plan.n_devices; d++) {
      if (plan.devices[d] < 0 || plan.devices[d] >= n_dev) { t_err = "batch: device ordinal out of range"; return RS_ERROR_CUDA; }
      devices.push_back(plan.devices[d]);
    }
  } else {
    if (int e = ensure_device()) return e;
    devices.push_back(t_device >= 0 ? t_device : 0);
  }
  int slots = plan.slots < 1 ? 1 : plan.slots;
  const int threads_wanted = (int)devices.size() * slots;
  if (threads_wanted > n_jobs) slots = (n_jobs + (int)devices.size() - 1) / (int)devices.size();
  // queue order: longest-processing-time first (stable), the classic greedy for unequal independent jobs
  std::vector<int> order(n_jobs);
  for (int i = 0; i < n_jobs; i++) order[i] = i;
  if (!plan.cost.empty())
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return plan.cost[a] > plan.cost[b]; });
  const uint32_t seed = t_seed;
  static std::atomic<unsigned long long> batch_counter{0};
  const unsigned long long batch = plan.share_corpora ? ++batch_counter : 0ull;
  std::vector<int> errs(n_jobs, 0);
  std::vector<std::string> texts(n_jobs);
  std::atomic<int> next{0};
  rs_cuda_set_job_slots(plan.share_sms ? slots : 1);
  auto worker = [&](int device) {
    if (rs_set_device(device)) {  // this device takes no jobs; the others drain the queue
      return;
    }
    rs_set_seed(seed);
    t_batch = batch;
    t_quiet = std::getenv("RS_BATCH_POLL") == nullptr;
    for (int k = next.fetch_add(1); k < n_jobs; k = next.fetch_add(1)) {
      const int i = order[k];
      errs[i] = run_one(i);
      if (errs[i]) texts[i] = t_err;  // thread-local text: keep it past the thread's end
    }
  };
  const int caller_device = t_device;
  const bool caller_chosen = t_device_chosen;
  std::vector<std::thread> pool;
  for (size_t d = 0; d < devices.size(); d++)
    for (int t = 0; t < slots; t++)
      if (!(d == 0 && t == 0)) pool.emplace_back(worker, devices[d]);
  worker(devices[0]);
  for (auto &t : pool) t.join();
  t_batch = 0;
  t_quiet = false;
  if (batch) rs_cuda_drop_shared_corpora(batch);  // the shared corpora live as long as the batch call
  rs_cuda_set_job_slots(1);
  if (caller_chosen && caller_device >= 0) rs_set_device(caller_device);  // the calling thread keeps the device it had
  int first = 0;
  if (next.load() < n_jobs) { t_err = "batch: no usable device"; first = RS_ERROR_CUDA; }
  for (int i = 0; i < n_jobs; i++) {
    if (errors_out) errors_out[i] = errs[i];
    if (!first && errs[i]) { first = errs[i]; t_err = texts[i]; }
  }
  return first;
}
}  // namespace

extern "C" int rs_engine_batch_multi(int n_jobs, const TImageSynthParameters *params, TFormatIndices *const *indices,
                                     Map *const *targetMaps, Map *const *corpusMaps, int n_devices, const int *devices,
                                     int slots, int *errors_out) {
  if (n_jobs <= 0) return 0;
  BatchPlan plan;
  plan.n_jobs = n_jobs; plan.n_devices = n_devices; plan.devices = devices;
  plan.cost.resize(n_jobs);
  size_t n_max = 0;
  bool equal = true;
  for (int i = 0; i < n_jobs; i++) {
    const Map *m = targetMaps[i];
    const size_t n_est = estimate_targets(reinterpret_cast<const uint8_t *>(m->data->data), (size_t)m->width * m->height, m->depth);
    plan.cost[i] = (double)n_est * (double)(params[i].patchSize + params[i].maxProbeCount);
    if (n_est > n_max) n_max = n_est;
    if (plan.cost[i] != plan.cost[0]) equal = false;
  }
  if (equal) plan.cost.clear();
  for (int i = 1; i < n_jobs && !plan.share_corpora; i++)   // one corpus, many targets?
    for (int k = 0; k < i; k++)
      if (corpusMaps[i]->data->data == corpusMaps[k]->data->data) { plan.share_corpora = true; break; }
  plan.slots = cap_slots(slots, n_max);
  plan.share_sms = share_sms_for(n_max);
  return run_batch(plan, [&](int i) {
    int dummy_cancel = 0;
    return engine(params[i], indices[i], targetMaps[i], corpusMaps[i], [](int, void *) {}, nullptr, &dummy_cancel);
  }, errors_out);
}
extern "C" int rs_engine_batch(int n_jobs, const TImageSynthParameters *params, TFormatIndices *const *indices,
                               Map *const *targetMaps, Map *const *corpusMaps, int slots, int *errors_out) {
  return rs_engine_batch_multi(n_jobs, params, indices, targetMaps, corpusMaps, 0, nullptr, slots, errors_out);
}
// The same for simple-API jobs: imageSynth() (masks2 == NULL) or imageSynth2() per image.
extern "C" int rs_image_synth_batch(int n_jobs, ImageBuffer *const *images, ImageBuffer *const *masks,
                                    ImageBuffer *const *masks2, TImageFormat format, const TImageSynthParameters *params,
                                    int n_devices, const int *devices, int slots, int *errors_out) {
  if (n_jobs <= 0) return 0;
  TImageSynthParameters prm;
  if (params) prm = *params; else setDefaultParams(&prm);
  BatchPlan plan;
  plan.n_jobs = n_jobs; plan.n_devices = n_devices; plan.devices = devices;
  plan.cost.resize(n_jobs);
  size_t n_max = 0;
  bool equal = true;
  for (int i = 0; i < n_jobs; i++) {
    const ImageBuffer *m = masks[i];
    size_t hits = 0;  // sparse sample, row by row (rows may be padded)
    for (unsigned y = 0; y < m->height; y += 7)
      for (unsigned x = y % 5; x < m->width; x += 9) hits += m->data[(size_t)y * m->rowBytes + x] != 0;
    const size_t n_est = hits * 63;
    plan.cost[i] = (double)n_est;
    if (n_est > n_max) n_max = n_est;
    if (plan.cost[i] != plan.cost[0]) equal = false;
  }
  if (equal) plan.cost.clear();
  plan.slots = cap_slots(slots, n_max);
  plan.share_sms = share_sms_for(n_max);
  return run_batch(plan, [&](int i) {
    int dummy_cancel = 0;
    TImageSynthParameters p = prm;
    return masks2 && masks2[i] ? imageSynth2(images[i], masks[i], masks2[i], format, &p, [](int, void *) {}, nullptr, &dummy_cancel)
                               : imageSynth(images[i], masks[i], format, &p, [](int, void *) {}, nullptr, &dummy_cancel);
  }, errors_out);
}

// ------------------------------------------------------------------------------- simple API (one image)
namespace {
// lib/imageSynth.c:61-216 + lib/adaptSimple.h: one image, its selection mask, optionally an explicit corpus mask.  The
// reference repacks them into two internal pixmaps on the host and copies every pixel back; here the caller's planes
// are staged as they are, the pixmaps are built on the device and only the rows with target points come back
// (the other pixels are unchanged by definition).
int simple_api(ImageBuffer *img, ImageBuffer *mask, ImageBuffer *mask2, TImageFormat fmt, TImageSynthParameters *prm,
               void (*cb)(int, void *), void *ctx, int *cancel) {
  if (img->width != mask->width || img->height != mask->height) return IMAGE_SYNTH_ERROR_IMAGE_MASK_MISMATCH;
  static TImageSynthParameters defaults;  // function-static like lib/imageSynth.c:82-86
  if (!prm) { setDefaultParams(&defaults); prm = &defaults; }
  TFormatIndices fi;
  if (int e = prepareImageFormatIndicesFromFormatType(&fi, fmt)) return e;
  PixelSource src;
  src.tw = src.cw = (int)img->width; src.th = src.ch = (int)img->height;
  src.bpp = (int)countPixelelsPerPixelForFormat(fmt) + 1;
  src.img = img; src.mask = mask; src.mask2 = mask2;
  return synth_core(*prm, &fi, src, cb, ctx, cancel, true);
}
}  // namespace

extern "C" int imageSynth(ImageBuffer *img, ImageBuffer *mask, TImageFormat fmt, TImageSynthParameters *prm,
                          void (*cb)(int, void *), void *ctx, int *cancel) {
  return simple_api(img, mask, nullptr, fmt, prm, cb, ctx, cancel);
}
extern "C" int imageSynth2(ImageBuffer *img, ImageBuffer *mask, ImageBuffer *mask2, TImageFormat fmt,
                           TImageSynthParameters *prm, void (*cb)(int, void *), void *ctx, int *cancel) {
  return simple_api(img, mask, mask2, fmt, prm, cb, ctx, cancel);
}

// ------------------------------------------------------------------- Map helpers (lib/mapOps.h:38-168)
namespace {
struct ArrayBox { GArray head; };  // GArray head followed by nothing: data owned separately
void alloc_map(Map *m, unsigned w, unsigned h, unsigned depth, unsigned elt) {
  m->width = w; m->height = h; m->depth = depth;
  ArrayBox *b = static_cast<ArrayBox *>(std::calloc(1, sizeof(ArrayBox)));
  b->head.data = static_cast<char *>(std::calloc((size_t)w * h, elt));
  b->head.len = 0;  // the reference never appends to map arrays either (g_array_sized_new reserves only)
  m->data = &b->head;
}
}  // namespace
extern "C" void free_map(Map *m) {
  if (!m || !m->data) return;
  std::free(m->data->data);
  std::free(m->data);
  m->data = nullptr;
}
extern "C" void new_pixmap(Map *m, unsigned w, unsigned h, unsigned depth) { alloc_map(m, w, h, depth, depth); }
extern "C" void new_bytemap(Map *m, unsigned w, unsigned h) { alloc_map(m, w, h, 1, 1); }
extern "C" void new_intmap(Map *m, unsigned w, unsigned h) { alloc_map(m, w, h, sizeof(unsigned), sizeof(unsigned)); }
extern "C" void new_coordmap(Map *m, unsigned w, unsigned h) { alloc_map(m, w, h, sizeof(Coordinates), sizeof(Coordinates)); }
extern "C" void set_bytemap(Map *m, unsigned char v) { std::memset(m->data->data, v, (size_t)m->width * m->height); }
extern "C" void invert_bytemap(Map *m) {
  unsigned char *p = reinterpret_cast<unsigned char *>(m->data->data);
  for (size_t i = 0, n = (size_t)m->width * m->height; i < n; i++) p[i] = (unsigned char)~p[i];
}
extern "C" void interleave_mask(Map *pixmap, Map *mask) {
  unsigned char *d = reinterpret_cast<unsigned char *>(pixmap->data->data);
  const unsigned char *s = reinterpret_cast<const unsigned char *>(mask->data->data);
  for (size_t i = 0, n = (size_t)pixmap->width * pixmap->height; i < n; i++) d[i * pixmap->depth] = s[i * mask->depth];
}

// Visit order and final best corpus point of every target of the last engine() call on this thread
// (packed x | y << 16; 0xFFFFFFFF = none).  Returns the number of targets.
extern "C" uint32_t rs_get_last_result(uint32_t *targets_out, uint32_t *sources_out, uint32_t cap) {
  const uint32_t n = (uint32_t)t_last_sources.size();
  for (uint32_t i = 0; i < n && i < cap; i++) {
    if (targets_out) targets_out[i] = t_last_targets[i];
    if (sources_out) sources_out[i] = t_last_sources[i];
  }
  return n;
}

extern "C" unsigned int rs_get_timeline(unsigned int pass, unsigned long long *out_ns, unsigned int cap) {
  if (pass >= 6) return 0;
  const unsigned int n = (unsigned int)t_timeline[pass].size() < cap ? (unsigned int)t_timeline[pass].size() : cap;
  for (unsigned int i = 0; i < n; i++) out_ns[i] = t_timeline[pass][i];
  return n;
}

// ------------------------------------------------ host-prep pieces exported for parity tests (include/rs_host.h)
extern "C" void rs_host_metric_tables(double sensitivity, double map_weight, uint16_t *color512, uint32_t *map512) {
  rs::build_metric_tables(sensitivity, map_weight, color512, map512);
}
extern "C" uint32_t rs_host_sorted_offsets(int tw, int th, int cw, int ch, int32_t *xy, uint32_t cap) {
  std::vector<uint32_t> o;
  rs::build_sorted_offsets(tw, th, cw, ch, o);
  const uint32_t n = (uint32_t)o.size() < cap ? (uint32_t)o.size() : cap;
  for (uint32_t i = 0; i < n; i++) { xy[2 * i] = (int16_t)(o[i] & 0xFFFFu); xy[2 * i + 1] = ((int32_t)o[i]) >> 16; }
  return (uint32_t)o.size();
}
extern "C" int rs_host_order_targets(int mode, int32_t *xy, uint32_t n, uint32_t seed) {
  std::vector<uint32_t> p(n);
  for (uint32_t i = 0; i < n; i++) p[i] = rs::pack_xy(xy[2 * i], xy[2 * i + 1]);
  rs::GRandMT prng(seed);
  const int e = rs::order_target_points(mode, p, prng);
  for (uint32_t i = 0; i < n; i++) { xy[2 * i] = rs::unpack_x(p[i]); xy[2 * i + 1] = rs::unpack_y(p[i]); }
  return e;
}
extern "C" void rs_host_draws(uint32_t seed, uint32_t n, uint32_t count, uint32_t *out, int via_raw_stream) {
  if (via_raw_stream) {
    rs::RawStream r(seed, (size_t)count + count / 32 + 65536);
    r.reduce(n, out, count);
  } else {
    rs::GRandMT g(seed);
    g.fill_int_range(n, out, count);
  }
}
extern "C" uint32_t rs_host_pass_schedule(uint32_t n, uint32_t *ends6) { return rs::pass_schedule(n, ends6); }
extern "C" uint32_t rs_host_mt_jump_poly(uint32_t q, uint32_t jump_words, uint16_t *idx) {
  const std::vector<uint16_t> &v = rs::mt_jump_poly(q, jump_words);
  if (idx) std::memcpy(idx, v.data(), v.size() * sizeof(uint16_t));
  return (uint32_t)v.size();
}
