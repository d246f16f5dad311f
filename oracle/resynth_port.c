/*
 * TEST INFRASTRUCTURE (oracle/) -- never linked, imported or executed by the
 * product path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it, and only as the checker.
 *
 * A plain-C restatement of the reference's synthesis path, written from the
 * behaviour of /root/reference (file:line cited at each function), with the
 * same C ABI (imageSynth / imageSynth2 / engine).  Parity is PINNED: in
 * "reference mode" this file reproduces, bit for bit, the compiled reference
 * (oracle/_ref/libref_mt_1t.so) and through it 17 of the reference's own
 * golden images (tests/test_oracle_goldens.py, tests/test_port_vs_ref.py).
 *
 * Two switchable semantics (port_set_mode):
 *   rng   0  GLib GRand (MT19937) sequential stream          -- reference product build
 *         1  libc rand() formula of glibProxy.c:36-49        -- reference standalone build
 *         2  counter-based hash keyed (seed,pass,index,probe) -- what the CUDA engine draws
 *   prober 0 live recentProber map (synthesize.h:556,577)     -- reference
 *          1 snapshot per epoch of port_set_epoch_len() visits (0 = per pass) -- experiments
 *          2 lagged epochs of max(64, ceil(n/32)) visits -- what the CUDA engine does: a visit of epoch e
 *            sees the stamps of earlier passes and of epochs <= e-2 of its pass (and its own visit's)
 * Mode (2,2) is the sequential definition of the GPU engine's semantics: the
 * CUDA path must equal it bit for bit on whole images.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- ABI types */
typedef struct { unsigned char *data; unsigned int width, height; size_t rowBytes; } ImageBuffer; /* imageBuffer.h:12-19 */
typedef struct {                      /* engineParams.h:31-86 */
  int htile, vtile, matchContextType;
  double mapWeight, sensitivityToOutliers;
  unsigned int patchSize, maxProbeCount;
} TImageSynthParameters;
typedef struct {                      /* imageFormatIndicies.h:46-58 */
  unsigned char colorEndBip, alpha_bip, map_start_bip, map_end_bip, img_match_bpp, map_match_bpp, total_bpp;
  int isAlphaTarget, isAlphaSource;
} TFormatIndices;
typedef struct { char *data; unsigned int len; } GArrayHead;          /* glibProxy.h:86-92 */
typedef struct { unsigned int width, height, depth; GArrayHead *data; } Map; /* map.h:28-33 */
typedef struct { int x, y; } Pt;                                      /* map.h:43-46 */
typedef void (*ProgressFn)(int, void *);

enum { ERR_FORMAT = 1, ERR_MASK_MISMATCH = 2, ERR_PATCH = 3, ERR_CTX = 4, ERR_EMPTY_TARGET = 5, ERR_EMPTY_CORPUS = 6 };
#define MAX_NB 64
#define MAX_PASSES 6

/* ------------------------------------------------------------ mode + stats */
static int g_rng_mode = 0, g_prober_mode = 0;
static unsigned int g_seed = 1198472u; /* engine.c:643 */
static unsigned int g_epoch_len = 0;    /* prober mode 1: visits per snapshot epoch (0 = one epoch per pass) */
void port_set_epoch_len(unsigned int n) { g_epoch_len = n; }

typedef struct {
  unsigned long long visits, evals, compares, offset_scans, heur_evals, heur_skips, perfect;
  unsigned long long betters[MAX_PASSES], pass_visits[MAX_PASSES], sum_best[MAX_PASSES];
  unsigned int passes_run, n_targets, n_corpus;
} PortStats;
static PortStats g_stats;

/* visit order + final sourceOf of the last engine() call (test instrumentation) */
static int *g_last_xy; static int *g_last_src; static unsigned int g_last_n;
unsigned int port_last_result(int *target_xy, int *source_xy, unsigned int cap) {
  for (unsigned int i = 0; i < g_last_n && i < cap; i++) {
    target_xy[2 * i] = g_last_xy[2 * i]; target_xy[2 * i + 1] = g_last_xy[2 * i + 1];
    source_xy[2 * i] = g_last_src[2 * i]; source_xy[2 * i + 1] = g_last_src[2 * i + 1];
  }
  return g_last_n;
}
void port_set_mode(int rng_mode, int prober_mode) { g_rng_mode = rng_mode; g_prober_mode = prober_mode; }
void port_set_seed(unsigned int seed) { g_seed = seed; }
void port_get_stats(PortStats *out) { *out = g_stats; }

/* ------------------------------------------------------------------- trace */
/* Per-visit dump for kernel-level parity tests: a flat byte stream
 *   header  u32 x11: pass, index, x, y, K, nCand, best, bestx, besty, bettered, nHeur
 *   K   x { i32 ox, oy; u8 px[8]; i32 sx, sy }      (24 B)
 *   nCand x { i32 x, y }                             (8 B)  candidates in evaluation order
 */
static unsigned char *g_trace; static size_t g_trace_len, g_trace_cap;
static unsigned int g_trace_stride = 0, g_trace_max = 0, g_trace_count = 0;
void port_trace_enable(unsigned int stride, unsigned int max_visits) {
  g_trace_stride = stride; g_trace_max = max_visits; g_trace_count = 0; g_trace_len = 0;
}
size_t port_trace_size(void) { return g_trace_len; }
unsigned int port_trace_count(void) { return g_trace_count; }
void port_trace_copy(unsigned char *dst) { memcpy(dst, g_trace, g_trace_len); }
static void trace_put(const void *p, size_t n) {
  if (g_trace_len + n > g_trace_cap) {
    g_trace_cap = (g_trace_len + n) * 2 + 4096;
    g_trace = (unsigned char *)realloc(g_trace, g_trace_cap);
  }
  memcpy(g_trace + g_trace_len, p, n); g_trace_len += n;
}

/* --------------------------------------------------------------------- RNG */
typedef struct { uint32_t mt[624]; int mti; } MT;
static void mt_seed(MT *r, uint32_t s) {
  r->mt[0] = s;
  for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = 624;
}
static uint32_t mt_next(MT *r) {
  if (r->mti >= 624) {
    for (int k = 0; k < 624; k++) {
      uint32_t y = (r->mt[k] & 0x80000000u) | (r->mt[(k + 1) % 624] & 0x7fffffffu);
      r->mt[k] = r->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    r->mti = 0;
  }
  uint32_t y = r->mt[r->mti++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
/* GLib g_rand_int_range(0, n) (GLib >= 2.2 rules; see grand_mt19937.c) or glibProxy.c:36-49 */
static unsigned int seq_range(MT *r, unsigned int n) {
  if (g_rng_mode == 1) {
    if (n < 1) return 0;
    return (unsigned int)(rand() / (RAND_MAX / (n - 1 + 1) + 1));
  }
  if (n == 0) return 0;
  uint32_t v;
  if (n <= 0x80000000u) {
    uint32_t left = (0x80000000u % n) * 2u;
    if (left >= n) left -= n;
    uint32_t maxv = 0xffffffffu - left;
    do v = mt_next(r); while (v > maxv);
  } else { do v = mt_next(r); while (v >= n); }
  return v % n;
}
/* Counter-based draw of the CUDA engine (resynthesizer_b200/csrc/rs_device.cuh: rs_probe_hash). */
static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
static inline uint32_t probe_hash(uint32_t seed, uint32_t pass, uint32_t index, uint32_t probe) {
  uint32_t h = mix32(seed + 0x9E3779B9u * (pass + 1u));
  h = mix32(h ^ (index * 0x85EBCA6Bu + 0x165667B1u));
  return mix32(h + probe * 0xC2B2AE35u);
}
static inline uint32_t counter_range(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

/* --------------------------------------------------------------- engine state */
typedef struct {
  TImageSynthParameters prm; TFormatIndices fi;
  int tw, th, cw, ch, bpp;
  unsigned char *tpix; const unsigned char *cpix;
  unsigned char *hasValue;      /* engine.c:152-168 */
  Pt *sourceOf;                 /* engine.c:185-224 */
  unsigned int *prober;         /* engine.c:314-327 (live), or snapshot A in prober mode 1 */
  unsigned int *proberNext;     /* prober mode 1: state being built during the pass */
  Pt *targets; unsigned int nT;
  Pt *corpus; unsigned int nC;
  Pt *offsets; unsigned int nOff;
  unsigned short cLUT[512]; unsigned int mLUT[512];
  MT mt;
  unsigned int epoch_len;       /* prober modes 1,2: visits per snapshot epoch (0 = whole pass) */
  unsigned int *proberPrev;     /* prober mode 2: stamps of the previous epoch, not yet visible (0xFFFFFFFF = none) */
} Eng;

typedef struct { Pt off; unsigned char px[8]; Pt src; } Nb; /* synthesize.h:120-125 */

/* matchWeighting.h:49-60,142-178,194-204: parameters narrowed to float first */
static void build_luts(Eng *e) {
  float cp = (float)e->prm.sensitivityToOutliers, mw = (float)e->prm.mapWeight;
  double den = (double)(cp * 256);
  double top = log((256.0 / den) * (256.0 / den) + 1.0);
  for (int d = -256; d < 256; d++) {
    double v = log(((double)d / den) * ((double)d / den) + 1.0) / top * (float)65535;
    e->cLUT[256 + d] = (unsigned short)v;
    e->mLUT[256 + d] = (unsigned int)(d * d * mw * 4.0);
  }
}

/* engine.c:465-497 with glibc merge-sort tie behaviour: ascending x^2+y^2,
 * ties in reverse row-major order.  Counting sort fed in reverse row-major order. */
static void build_offsets(Eng *e) {
  int w = e->cw < e->tw ? e->cw : e->tw, h = e->ch < e->th ? e->ch : e->th;
  unsigned int n = (unsigned int)(2 * w - 1) * (unsigned int)(2 * h - 1);
  unsigned int maxd = (unsigned int)((w - 1) * (w - 1) + (h - 1) * (h - 1));
  unsigned int *start = (unsigned int *)calloc((size_t)maxd + 2, sizeof(unsigned int));
  for (int y = -h + 1; y < h; y++) for (int x = -w + 1; x < w; x++) start[(unsigned int)(x * x + y * y) + 1]++;
  for (unsigned int d = 0; d <= maxd; d++) start[d + 1] += start[d];
  e->offsets = (Pt *)malloc((size_t)n * sizeof(Pt)); e->nOff = n;
  for (int y = h - 1; y > -h; y--) for (int x = w - 1; x > -w; x--) {
    Pt p = {x, y}; e->offsets[start[(unsigned int)(x * x + y * y)]++] = p;
  }
  free(start);
}

/* ------------------------------------------------ target ordering (orderTarget.h, brushfire.h) */
typedef struct { Pt p; float key; } SortEl; /* engineTypes.h:30-33 */
/* stable merge sort on (key) with the two tie rules the never-equal comparators produce under
 * glibc's merge sort (engineTypes.h:51-56): "less" => ascending, ties reversed;
 * "more" => descending, ties kept. */
typedef struct { double key; unsigned int orig; } Rank;
static int rank_cmp_asc_rev(const void *a, const void *b) {
  const Rank *x = (const Rank *)a, *y = (const Rank *)b;
  if (x->key < y->key) return -1; if (x->key > y->key) return 1;
  return x->orig > y->orig ? -1 : 1; /* ties: later original first */
}
static int rank_cmp_desc_keep(const void *a, const void *b) {
  const Rank *x = (const Rank *)a, *y = (const Rank *)b;
  if (x->key > y->key) return -1; if (x->key < y->key) return 1;
  return x->orig < y->orig ? -1 : 1;
}
static void band_shuffle(Eng *e) { /* orderTarget.h:58-79 */
  int last = (int)e->nT - 1, half = (int)(e->nT * 0.1);
  for (int i = 0; i <= last; i++) {
    int bs = i - half > 0 ? i - half : 0, be = i + half < last ? i + half : last;
    int j = bs + (int)seq_range(&e->mt, (unsigned int)(be - bs));
    Pt t = e->targets[i]; e->targets[i] = e->targets[j]; e->targets[j] = t;
  }
}
static Pt bbox_center(const Pt *p, unsigned int n) { /* engineTypes.h:192-226 */
  int ulx = INT_MAX, uly = INT_MAX, lrx = 0, lry = 0;
  for (unsigned int i = 0; i < n; i++) {
    if (p[i].x < ulx) ulx = p[i].x; if (p[i].y < uly) uly = p[i].y;
    if (p[i].x > lrx) lrx = p[i].x; if (p[i].y > lry) lry = p[i].y;
  }
  Pt c = {(lrx - ulx) / 2 + ulx, (lry - uly) / 2 + uly}; return c;
}
static unsigned int ray_of(Pt a) { /* brushfire.h:62-67 */
  return (unsigned int)(atan2((float)a.y, (float)a.x) * 200 / 3.1415926535897932384626433832795028841971693993751 + 200);
}
static int order_targets(Eng *e) { /* orderTarget.h:268-343 */
  int mode = e->prm.matchContextType; unsigned int n = e->nT;
  if (mode < 0 || mode > 8) return ERR_CTX;
  if (mode <= 1) { /* orderTarget.h:35-47 */
    for (unsigned int i = 0; i < n; i++) {
      unsigned int j = seq_range(&e->mt, n);
      Pt t = e->targets[i]; e->targets[i] = e->targets[j]; e->targets[j] = t;
    }
    return 0;
  }
  Pt c = bbox_center(e->targets, n);
  Rank *r = (Rank *)malloc((size_t)n * sizeof(Rank));
  Pt *off = (Pt *)malloc((size_t)n * sizeof(Pt));
  for (unsigned int i = 0; i < n; i++) { off[i].x = e->targets[i].x - c.x; off[i].y = e->targets[i].y - c.y; r[i].orig = i; }
  int descending;
  if (mode == 2 || mode == 5 || mode == 8) { /* brushfire.h:72-148; mode 8 nets out to mode 2's sort (orderTarget.h:209-262) */
    unsigned int maxray[401]; memset(maxray, 0, sizeof maxray);
    for (unsigned int i = 0; i < n; i++) {
      unsigned int d = (unsigned int)(off[i].x * off[i].x + off[i].y * off[i].y), g = ray_of(off[i]);
      if (d > maxray[g]) maxray[g] = d;
    }
    for (unsigned int i = 0; i < n; i++) {
      float k = (float)(off[i].y * off[i].y + off[i].x * off[i].x) / maxray[ray_of(off[i])];
      r[i].key = k; /* NaN only when n == 1 */
    }
    descending = (mode != 5);
  } else {
    int by_y = (mode == 4 || mode == 7);
    for (unsigned int i = 0; i < n; i++) r[i].key = by_y ? (double)(off[i].y * off[i].y) : (double)(off[i].x * off[i].x);
    descending = (mode == 3 || mode == 4);
  }
  if (n > 1) qsort(r, n, sizeof(Rank), descending ? rank_cmp_desc_keep : rank_cmp_asc_rev);
  for (unsigned int i = 0; i < n; i++) { e->targets[i].x = off[r[i].orig].x + c.x; e->targets[i].y = off[r[i].orig].y + c.y; }
  free(r); free(off);
  band_shuffle(e);
  return 0;
}

/* ------------------------------------------------------------- the hot path */
static inline const unsigned char *tpx(const Eng *e, int x, int y) { return e->tpix + ((size_t)y * e->tw + x) * e->bpp; }
static inline const unsigned char *cpx(const Eng *e, int x, int y) { return e->cpix + ((size_t)y * e->cw + x) * e->bpp; }
/* engine.c:505-517 */
static inline int corpus_out(const Eng *e, int x, int y) {
  return x < 0 || y < 0 || x >= e->cw || y >= e->ch || cpx(e, x, y)[0] != 0xFF;
}

/* synthesize.h:189-241 (+ :81-113 wrap/clip) */
static unsigned int gather_patch(Eng *e, Pt pos, Nb *nb) {
  unsigned int k = 0;
  nb[0].off.x = 0; nb[0].off.y = 0;
  nb[0].src = e->sourceOf[(size_t)pos.y * e->tw + pos.x];
  memcpy(nb[0].px, tpx(e, pos.x, pos.y), (size_t)e->bpp);
  k = 1;
  for (unsigned int j = 1; j < e->nOff; j++) {
    Pt o = e->offsets[j]; int x = pos.x + o.x, y = pos.y + o.y;
    g_stats.offset_scans++;
    if (x < 0 || x >= e->tw) { if (!e->prm.htile) continue; while (x < 0) x += e->tw; while (x >= e->tw) x -= e->tw; }
    if (y < 0 || y >= e->th) { if (!e->prm.vtile) continue; while (y < 0) y += e->th; while (y >= e->th) y -= e->th; }
    if (!e->hasValue[(size_t)y * e->tw + x]) continue;
    nb[k].off = o; nb[k].src = e->sourceOf[(size_t)y * e->tw + x];
    memcpy(nb[k].px, tpx(e, x, y), (size_t)e->bpp);
    k++;
    if (k >= e->prm.patchSize) break;
  }
  return k;
}

/* synthesize.h:266-400.  Returns 1 on perfect match. */
static int eval_candidate(Eng *e, Pt cand, const Nb *nb, unsigned int K, unsigned int *best, Pt *bestPt, int *bettered) {
  unsigned int sum = 0;
  g_stats.evals++;
  for (unsigned int i = 0; i < K; i++) {
    int x = cand.x + nb[i].off.x, y = cand.y + nb[i].off.y;
    g_stats.compares++;
    if (corpus_out(e, x, y)) {
      sum += 65535u * e->fi.img_match_bpp + e->mLUT[0] * e->fi.map_match_bpp;
    } else {
      const unsigned char *c = cpx(e, x, y), *t = nb[i].px;
      if (i) for (int b = 1; b < e->fi.colorEndBip; b++) sum += e->cLUT[256u + t[b] - c[b]];
      if (e->fi.map_match_bpp > 0) for (int b = e->fi.map_start_bip; b < e->fi.map_end_bip; b++) sum += e->mLUT[256u + t[b] - c[b]];
    }
    if (sum >= *best) return 0;
  }
  *best = sum; *bestPt = cand; *bettered = 1;
  return sum == 0;
}

/* prober mode 1: proberNext collects the current epoch, merged into prober at the epoch boundary.
 * prober mode 2: proberNext = current epoch, proberPrev = previous epoch, prober = everything older (visible). */
static void prober_begin_pass(Eng *e) {
  size_t n = (size_t)e->cw * e->ch;
  if (g_prober_mode == 1) memcpy(e->proberNext, e->prober, n * sizeof(unsigned int));
  if (g_prober_mode == 2) { memset(e->proberNext, 0xFF, n * sizeof(unsigned int)); memset(e->proberPrev, 0xFF, n * sizeof(unsigned int)); }
}
static void prober_advance_epoch(Eng *e) {
  size_t n = (size_t)e->cw * e->ch;
  if (g_prober_mode == 1) { unsigned int *t = e->prober; e->prober = e->proberNext; e->proberNext = t; memcpy(e->proberNext, e->prober, n * sizeof(unsigned int)); }
  if (g_prober_mode == 2) {
    for (size_t i = 0; i < n; i++) if (e->proberPrev[i] != 0xFFFFFFFFu) e->prober[i] = e->proberPrev[i];
    unsigned int *t = e->proberPrev; e->proberPrev = e->proberNext; e->proberNext = t;
    memset(e->proberNext, 0xFF, n * sizeof(unsigned int));
  }
}
static void prober_end_pass(Eng *e) {
  size_t n = (size_t)e->cw * e->ch;
  if (g_prober_mode == 1) { unsigned int *t = e->prober; e->prober = e->proberNext; e->proberNext = t; }
  if (g_prober_mode == 2) {  /* a pass boundary makes everything visible */
    for (size_t i = 0; i < n; i++) if (e->proberPrev[i] != 0xFFFFFFFFu) e->prober[i] = e->proberPrev[i];
    for (size_t i = 0; i < n; i++) if (e->proberNext[i] != 0xFFFFFFFFu) e->prober[i] = e->proberNext[i];
  }
}

/* synthesize.h:426-642 */
static unsigned int run_pass(Eng *e, unsigned int pass, unsigned int end, ProgressFn tick_cb, void *tick_ctx,
                             unsigned int *completed, unsigned int estimated, unsigned int *prior_pct, int *cancel) {
  unsigned int betters = 0; Nb nb[MAX_NB]; Pt bestPt = {0, 0};
  Pt cands[MAX_NB + 4096]; /* trace only */
  prober_begin_pass(e);
  for (unsigned int ti = 0; ti < end; ti++) {
    if ((ti & 4095u) == 0) { /* synthesize.h:493-497 + progress.c:53-65 */
      *completed += 4095u;
      unsigned int pct = (unsigned int)(((float)*completed / estimated) * 100);
      if (pct > *prior_pct) { tick_cb((int)pct, tick_ctx); *prior_pct = pct; }
      if (*cancel) break;
    }
    if (g_prober_mode >= 1 && e->epoch_len && ti && (ti % e->epoch_len) == 0) prober_advance_epoch(e);
    Pt pos = e->targets[ti];
    unsigned int K = gather_patch(e, pos, nb);
    unsigned int best = UINT_MAX; int bettered = 0, perfect = 0;
    unsigned int nCand = 0, nHeur = 0;
    int tracing = g_trace_stride && (g_stats.visits % g_trace_stride == 0) && g_trace_count < g_trace_max && e->prm.maxProbeCount <= 4096;
    g_stats.visits++; g_stats.pass_visits[pass]++;
    /* heuristic 1 + 2: synthesize.h:537-580 */
    Pt stamped[MAX_NB]; unsigned int nStamped = 0;
    for (unsigned int j = 0; j < K && best != 0; j++) {
      if (nb[j].src.x == -1) continue;
      Pt c = {nb[j].src.x - nb[j].off.x, nb[j].src.y - nb[j].off.y};
      if (corpus_out(e, c.x, c.y)) continue;
      unsigned int *slot = &e->prober[(size_t)c.y * e->cw + c.x];
      if (g_prober_mode == 0) {
        if (*slot == ti) { g_stats.heur_skips++; continue; }
      } else {
        int dup = (*slot == ti);
        for (unsigned int s = 0; s < nStamped && !dup; s++) dup = (stamped[s].x == c.x && stamped[s].y == c.y);
        if (dup) { g_stats.heur_skips++; continue; }
      }
      if (tracing) cands[nCand] = c;
      nCand++; nHeur++; g_stats.heur_evals++;
      perfect = eval_candidate(e, c, nb, K, &best, &bestPt, &bettered);
      if (perfect) break;
      if (g_prober_mode == 0) *slot = ti;
      else { e->proberNext[(size_t)c.y * e->cw + c.x] = ti; stamped[nStamped++] = c; }
    }
    /* random probes: synthesize.h:583-604, engine.c:434-443 */
    if (!perfect) {
      for (unsigned int j = 0; j < e->prm.maxProbeCount; j++) {
        unsigned int idx = (g_rng_mode == 2) ? counter_range(probe_hash(g_seed, pass, ti, j), e->nC) : seq_range(&e->mt, e->nC);
        Pt c = e->corpus[idx];
        if (tracing) cands[nCand] = c;
        nCand++;
        perfect = eval_candidate(e, c, nb, K, &best, &bestPt, &bettered);
        if (perfect) break;
      }
    }
    if (perfect) g_stats.perfect++;
    if (tracing) {
      uint32_t hd[11] = {pass, ti, (uint32_t)pos.x, (uint32_t)pos.y, K, nCand, best, (uint32_t)bestPt.x, (uint32_t)bestPt.y, (uint32_t)bettered, nHeur};
      trace_put(hd, sizeof hd);
      for (unsigned int k = 0; k < K; k++) { int32_t a[2] = {nb[k].off.x, nb[k].off.y}; trace_put(a, 8); trace_put(nb[k].px, 8); int32_t s[2] = {nb[k].src.x, nb[k].src.y}; trace_put(s, 8); }
      for (unsigned int k = 0; k < nCand; k++) { int32_t a[2] = {cands[k].x, cands[k].y}; trace_put(a, 8); }
      g_trace_count++;
    }
    /* commit: synthesize.h:620-639 */
    if (bettered) {
      Pt *so = &e->sourceOf[(size_t)pos.y * e->tw + pos.x];
      g_stats.sum_best[pass] += best;
      if (so->x != bestPt.x || so->y != bestPt.y) {
        betters++;
        unsigned char *t = e->tpix + ((size_t)pos.y * e->tw + pos.x) * e->bpp; const unsigned char *c = cpx(e, bestPt.x, bestPt.y);
        for (int b = 1; b < e->fi.colorEndBip; b++) t[b] = c[b];
        *so = bestPt;
      }
    }
    e->hasValue[(size_t)pos.y * e->tw + pos.x] = 1;
  }
  prober_end_pass(e);
  return betters;
}

int engine(TImageSynthParameters prm, TFormatIndices *fi, Map *targetMap, Map *corpusMap,
           ProgressFn cb, void *ctx, int *cancel) { /* engine.c:539-690 */
  Eng e; memset(&e, 0, sizeof e); memset(&g_stats, 0, sizeof g_stats);
  e.prm = prm; e.fi = *fi; e.bpp = fi->total_bpp;
  e.tw = (int)targetMap->width; e.th = (int)targetMap->height; e.cw = (int)corpusMap->width; e.ch = (int)corpusMap->height;
  e.tpix = (unsigned char *)targetMap->data->data; e.cpix = (const unsigned char *)corpusMap->data->data;
  if (prm.patchSize > MAX_NB) return ERR_PATCH;
  size_t tn = (size_t)e.tw * e.th, cn = (size_t)e.cw * e.ch;
  /* engine.c:338-391 */
  e.hasValue = (unsigned char *)calloc(tn, 1);
  e.targets = (Pt *)malloc((tn ? tn : 1) * sizeof(Pt));
  for (int y = 0; y < e.th; y++) for (int x = 0; x < e.tw; x++) {
    const unsigned char *p = tpx(&e, x, y); int sel = p[0] != 0;
    e.hasValue[(size_t)y * e.tw + x] = (unsigned char)(prm.matchContextType && !sel && (fi->isAlphaTarget ? p[fi->alpha_bip] != 0 : 1));
    if (sel) { Pt q = {x, y}; e.targets[e.nT++] = q; }
  }
  if (!e.nT) { free(e.hasValue); free(e.targets); return ERR_EMPTY_TARGET; }
  /* engine.c:400-431 */
  e.corpus = (Pt *)malloc((cn ? cn : 1) * sizeof(Pt));
  for (int y = 0; y < e.ch; y++) for (int x = 0; x < e.cw; x++) {
    const unsigned char *p = cpx(&e, x, y);
    if (p[0] == 0xFF && (fi->isAlphaSource ? p[fi->alpha_bip] != 0 : 1)) { Pt q = {x, y}; e.corpus[e.nC++] = q; }
  }
  if (!e.nC) { free(e.hasValue); free(e.targets); free(e.corpus); return ERR_EMPTY_CORPUS; }
  e.sourceOf = (Pt *)malloc(tn * sizeof(Pt));
  for (size_t i = 0; i < tn; i++) { e.sourceOf[i].x = -1; e.sourceOf[i].y = -1; }
  build_offsets(&e); build_luts(&e);
  if (g_rng_mode == 1) srand(g_seed); else mt_seed(&e.mt, g_seed);
  int err = order_targets(&e);
  if (!err) {
    e.prober = (unsigned int *)malloc(cn * sizeof(unsigned int)); memset(e.prober, 0xFF, cn * sizeof(unsigned int));
    if (g_prober_mode >= 1) e.proberNext = (unsigned int *)malloc(cn * sizeof(unsigned int));
    if (g_prober_mode == 2) e.proberPrev = (unsigned int *)malloc(cn * sizeof(unsigned int));
    e.epoch_len = g_prober_mode == 1 ? g_epoch_len : (g_prober_mode == 2 ? ((e.nT + 31u) / 32u < 64u ? 64u : (e.nT + 31u) / 32u) : 0u);
    /* refiner.h:42-122, passes.h:67-93 */
    unsigned int ends[MAX_PASSES], est = 0, n = e.nT;
    ends[0] = n; est = n;
    for (int p = 1; p < MAX_PASSES; p++) { ends[p] = n; est += n; n = n * 3 / 4; }
    unsigned int completed = 0, prior = 0;
    g_stats.n_targets = e.nT; g_stats.n_corpus = e.nC;
    for (unsigned int p = 0; p < MAX_PASSES; p++) {
      unsigned int b = run_pass(&e, p, ends[p], cb, ctx, &completed, est, &prior, cancel);
      g_stats.betters[p] = b; g_stats.passes_run = p + 1;
      if ((float)b / e.nT < 0.1) break;
    }
    free(e.prober); free(e.proberNext); free(e.proberPrev);
    g_last_xy = (int *)realloc(g_last_xy, (size_t)e.nT * 8); g_last_src = (int *)realloc(g_last_src, (size_t)e.nT * 8); g_last_n = e.nT;
    for (unsigned int i = 0; i < e.nT; i++) {
      Pt t = e.targets[i], so = e.sourceOf[(size_t)t.y * e.tw + t.x];
      g_last_xy[2 * i] = t.x; g_last_xy[2 * i + 1] = t.y; g_last_src[2 * i] = so.x; g_last_src[2 * i + 1] = so.y;
    }
  }
  free(e.hasValue); free(e.targets); free(e.corpus); free(e.sourceOf); free(e.offsets);
  return err;
}

/* ------------------------------------------------------- format + simple API */
unsigned int countPixelelsPerPixelForFormat(int f) { /* imageFormat.c:39-56 */
  switch (f) { case 0: return 3; case 1: return 4; case 2: return 1; case 3: return 2; default: return 0; }
}
void prepareImageFormatIndices(TFormatIndices *o, unsigned int nColor, unsigned int nMap, int aT, int aS, int isMap) { /* imageFormat.c:116-207 */
  o->img_match_bpp = (unsigned char)nColor; o->colorEndBip = (unsigned char)(1 + nColor);
  if (aT || aS) { o->alpha_bip = o->colorEndBip; o->map_start_bip = (unsigned char)(1 + o->colorEndBip); }
  else o->map_start_bip = o->colorEndBip;
  o->map_match_bpp = (unsigned char)(isMap ? nMap : 0);
  o->map_end_bip = (unsigned char)(o->map_start_bip + o->map_match_bpp); o->total_bpp = o->map_end_bip;
  o->isAlphaTarget = aT; o->isAlphaSource = aS;
}
int prepareImageFormatIndicesFromFormatType(TFormatIndices *o, int f) { /* imageFormat.c:65-113 */
  switch (f) {
    case 0: prepareImageFormatIndices(o, 3, 0, 0, 0, 0); return 0;
    case 1: prepareImageFormatIndices(o, 3, 0, 1, 1, 0); return 0;
    case 2: prepareImageFormatIndices(o, 1, 0, 0, 0, 0); return 0;
    case 3: prepareImageFormatIndices(o, 1, 0, 1, 1, 0); return 0;
    default: return ERR_FORMAT;
  }
}
void setDefaultParams(TImageSynthParameters *p) { /* engineParams.c:8-19 */
  p->htile = 0; p->vtile = 0; p->matchContextType = 1; p->mapWeight = 0.5; p->sensitivityToOutliers = 0.117; p->patchSize = 30; p->maxProbeCount = 200;
}
/* imageSynth.c:61-216 + adaptSimple.h:47-307 */
static int simple(ImageBuffer *img, ImageBuffer *mask, ImageBuffer *mask2, int fmt, TImageSynthParameters *prm,
                  ProgressFn cb, void *ctx, int *cancel) {
  if (img->width != mask->width || img->height != mask->height) return ERR_MASK_MISMATCH;
  static TImageSynthParameters dflt;
  if (!prm) { setDefaultParams(&dflt); prm = &dflt; }
  TFormatIndices fi; int err = prepareImageFormatIndicesFromFormatType(&fi, fmt); if (err) return err;
  unsigned int nc = countPixelelsPerPixelForFormat(fmt), d = nc + 1, w = img->width, h = img->height;
  unsigned char *t = (unsigned char *)calloc((size_t)w * h, d), *c = (unsigned char *)calloc((size_t)w * h, d);
  for (unsigned int y = 0; y < h; y++) for (unsigned int x = 0; x < w; x++) {
    size_t o = ((size_t)y * w + x) * d; unsigned char m = mask->data[y * mask->rowBytes + x];
    t[o] = m; c[o] = mask2 ? mask2->data[y * mask2->rowBytes + x] : (unsigned char)~m;
    for (unsigned int k = 0; k < nc; k++) t[o + 1 + k] = c[o + 1 + k] = img->data[y * img->rowBytes + x * nc + k];
  }
  GArrayHead ta = {(char *)t, w * h}, ca = {(char *)c, w * h};
  Map tm = {w, h, d, &ta}, cm = {w, h, d, &ca};
  err = engine(*prm, &fi, &tm, &cm, cb, ctx, cancel);
  if (!err && !*cancel)
    for (unsigned int y = 0; y < h; y++) for (unsigned int x = 0; x < w; x++)
      for (unsigned int k = 0; k < nc; k++) img->data[y * img->rowBytes + x * nc + k] = t[((size_t)y * w + x) * d + 1 + k];
  free(t); free(c);
  return err;
}
int imageSynth(ImageBuffer *img, ImageBuffer *mask, int fmt, TImageSynthParameters *prm, ProgressFn cb, void *ctx, int *cancel) {
  return simple(img, mask, NULL, fmt, prm, cb, ctx, cancel);
}
int imageSynth2(ImageBuffer *img, ImageBuffer *mask, ImageBuffer *mask2, int fmt, TImageSynthParameters *prm, ProgressFn cb, void *ctx, int *cancel) {
  return simple(img, mask, mask2, fmt, prm, cb, ctx, cancel);
}

/* ---------------------------------------------- pieces exported for host-prep parity tests */
/* Each fills caller buffers from the same code the engine above runs. */
void port_luts(double sens, double mapWeight, unsigned short *c512, unsigned int *m512) {
  Eng e; memset(&e, 0, sizeof e); e.prm.sensitivityToOutliers = sens; e.prm.mapWeight = mapWeight; build_luts(&e);
  memcpy(c512, e.cLUT, sizeof e.cLUT); memcpy(m512, e.mLUT, sizeof e.mLUT);
}
unsigned int port_offsets(int tw, int th, int cw, int ch, int *xy, unsigned int cap) {
  Eng e; memset(&e, 0, sizeof e); e.tw = tw; e.th = th; e.cw = cw; e.ch = ch; build_offsets(&e);
  unsigned int n = e.nOff < cap ? e.nOff : cap;
  for (unsigned int i = 0; i < n; i++) { xy[2 * i] = e.offsets[i].x; xy[2 * i + 1] = e.offsets[i].y; }
  free(e.offsets); return e.nOff;
}
/* Orders `n` points (xy pairs, row-major scan order as the engine builds them) in place. */
int port_order(int mode, int *xy, unsigned int n, unsigned int seed) {
  Eng e; memset(&e, 0, sizeof e); e.prm.matchContextType = mode; e.nT = n; e.targets = (Pt *)xy;
  if (g_rng_mode == 1) srand(seed); else mt_seed(&e.mt, seed);
  return order_targets(&e);
}
/* Best fit over an explicit candidate list (kernel-level vectors): returns best sum, writes best point/index. */
unsigned int port_bestfit(const TFormatIndices *fi, const unsigned char *corpus, int cw, int ch,
                          const unsigned short *c512, const unsigned int *m512,
                          unsigned int K, const int *off_xy, const unsigned char *px8,
                          unsigned int nCand, const int *cand_xy, int *best_xy, int *best_index) {
  Eng e; memset(&e, 0, sizeof e); e.fi = *fi; e.bpp = fi->total_bpp; e.cpix = corpus; e.cw = cw; e.ch = ch;
  memcpy(e.cLUT, c512, sizeof e.cLUT); memcpy(e.mLUT, m512, sizeof e.mLUT);
  Nb nb[MAX_NB];
  for (unsigned int k = 0; k < K; k++) { nb[k].off.x = off_xy[2 * k]; nb[k].off.y = off_xy[2 * k + 1]; memcpy(nb[k].px, px8 + 8 * k, 8); }
  unsigned int best = UINT_MAX; Pt bp = {-1, -1}; int bettered = 0; *best_index = -1;
  for (unsigned int c = 0; c < nCand; c++) {
    Pt cand = {cand_xy[2 * c], cand_xy[2 * c + 1]}; unsigned int before = best;
    int perfect = eval_candidate(&e, cand, nb, K, &best, &bp, &bettered);
    if (best != before) *best_index = (int)c;
    if (perfect) break;
  }
  best_xy[0] = bp.x; best_xy[1] = bp.y;
  return best;
}
