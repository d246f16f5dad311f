// Device-side building blocks of the synthesis pass (sm_100a).
//
// Data layout in HBM (one job; see DESIGN.md "Data layout"):
//   corpus   canonical pixels, 4 B [c0,c1,c2,mask] (no map channels) or 8 B [c0,c1,c2,mask | m0,m1,m2,0]
//            -> one aligned 32/64-bit gather per neighbour compare (lib/mapOps.h:147-152 interleaves the
//            mask for the same reason); mask in the top byte makes "selected" one unsigned compare.
//            One sentinel pixel (all zero: not selected) follows the image at index cw*ch; compares that fall
//            outside the corpus read it instead of branching.
//   W        dynamic state of every target-image pixel, 2 version slots x 64 bit:
//            [c0,c1,c2 | ver | srcx16 | srcy16]; a visit of pass p publishes version p+1 into slot (p+1)&1
//            with ONE 64-bit store, so readers need no fence (replaces targetMap colour bytes, sourceOfMap
//            and hasValueMap: lib/engine.c:152-224)
//   meta     per target-image pixel: visit-order index of a target pixel, RS_CTX_VALUED for a context pixel
//            usable as neighbour, RS_NEVER otherwise
//   prober   recentProberMap (lib/engine.c:314-327), deterministic bounded-staleness version of the reference's
//            live map.  A pass is cut into epochs of epoch_len visits; a visit of epoch e sees the stamps of
//            earlier passes and of epochs <= e-2 of its own pass.  Three arrays (epoch mod 3; interleaved per corpus
//            pixel, 32 bytes = one sector for all three), each one 64-bit
//            word per corpus pixel holding two stamps ((pass+1) << 29 | index; 0 = never): hi = newest stamp
//            written to this array, lo = newest stamp of an earlier epoch than hi's.  A visit of epoch e stamps
//            only array e%3, by 64-bit CAS, and only once every visit of epochs <= e-2 has completed; so each
//            array has one writing epoch at a time, and a reader of epoch e (which hides every stamp of its pass
//            with index >= (e-1)*len: epochs e-1, e and the already-running e+1) finds the newest visible
//            stamp of each array in hi or lo.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define RS_CTX_VALUED 0xFFFFFFFEu
#define RS_NEVER 0xFFFFFFFFu
#define RS_PENDING 0xFFFFFFFDu
#define RS_NO_SRC 0xFFFFFFFFu
#define RS_IDX_MASK 0x1FFFFFFFu
#define RS_FULL 0xFFFFFFFFu
#define RS_MAX_NB 64
// Neighbours gathered per lane between two early-out checks (template parameter CH of the distance code).  More per
// check = more gathers in flight and fewer rounds for long patches, but more wasted compares for short ones; B200
// sweeps (profiles/): 2 for patches below RS_CHUNK_SWITCH_K neighbours, 3 from there on.
#ifndef RS_CHUNK_SMALL
#define RS_CHUNK_SMALL 2
#endif
#ifndef RS_CHUNK_LARGE
#define RS_CHUNK_LARGE 3
#endif
#define RS_CHUNK_MAX 8
// Latency mode (team kernel): after a probe's first chunk the rest of its patch is walked RS_CHUNK_CONT neighbours at a
// time -- lanes are plentiful there and what counts is the number of dependent gather rounds, not wasted compares.
#ifndef RS_CHUNK_CONT
#define RS_CHUNK_CONT 6
#endif
#ifndef RS_CHUNK_SWITCH_K
#define RS_CHUNK_SWITCH_K 16
#endif
// Copies of each metric-table entry, one per lane group (lane % RS_LUT_REP picks the column): 32 = no two lanes ever
// share a bank, 8 = a quarter of the shared memory for occasional 2-way conflicts.
#ifndef RS_LUT_REP
#define RS_LUT_REP 32
#endif
#define RS_LUT_WORDS (256 * RS_LUT_REP)
#ifndef RS_SELECT_MIN_POINTS
#define RS_SELECT_MIN_POINTS (8u << 20)   // corpus points from which rs_corpus_point selects from the bitmap (32 MB of table)
#endif
#define RS_MAX_LAUNCHES 16   // pass-kernel launches per job: 6 passes, the first ones cut into up to 4 segments
#define RS_TIMELINE 320      // progress ticks per pass whose start time is kept (4096 visits each)
#define RS_MAX_EPOCHS 40     // epochs per pass: ceil(n / max(64, ceil(n/32))) <= 32

struct __align__(128) RsLine {  // a counter alone on its 128-byte line: hot words must not share an L2 slice
  unsigned int v;
  unsigned int pad[31];
};
struct RsCtrl {             // device-resident control block of one job (zeroed at upload)
  // ---- results and statistics: the part copied back to the host after a run (up to RS_CTRL_COPY_BYTES)
  unsigned int betters[6];
  unsigned int n_corpus;    // number of corpus points (written at upload: host value or device compaction count)
  unsigned int stop;        // set by the last CTA of a pass when betters/n < fraction, or on cancel
  unsigned int passes_run;
  unsigned int fault;       // set by a warp whose wait for another visit outlived RS_SPIN_LIMIT_NS: the job is invalid
  unsigned long long dg_h1, dg_h2;        // digest of the target selection (k_target_digest), layout = RsTargetDigest
  unsigned int dg_n, dg_ymin, dg_ymax, dg_xmin, dg_xmax, dg_pad;  // count, rows and columns that hold target points
  unsigned int dg_sel;      // scratch counter of the target-point compaction (rs_job_shuffle_order)
  unsigned int dg_acc;      // number of raw PRNG words the rejection rule accepted (rs_job_shuffle_order_raw)
  unsigned long long visits, evals, evals_issued, compares, offset_scans, heur_evals, heur_skips, perfect;
  unsigned long long pass_visits[6], sum_best[6];
  unsigned long long pass_end_ns[6];          // globaltimer when the last CTA of a pass left
  unsigned long long tick_ns[6][RS_TIMELINE];  // globaltimer when visit 4096 * i of a pass was claimed ([0] = pass start)
  // ---- synchronisation words, one per line
  RsLine next[RS_MAX_LAUNCHES];       // visits claimed so far, per pass-kernel launch (one atomicAdd per visit)
  RsLine done_ctas[RS_MAX_LAUNCHES];  // CTAs that have left, per pass-kernel launch
  RsLine epoch_wm[6];       // per pass: number of leading epochs that are complete (what waiters poll)
  RsLine epoch_done[6][RS_MAX_EPOCHS];  // per pass and epoch: visits completed (state word + stamps published)
  RsLine claims[3];         // work counters of the pass-0 gather kernels
  RsLine n_ctx;             // context pixels usable as neighbours in the whole target image (k_ctx_blocks)
};
#define RS_CTRL_COPY_BYTES offsetof(RsCtrl, next)

struct RsDev {              // kernel argument (by value)
  const uint32_t *corpus4;
  const uint2 *corpus8;
  unsigned long long *W;
  const uint32_t *meta;
  const uint32_t *tmaps;
  const uint32_t *targets;
  const uint32_t *corpus_pts;
  const uint32_t *cbits;    // bitmap of the usable corpus pixels, row-major, 32 per word (or nullptr: table only)
  const uint2 *csamples;    // {linear index p of usable corpus pixel 16 * j, bitmap window of pixels p .. p + 31}
  const uint32_t *offsets;
  const uint32_t *lut_rep;  // [2][256][32] colour then map metric, replicated per lane (bank-conflict free)
  unsigned long long *prober;     // [cw*ch][4]: stamps of epochs = 0, 1, 2 (mod 3) of a corpus pixel (+ 1 pad word), see above
  const uint2 *nb_lists;    // pass-0 patches gathered up front by k_gather_pass0: [nT][kmax-1]
  const uint8_t *nb_counts; // [nT] patch size of each pass-0 visit
  const uint2 *nb_later;    // patches of the passes >= 1 (every target point has a value by then, so they are the same in
  const uint8_t *nb_later_counts;  // all of them): [nT][kmax-1] entries {offset, meta word of the pixel}, and sizes
  const uint32_t *ctx_blocks;  // usable context pixels per 32x32 block of the target image, [gh][gw] (k_ctx_blocks), or nullptr
  RsCtrl *ctrl;
  volatile unsigned int *host_ticks;  // mapped pinned: [6] highest tick index started per pass (+1)
  const volatile int *host_cancel;    // mapped pinned
  int tw, th, cw, ch;
  int ow, oh;               // the offsets table holds every (x, y) with |x| < ow, |y| < oh (0: a caller's partial table)
  int gw, gh;               // blocks per row / column of ctx_blocks
  uint32_t cn;              // cw * ch = index of the sentinel corpus pixel
  uint32_t nT, nOff;
  uint32_t kmax, probes, seed, penalty;
  uint32_t pass, pass_end;
  uint32_t seg_begin, seg_end;  // this launch claims the visits [seg_begin, seg_end) of the pass, in order
  uint32_t slot, last_seg;      // index of this launch's counters in RsCtrl; last launch of its pass?
  uint32_t chunk;               // CH the launched kernel was instantiated with (patch padding follows it)
  uint32_t epoch_len;       // visits per recentProber epoch: max(64, ceil(nT/32))
  uint32_t epoch_inv;       // floor(2^32 / epoch_len)
  uint32_t regular_r;       // != 0: passes >= 1 may take the head of the offsets table as the patch of a point that is at least
                            // this far from a clipping border, provided RsCtrl::n_ctx + nT == tw * th (no unusable pixel)
  uint32_t ends[6];
  int htile, vtile;
  uint32_t sc_slice;        // corpus pixels per CTA slice when the corpus is staged into shared memory (k_synth_pass<..., true>)
  uint32_t select_min;      // corpus points from which rs_corpus_point selects from the bitmap (RS_SELECT_MIN_POINTS)
  uint32_t cw_inv;          // floor(2^32 / cw): quotient estimate for rs_corpus_point (one correction step makes it exact)
  double terminate_fraction;
};

// ---- counter-based probe draw: identical to oracle/resynth_port.c probe_hash/counter_range ----
__host__ __device__ __forceinline__ uint32_t rs_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
__host__ __device__ __forceinline__ uint32_t rs_probe_hash(uint32_t seed, uint32_t pass, uint32_t index, uint32_t probe) {
  uint32_t h = rs_mix32(seed + 0x9E3779B9u * (pass + 1u));
  h = rs_mix32(h ^ (index * 0x85EBCA6Bu + 0x165667B1u));
  return rs_mix32(h + probe * 0xC2B2AE35u);
}
// rs_probe_hash(seed,pass,index,probe) == rs_mix32(rs_probe_hash_visit(seed,pass,index) + probe * 0xC2B2AE35u)
__host__ __device__ __forceinline__ uint32_t rs_probe_hash_visit(uint32_t seed, uint32_t pass, uint32_t index) {
  const uint32_t h = rs_mix32(seed + 0x9E3779B9u * (pass + 1u));
  return rs_mix32(h ^ (index * 0x85EBCA6Bu + 0x165667B1u));
}
__device__ __forceinline__ uint32_t rs_range(uint32_t r, uint32_t n) { return __umulhi(r, n); }
__device__ __forceinline__ uint32_t rs_nth_set_bit(uint32_t x, uint32_t r) {  // position of the r-th (0-based) set bit of x
  uint32_t pos = 0, c;
  c = __popc(x & 0xFFFFu); if (r >= c) { r -= c; pos += 16; x >>= 16; }
  c = __popc(x & 0xFFu);   if (r >= c) { r -= c; pos += 8;  x >>= 8; }
  c = __popc(x & 0xFu);    if (r >= c) { r -= c; pos += 4;  x >>= 4; }
  c = __popc(x & 0x3u);    if (r >= c) { r -= c; pos += 2;  x >>= 2; }
  c = x & 1u;              if (r >= c) { pos += 1; }
  return pos;
}
// The idx-th corpus point (lib/engine.c:400-431: row-major list of the usable corpus pixels), packed x | y << 16.  When
// EVERY corpus pixel is usable (nC == cw * ch: texture tiles, whole-image corpora) the list is the identity and the
// point follows from the index -- no table lookup, and the table stays out of L1.
__device__ __forceinline__ uint32_t rs_corpus_point(const RsDev &J, uint32_t nC, uint32_t idx) {
  if (nC == J.cn) {
    uint32_t y = __umulhi(idx, J.cw_inv), x = idx - y * (uint32_t)J.cw;  // y is exact or one short (idx < 2^32)
    if (x >= (uint32_t)J.cw) { x -= (uint32_t)J.cw; y++; }
    return x | (y << 16);
  }
  if (J.cbits != nullptr && nC >= (J.cn >> 2) && nC >= J.select_min) {
    // Dense selections (a quarter or more of the pixels usable: an image minus its hole) of RS_SELECT_MIN_POINTS points or
    // more: the point table is tens of megabytes of one-sector DRAM misses, the bitmap and its samples stay in L2
    // (B200: 4096^2 inpaint 59.1 -> 54.0 ms of kernels; a 16 MB table still lives in L2 and the lookup wins by 3 %).
    uint32_t r = idx & 15u, pos;
    const uint2 e = __ldg(J.csamples + (idx >> 4));
    uint32_t c = __popc(e.y);
    if (e.y == 0xFFFFFFFFu) {  // away from the edges of the selection every window is full: nothing to search
      pos = e.x + r;
    } else if (r < c) {
      pos = e.x + rs_nth_set_bit(e.y, r);
    } else {  // sparser than 16 usable pixels in 32 here: go on in the bitmap behind the window
      const uint32_t q = e.x + 32u;
      uint32_t wi = q >> 5, x = __ldg(J.cbits + wi) & (0xFFFFFFFFu << (q & 31u));
      r -= c;
      c = __popc(x);
      while (r >= c) {
        r -= c;
        x = __ldg(J.cbits + ++wi);
        c = __popc(x);
      }
      pos = (wi << 5) + rs_nth_set_bit(x, r);
    }
    uint32_t y = __umulhi(pos, J.cw_inv), xx = pos - y * (uint32_t)J.cw;
    if (xx >= (uint32_t)J.cw) { xx -= (uint32_t)J.cw; y++; }
    return xx | (y << 16);
  }
  return __ldg(J.corpus_pts + idx);
}

// ---- single-copy-atomic 64-bit state word access (coherent at L2, no fence needed) ----
__device__ __forceinline__ unsigned long long rs_ld_state(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 rs_ld_state2(const unsigned long long *p) {  // two adjacent words (16-byte aligned), each atomic
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void rs_st_state(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int rs_ld_u32_relaxed(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long rs_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A wait on another visit normally lasts microseconds.  If the inputs are inconsistent (a corrupted order, a bug) it
// would last forever and take the GPU with it; instead the waiter gives up after RS_SPIN_LIMIT_NS, flags the job as
// faulted (rs_job_run then returns an error) and every other waiter follows within a few thousand polls.
#define RS_SPIN_LIMIT_NS 10000000000ull
struct RsSpinGuard {
  unsigned long long t0 = 0;
  unsigned int polls = 0;
  __device__ __forceinline__ bool expired(RsCtrl *ctrl) {
    if ((++polls & 4095u) != 0u) return false;
    const unsigned long long now = rs_globaltimer();
    if (t0 == 0) t0 = now;
    if (*(volatile unsigned int *)&ctrl->fault) return true;
    if (now - t0 > RS_SPIN_LIMIT_NS) { atomicExch(&ctrl->fault, 1u); return true; }
    return false;
  }
};

__device__ __forceinline__ int rs_off_x(uint32_t o) { return (int)(short)(o & 0xFFFFu); }
__device__ __forceinline__ int rs_off_y(uint32_t o) { return ((int)o) >> 16; }

// ---- mbarrier + TMA 1-D bulk copy (global -> shared), used to stage the metric tables ----
__device__ __forceinline__ void rs_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rs_tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// ---- the corpus on chip (throughput kernel, corpora without map channels that fit) ----
// A corpus tile of up to ~80 k pixels (render-texture's 256x256 tile is 256 KB) is staged ONCE per CTA into shared memory
// by TMA bulk copies -- whole when it fits one CTA, else split over the two CTAs of a thread-block cluster, each holding
// a slice and reading the other's through distributed shared memory (ld.shared::cluster).  Every neighbour compare is
// then an on-chip load instead of an L1/L2 sector gather.  lo / hi are shared::cluster addresses such that pixel a lives
// at (a < split ? lo : hi) + 4 * a.
struct CorpusSmem {
  unsigned lo = 0, hi = 0;
  uint32_t split = 0xFFFFFFFFu;
};
template <bool SMEMC>
__device__ __forceinline__ uint32_t rs_corpus4(const RsDev &J, const CorpusSmem &cs, uint32_t a) {
  if (SMEMC) {
    uint32_t v;
    const unsigned addr = (a < cs.split ? cs.lo : cs.hi) + a * 4u;
    asm("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  }
  return __ldg(J.corpus4 + a);
}
__device__ __forceinline__ unsigned rs_cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned rs_cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned rs_mapa(unsigned shared_addr, unsigned rank) {  // the same shared-memory offset in CTA `rank` of the cluster
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(shared_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void rs_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- one neighbour of the patch as the distance loop reads it (shared memory, one LDS.128, broadcast) ----
// lin = dy * cw + dx and dx: the corpus pixel compared with this neighbour for candidate (cx, cy) is
// clin + lin, inside the corpus iff (unsigned)(cx + dx) < cw and (unsigned)(clin + lin) < cw * ch.
// pix = neighbour colours [c0,c1,c2,0]; pen = what an unusable corpus pixel costs (lib/synthesize.h:291-307).
// Records past the patch (padding up to a whole chunk) have dx = RS_PAD_DX (never inside) and pen = 0.
struct __align__(16) RsNb {
  int32_t lin, dx;
  uint32_t pix, pen;
};
#define RS_PAD_DX 0x40000000
#define RS_NB_SLOTS (RS_MAX_NB + RS_CHUNK_MAX)

__device__ __forceinline__ uint32_t rs_lds_u32(unsigned shared_addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr));
  return v;
}
// Sum of the three table entries selected by bytes 0..2 of d; `col` = shared-space address of THIS LANE's column of
// a replicated table (row stride RS_LUT_REP words; with 32 copies the lanes of a warp never share a bank).
__device__ __forceinline__ uint32_t rs_lut3(unsigned col, uint32_t d) {
  return rs_lds_u32(col + __byte_perm(d, 0, 0x4440) * (RS_LUT_REP * 4u)) + rs_lds_u32(col + __byte_perm(d, 0, 0x4441) * (RS_LUT_REP * 4u)) +
         rs_lds_u32(col + __byte_perm(d, 0, 0x4442) * (RS_LUT_REP * 4u));
}

// ---- CH neighbour compares of one candidate: the body of computeBestFit's loop for neighbours k0.. ----
// (lib/synthesize.h:288-383), k0 >= 1: the target point itself (k = 0) carries no colour term (synthesize.h:328) and
// its corpus pixel is the candidate, always inside and selected; its map terms are added by the caller.
// Branch-free: every gather is issued before the first table lookup; a pixel outside the corpus reads the sentinel
// pixel cn (mask 0), so "usable" is one compare on the loaded word; padding records cost 0.
// The two halves of a chunk, usable apart: the gathers need only the patch GEOMETRY (lin, dx of the records), the
// table lookups need the neighbour colours.  The team kernel issues the gathers of a visit's probes while the visit
// is still waiting for its neighbours to be synthesised.
template <bool MAPS, int CH, bool SMEMC = false>
__device__ __forceinline__ void rs_chunk_gather(const RsDev &J, const RsNb *nb, int cx, uint32_t clin, uint32_t k0,
                                                uint32_t (&cp)[CH], uint32_t (&cm)[CH], const CorpusSmem &cs = CorpusSmem()) {
#pragma unroll
  for (int u = 0; u < CH; u++) {
    const int2 g = *reinterpret_cast<const int2 *>(&nb[k0 + u]);  // lin, dx
    const uint32_t lin = clin + (uint32_t)g.x;
    const bool in = (unsigned)(cx + g.y) < (unsigned)J.cw && lin < J.cn;
    const uint32_t a = in ? lin : J.cn;
    if (MAPS) {
      const uint2 t = __ldg(J.corpus8 + a);
      cp[u] = t.x;
      cm[u] = t.y;
    } else {
      cp[u] = rs_corpus4<SMEMC>(J, cs, a);
      cm[u] = 0u;
    }
  }
}
template <bool MAPS, int CH>
__device__ __forceinline__ uint32_t rs_chunk_reduce(unsigned lutc, unsigned lutm, const RsNb *nb, const uint32_t *nmap,
                                                    uint32_t k0, const uint32_t (&cp)[CH], const uint32_t (&cm)[CH]) {
  uint32_t sum = 0;
#pragma unroll
  for (int u = 0; u < CH; u++) {
    const uint2 pv = *reinterpret_cast<const uint2 *>(&nb[k0 + u].pix);  // pix, pen
    uint32_t t = rs_lut3(lutc, __vabsdiffu4(cp[u], pv.x));
    if (MAPS) t += rs_lut3(lutm, __vabsdiffu4(cm[u], nmap[k0 + u]));
    sum += (cp[u] >= 0xFF000000u) ? t : pv.y;
  }
  return sum;
}
template <bool MAPS, int CH, bool SMEMC = false>
__device__ __forceinline__ uint32_t rs_chunk_sum(const RsDev &J, unsigned lutc, unsigned lutm, const RsNb *nb,
                                                 const uint32_t *nmap, int cx, uint32_t clin, uint32_t k0,
                                                 const CorpusSmem &cs = CorpusSmem()) {
  RsNb r[CH];
  uint32_t cp[CH], cm[CH];
#pragma unroll
  for (int u = 0; u < CH; u++) {
    r[u] = nb[k0 + u];
    const uint32_t lin = clin + (uint32_t)r[u].lin;
    const bool in = (unsigned)(cx + r[u].dx) < (unsigned)J.cw && lin < J.cn;
    const uint32_t a = in ? lin : J.cn;
    if (MAPS) {
      const uint2 t = __ldg(J.corpus8 + a);
      cp[u] = t.x;
      cm[u] = t.y;
    } else {
      cp[u] = rs_corpus4<SMEMC>(J, cs, a);
      cm[u] = 0u;
    }
  }
  uint32_t sum = 0;
#pragma unroll
  for (int u = 0; u < CH; u++) {
    uint32_t t = rs_lut3(lutc, __vabsdiffu4(cp[u], r[u].pix));
    if (MAPS) t += rs_lut3(lutm, __vabsdiffu4(cm[u], nmap[k0 + u]));
    sum += (cp[u] >= 0xFF000000u) ? t : r[u].pen;  // selected (mask 0xFF)? else the maximum weighted difference
  }
  return sum;
}

// ---- the patch distance with early-out, one candidate per lane, lanes refilled dynamically ----
// Restates computeBestFit (lib/synthesize.h:266-400) for a whole candidate list at once.  Sequential
// semantics = the FIRST candidate in list order with the minimum full sum wins (strict '<' to better,
// synthesize.h:382); a lane therefore abandons only when (partial, index) > (best, bestIndex)
// lexicographically.  Result is independent of scheduling, hence bit-exact.
// bestLin/bestCx return the winner's corpus pixel (linear index, x) when a candidate of this range wins.
// Candidates are fetched a window of 32 ahead (one per lane, all lanes at once) and handed to the lanes that need
// one by shuffle, so the dependent table load of cand_of() is off the critical path of a round.
// K = patch size (>= 1); nb/nmap hold 1 + ceil((K-1)/CH)*CH records.
template <bool MAPS, int CH, bool SMEMC = false, class CandFn>
__device__ __forceinline__ void rs_eval_range(const RsDev &J, unsigned lutc, unsigned lutm, const RsNb *nb,
                                              const uint32_t *nmap, uint32_t K, int begin, int end, CandFn cand_of,
                                              uint32_t &bestSum, int &bestIdx, uint32_t &bestLin, int &bestCx,
                                              uint32_t &nCompares, uint32_t &nIssued, const CorpusSmem &cs = CorpusSmem()) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  int next = begin;  // warp-uniform
  int win0 = begin;  // candidate index held by lane 0 in wc; wn holds the 32 after
  uint32_t wc = (win0 + (int)lane < end) ? cand_of(win0 + (int)lane) : 0u;
  uint32_t wn = (win0 + 32 + (int)lane < end) ? cand_of(win0 + 32 + (int)lane) : 0u;
  int myIdx = -1;
  int cx = 0;
  uint32_t clin = 0, k = 1, partial = 0, m0 = 0;
  const uint32_t selfmap = MAPS ? nmap[0] : 0u;
  while (true) {
    if (bestSum == 0u && end > next) end = next;  // perfect match found: hand out nothing later (synthesize.h:565,599)
    const bool need = myIdx < 0;
    const unsigned nbm = __ballot_sync(RS_FULL, need);
    if (nbm && next < end) {
      const int take = next + __popc(nbm & lt);
      const int s = take - win0;  // 0..63
      const uint32_t c0 = __shfl_sync(RS_FULL, wc, s & 31), c1 = __shfl_sync(RS_FULL, wn, s & 31);
      if (need && take < end) {
        const uint32_t c = s < 32 ? c0 : c1;
        myIdx = take;
        cx = (int)(c & 0xFFFFu);
        clin = (c >> 16) * (uint32_t)J.cw + (uint32_t)cx;
        if (MAPS) m0 = __ldg(&J.corpus8[clin].y);  // consumed after the first chunk's gathers are in flight
        k = 1;
        partial = 0;
      }
      next += __popc(nbm);
      if (next > end) next = end;
      if (next - win0 >= 32) {
        win0 += 32;
        wc = wn;
        wn = (win0 + 32 + (int)lane < end) ? cand_of(win0 + 32 + (int)lane) : 0u;
      }
    }
    const bool active = myIdx >= 0;
    if (!__any_sync(RS_FULL, active)) {
      nIssued += (lane == 0u) ? (uint32_t)(next - begin) : 0u;  // every candidate handed out, counted once
      break;
    }
    bool finished = false;
    if (active) {
      partial += rs_chunk_sum<MAPS, CH, SMEMC>(J, lutc, lutm, nb, nmap, cx, clin, k, cs);
      if (MAPS && k == 1u) partial += rs_lut3(lutm, __vabsdiffu4(m0, selfmap));  // map terms of the target point itself (synthesize.h:342-355)
      k += CH;
      finished = (k >= K);
    }
    const bool worse = active && (partial > bestSum || (partial == bestSum && myIdx > bestIdx));
    const bool propose = active && finished && !worse;
    if (__ballot_sync(RS_FULL, propose)) {
      const uint32_t m = __reduce_min_sync(RS_FULL, propose ? partial : 0xFFFFFFFFu);
      const int mi = __reduce_min_sync(RS_FULL, (propose && partial == m) ? myIdx : 0x7FFFFFFF);
      if (m < bestSum || (m == bestSum && mi < bestIdx)) {
        bestSum = m;
        bestIdx = mi;
        // the winner's corpus pixel travels with it, so the commit needs no point lookup
        const int src = __ffs(__ballot_sync(RS_FULL, propose && myIdx == mi)) - 1;
        bestLin = __shfl_sync(RS_FULL, clin, src);
        bestCx = __shfl_sync(RS_FULL, cx, src);
      }
    }
    if (active && (finished || worse)) {
      nCompares += min(k, K);
      myIdx = -1;
    }
  }
}
