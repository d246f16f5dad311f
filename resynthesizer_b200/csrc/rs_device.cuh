// Device-side building blocks of the synthesis pass (sm_100a).
//
// Data layout in HBM (one job; see DESIGN.md "Data layout"):
//   corpus   canonical pixels, 4 B [mask,c0,c1,c2] (no map channels) or 8 B [mask,c0,c1,c2 | m0,m1,m2,0]
//            -> one aligned 32/64-bit gather per neighbour compare (lib/mapOps.h:147-152 interleaves the
//            mask for the same reason)
//   W        dynamic state of every target-image pixel, 2 version slots x 64 bit:
//            [c0,c1,c2 | ver | srcx16 | srcy16]; a visit of pass p publishes version p+1 into slot (p+1)&1
//            with ONE 64-bit store, so readers need no fence (replaces targetMap colour bytes, sourceOfMap
//            and hasValueMap: lib/engine.c:152-224)
//   meta     per target-image pixel: visit-order index of a target pixel, RS_CTX_VALUED for a context pixel
//            usable as neighbour, RS_NEVER otherwise
//   prober   recentProberMap (lib/engine.c:314-327), deterministic bounded-staleness version of the reference's
//            live map.  A pass is cut into epochs of epoch_len visits; a visit of epoch e sees the stamps of
//            earlier passes and of epochs <= e-2 of its own pass.  Three arrays (epoch mod 3), each one 64-bit
//            word per corpus pixel holding two stamps ((pass+1) << 29 | index; 0 = never): hi = newest stamp
//            written to this array, lo = newest stamp of an earlier epoch than hi's.  A visit of epoch e stamps
//            only array e%3, by 64-bit CAS, and only once every visit of epochs <= e-2 has completed; so each
//            array has one writing epoch at a time, and a reader of epoch e (which hides every stamp of its pass
//            with index >= (e-1)*len: epochs e-1, e and the already-running e+1) finds the newest visible
//            stamp of each array in hi or lo.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RS_CTX_VALUED 0xFFFFFFFEu
#define RS_NEVER 0xFFFFFFFFu
#define RS_PENDING 0xFFFFFFFDu
#define RS_NO_SRC 0xFFFFFFFFu
#define RS_IDX_MASK 0x1FFFFFFFu
#define RS_FULL 0xFFFFFFFFu
#define RS_MAX_NB 64
#ifndef RS_CHUNK
#define RS_CHUNK 3          // neighbours gathered per lane between two early-out checks (sweep: profiles/)
#endif
#define RS_LUT_WORDS (256 * 32)
#define RS_MAX_EPOCHS 40     // epochs per pass: ceil(n / max(64, ceil(n/32))) <= 32

struct RsCtrl {             // device-resident control block of one job (zeroed at upload)
  unsigned int next[6];     // next visit index to claim, per pass
  unsigned int betters[6];
  unsigned int done_ctas[6];
  unsigned int epoch_done[6][40];  // per pass and epoch: visits completed (state word + stamps published)
  unsigned int epoch_wm[6];        // per pass: number of leading epochs that are complete (what waiters poll)
  unsigned int n_corpus;    // number of corpus points (written at upload: host value or device compaction count)
  unsigned int stop;        // set by the last CTA of a pass when betters/n < fraction, or on cancel
  unsigned int passes_run;
  unsigned long long visits, evals, evals_issued, compares, offset_scans, heur_evals, heur_skips, perfect;
  unsigned long long pass_visits[6], sum_best[6];
};

struct RsDev {              // kernel argument (by value)
  const uint32_t *corpus4;
  const uint2 *corpus8;
  unsigned long long *W;
  const uint32_t *meta;
  const uint32_t *tmaps;
  const uint32_t *targets;
  const uint32_t *corpus_pts;
  const uint32_t *offsets;
  const uint32_t *lut_rep;  // [2][256][32] colour then map metric, replicated per lane (bank-conflict free)
  unsigned long long *prober[3];  // [cw*ch] each: stamps of epochs = 0, 1, 2 (mod 3), see above
  const uint2 *nb_lists;    // pass-0 patches gathered up front by k_gather_pass0: [nT][kmax-1]
  const uint8_t *nb_counts; // [nT] patch size of each pass-0 visit
  RsCtrl *ctrl;
  volatile unsigned int *host_ticks;  // mapped pinned: [6] highest tick index started per pass (+1)
  const volatile int *host_cancel;    // mapped pinned
  int tw, th, cw, ch;
  uint32_t nT, nOff;
  uint32_t kmax, probes, seed, penalty;
  uint32_t pass, pass_end;
  uint32_t epoch_len;       // visits per recentProber epoch: max(64, ceil(nT/32))
  uint32_t ends[6];
  int htile, vtile;
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

// ---- single-copy-atomic 64-bit state word access (coherent at L2, no fence needed) ----
__device__ __forceinline__ unsigned long long rs_ld_state(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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

// ---- up to RS_CHUNK neighbour compares of one candidate: the body of computeBestFit's loop ----
// (lib/synthesize.h:288-383).  Branch-free: every gather is issued (clamped to pixel 0 when the neighbour falls
// outside the corpus or past the patch) before the first table lookup; validity is applied by selects.
// lutc/lutm point at THIS LANE's column of the replicated tables (stride 32 words per table row).
template <bool MAPS>
__device__ __forceinline__ uint32_t rs_chunk_sum(const RsDev &J, const uint32_t *lutc, const uint32_t *lutm,
                                                 const uint32_t *s_off, const uint32_t *s_pix, const uint32_t *s_map,
                                                 uint32_t K, int cx, int cy, uint32_t k0, uint32_t &nCompares) {
  uint32_t cp[RS_CHUNK], cm[RS_CHUNK];
  bool inb[RS_CHUNK];
#pragma unroll
  for (int u = 0; u < RS_CHUNK; u++) {
    const uint32_t kk = k0 + u;
    const bool valid = kk < K;
    const uint32_t o = s_off[valid ? kk : 0u];
    const int x = cx + rs_off_x(o), y = cy + rs_off_y(o);
    inb[u] = valid && (unsigned)x < (unsigned)J.cw && (unsigned)y < (unsigned)J.ch;
    const uint32_t a = inb[u] ? (uint32_t)y * (uint32_t)J.cw + (uint32_t)x : 0u;
    if (MAPS) {
      const uint2 t = __ldg(J.corpus8 + a);
      cp[u] = t.x;
      cm[u] = t.y;
    } else {
      cp[u] = __ldg(J.corpus4 + a);
      cm[u] = 0u;
    }
  }
  uint32_t sum = 0;
#pragma unroll
  for (int u = 0; u < RS_CHUNK; u++) {
    const uint32_t kk = k0 + u;
    const bool valid = kk < K;
    const uint32_t ks = valid ? kk : 0u;
    // inside the corpus and selected (mask 0xFF)? else the maximum weighted difference (synthesize.h:291-307)
    const bool usable = inb[u] && (cp[u] & 0xFFu) == 0xFFu;
    const uint32_t d = __vabsdiffu4(cp[u], s_pix[ks]);
    uint32_t t = lutc[((d >> 8) & 0xFFu) * 32u] + lutc[((d >> 16) & 0xFFu) * 32u] + lutc[(d >> 24) * 32u];
    t = kk ? t : 0u;  // the target point itself carries no colour term (synthesize.h:328)
    if (MAPS) {       // map terms also for the target point itself (synthesize.h:342-355)
      const uint32_t dm = __vabsdiffu4(cm[u], s_map[ks]);
      t += lutm[(dm & 0xFFu) * 32u] + lutm[((dm >> 8) & 0xFFu) * 32u] + lutm[((dm >> 16) & 0xFFu) * 32u];
    }
    sum += valid ? (usable ? t : J.penalty) : 0u;
    nCompares += valid ? 1u : 0u;
  }
  return sum;
}

// ---- the patch distance with early-out, one candidate per lane, lanes refilled dynamically ----
// Restates computeBestFit (lib/synthesize.h:266-400) for a whole candidate list at once.  Sequential
// semantics = the FIRST candidate in list order with the minimum full sum wins (strict '<' to better,
// synthesize.h:382); a lane therefore abandons only when (partial, index) > (best, bestIndex)
// lexicographically.  Result is independent of scheduling, hence bit-exact.
template <bool MAPS, class CandFn>
__device__ __forceinline__ void rs_eval_range(const RsDev &J, const uint32_t *lutc, const uint32_t *lutm,
                                              const uint32_t *s_off, const uint32_t *s_pix, const uint32_t *s_map,
                                              uint32_t K, int begin, int end, CandFn cand_of, uint32_t &bestSum,
                                              int &bestIdx, uint32_t &nCompares, uint32_t &nIssued) {
  const unsigned lt = (1u << (threadIdx.x & 31u)) - 1u;
  int next = begin;  // warp-uniform
  int myIdx = -1;
  int cx = 0, cy = 0;
  uint32_t k = 0, partial = 0;
  while (true) {
    if (bestSum == 0u && end > next) end = next;  // perfect match found: hand out nothing later (synthesize.h:565,599)
    const bool need = myIdx < 0;
    const unsigned nb = __ballot_sync(RS_FULL, need);
    if (nb && next < end) {
      const int take = next + __popc(nb & lt);
      if (need && take < end) {
        myIdx = take;
        const uint32_t c = cand_of(take);
        cx = (int)(c & 0xFFFFu);
        cy = (int)(c >> 16);
        k = 0;
        partial = 0;
        nIssued++;
      }
      next += __popc(nb);
      if (next > end) next = end;
    }
    const bool active = myIdx >= 0;
    if (!__any_sync(RS_FULL, active)) break;
    bool finished = false;
    if (active) {
      partial += rs_chunk_sum<MAPS>(J, lutc, lutm, s_off, s_pix, s_map, K, cx, cy, k, nCompares);
      k += RS_CHUNK;
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
      }
    }
    if (active && (finished || worse)) myIdx = -1;
  }
}
