// sm_100a kernels of the synthesis engine and the thin C-ABI over them (include/rs_cuda.h).
//
// One persistent kernel per pass.  Every warp claims target visits IN ORDER (atomic counter), so a visit
// can only ever wait on visits claimed before it by warps that are already running: the dependency
// wavefront of the reference's sequential loop (lib/synthesize.h:480-640) is respected exactly, with no
// barrier between "waves".  A visit waits only for the neighbours it actually reads (RAW); write-after-read
// hazards are removed by the two version slots of the state word (rs_device.cuh).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rs_cuda.h"
#include "rs_device.cuh"

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
#define RS_CHECK(call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char b_[512];                                                                           \
      snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      g_err = b_;                                                                             \
      return 100;                                                                             \
    }                                                                                         \
  } while (0)

extern "C" const char *rs_cuda_last_error(void) { return g_err.c_str(); }
extern "C" int rs_cuda_set_device(int ordinal) {
  RS_CHECK(cudaSetDevice(ordinal));
  return 0;
}
extern "C" int rs_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ------------------------------------------------------------------------------------------- init kernels
// Raw internal pixel [mask][colours][alpha?][maps] -> canonical corpus pixel.
__global__ void k_canon_corpus(const uint8_t *__restrict__ raw, int n_px, int bpp, int n_color, int n_map, int map_bip,
                               uint32_t *__restrict__ out4, uint2 *__restrict__ out8) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const uint8_t *p = raw + (size_t)i * bpp;
  uint32_t lo = p[0];
  for (int c = 0; c < n_color; c++) lo |= (uint32_t)p[1 + c] << (8 * (c + 1));
  if (out8) {
    uint32_t hi = 0;
    for (int c = 0; c < n_map; c++) hi |= (uint32_t)p[map_bip + c] << (8 * c);
    out8[i] = make_uint2(lo, hi);
  } else {
    out4[i] = lo;
  }
}

// Target image -> state words, meta, map bytes.  (lib/engine.c:338-391 hasValue rule, :207-224 sourceOf := none)
__global__ void k_init_target(const uint8_t *__restrict__ raw, int n_px, int bpp, int n_color, int n_map, int map_bip,
                              int alpha_bip, int alpha_target, int use_context, unsigned long long *__restrict__ W,
                              uint32_t *__restrict__ meta, uint32_t *__restrict__ tmaps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const uint8_t *p = raw + (size_t)i * bpp;
  uint32_t col = 0;
  for (int c = 0; c < n_color; c++) col |= (uint32_t)p[1 + c] << (8 * c);
  const unsigned long long none = (unsigned long long)RS_NO_SRC << 32;
  W[2 * (size_t)i] = none | col;                      // version 0
  W[2 * (size_t)i + 1] = none | col | (0xFFull << 24);  // invalid until version 1 is published
  const bool selected = p[0] != 0;
  const bool opaque = alpha_target ? (p[alpha_bip] != 0) : true;
  meta[i] = selected ? RS_PENDING : ((use_context && opaque) ? RS_CTX_VALUED : RS_NEVER);
  if (tmaps) {
    uint32_t m = 0;
    for (int c = 0; c < n_map; c++) m |= (uint32_t)p[map_bip + c] << (8 * c);
    tmaps[i] = m;
  }
}

__global__ void k_scatter_order(const uint32_t *__restrict__ targets, uint32_t n, int tw, uint32_t *__restrict__ meta) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = targets[i];
  meta[(size_t)(t >> 16) * tw + (t & 0xFFFFu)] = i;
}

__global__ void k_replicate_lut(const uint32_t *__restrict__ c256, const uint32_t *__restrict__ m256,
                                uint32_t *__restrict__ rep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < RS_LUT_WORDS) {
    rep[i] = c256[i >> 5];
    rep[RS_LUT_WORDS + i] = m256[i >> 5];
  }
}

__global__ void k_copy_u32(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t n, const RsCtrl *ctrl) {
  if (ctrl && ctrl->stop) return;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}

// Final colours and sources of the target points, from the newest published version of each.
__global__ void k_extract(const unsigned long long *__restrict__ W, const uint32_t *__restrict__ targets, uint32_t n,
                          int tw, uint32_t *__restrict__ colours, uint32_t *__restrict__ sources) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = targets[i];
  const size_t q = (size_t)(t >> 16) * tw + (t & 0xFFFFu);
  const unsigned long long a = W[2 * q], b = W[2 * q + 1];
  const unsigned va = (unsigned)(a >> 24) & 0xFFu, vb = (unsigned)(b >> 24) & 0xFFu;
  const unsigned long long w = (vb != 0xFFu && vb > va) ? b : a;
  colours[i] = (uint32_t)(w & 0xFFFFFFull);
  sources[i] = (uint32_t)(w >> 32);
}

// --------------------------------------------------------------------------------------- the pass kernel
#define RS_WARPS_PER_CTA 16
#define RS_THREADS (RS_WARPS_PER_CTA * 32)

struct WarpScratch {
  uint32_t off[RS_MAX_NB];   // neighbour offsets (packed int16 pair), ascending distance; [0] = (0,0)
  uint32_t pix[RS_MAX_NB];   // neighbour colours [0,c0,c1,c2]
  uint32_t map[RS_MAX_NB];   // neighbour map bytes
  uint32_t q[RS_MAX_NB];     // neighbour pixel index, later: packed heuristic candidate or RS_NO_SRC
  uint32_t aux[RS_MAX_NB];   // neighbour meta, later: neighbour source, later: compacted candidate list
};

template <bool MAPS>
__global__ void __launch_bounds__(RS_THREADS, 2) k_synth_pass(const RsDev J) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t *lutc = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *lutm = lutc + RS_LUT_WORDS;  // only staged when MAPS
  const unsigned lut_bytes = (MAPS ? 2u : 1u) * RS_LUT_WORDS * 4u;
  WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw + lut_bytes);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + lut_bytes + sizeof(WarpScratch) * RS_WARPS_PER_CTA);

  RsCtrl *ctrl = J.ctrl;
  if (rs_ld_u32_relaxed(&ctrl->stop)) return;

  // Stage the replicated metric tables with one TMA bulk copy per table.
  if (threadIdx.x == 0) {
    rs_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rs_mbar_expect_tx(bar, lut_bytes);
    rs_tma_load_1d(lutc, J.lut_rep, RS_LUT_WORDS * 4u, bar);
    if (MAPS) rs_tma_load_1d(lutm, J.lut_rep + RS_LUT_WORDS, RS_LUT_WORDS * 4u, bar);
  }
  __syncthreads();
  rs_mbar_wait(bar, 0);

  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  WarpScratch &S = scratch[warp];
  const uint32_t pass = J.pass, pass_end = J.pass_end;
  const uint32_t tag = (pass + 1u) << 29;

  uint32_t st_visits = 0, st_compares = 0, st_issued = 0, st_scans = 0, st_heur = 0, st_skips = 0, st_perfect = 0,
           st_betters = 0;
  unsigned long long st_evals = 0, st_sumbest = 0;

  while (true) {
    // ---- claim the next visit, in order (lib/synthesize.h:480-482 with THREAD_LIMIT 1)
    uint32_t v = 0;
    if (lane == 0) {
      v = atomicAdd(&ctrl->next[pass], 1u);
      if (v < pass_end && (v & 4095u) == 0u) {  // progress tick + cancel poll (synthesize.h:493-497)
        J.host_ticks[pass] = v + 1u;
        if (*J.host_cancel) {
          atomicExch(&ctrl->stop, 1u);
          atomicAdd(&ctrl->next[pass], 0x40000000u);
          v = 0xFFFFFFFFu;
        }
      }
    }
    v = __shfl_sync(RS_FULL, v, 0);
    if (v >= pass_end) break;
    st_visits++;

    const uint32_t tpos = __ldg(J.targets + v);
    const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
    const uint32_t selfq = (uint32_t)py * (uint32_t)J.tw + (uint32_t)px;

    // ---- gather the patch: self + nearest valued pixels (lib/synthesize.h:189-241)
    if (lane == 0) {
      S.off[0] = 0u;
      S.q[0] = selfq;
      S.aux[0] = v;
    }
    uint32_t count = 1;
    for (uint32_t base = 1; base < J.nOff && count < J.kmax; base += 32) {
      const uint32_t j = base + lane;
      bool ok = false;
      uint32_t o = 0, q = 0, m = 0;
      if (j < J.nOff) {
        o = __ldg(J.offsets + j);
        int x = px + rs_off_x(o), y = py + rs_off_y(o);
        bool in = true;  // wrap when tiling, else clip (lib/synthesize.h:81-113); |offset| < image size
        if (x < 0) { if (J.htile) x += J.tw; else in = false; }
        else if (x >= J.tw) { if (J.htile) x -= J.tw; else in = false; }
        if (y < 0) { if (J.vtile) y += J.th; else in = false; }
        else if (y >= J.th) { if (J.vtile) y -= J.th; else in = false; }
        if (in) {
          q = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
          m = __ldg(J.meta + q);
          // valued: usable context, or a target point already synthesised (in pass 0: visited before me)
          ok = (pass == 0u) ? (m == RS_CTX_VALUED || m < v) : (m != RS_NEVER);
        }
      }
      const unsigned b = __ballot_sync(RS_FULL, ok);
      const uint32_t slot = count + __popc(b & lt);
      if (ok && slot < J.kmax) {
        S.off[slot] = o;
        S.q[slot] = q;
        S.aux[slot] = m;
      }
      count += __popc(b);
      st_scans += (lane == 0) ? min(32u, J.nOff - base) : 0u;
    }
    const uint32_t K = min(count, J.kmax);
    __syncwarp();

    // ---- wait for exactly the versions the sequential order would see, then read them (one 64-bit load each)
    unsigned long long selfw = 0;
    for (uint32_t k = lane; k < K; k += 32) {
      const uint32_t q = S.q[k], m = S.aux[k];
      uint32_t r = 0;
      if (k == 0) r = pass;
      else if (m != RS_CTX_VALUED) {
        if (m < v && m < pass_end) r = pass + 1u;
        else for (uint32_t p2 = 0; p2 < pass; p2++) r += (m < J.ends[p2]) ? 1u : 0u;
      }
      const unsigned long long *wp = J.W + 2 * (size_t)q + (r & 1u);
      unsigned long long w = rs_ld_state(wp);
      while (((unsigned)(w >> 24) & 0xFFu) != r) {
        __nanosleep(40);
        w = rs_ld_state(wp);
      }
      if (k == 0) selfw = w;
      S.pix[k] = ((uint32_t)w & 0xFFFFFFu) << 8;
      if (MAPS) S.map[k] = __ldg(J.tmaps + q);
      S.aux[k] = (uint32_t)(w >> 32);  // source of this neighbour, or RS_NO_SRC
    }
    __syncwarp();

    // ---- heuristic 1 + 2 candidates (lib/synthesize.h:537-580): source of neighbour minus its offset,
    //      dropped if outside/masked corpus, if this target index was the last prober of that corpus point
    //      (snapshot at pass start), or if an earlier neighbour proposes the same point.
    uint32_t mycand[2];
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
      const uint32_t k = lane + 32u * rnd;
      uint32_t c = RS_NO_SRC;
      if (k < K) {
        const uint32_t src = S.aux[k];
        if (src != RS_NO_SRC) {
          const uint32_t o = S.off[k];
          const int x = (int)(src & 0xFFFFu) - rs_off_x(o), y = (int)(src >> 16) - rs_off_y(o);
          if ((unsigned)x < (unsigned)J.cw && (unsigned)y < (unsigned)J.ch) {
            const size_t a = (size_t)y * J.cw + x;
            const uint32_t cm = MAPS ? __ldg(&J.corpus8[a].x) : __ldg(J.corpus4 + a);
            if ((cm & 0xFFu) == 0xFFu) c = (uint32_t)x | ((uint32_t)y << 16);
          }
        }
      }
      mycand[rnd] = c;
      if (k < RS_MAX_NB) S.q[k] = c;
    }
    __syncwarp();
    uint32_t nHeur = 0;
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
      const uint32_t k = lane + 32u * rnd;
      const uint32_t c = mycand[rnd];
      bool valid = (k < K) && (c != RS_NO_SRC);
      if (valid) {
        const size_t a = (size_t)(c >> 16) * J.cw + (c & 0xFFFFu);
        const uint32_t pr = __ldg(J.proberA + a);  // 0 = never probed
        bool skip = (pr >> 29) != 0u && (pr & RS_IDX_MASK) == v;
        for (uint32_t k2 = 0; k2 < k && !skip; k2++) skip = (S.q[k2] == c);
        if (skip) { valid = false; st_skips++; }
      }
      const unsigned b = __ballot_sync(RS_FULL, valid);
      if (valid) S.aux[nHeur + __popc(b & lt)] = c;   // aux[] of lanes k>=32 is read above only in round 1 (own entry)
      nHeur += __popc(b);
      __syncwarp();
    }
    // NOTE: S.aux (sources) is overwritten by the compacted candidate list; sources are no longer needed,
    // and round 1 lanes read their own S.aux[k] before any lane writes (mycand computed above).

    // ---- evaluate: heuristic candidates first, then the random probes (lib/synthesize.h:583-604)
    uint32_t bestSum = 0xFFFFFFFFu;
    int bestIdx = 0x7FFFFFFF;
    const uint32_t *candlist = S.aux;
    rs_eval_range<MAPS>(J, lutc, lutm, S.off, S.pix, S.map, K, 0, (int)nHeur,
                        [&](int i) { return candlist[i]; }, bestSum, bestIdx, st_compares, st_issued);
    const uint32_t seed = J.seed, nC = J.nC;
    const uint32_t *cpts = J.corpus_pts;
    if (bestSum != 0u)
      rs_eval_range<MAPS>(J, lutc, lutm, S.off, S.pix, S.map, K, (int)nHeur, (int)(nHeur + J.probes),
                          [&](int i) { return __ldg(cpts + rs_range(rs_probe_hash(seed, pass, v, (uint32_t)i - nHeur), nC)); },
                          bestSum, bestIdx, st_compares, st_issued);

    // ---- commit (lib/synthesize.h:620-639): new colour + source only if the source changed; always publish
    const bool bettered = bestIdx != 0x7FFFFFFF;
    const uint32_t total = nHeur + J.probes;
    const uint32_t seq_evals = !bettered ? 0u : (bestSum == 0u ? (uint32_t)bestIdx + 1u : total);
    if (lane == 0) {
      uint32_t colour = (uint32_t)selfw & 0xFFFFFFu, src = (uint32_t)(selfw >> 32);
      if (bettered) {
        const uint32_t bp = ((uint32_t)bestIdx < nHeur)
                                ? candlist[bestIdx]
                                : __ldg(cpts + rs_range(rs_probe_hash(seed, pass, v, (uint32_t)bestIdx - nHeur), nC));
        if (bp != src) {
          const size_t a = (size_t)(bp >> 16) * J.cw + (bp & 0xFFFFu);
          const uint32_t cpx = MAPS ? __ldg(&J.corpus8[a].x) : __ldg(J.corpus4 + a);
          colour = cpx >> 8;
          src = bp;
          st_betters++;
        }
        st_sumbest += bestSum;
      }
      rs_st_state(J.W + 2 * (size_t)selfq + ((pass + 1u) & 1u),
                  ((unsigned long long)src << 32) | ((unsigned long long)(pass + 1u) << 24) | colour);
      st_evals += seq_evals;
      st_heur += min(nHeur, seq_evals);
      st_perfect += (bettered && bestSum == 0u) ? 1u : 0u;
    }
    // ---- heuristic 2 bookkeeping: stamp the evaluated heuristic candidates before the perfect one, if any
    const uint32_t stampEnd = (bettered && bestSum == 0u && (uint32_t)bestIdx < nHeur) ? (uint32_t)bestIdx : nHeur;
    for (uint32_t i = lane; i < stampEnd; i += 32) {
      const uint32_t c = candlist[i];
      atomicMax(J.proberB + (size_t)(c >> 16) * J.cw + (c & 0xFFFFu), tag | v);
    }
    __syncwarp();
  }

  // ---- flush per-warp statistics
  st_compares = __reduce_add_sync(RS_FULL, st_compares);
  st_issued = __reduce_add_sync(RS_FULL, st_issued);
  st_skips = __reduce_add_sync(RS_FULL, st_skips);
  if (lane == 0 && st_visits) {
    atomicAdd(&ctrl->visits, (unsigned long long)st_visits);
    atomicAdd(&ctrl->pass_visits[pass], (unsigned long long)st_visits);
    atomicAdd(&ctrl->evals, st_evals);
    atomicAdd(&ctrl->evals_issued, (unsigned long long)st_issued);
    atomicAdd(&ctrl->compares, (unsigned long long)st_compares);
    atomicAdd(&ctrl->offset_scans, (unsigned long long)st_scans);
    atomicAdd(&ctrl->heur_evals, (unsigned long long)st_heur);
    atomicAdd(&ctrl->heur_skips, (unsigned long long)st_skips);
    atomicAdd(&ctrl->perfect, (unsigned long long)st_perfect);
    atomicAdd(&ctrl->sum_best[pass], st_sumbest);
    atomicAdd(&ctrl->betters[pass], st_betters);
  }
  // ---- last CTA out decides whether later passes run (lib/refiner.h:111): (float)betters/n < 0.1
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&ctrl->done_ctas[pass], 1u) + 1u;
    if (done == gridDim.x) {
      __threadfence();
      const unsigned b = atomicAdd(&ctrl->betters[pass], 0u);
      ctrl->passes_run = pass + 1u;
      if ((double)((float)b / (float)J.nT) < J.terminate_fraction) atomicExch(&ctrl->stop, 1u);
    }
  }
}

// ------------------------------------------------------------------------------ standalone best-fit kernel
template <bool MAPS>
__global__ void __launch_bounds__(RS_THREADS, 2)
    k_bestfit_batch(const RsDev J, uint32_t n_visits, const uint32_t *__restrict__ nb_begin,
                    const uint32_t *__restrict__ nb_offsets, const uint8_t *__restrict__ nb_pixels, int n_color,
                    int n_map, int map_bip, const uint32_t *__restrict__ cand_begin, const uint32_t *__restrict__ cands,
                    uint32_t *__restrict__ best_sum, int32_t *__restrict__ best_index) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t *lutc = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *lutm = lutc + RS_LUT_WORDS;
  const unsigned lut_bytes = (MAPS ? 2u : 1u) * RS_LUT_WORDS * 4u;
  WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw + lut_bytes);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + lut_bytes + sizeof(WarpScratch) * RS_WARPS_PER_CTA);
  if (threadIdx.x == 0) {
    rs_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rs_mbar_expect_tx(bar, lut_bytes);
    rs_tma_load_1d(lutc, J.lut_rep, RS_LUT_WORDS * 4u, bar);
    if (MAPS) rs_tma_load_1d(lutm, J.lut_rep + RS_LUT_WORDS, RS_LUT_WORDS * 4u, bar);
  }
  __syncthreads();
  rs_mbar_wait(bar, 0);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  WarpScratch &S = scratch[warp];
  for (uint32_t v = blockIdx.x * RS_WARPS_PER_CTA + warp; v < n_visits; v += gridDim.x * RS_WARPS_PER_CTA) {
    const uint32_t nb0 = nb_begin[v], K = min(nb_begin[v + 1] - nb0, (uint32_t)RS_MAX_NB);
    for (uint32_t k = lane; k < K; k += 32) {
      S.off[k] = nb_offsets[nb0 + k];
      const uint8_t *p = nb_pixels + (size_t)(nb0 + k) * 8;
      uint32_t col = 0, mp = 0;
      for (int c = 0; c < n_color; c++) col |= (uint32_t)p[1 + c] << (8 * (c + 1));
      for (int c = 0; c < n_map; c++) mp |= (uint32_t)p[map_bip + c] << (8 * c);
      S.pix[k] = col;
      S.map[k] = mp;
    }
    __syncwarp();
    const uint32_t c0 = cand_begin[v], nc = cand_begin[v + 1] - c0;
    uint32_t bestSum = 0xFFFFFFFFu, cmp = 0, iss = 0;
    int bestIdx = 0x7FFFFFFF;
    rs_eval_range<MAPS>(J, lutc, lutm, S.off, S.pix, S.map, K, 0, (int)nc,
                        [&](int i) { return __ldg(cands + c0 + i); }, bestSum, bestIdx, cmp, iss);
    if (lane == 0) {
      best_sum[v] = bestSum;
      best_index[v] = (bestIdx == 0x7FFFFFFF) ? -1 : bestIdx;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ the job
struct RsJob {
  RsJobDesc d;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool maps = false;
  // device buffers
  uint8_t *d_target_raw = nullptr, *d_corpus_raw = nullptr;
  uint32_t *d_corpus4 = nullptr;
  uint2 *d_corpus8 = nullptr;
  unsigned long long *d_W = nullptr;
  uint32_t *d_meta = nullptr, *d_tmaps = nullptr, *d_targets = nullptr, *d_cpts = nullptr, *d_offsets = nullptr;
  uint32_t *d_lut256 = nullptr, *d_lut_rep = nullptr, *d_prober[2] = {nullptr, nullptr};
  uint32_t *d_colours = nullptr, *d_sources = nullptr;
  RsCtrl *d_ctrl = nullptr;
  // mapped pinned host words
  unsigned int *h_ticks = nullptr;  // [6]
  int *h_cancel = nullptr;
  RsCtrl *h_ctrl = nullptr;         // pinned copy of the control block, read back after the run
  uint32_t nT = 0, nC = 0, nOff = 0, penalty = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evDone = nullptr;
  float ms_passes = 0.f;
  int grid = 0;
  size_t smem = 0;
};

static size_t pass_smem(bool maps) {
  return (maps ? 2u : 1u) * RS_LUT_WORDS * 4u + sizeof(WarpScratch) * RS_WARPS_PER_CTA + 16;
}

template <bool MAPS>
static int configure_pass_kernel(RsJob *j) {
  size_t smem = pass_smem(MAPS);
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<MAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, sms = 0;
  RS_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_synth_pass<MAPS>, RS_THREADS, smem));
  RS_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, j->device));
  if (per_sm < 1) { g_err = "k_synth_pass does not fit on an SM"; return 100; }
  j->grid = per_sm * sms;
  j->smem = smem;
  return 0;
}

extern "C" void rs_job_destroy(RsJob *j) {
  if (!j) return;
  cudaSetDevice(j->device);
  if (j->stream) cudaStreamSynchronize(j->stream);
  cudaFree(j->d_target_raw); cudaFree(j->d_corpus_raw); cudaFree(j->d_corpus4); cudaFree(j->d_corpus8);
  cudaFree(j->d_W); cudaFree(j->d_meta); cudaFree(j->d_tmaps); cudaFree(j->d_targets); cudaFree(j->d_cpts);
  cudaFree(j->d_offsets); cudaFree(j->d_lut256); cudaFree(j->d_lut_rep); cudaFree(j->d_prober[0]);
  cudaFree(j->d_prober[1]); cudaFree(j->d_colours); cudaFree(j->d_sources); cudaFree(j->d_ctrl);
  if (j->h_ticks) cudaFreeHost(j->h_ticks);
  if (j->h_cancel) cudaFreeHost(j->h_cancel);
  if (j->h_ctrl) cudaFreeHost(j->h_ctrl);
  if (j->ev0) cudaEventDestroy(j->ev0);
  if (j->ev1) cudaEventDestroy(j->ev1);
  if (j->evDone) cudaEventDestroy(j->evDone);
  if (j->stream) cudaStreamDestroy(j->stream);
  delete j;
}

extern "C" int rs_job_create(const RsJobDesc *desc, RsJob **out) {
  *out = nullptr;
  if (desc->tw <= 0 || desc->th <= 0 || desc->cw <= 0 || desc->ch <= 0 || desc->tw > 65535 || desc->th > 65535 ||
      desc->cw > 65535 || desc->ch > 65535 || desc->bpp < 2 || desc->bpp > 8 || desc->n_color < 1 || desc->n_color > 3 ||
      desc->n_map < 0 || desc->n_map > 3 || desc->n_passes < 1 || desc->n_passes > 6) {
    g_err = "rs_job_create: descriptor out of range";
    return 100;
  }
  RsJob *j = new RsJob();
  j->d = *desc;
  j->maps = desc->n_map > 0;
  if (cudaGetDevice(&j->device) != cudaSuccess) { g_err = "no CUDA device"; delete j; return 100; }
  int rc = j->maps ? configure_pass_kernel<true>(j) : configure_pass_kernel<false>(j);
  if (rc) { delete j; return rc; }
#define JCHK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); rs_job_destroy(j); return 100; } } while (0)
  JCHK(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
  JCHK(cudaEventCreate(&j->ev0)); JCHK(cudaEventCreate(&j->ev1));
  JCHK(cudaEventCreateWithFlags(&j->evDone, cudaEventDisableTiming));
  const size_t tn = (size_t)desc->tw * desc->th, cn = (size_t)desc->cw * desc->ch;
  JCHK(cudaMalloc(&j->d_target_raw, tn * desc->bpp));
  JCHK(cudaMalloc(&j->d_corpus_raw, cn * desc->bpp));
  if (j->maps) JCHK(cudaMalloc(&j->d_corpus8, cn * sizeof(uint2))); else JCHK(cudaMalloc(&j->d_corpus4, cn * 4));
  JCHK(cudaMalloc(&j->d_W, tn * 16));
  JCHK(cudaMalloc(&j->d_meta, tn * 4));
  if (j->maps) JCHK(cudaMalloc(&j->d_tmaps, tn * 4));
  JCHK(cudaMalloc(&j->d_lut256, 512 * 4));
  JCHK(cudaMalloc(&j->d_lut_rep, 2 * RS_LUT_WORDS * 4));
  JCHK(cudaMalloc(&j->d_prober[0], cn * 4)); JCHK(cudaMalloc(&j->d_prober[1], cn * 4));
  JCHK(cudaMalloc(&j->d_ctrl, sizeof(RsCtrl)));
  JCHK(cudaHostAlloc(&j->h_ticks, 6 * sizeof(unsigned int), cudaHostAllocMapped));
  JCHK(cudaHostAlloc(&j->h_cancel, sizeof(int), cudaHostAllocMapped));
  JCHK(cudaHostAlloc(&j->h_ctrl, sizeof(RsCtrl), cudaHostAllocDefault));
#undef JCHK
  *out = j;
  return 0;
}

extern "C" int rs_job_upload(RsJob *j, const uint8_t *target_raw, const uint8_t *corpus_raw, const uint32_t *targets,
                             uint32_t n_targets, const uint32_t *corpus_points, uint32_t n_corpus,
                             const uint32_t *offsets, uint32_t n_offsets, const uint32_t *color_lut256,
                             const uint32_t *map_lut256, uint32_t map_lut_max) {
  RS_CHECK(cudaSetDevice(j->device));
  const RsJobDesc &d = j->d;
  if (n_targets == 0 || n_corpus == 0 || n_offsets == 0 || n_targets >= RS_IDX_MASK) {
    g_err = "rs_job_upload: empty or oversized point list";
    return 100;
  }
  const size_t tn = (size_t)d.tw * d.th, cn = (size_t)d.cw * d.ch;
  cudaStream_t s = j->stream;
  j->nT = n_targets; j->nC = n_corpus; j->nOff = n_offsets;
  j->penalty = 65535u * (uint32_t)d.n_color + map_lut_max * (uint32_t)d.n_map;
  cudaFree(j->d_targets); cudaFree(j->d_cpts); cudaFree(j->d_offsets); cudaFree(j->d_colours); cudaFree(j->d_sources);
  j->d_targets = j->d_cpts = j->d_offsets = j->d_colours = j->d_sources = nullptr;
  RS_CHECK(cudaMalloc(&j->d_targets, (size_t)n_targets * 4));
  RS_CHECK(cudaMalloc(&j->d_cpts, (size_t)n_corpus * 4));
  RS_CHECK(cudaMalloc(&j->d_offsets, (size_t)n_offsets * 4));
  RS_CHECK(cudaMalloc(&j->d_colours, (size_t)n_targets * 4));
  RS_CHECK(cudaMalloc(&j->d_sources, (size_t)n_targets * 4));
  RS_CHECK(cudaMemcpyAsync(j->d_target_raw, target_raw, tn * d.bpp, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_corpus_raw, corpus_raw, cn * d.bpp, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_targets, targets, (size_t)n_targets * 4, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_cpts, corpus_points, (size_t)n_corpus * 4, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_offsets, offsets, (size_t)n_offsets * 4, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_lut256, color_lut256, 256 * 4, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemcpyAsync(j->d_lut256 + 256, map_lut256, 256 * 4, cudaMemcpyHostToDevice, s));
  RS_CHECK(cudaMemsetAsync(j->d_ctrl, 0, sizeof(RsCtrl), s));
  RS_CHECK(cudaMemsetAsync(j->d_prober[0], 0, cn * 4, s));
  RS_CHECK(cudaMemsetAsync(j->d_prober[1], 0, cn * 4, s));
  const int T = 256;
  k_canon_corpus<<<(unsigned)((cn + T - 1) / T), T, 0, s>>>(j->d_corpus_raw, (int)cn, d.bpp, d.n_color, d.n_map, d.map_bip,
                                                          j->d_corpus4, j->d_corpus8);
  k_init_target<<<(unsigned)((tn + T - 1) / T), T, 0, s>>>(j->d_target_raw, (int)tn, d.bpp, d.n_color, d.n_map, d.map_bip,
                                                         d.alpha_bip, d.alpha_target, d.use_context, j->d_W, j->d_meta,
                                                         j->d_tmaps);
  k_scatter_order<<<(n_targets + T - 1) / T, T, 0, s>>>(j->d_targets, n_targets, d.tw, j->d_meta);
  k_replicate_lut<<<(RS_LUT_WORDS + T - 1) / T, T, 0, s>>>(j->d_lut256, j->d_lut256 + 256, j->d_lut_rep);
  RS_CHECK(cudaGetLastError());
  for (int p = 0; p < 6; p++) j->h_ticks[p] = 0;
  *j->h_cancel = 0;
  return 0;
}

static RsDev make_dev(const RsJob *j, uint32_t pass) {
  RsDev D;
  memset(&D, 0, sizeof D);
  const RsJobDesc &d = j->d;
  D.corpus4 = j->d_corpus4; D.corpus8 = j->d_corpus8; D.W = j->d_W; D.meta = j->d_meta; D.tmaps = j->d_tmaps;
  D.targets = j->d_targets; D.corpus_pts = j->d_cpts; D.offsets = j->d_offsets; D.lut_rep = j->d_lut_rep;
  D.proberA = j->d_prober[pass & 1]; D.proberB = j->d_prober[(pass + 1) & 1];
  D.ctrl = j->d_ctrl; D.host_ticks = j->h_ticks; D.host_cancel = j->h_cancel;
  D.tw = d.tw; D.th = d.th; D.cw = d.cw; D.ch = d.ch;
  D.nT = j->nT; D.nC = j->nC; D.nOff = j->nOff;
  uint32_t kmax = d.patch_size < 2 ? 2 : d.patch_size;  // the size test follows the append (synthesize.h:222-224)
  D.kmax = kmax > RS_MAX_NB ? RS_MAX_NB : kmax;
  D.probes = d.max_probes; D.seed = d.seed; D.penalty = j->penalty;
  D.pass = pass; D.pass_end = d.pass_end[pass];
  for (int p = 0; p < 6; p++) D.ends[p] = d.pass_end[p];
  D.htile = d.htile; D.vtile = d.vtile; D.terminate_fraction = d.terminate_fraction;
  return D;
}

extern "C" int rs_job_run(RsJob *j, RsTickFn tick, void *tick_ctx) {
  RS_CHECK(cudaSetDevice(j->device));
  cudaStream_t s = j->stream;
  const size_t cn = (size_t)j->d.cw * j->d.ch;
  RS_CHECK(cudaEventRecord(j->ev0, s));
  for (uint32_t p = 0; p < j->d.n_passes; p++) {
    if (p > 0)  // B := A before the pass stamps into B (pass-snapshot semantics of heuristic 2)
      k_copy_u32<<<1184, 256, 0, s>>>(j->d_prober[p & 1], j->d_prober[(p + 1) & 1], cn, j->d_ctrl);
    RsDev D = make_dev(j, p);
    if (j->maps) k_synth_pass<true><<<j->grid, RS_THREADS, j->smem, s>>>(D);
    else k_synth_pass<false><<<j->grid, RS_THREADS, j->smem, s>>>(D);
  }
  RS_CHECK(cudaGetLastError());
  RS_CHECK(cudaEventRecord(j->ev1, s));
  k_extract<<<(j->nT + 255) / 256, 256, 0, s>>>(j->d_W, j->d_targets, j->nT, j->d.tw, j->d_colours, j->d_sources);
  RS_CHECK(cudaMemcpyAsync(j->h_ctrl, j->d_ctrl, sizeof(RsCtrl), cudaMemcpyDeviceToHost, s));
  RS_CHECK(cudaEventRecord(j->evDone, s));
  // Host side of the progress/cancel contract: replay ticks in order while the device runs.
  uint32_t emitted[6] = {0, 0, 0, 0, 0, 0};  // number of ticks already forwarded per pass
  bool cancelled = false;
  auto drain = [&](bool final_) {
    for (uint32_t p = 0; p < j->d.n_passes; p++) {
      const unsigned int seen = ((volatile unsigned int *)j->h_ticks)[p];  // highest started tick index + 1
      if (seen == 0) { if (!final_) break; else continue; }
      const uint32_t upto = (seen - 1u) / 4096u + 1u;  // ticks 0..upto-1 have started
      while (emitted[p] < upto) {
        const uint32_t idx = emitted[p] * 4096u;
        emitted[p]++;
        if (tick && !cancelled && tick(tick_ctx, p, idx)) {
          cancelled = true;
          *(volatile int *)j->h_cancel = 1;
        }
      }
    }
  };
  while (true) {
    cudaError_t q = cudaEventQuery(j->evDone);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) { g_err = std::string("rs_job_run: ") + cudaGetErrorString(q); return 100; }
    drain(false);
  }
  drain(true);
  RS_CHECK(cudaStreamSynchronize(s));
  RS_CHECK(cudaEventElapsedTime(&j->ms_passes, j->ev0, j->ev1));
  return 0;
}

extern "C" int rs_job_download(RsJob *j, uint32_t *colours_out, uint32_t *sources_out) {
  RS_CHECK(cudaSetDevice(j->device));
  RS_CHECK(cudaMemcpyAsync(colours_out, j->d_colours, (size_t)j->nT * 4, cudaMemcpyDeviceToHost, j->stream));
  if (sources_out)
    RS_CHECK(cudaMemcpyAsync(sources_out, j->d_sources, (size_t)j->nT * 4, cudaMemcpyDeviceToHost, j->stream));
  RS_CHECK(cudaStreamSynchronize(j->stream));
  return 0;
}

extern "C" int rs_job_counters(RsJob *j, RsJobCounters *out) {
  const RsCtrl &c = *j->h_ctrl;
  memset(out, 0, sizeof *out);
  out->visits = c.visits; out->evals = c.evals; out->evals_issued = c.evals_issued; out->compares = c.compares;
  out->offset_scans = c.offset_scans; out->heur_evals = c.heur_evals; out->heur_skips = c.heur_skips;
  out->perfect = c.perfect;
  for (int p = 0; p < 6; p++) { out->betters[p] = c.betters[p]; out->pass_visits[p] = c.pass_visits[p]; out->sum_best[p] = c.sum_best[p]; }
  out->passes_run = c.passes_run;
  out->ms_passes = j->ms_passes;
  return 0;
}

// ------------------------------------------------------------------------------------- rs_bestfit_batch
extern "C" int rs_bestfit_batch(const RsJobDesc *desc, const uint8_t *corpus_raw, const uint32_t *color_lut256,
                                const uint32_t *map_lut256, uint32_t map_lut_max, uint32_t n_visits,
                                const uint32_t *nb_begin, const uint32_t *nb_offsets, const uint8_t *nb_pixels,
                                const uint32_t *cand_begin, const uint32_t *cands, uint32_t *best_sum_out,
                                int32_t *best_index_out) {
  if (n_visits == 0) return 0;
  const bool maps = desc->n_map > 0;
  const size_t cn = (size_t)desc->cw * desc->ch;
  const uint32_t n_nb = nb_begin[n_visits], n_cand = cand_begin[n_visits];
  uint8_t *d_raw = nullptr, *d_nbpix = nullptr;
  uint32_t *d_c4 = nullptr, *d_lut = nullptr, *d_rep = nullptr, *d_nbb = nullptr, *d_nbo = nullptr, *d_cb = nullptr,
           *d_c = nullptr, *d_bs = nullptr;
  int32_t *d_bi = nullptr;
  uint2 *d_c8 = nullptr;
  int rc = 0;
#define BCHK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); rc = 100; goto done; } } while (0)
  BCHK(cudaMalloc(&d_raw, cn * desc->bpp));
  if (maps) BCHK(cudaMalloc(&d_c8, cn * 8)); else BCHK(cudaMalloc(&d_c4, cn * 4));
  BCHK(cudaMalloc(&d_lut, 512 * 4)); BCHK(cudaMalloc(&d_rep, 2 * RS_LUT_WORDS * 4));
  BCHK(cudaMalloc(&d_nbb, (size_t)(n_visits + 1) * 4)); BCHK(cudaMalloc(&d_cb, (size_t)(n_visits + 1) * 4));
  BCHK(cudaMalloc(&d_nbo, (size_t)(n_nb + 1) * 4)); BCHK(cudaMalloc(&d_nbpix, (size_t)(n_nb + 1) * 8));
  BCHK(cudaMalloc(&d_c, (size_t)(n_cand + 1) * 4));
  BCHK(cudaMalloc(&d_bs, (size_t)n_visits * 4)); BCHK(cudaMalloc(&d_bi, (size_t)n_visits * 4));
  BCHK(cudaMemcpy(d_raw, corpus_raw, cn * desc->bpp, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_lut, color_lut256, 256 * 4, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_lut + 256, map_lut256, 256 * 4, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_nbb, nb_begin, (size_t)(n_visits + 1) * 4, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_cb, cand_begin, (size_t)(n_visits + 1) * 4, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_nbo, nb_offsets, (size_t)n_nb * 4, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_nbpix, nb_pixels, (size_t)n_nb * 8, cudaMemcpyHostToDevice));
  BCHK(cudaMemcpy(d_c, cands, (size_t)n_cand * 4, cudaMemcpyHostToDevice));
  {
    k_canon_corpus<<<(unsigned)((cn + 255) / 256), 256>>>(d_raw, (int)cn, desc->bpp, desc->n_color, desc->n_map,
                                                        desc->map_bip, d_c4, d_c8);
    k_replicate_lut<<<(RS_LUT_WORDS + 255) / 256, 256>>>(d_lut, d_lut + 256, d_rep);
    RsDev D;
    memset(&D, 0, sizeof D);
    D.corpus4 = d_c4; D.corpus8 = d_c8; D.lut_rep = d_rep; D.cw = desc->cw; D.ch = desc->ch;
    D.penalty = 65535u * (uint32_t)desc->n_color + map_lut_max * (uint32_t)desc->n_map;
    const size_t smem = pass_smem(maps);
    const unsigned grid = (n_visits + RS_WARPS_PER_CTA - 1) / RS_WARPS_PER_CTA;
    if (maps) {
      BCHK(cudaFuncSetAttribute(k_bestfit_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_bestfit_batch<true><<<grid > 1184 ? 1184 : grid, RS_THREADS, smem>>>(D, n_visits, d_nbb, d_nbo, d_nbpix, desc->n_color, desc->n_map,
                                                       desc->map_bip, d_cb, d_c, d_bs, d_bi);
    } else {
      BCHK(cudaFuncSetAttribute(k_bestfit_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_bestfit_batch<false><<<grid > 1184 ? 1184 : grid, RS_THREADS, smem>>>(D, n_visits, d_nbb, d_nbo, d_nbpix, desc->n_color, desc->n_map,
                                                        desc->map_bip, d_cb, d_c, d_bs, d_bi);
    }
    BCHK(cudaGetLastError());
    BCHK(cudaMemcpy(best_sum_out, d_bs, (size_t)n_visits * 4, cudaMemcpyDeviceToHost));
    BCHK(cudaMemcpy(best_index_out, d_bi, (size_t)n_visits * 4, cudaMemcpyDeviceToHost));
  }
done:
#undef BCHK
  cudaFree(d_raw); cudaFree(d_c4); cudaFree(d_c8); cudaFree(d_lut); cudaFree(d_rep); cudaFree(d_nbb); cudaFree(d_cb);
  cudaFree(d_nbo); cudaFree(d_nbpix); cudaFree(d_c); cudaFree(d_bs); cudaFree(d_bi);
  return rc;
}
