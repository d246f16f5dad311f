// sm_100a kernels of the synthesis engine and the thin C-ABI over them (include/rs_cuda.h).
//
// A pass is one or a few persistent launches (segments of shrinking team width, see plan_segments).  Every warp -- or
// team of 2/4/8 warps, or half of a warp for small patches -- claims target visits IN ORDER (atomic counter), so a visit can only ever wait on visits
// claimed before it by warps that are already running: the dependency wavefront of the reference's sequential loop
// (lib/synthesize.h:480-640) is respected exactly, with no barrier between "waves".  A visit waits only for the
// neighbours it actually reads (RAW); write-after-read hazards are removed by the two version slots of the state word
// (rs_device.cuh).  File order: init kernels, pass-0 / later-pass patch gathers (scan, and search for the first visits),
// visit steps (geometry, values, candidates, finish; templated on the lanes that work on a visit), the two pass kernels, best-fit test kernel, workspaces, visit-order machinery (digest, cache,
// exact device shuffle, pair sort), staging / upload, rs_job_run, download, counters.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "../../include/rs_cuda.h"
#include "rs_device.cuh"

namespace rs {
// csrc/host_prep.cpp: set-bit positions of z^(q * jump_words) mod MT19937's characteristic polynomial (jump-ahead)
const std::vector<uint16_t> &mt_jump_poly(uint32_t q, uint32_t jump_words);
}

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
extern "C" const char *rs_cuda_peek_error(void) {
  const cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
extern "C" int rs_cuda_set_device(int ordinal) {
  RS_CHECK(cudaSetDevice(ordinal));
  return 0;
}
// Frees and evictions may run on a thread that is working on another device: they select the owning device for the
// call and put the caller's back (the current device is per-thread state that ws_acquire and the launches rely on).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) cudaSetDevice(device); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// Host cores this process may use for its own helper threads (staging copies, ordering keys): the machine's cores
// divided among the processes that share it -- one process per GPU under torchrun (LOCAL_WORLD_SIZE), where 8 ranks
// x 4 staging threads + 8 PRNG producers on 32 cores slowed one another's copies (SCALE_r01: e2e 0.93 at 8 GPUs).
// RS_HOST_THREADS overrides.
extern "C" unsigned rs_host_cores(void) {
  static const unsigned cached = []() -> unsigned {
    if (const char *e = getenv("RS_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return (unsigned)v; }
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    unsigned procs = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) procs = (unsigned)v; }
    const unsigned c = hw / procs;
    return c ? c : 1u;
  }();
  return cached;
}
// Most threads one staging copy (host buffer <-> pinned memory) is spread over; RS_COPY_THREADS overrides (sweeps).
static size_t rs_copy_threads_max() {
  static const size_t cached = []() -> size_t {
    if (const char *e = getenv("RS_COPY_THREADS")) { const int v = atoi(e); if (v > 0) return (size_t)v; }
    return 4;
  }();
  return cached;
}
extern "C" int rs_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ------------------------------------------------------------------------------------------- init kernels
// Raw internal pixel [mask][colours][alpha?][maps] -> canonical corpus pixel [c0,c1,c2,mask | maps]; thread n_px
// writes the sentinel pixel (not selected) that out-of-corpus compares read.
__global__ void k_canon_corpus(const uint8_t *__restrict__ raw, int n_px, int bpp, int n_color, int n_map, int map_bip,
                               uint32_t *__restrict__ out4, uint2 *__restrict__ out8) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_px) return;
  uint32_t lo = 0, hi = 0;
  if (i < n_px) {
    const uint8_t *p = raw + (size_t)i * bpp;
    lo = (uint32_t)p[0] << 24;
    for (int c = 0; c < n_color; c++) lo |= (uint32_t)p[1 + c] << (8 * c);
    if (out8)
      for (int c = 0; c < n_map; c++) hi |= (uint32_t)p[map_bip + c] << (8 * c);
  }
  if (out8) out8[i] = make_uint2(lo, hi);
  else out4[i] = lo;
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
    rep[i] = c256[i / RS_LUT_REP];
    rep[RS_LUT_WORDS + i] = m256[i / RS_LUT_REP];
  }
}

// Final colours of the target points, from the newest published version of each, written into the raw target
// pixmap (colour bytes only: alpha and maps are never synthesised, lib/synthesize.h:403-419); sources optional.
__global__ void k_writeback(const unsigned long long *__restrict__ W, const uint32_t *__restrict__ targets, uint32_t n,
                            int tw, int bpp, int n_color, uint8_t *__restrict__ raw, uint32_t *__restrict__ sources) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = targets[i];
  const size_t q = (size_t)(t >> 16) * tw + (t & 0xFFFFu);
  const unsigned long long a = W[2 * q], b = W[2 * q + 1];
  const unsigned va = (unsigned)(a >> 24) & 0xFFu, vb = (unsigned)(b >> 24) & 0xFFu;
  const unsigned long long w = (vb != 0xFFu && vb > va) ? b : a;
  uint8_t *p = raw + q * bpp;
  for (int c = 0; c < n_color; c++) p[1 + c] = (uint8_t)(w >> (8 * c));
  if (sources) sources[i] = (uint32_t)(w >> 32);
}

// ------------------------------------------------------------------------------ pass-0 patch precompute
// In pass 0 a target point may only use context pixels and target points visited BEFORE it (hasValue is set
// after the visit, lib/synthesize.h:639), so which pixels form the patch of visit v depends on the visit order
// alone, not on any synthesised colour.  All pass-0 patches are therefore gathered up front by a kernel with no
// dependencies at all; the (long, deep-hole) offset scans leave the critical path of the ordered pass.
// Output per visit: count-1 entries {packed offset, pixel index | target flag << 31}.
#define RS_TARGET_FLAG 0x80000000u
// One step of a pass-0 scan by one warp: U table entries per lane from `base` on (all of their loads in flight together),
// the valued ones appended in table order to out[count - 1 ...]; returns the new count.
template <int U>
__device__ __forceinline__ uint32_t rs_scan_step(const RsDev &J, const uint32_t v, const int px, const int py, const uint32_t base,
                                                 uint32_t count, uint2 *__restrict__ out) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t o[U], q[U], m[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t j = base + 32u * u + lane;
    o[u] = 0; q[u] = 0; m[u] = RS_NEVER;
    if (j < J.nOff) {
      o[u] = __ldg(J.offsets + j);
      int x = px + rs_off_x(o[u]), y = py + rs_off_y(o[u]);
      bool in = true;
      if (x < 0) { if (J.htile) x += J.tw; else in = false; }
      else if (x >= J.tw) { if (J.htile) x -= J.tw; else in = false; }
      if (y < 0) { if (J.vtile) y += J.th; else in = false; }
      else if (y >= J.th) { if (J.vtile) y -= J.th; else in = false; }
      if (in) {
        q[u] = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
        m[u] = __ldg(J.meta + q[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const bool ok = (m[u] == RS_CTX_VALUED) || (m[u] < v);
    const unsigned b = __ballot_sync(RS_FULL, ok);
    const uint32_t slot = count + __popc(b & lt);
    if (ok && slot < J.kmax) out[slot - 1u] = make_uint2(o[u], q[u] | (m[u] == RS_CTX_VALUED ? 0u : RS_TARGET_FLAG));
    count += __popc(b);
  }
  return count;
}
#define RS_GATHER_STRIPES 32768u
__global__ void __launch_bounds__(256) k_gather_pass0(const RsDev J, uint2 *__restrict__ lists, uint8_t *__restrict__ counts,
                                                      const uint32_t v_begin, const uint32_t v_end, unsigned int *__restrict__ claim) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t stride = J.kmax - 1u;
  unsigned long long scans = 0;
  // The long scans of the first visits are not in this range (k_gather_pass0_sparse / _coop).  The rest is cut into
  // RS_GATHER_STRIPES stripes -- stripe s = the visits v_begin + s + k * RS_GATHER_STRIPES, the same mix of long and short
  // scans in each -- that the warps claim one at a time: CTAs that only find room on an SM once the search kernel beside
  // them has left take fewer stripes, not a fixed share that would finish late.
  while (true) {
    uint32_t sidx = 0;
    if (lane == 0) sidx = atomicAdd(claim, 1u);
    sidx = __shfl_sync(RS_FULL, sidx, 0);
    if (sidx >= RS_GATHER_STRIPES) break;
    for (uint32_t v = v_begin + sidx; v < v_end; v += RS_GATHER_STRIPES) {
    const uint32_t tpos = __ldg(J.targets + v);
    const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
    uint2 *out = lists + (size_t)v * stride;
    // The kernel is bound by the L2 sectors of its scattered meta-word loads, so what a step tests beyond the patch is what
    // it costs: steps of 32, 32 and 64 entries first (the second half of a pass needs fewer than 64), then 128-wide ones.
    uint32_t count = rs_scan_step<1>(J, v, px, py, 1u, 1u, out), base = 33u;
    if (count < J.kmax && base < J.nOff) { count = rs_scan_step<1>(J, v, px, py, base, count, out); base += 32u; }
    if (count < J.kmax && base < J.nOff) { count = rs_scan_step<2>(J, v, px, py, base, count, out); base += 64u; }
    for (; base < J.nOff && count < J.kmax; base += 128u) count = rs_scan_step<4>(J, v, px, py, base, count, out);
    scans += base - 1u;
    if (lane == 0) counts[v] = (uint8_t)min(count, J.kmax);
    }
  }
  if (lane == 0 && scans) atomicAdd(&J.ctrl->offset_scans, scans);
}

// From pass 1 on every target point has a value, so the patch of a target is the same in all later passes: the nearest
// pixels that are usable context or target points at all.  Gathered once (beside pass 0, on the side stream); the pass
// kernels then read 8 bytes per neighbour in one contiguous piece instead of scanning offsets and gathering a meta
// word for each.  Entries {offset, meta word}: the meta word (visit index or context) drives the version arithmetic.
__global__ void __launch_bounds__(256) k_gather_later(const RsDev J, uint2 *__restrict__ lists, uint8_t *__restrict__ counts) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t stride = J.kmax - 1u;
  const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < J.nT; v += nwarps) {
    const uint32_t tpos = __ldg(J.targets + v);
    const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
    uint2 *out = lists + (size_t)v * stride;
    uint32_t count = 1;
    for (uint32_t base = 1; base < J.nOff && count < J.kmax; base += 32) {
      const uint32_t j = base + lane;
      uint32_t o = 0, m = RS_NEVER;
      if (j < J.nOff) {
        o = __ldg(J.offsets + j);
        int x = px + rs_off_x(o), y = py + rs_off_y(o);
        bool in = true;
        if (x < 0) { if (J.htile) x += J.tw; else in = false; }
        else if (x >= J.tw) { if (J.htile) x -= J.tw; else in = false; }
        if (y < 0) { if (J.vtile) y += J.th; else in = false; }
        else if (y >= J.th) { if (J.vtile) y -= J.th; else in = false; }
        if (in) m = __ldg(J.meta + (uint32_t)y * (uint32_t)J.tw + (uint32_t)x);
      }
      const bool ok = m != RS_NEVER;
      const unsigned b = __ballot_sync(RS_FULL, ok);
      const uint32_t slot = count + __popc(b & lt);
      if (ok && slot < J.kmax) out[slot - 1u] = make_uint2(o, m);
      count += __popc(b);
    }
    if (lane == 0) counts[v] = (uint8_t)min(count, J.kmax);
  }
}

// The first visits of pass 0 see almost no valued pixels and scan a long way down the offset table (all of it
// when there is no context).  A whole CTA scans for one such visit: 1024 table entries per step, compacted in
// table order through a shared-memory prefix over the 16 warps.
#define RS_COOP_THREADS 512
#ifndef RS_COOP_U
#define RS_COOP_U 4       // table entries per thread and step (2, 4 and 8 are within 1 % of each other on B200)
#endif
__global__ void __launch_bounds__(RS_COOP_THREADS) k_gather_pass0_coop(const RsDev J, uint2 *__restrict__ lists,
                                                                     uint8_t *__restrict__ counts, uint32_t v_end,
                                                                     unsigned int *__restrict__ claim, const int only_if_ctx) {
  __shared__ uint32_t s_cnt[2][RS_COOP_THREADS / 32];
  if (only_if_ctx && J.ctrl->n_ctx.v == 0u) return;  // (k_gather_pass0_sparse took these visits)
  __shared__ uint32_t s_v;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t stride = J.kmax - 1u;
  unsigned long long scans = 0;
  while (true) {
    if (threadIdx.x == 0) s_v = atomicAdd(claim, 1u);
    __syncthreads();
    const uint32_t v = s_v;
    if (v >= v_end) break;
    const uint32_t tpos = __ldg(J.targets + v);
    const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
    uint2 *out = lists + (size_t)v * stride;
    uint32_t count = 1;  // CTA-uniform
    // One step = RS_COOP_U entries per thread, all of their loads in flight together: the first visits scan the whole
    // table and a step's duration is two dependent round trips plus a barrier, however many entries it covers.
    for (uint32_t base = 1; base < J.nOff && count < J.kmax; base += RS_COOP_U * RS_COOP_THREADS) {
      uint32_t o[RS_COOP_U], q[RS_COOP_U], m[RS_COOP_U];
#pragma unroll
      for (int u = 0; u < RS_COOP_U; u++) {
        const uint32_t j = base + warp * (32u * RS_COOP_U) + 32u * u + lane;
        o[u] = 0; q[u] = 0; m[u] = RS_NEVER;
        if (j < J.nOff) {
          o[u] = __ldg(J.offsets + j);
          int x = px + rs_off_x(o[u]), y = py + rs_off_y(o[u]);
          bool in = true;
          if (x < 0) { if (J.htile) x += J.tw; else in = false; }
          else if (x >= J.tw) { if (J.htile) x -= J.tw; else in = false; }
          if (y < 0) { if (J.vtile) y += J.th; else in = false; }
          else if (y >= J.th) { if (J.vtile) y -= J.th; else in = false; }
          if (in) {
            q[u] = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
            m[u] = __ldg(J.meta + q[u]);
          }
        }
      }
      unsigned bal[RS_COOP_U];
      uint32_t mine = 0;
#pragma unroll
      for (int u = 0; u < RS_COOP_U; u++) {
        bal[u] = __ballot_sync(RS_FULL, (m[u] == RS_CTX_VALUED) || (m[u] < v));
        mine += __popc(bal[u]);
      }
      const int buf = (int)((base / (RS_COOP_U * RS_COOP_THREADS)) & 1u);
      if (lane == 0) s_cnt[buf][warp] = mine;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w2 = 0; w2 < RS_COOP_THREADS / 32; w2++) {
        const uint32_t c = s_cnt[buf][w2];
        before += (w2 < (int)warp) ? c : 0u;
        total += c;
      }
      uint32_t slot = count + before;  // table order within the step: warp, then u, then lane
#pragma unroll
      for (int u = 0; u < RS_COOP_U; u++) {
        const bool ok = (bal[u] >> lane) & 1u;
        const uint32_t sl = slot + __popc(bal[u] & lt);
        if (ok && sl < J.kmax) out[sl - 1u] = make_uint2(o[u], q[u] | (m[u] == RS_CTX_VALUED ? 0u : RS_TARGET_FLAG));
        slot += __popc(bal[u]);
      }
      count += total;
      scans += (threadIdx.x == 0) ? (unsigned long long)RS_COOP_U * RS_COOP_THREADS : 0ull;
    }
    if (threadIdx.x == 0) counts[v] = (uint8_t)min(count, J.kmax);
    __syncthreads();  // s_v is rewritten by the next claim
  }
  if (threadIdx.x == 0 && scans) atomicAdd(&J.ctrl->offset_scans, scans);
}

// ---- the same patches without the scan: nearest valued pixels by search, for the first visits of pass 0 ----
// The scan above tests EVERY offset in table order until K-1 valued pixels have turned up.  For the first visits of a
// large hole that is a disc of millions of empty pixels (all 16 Mi table entries for the first visits of a job without
// context: render-texture, map-style), although the valued pixels are few and known: the v points visited before v, and
// the context.  k_gather_pass0_sparse finds the same K-1 entries -- the K-1 smallest offsets in TABLE ORDER whose pixel
// is valued -- by search:
//   table order = ascending (x^2 + y^2, rank of (x, y) in reverse row-major order), k_gen_offsets below; that pair is the
//   64-bit key of an offset, and the patch is the K-1 smallest keys among the valued pixels in reach (|x| < ow, |y| < oh);
//   (A) a short scan of the first RS_SPARSE_PROBE table entries settles the visits whose surroundings are already dense;
//   (B) every earlier target point (and, when tiling, its wrapped aliases: the scan meets a pixel once per offset that
//       reaches it) goes through a sorted top-(K-1) list kept in shared memory by the warp;
//   (C) context pixels: 32x32 blocks of the target image in rings around the visit's block, nearest ring first; a block
//       without context (k_ctx_blocks) or farther away than the list's last entry is skipped, a ring that lies farther
//       away ends the search.  Tiling with context is left to the scan (k_gather_pass0_coop).
// Same lists bit for bit (the parity suite compares whole images); cfg4's 27 ms and cfg3's 3.8 ms of scanning become < 0.3.
#define RS_SPARSE_WARPS 8
#define RS_SPARSE_PROBE 512u  // table entries scanned first (RS_SPARSE_PROBE=n overrides at run time; 0: always search)
struct SparseList {  // the K-1 best entries so far, ascending key
  unsigned long long key[64];
  uint2 pay[64];     // {packed offset, pixel index | target flag}
};
__device__ __forceinline__ unsigned long long rs_offset_key(const RsDev &J, int dx, int dy) {
  const uint32_t d2 = (uint32_t)(dx * dx + dy * dy);
  const uint32_t tie = (uint32_t)(J.oh - 1 - dy) * (uint32_t)(2 * J.ow - 1) + (uint32_t)(J.ow - 1 - dx);
  return ((unsigned long long)d2 << 32) | tie;
}
// One warp, all lanes with the same entry: insert it; n = entries held, tau = key of the (K-1)-th entry once n == K1.
__device__ __forceinline__ void rs_sparse_insert(SparseList &L, uint32_t &n, unsigned long long &tau, const uint32_t K1,
                                                 const unsigned long long nk, const uint2 np) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long e0 = lane < n ? L.key[lane] : ~0ull, e1 = lane + 32u < n ? L.key[lane + 32u] : ~0ull;
  const uint32_t pos = __popc(__ballot_sync(RS_FULL, e0 < nk)) + __popc(__ballot_sync(RS_FULL, e1 < nk));
  if (pos >= K1) return;
  const uint2 p0 = L.pay[lane], p1 = L.pay[lane + 32u];
  __syncwarp();
  if (lane >= pos && lane < n && lane + 1u < K1) { L.key[lane + 1u] = e0; L.pay[lane + 1u] = p0; }
  if (lane + 32u >= pos && lane + 32u < n && lane + 33u < K1) { L.key[lane + 33u] = e1; L.pay[lane + 33u] = p1; }
  if (lane == 0) { L.key[pos] = nk; L.pay[pos] = np; }
  __syncwarp();
  n = min(n + 1u, K1);
  tau = (n == K1) ? L.key[K1 - 1u] : ~0ull;
}
// One warp: the lanes with `cand` offer (key, pay); those still under tau enter the list, in lane order.
__device__ __forceinline__ void rs_sparse_offer(SparseList &L, uint32_t &n, unsigned long long &tau, const uint32_t K1,
                                                bool cand, unsigned long long key, uint2 pay) {
  unsigned b = __ballot_sync(RS_FULL, cand && key < tau);
  while (b) {
    const int src = __ffs(b) - 1;
    b &= b - 1u;
    const unsigned long long nk = __shfl_sync(RS_FULL, key, src);
    const uint2 np = make_uint2(__shfl_sync(RS_FULL, pay.x, src), __shfl_sync(RS_FULL, pay.y, src));
    if (nk < tau) rs_sparse_insert(L, n, tau, K1, nk, np);
  }
}
// Usable context pixels per 32x32 block of the target image, and their total.
__global__ void __launch_bounds__(256) k_ctx_blocks(const uint32_t *__restrict__ meta, int tw, int th, int gw, int gh,
                                                    uint32_t *__restrict__ blocks, unsigned int *__restrict__ total) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t nb = (uint32_t)gw * (uint32_t)gh, nwarps = gridDim.x * (blockDim.x >> 5);
  uint32_t sum = 0;
  for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += nwarps) {
    const int bx = (int)(b % (uint32_t)gw), by = (int)(b / (uint32_t)gw), x = 32 * bx + (int)lane;
    uint32_t c = 0;
    for (int r = 0; r < 32; r++) {
      const int y = 32 * by + r;
      if (y < th && x < tw) c += (__ldg(meta + (size_t)y * tw + x) == RS_CTX_VALUED) ? 1u : 0u;
    }
    c = __reduce_add_sync(RS_FULL, c);
    if (lane == 0) blocks[b] = c;
    sum += c;
  }
  if (lane == 0 && sum) atomicAdd(total, sum);
}
__global__ void __launch_bounds__(RS_SPARSE_WARPS * 32) k_gather_pass0_sparse(const RsDev J, uint2 *__restrict__ lists,
                                                                              uint8_t *__restrict__ counts, uint32_t v_end,
                                                                              unsigned int *__restrict__ claim, const uint32_t probe) {
  __shared__ SparseList s_list[RS_SPARSE_WARPS];
  const unsigned lane = threadIdx.x & 31u;
  SparseList &L = s_list[threadIdx.x >> 5];
  const uint32_t K1 = J.kmax - 1u, stride = K1;
  const uint32_t n_ctx = J.ctx_blocks != nullptr ? J.ctrl->n_ctx.v : 0u;
  const bool tiled = J.htile || J.vtile;
  if (tiled && n_ctx) return;  // wrapped context: k_gather_pass0_coop scans for these visits
  unsigned long long scans = 0;
  while (true) {
    uint32_t v = 0;
    if (lane == 0) v = atomicAdd(claim, 1u);
    v = __shfl_sync(RS_FULL, v, 0);
    if (v >= v_end) break;
    v = (v & 1u) ? v_end - 1u - (v >> 1) : (v >> 1);  // the ends first: the fewest earlier points (a far walk to the context
                                                      // while the list is not full) and the most
    const uint32_t tpos = __ldg(J.targets + v);
    const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
    uint2 *out = lists + (size_t)v * stride;
    // ---- (A) the head of the table, scanned
    {
      uint32_t count = 1;
      for (uint32_t base = 1; base < probe && base < J.nOff && count < J.kmax; base += 128) {
        count = rs_scan_step<4>(J, v, px, py, base, count, out);
        scans += 128;
      }
      if (count >= J.kmax || (probe >= J.nOff && probe > 1u)) {
        if (lane == 0) counts[v] = (uint8_t)min(count, J.kmax);
        continue;
      }
    }
    // ---- (B) the earlier target points
    uint32_t n = 0;
    unsigned long long tau = ~0ull;
    __syncwarp();
    for (uint32_t i0 = 0; i0 < v; i0 += 32) {
      const uint32_t i = i0 + lane;
      const bool have = i < v;
      const uint32_t t = have ? __ldg(J.targets + i) : 0u;
      const int qx = (int)(t & 0xFFFFu), qy = (int)(t >> 16);
      const int dx1 = qx - px, dy1 = qy - py;
      const uint32_t qf = ((uint32_t)qy * (uint32_t)J.tw + (uint32_t)qx) | RS_TARGET_FLAG;
      for (int ay = 0; ay <= (J.vtile ? 1 : 0); ay++)
        for (int ax = 0; ax <= (J.htile ? 1 : 0); ax++) {
          const int dx = ax ? (dx1 > 0 ? dx1 - J.tw : dx1 + J.tw) : dx1, dy = ay ? (dy1 > 0 ? dy1 - J.th : dy1 + J.th) : dy1;
          const bool ok = have && !(ax && dx1 == 0) && !(ay && dy1 == 0) && abs(dx) < J.ow && abs(dy) < J.oh;
          rs_sparse_offer(L, n, tau, K1, ok, rs_offset_key(J, dx, dy),
                          make_uint2(((uint32_t)dx & 0xFFFFu) | ((uint32_t)dy << 16), qf));
        }
    }
    scans += v;
    // ---- (C) context pixels, block rings outwards
    if (n_ctx) {
      const int bx0 = px >> 5, by0 = py >> 5;
      const int rmax = max(max(bx0, J.gw - 1 - bx0), max(by0, J.gh - 1 - by0));
      for (int r = 0; r <= rmax; r++) {
        if (r >= 1 && n == K1) {  // every pixel of ring r is at least 32 (r - 1) + 1 away
          const unsigned long long md = 32ull * (unsigned)(r - 1) + 1ull;
          if (md * md > (tau >> 32)) break;
        }
        const int nblk = r ? 8 * r : 1, top = 2 * r + 1, side = 2 * r - 1;
        for (int t0 = 0; t0 < nblk; t0 += 32) {
          const int t = t0 + (int)lane;
          int bx = bx0, by = by0;
          if (r) {
            if (t < top) { by = by0 - r; bx = bx0 - r + t; }
            else if (t < 2 * top) { by = by0 + r; bx = bx0 - r + (t - top); }
            else if (t < 2 * top + side) { bx = bx0 - r; by = by0 - r + 1 + (t - 2 * top); }
            else { bx = bx0 + r; by = by0 - r + 1 + (t - 2 * top - side); }
          }
          uint32_t md2 = 0;
          bool want = t < nblk && bx >= 0 && bx < J.gw && by >= 0 && by < J.gh;
          if (want) want = __ldg(J.ctx_blocks + (size_t)by * J.gw + bx) != 0u;
          if (want) {
            const int ddx = max(0, max(32 * bx - px, px - (32 * bx + 31))), ddy = max(0, max(32 * by - py, py - (32 * by + 31)));
            md2 = (uint32_t)(ddx * ddx + ddy * ddy);
            want = n < K1 || md2 <= (uint32_t)(tau >> 32);
          }
          unsigned wb = __ballot_sync(RS_FULL, want);
          while (wb) {
            const int src = __ffs(wb) - 1;
            wb &= wb - 1u;
            const int sbx = __shfl_sync(RS_FULL, bx, src), sby = __shfl_sync(RS_FULL, by, src);
            const uint32_t smd2 = __shfl_sync(RS_FULL, md2, src);
            if (n == K1 && smd2 > (uint32_t)(tau >> 32)) continue;
            const bool up = 32 * sby + 31 < py;  // a block above the visit: its bottom rows are the near ones
            const int x = 32 * sbx + (int)lane, dx = x - px;
            for (int rr = 0; rr < 32; rr++) {
              const int y = 32 * sby + (up ? 31 - rr : rr), dy = y - py;
              if (y >= J.th) continue;
              if (n == K1 && (uint32_t)(dy * dy) > (uint32_t)(tau >> 32)) continue;
              const uint32_t q = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
              const bool ok = x < J.tw && abs(dx) < J.ow && abs(dy) < J.oh && __ldg(J.meta + q) == RS_CTX_VALUED;
              rs_sparse_offer(L, n, tau, K1, ok, rs_offset_key(J, dx, dy),
                              make_uint2(((uint32_t)dx & 0xFFFFu) | ((uint32_t)dy << 16), q));
            }
            scans += 1024;
          }
        }
      }
    }
    __syncwarp();
    for (uint32_t k = lane; k < n; k += 32) out[k] = L.pay[k];
    if (lane == 0) counts[v] = (uint8_t)(1u + n);
    __syncwarp();
  }
  if (lane == 0 && scans) atomicAdd(&J.ctrl->offset_scans, scans);
}

// --------------------------------------------------------------------------------------- the pass kernel
// CTA shapes. Throughput kernel: ONE 1024-thread CTA per SM (64 registers per thread), so the metric tables are
// staged once per SM and the shared-memory carve-out leaves >= 124 KB of L1 for the offset/meta/point tables that
// every visit re-reads (2 x 512 threads with map tables took the carve-out to 228 KB and cfg4 ran 1.8x slower).
// Team kernel: 2 x 512 threads, at most 8 teams (W >= 2) per CTA, one named barrier per team.
#ifndef RS_TP_WARPS
#define RS_TP_WARPS 32
#endif
#ifndef RS_TP_MIN_CTAS
#define RS_TP_MIN_CTAS 1
#endif
// The instantiation for large patches without map channels (the heal / inpaint jobs) compiles to 48 registers without a
// spill, so TWO CTAs of 20 warps fit an SM: 40 warps to hide the gathers' latency behind instead of 32 (cfg3 48.7 -> 45.8
// ms).  Not for the others: with map channels a second CTA doubles 64 KB of tables and takes the L1 with it (cfg4 46.5 ->
// 61.3), the small-patch kernels lose the same way (cfg2 3.49 -> 4.53), a corpus slice leaves room for one CTA only, and
// 2 x 24 warps at 40 registers spill (cfg3 56.1).
#ifndef RS_TP_SPLIT_WARPS
#define RS_TP_SPLIT_WARPS 20
#endif
template <bool MAPS, int NB, bool SMEMC, int LW>
struct TpShape {
  static constexpr bool split = !MAPS && NB > 16 && !SMEMC && LW == 32 && RS_TP_WARPS == 32 && RS_TP_MIN_CTAS == 1;
  static constexpr int warps = split ? RS_TP_SPLIT_WARPS : RS_TP_WARPS;
  static constexpr int ctas = split ? 2 : RS_TP_MIN_CTAS;
};
#define RS_TEAM_WARPS 16
#define RS_TEAM_SLOTS 8
#define RS_BF_WARPS 16

// Registers are the scarce resource of the pass kernels (64 per thread), and what the distance loop does with a spare
// one is keep another gather in flight.  So everything a visit needs only before or after that loop lives in the
// warp's shared-memory scratch (WarpScratch::vis, ::st), written and read by lane 0.
struct VisitShared {   // handed from rs_visit_prepare to rs_visit_finish
  unsigned long long selfw;  // the visit's own state word of version `pass`
  uint32_t selfq, epoch_idx, hide_from, my_base;
};
struct WarpStats {     // per-warp counters, flushed once at kernel end
  unsigned long long evals, sumbest;
  uint32_t visits, scans, heur, skips, perfect, betters;
};
struct LaneStats {     // per-lane counters of the distance loop (registers), reduced at kernel end
  uint32_t compares = 0, issued = 0;
};
struct Visit {         // the visit a warp is working on (warp-uniform registers)
  uint32_t v, K, nHeur;
};

// NB = the most neighbours a patch can have in this instantiation: 64 (the reference's limit, lib/engine.c:591) or 16 for
// the small patches of the texture scripts (9 neighbours: render-texture, map-style).  The small variant takes a warp's
// scratch from 2.0 KB to 0.85 KB, i.e. the throughput kernel's CTA from 96 KB to 59 KB of shared memory -- and what
// shared memory does not take is L1: a 256 KB corpus tile then mostly stays in the SM's own cache.
template <bool MAPS, int NB>
struct __align__(16) WarpScratch {
  static constexpr int kSlots = NB + RS_CHUNK_MAX;         // distance records: the patch padded to whole chunks
  static constexpr int kLaneSlots = NB < 64 ? 32 : 64;     // arrays that are also indexed by lane (+ 32)
  RsNb nb[kSlots];                        // the patch as the distance loop reads it (rs_device.cuh), padded to whole chunks
  uint32_t map[MAPS ? kSlots : 4];        // neighbour map bytes
  uint32_t off[kLaneSlots];  // neighbour offsets (packed int16 pair), ascending distance; [0] = (0,0).  Dead once the
                             // heuristic candidates exist: reused as the full patch distance of each of them (hsum)
  uint32_t q[kLaneSlots];    // neighbour pixel index, later: packed heuristic candidate or RS_NO_SRC
  uint32_t aux[kLaneSlots];  // neighbour meta, later: neighbour source, later: compacted candidate list
  VisitShared vis;
  WarpStats st;
};
#define RS_NB_SMALL 16
#define RS_NB_FULL RS_MAX_NB

// Stage the replicated metric tables with one TMA bulk copy per table (whole CTA calls this).
template <bool MAPS>
__device__ __forceinline__ void rs_stage_tables(const RsDev &J, uint32_t *lutc, uint32_t *lutm, uint64_t *bar) {
  const unsigned lut_bytes = (MAPS ? 2u : 1u) * RS_LUT_WORDS * 4u;
  if (threadIdx.x == 0) {
    rs_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rs_mbar_expect_tx(bar, lut_bytes);
    rs_tma_load_1d(lutc, J.lut_rep, RS_LUT_WORDS * 4u, bar);
    if (MAPS) rs_tma_load_1d(lutm, J.lut_rep + RS_LUT_WORDS, RS_LUT_WORDS * 4u, bar);
  }
  __syncthreads();
  rs_mbar_wait(bar, 0);
}

// Claim the next visit of this launch, in order (lib/synthesize.h:480-482 with THREAD_LIMIT 1).  Split in two so that
// the round trip of the atomic can overlap other work: rs_claim_issue returns lane 0's raw ticket,
// rs_claim_resolve turns it into the warp-uniform visit index (>= J.seg_end: nothing left) and does the
// progress tick + cancel poll of the reference (synthesize.h:493-497).
template <int NV = 1>  // visits per claim: consecutive ones
__device__ __forceinline__ uint32_t rs_claim_issue(const RsDev &J, RsCtrl *ctrl) {
  uint32_t raw = 0;
  if ((threadIdx.x & 31u) == 0u) raw = atomicAdd(&ctrl->next[J.slot].v, (unsigned)NV);
  return raw;
}
template <int NV = 1>
__device__ __forceinline__ uint32_t rs_claim_resolve(const RsDev &J, RsCtrl *ctrl, uint32_t raw) {
  uint32_t v = J.seg_begin + raw;
  if ((threadIdx.x & 31u) == 0u) {
#pragma unroll
    for (uint32_t x = 0; x < (uint32_t)NV; x++) {
      const uint32_t vx = v + x;
      if (vx < J.seg_end && (vx & 4095u) == 0u) {
        const uint32_t pass = J.pass;
        J.host_ticks[pass] = vx + 1u;
        if ((vx >> 12) < RS_TIMELINE) ctrl->tick_ns[pass][vx >> 12] = rs_globaltimer();
        if (*J.host_cancel) {  // no further claims succeed; this visit still runs (later ones may wait on it)
          atomicExch(&ctrl->stop, 1u);
          atomicAdd(&ctrl->next[J.slot].v, 0x40000000u);
        }
      }
    }
  }
  return __shfl_sync(RS_FULL, v, 0);
}

// A visit is prepared and committed by a GROUP of LW lanes: the whole warp (LW = 32), or one half of it (LW = 16: two
// visits of small patches side by side in one warp, k_synth_pass<..., 16>).  The two halves run CONVERGED -- one
// instruction stream, that is the point -- so every vote and barrier is a full-warp one that all 32 lanes reach (loops
// around them run while ANY group needs them, bodies predicated), and a group reads its own 16 bits of the result.
template <int LW>
struct Grp {
  static __device__ __forceinline__ unsigned lane() { return threadIdx.x & (unsigned)(LW - 1); }
  static __device__ __forceinline__ unsigned shift() { return LW == 32 ? 0u : (threadIdx.x & 31u & ~(unsigned)(LW - 1)); }
  static __device__ __forceinline__ unsigned bits(unsigned full) { return LW == 32 ? full : ((full >> shift()) & ((1u << (LW & 31)) - 1u)); }
  static __device__ __forceinline__ unsigned ballot(bool p) { return bits(__ballot_sync(RS_FULL, p)); }
  static __device__ __forceinline__ bool any(bool p) { return ballot(p) != 0u; }          // in my group
  static __device__ __forceinline__ bool any_group(bool p) { return __any_sync(RS_FULL, p) != 0; }  // in any group of the warp
  static __device__ __forceinline__ unsigned match_any(uint32_t c) { return bits(__match_any_sync(RS_FULL, c)); }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};
// v / J.epoch_len without the division: the estimate from floor(2^32 / len) is exact or one short.
__device__ __forceinline__ uint32_t rs_epoch_of(const RsDev &J, uint32_t v) {
  uint32_t q = __umulhi(v, J.epoch_inv);
  if (v - q * J.epoch_len >= J.epoch_len) q++;
  return q;
}
// t / n for n in [1, 32] and t < 2^16 (a candidate's chunks): multiply by ceil(2^32 / n) kept in constant memory.
__constant__ uint32_t c_inv32[33] = {
    0u, 0xFFFFFFFFu, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u, 0x2AAAAAABu, 0x24924925u, 0x20000000u, 0x1C71C71Du, 0x1999999Au,
    0x1745D175u, 0x15555556u, 0x13B13B14u, 0x12492493u, 0x11111112u, 0x10000000u, 0x0F0F0F10u, 0x0E38E38Fu, 0x0D79435Fu, 0x0CCCCCCDu,
    0x0C30C30Du, 0x0BA2E8BBu, 0x0B21642Du, 0x0AAAAAABu, 0x0A3D70A4u, 0x09D89D8Au, 0x097B425Fu, 0x0924924Au, 0x08D3DCB1u, 0x08888889u,
    0x08421085u, 0x08000000u};
// Lane 0: until every visit of the epochs <= epoch_idx - 2 of this pass has completed.  Finishing visits only count
// themselves (a fire-and-forget reduction); whoever waits moves the watermark over the leading complete epochs.
__device__ __forceinline__ void rs_wait_epochs(const RsDev &J, RsCtrl *ctrl, uint32_t epoch_idx) {
  const uint32_t pass = J.pass;
  RsSpinGuard guard;
  while (true) {
    const uint32_t wmk = rs_ld_u32_relaxed(&ctrl->epoch_wm[pass].v);
    if (wmk + 1u >= epoch_idx || guard.expired(ctrl)) break;
    const uint32_t first = wmk * J.epoch_len;
    if (rs_ld_u32_relaxed(&ctrl->epoch_done[pass][wmk].v) == min(J.epoch_len, J.pass_end - first))
      atomicCAS(&ctrl->epoch_wm[pass].v, wmk, wmk + 1u);
    else
      __nanosleep(100);
  }
}

// One warp: gather the patch of visit v (target point tpos), wait for exactly the neighbour versions the
// sequential loop would see, build the heuristic candidate list (S.aux[0..nHeur)).
// A visit is prepared in three steps.  (1) GEOMETRY: which pixels form the patch -- depends on nothing another visit
// of this pass computes.  (2) VALUES: wait for exactly the neighbour versions the sequential loop would see and read
// colours + sources.  (3) CANDIDATES: heuristic candidates from the neighbours' sources.  The throughput kernel runs
// them back to back; the team kernel lets the other warps of the team fetch the probes' corpus pixels between (1)
// and (2)/(3), because those gathers need the geometry only.
//
// (1) One warp: gather the patch of visit v (target point tpos): S.off / S.q / S.aux(meta) and the geometry half of
// the distance records S.nb[k].{lin,dx,pen}, padded to whole chunks.  Returns K.
template <int CH, int LW = 32, bool MAPS, int NB>
__device__ __forceinline__ uint32_t rs_visit_geometry(const RsDev &J, WarpScratch<MAPS, NB> &S, const uint32_t v, const uint32_t tpos,
                                                      const bool regular, const bool on = true) {
  typedef Grp<LW> G;
  const unsigned lane = G::lane();
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t pass = J.pass;
  const int px = (int)(tpos & 0xFFFFu), py = (int)(tpos >> 16);
  const uint32_t selfq = (uint32_t)py * (uint32_t)J.tw + (uint32_t)px;

  // ---- gather the patch: self + nearest valued pixels (lib/synthesize.h:189-241)
  // (`on`: this group has a visit; the groups of a warp take the same one of the four ways below)
  if (lane == 0 && on) {
    S.off[0] = 0u;
    S.q[0] = selfq;
    S.vis.selfq = selfq;
    S.aux[0] = v;
  }
  uint32_t count = 1;
  if (pass == 0u && J.nb_lists != nullptr) {
    // precomputed by k_gather_pass0; aux := 0 (< v+1) marks "target visited before me", else context
    if (on) count = J.nb_counts[v];
    const uint2 *lst = J.nb_lists + (size_t)v * (J.kmax - 1u);
    for (uint32_t k = 1u + lane; k < count; k += LW) {
      const uint2 e = __ldg(lst + (k - 1u));
      S.off[k] = e.x;
      S.q[k] = e.y & ~RS_TARGET_FLAG;
      S.aux[k] = (e.y & RS_TARGET_FLAG) ? 0u : RS_CTX_VALUED;
    }
  } else if (pass != 0u && J.regular_r != 0u && regular &&
             !G::any_group(on && !((J.htile || (unsigned)(px - (int)J.regular_r) < (unsigned)(J.tw - 2 * (int)J.regular_r)) &&
                                   (J.vtile || (unsigned)(py - (int)J.regular_r) < (unsigned)(J.th - 2 * (int)J.regular_r))))) {
    // Every pixel of the image is usable context or a target point (which all have a value from pass 1 on), and the point
    // is not near a border that clips: the patch is the head of the offsets table, whatever the pixels hold.
    if (on) count = J.kmax;
    for (uint32_t k = 1u + lane; k < count; k += LW) {
      const uint32_t o = __ldg(J.offsets + k);
      int x = px + rs_off_x(o), y = py + rs_off_y(o);
      if (x < 0) x += J.tw; else if (x >= J.tw) x -= J.tw;  // (tiling; a clipping axis was excluded above)
      if (y < 0) y += J.th; else if (y >= J.th) y -= J.th;
      const uint32_t q = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
      S.off[k] = o;
      S.q[k] = q;
      S.aux[k] = __ldg(J.meta + q);
    }
  } else if (pass != 0u && J.nb_later != nullptr) {
    // gathered once for all later passes by k_gather_later: offset + meta word; the pixel index follows from the offset
    if (on) count = J.nb_later_counts[v];
    const uint2 *lst = J.nb_later + (size_t)v * (J.kmax - 1u);
    for (uint32_t k = 1u + lane; k < count; k += LW) {
      const uint2 e = __ldcs(lst + (k - 1u));  // streamed: read once per pass
      int x = px + rs_off_x(e.x), y = py + rs_off_y(e.x);
      if (x < 0) x += J.tw; else if (x >= J.tw) x -= J.tw;  // only offsets that wrap (tiling) or stay inside were listed
      if (y < 0) y += J.th; else if (y >= J.th) y -= J.th;
      S.off[k] = e.x;
      S.q[k] = (uint32_t)y * (uint32_t)J.tw + (uint32_t)x;
      S.aux[k] = e.y;
    }
  } else {
    for (uint32_t base = 1;; base += LW) {
      const bool go = on && base < J.nOff && count < J.kmax;  // (a group that has its patch idles while another still scans)
      if (!G::any_group(go)) break;
      const uint32_t j = base + lane;
      bool ok = false;
      uint32_t o = 0, q = 0, m = 0;
      if (go && j < J.nOff) {
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
      const unsigned b = G::ballot(ok);
      const uint32_t slot = count + __popc(b & lt);
      if (ok && slot < J.kmax) {
        S.off[slot] = o;
        S.q[slot] = q;
        S.aux[slot] = m;
      }
      count += __popc(b);
      if (lane == 0 && go) S.st.scans += min((uint32_t)LW, J.nOff - base);
    }
  }
  const uint32_t K = on ? min(count, J.kmax) : 0u;
  G::sync();
  {  // geometry half of the distance records, padded to whole chunks with records that cost nothing
    // (whole chunks of the launched kernel's size, and a continuation chunk of the team kernel may start at any k < K)
    const uint32_t nch = (K + CH - 2u) / CH;  // (CH == J.chunk: a division by a constant)
    const uint32_t kpad = min((uint32_t)WarpScratch<MAPS, NB>::kSlots, max(1u + (nch ? nch : 1u) * CH, K + (uint32_t)RS_CHUNK_MAX));
    for (uint32_t k = lane; on && k < kpad; k += LW) {
      RsNb r;
      if (k < K) {
        const uint32_t o = S.off[k];
        r.dx = rs_off_x(o);
        r.lin = rs_off_y(o) * J.cw + r.dx;
        r.pix = 0u;
        r.pen = J.penalty;
      } else {
        r.lin = 0; r.dx = RS_PAD_DX; r.pix = 0u; r.pen = 0u;
        if (MAPS) S.map[k] = 0u;
      }
      S.nb[k] = r;
    }
  }
  G::sync();
  return K;
}

// (2) One warp: wait for exactly the versions the sequential order would see, then read them (one 64-bit load each):
// colours into S.nb[k].pix (+ S.map), sources into S.aux.
template <int LW = 32, bool MAPS, int NB>
__device__ __forceinline__ void rs_visit_values(const RsDev &J, WarpScratch<MAPS, NB> &S, const uint32_t v, const uint32_t K) {
  typedef Grp<LW> G;
  const unsigned lane = G::lane();
  const uint32_t pass = J.pass, pass_end = J.pass_end;
  for (uint32_t k = lane; k < K; k += LW) {
    const uint32_t q = S.q[k], m = S.aux[k];
    uint32_t r = 0;
    if (k == 0) r = pass;
    else if (m != RS_CTX_VALUED) {
      if (m < v && m < pass_end) r = pass + 1u;
      else {
#pragma unroll
        for (uint32_t p2 = 0; p2 < 5u; p2++) r += (p2 < pass && m < J.ends[p2]) ? 1u : 0u;  // (at most 6 passes)
      }
    }
    const unsigned long long *wp = J.W + 2 * (size_t)q + (r & 1u);
    unsigned long long w = rs_ld_state(wp);
    RsSpinGuard guard;
    while (((unsigned)(w >> 24) & 0xFFu) != r) {
      __nanosleep(40);
      w = rs_ld_state(wp);
      if (guard.expired(J.ctrl)) break;
    }
    if (k == 0) S.vis.selfw = w;
    S.nb[k].pix = (uint32_t)w & 0xFFFFFFu;
    if (MAPS) S.map[k] = __ldg(J.tmaps + q);
    S.aux[k] = (uint32_t)(w >> 32);  // source of this neighbour, or RS_NO_SRC
  }
  G::sync();
}

// (3) One warp: the heuristic candidate list S.aux[0..nHeur) of visit v, and what rs_visit_finish needs (S.vis).
template <bool SMEMC = false, int LW = 32, bool MAPS, int NB>
__device__ __forceinline__ void rs_visit_candidates(const RsDev &J, RsCtrl *ctrl, WarpScratch<MAPS, NB> &S, Visit &V,
                                                    const uint32_t v, const uint32_t K, const CorpusSmem &cs = CorpusSmem(),
                                                    const bool on = true) {  // (a group without a visit comes with K = 0)
  typedef Grp<LW> G;
  const unsigned lane = G::lane();
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t pass = J.pass;
  const uint32_t tag = (pass + 1u) << 29;

  // ---- heuristic 1 + 2 candidates (lib/synthesize.h:537-580): source of neighbour minus its offset,
  //      dropped if outside/masked corpus, if this target index was the last VISIBLE prober of that corpus
  //      point (rs_device.cuh: epochs), or if an earlier neighbour proposes the same point.
  //      The corpus pixel and the three stamp words of a candidate are fetched in one round trip.
  const uint32_t epoch_idx = rs_epoch_of(J, v), epoch0 = epoch_idx * J.epoch_len;
  const uint32_t hide_from = epoch_idx ? epoch0 - J.epoch_len : 0u;  // stamps of my pass from here on are hidden
  const uint32_t hide_base = tag | hide_from;
  constexpr int NR = NB > LW ? 2 : 1;  // rounds of LW neighbours
  static_assert(NB <= 2 * LW, "a group covers the patch in at most two rounds");
  uint32_t mycand[2];
#pragma unroll
  for (int rnd = 0; rnd < NR; rnd++) {
    const uint32_t k = lane + (unsigned)LW * rnd;
    uint32_t c = RS_NO_SRC;
    if (k < K) {
      const uint32_t src = S.aux[k];
      if (src != RS_NO_SRC) {
        const uint32_t o = S.off[k];
        const int x = (int)(src & 0xFFFFu) - rs_off_x(o), y = (int)(src >> 16) - rs_off_y(o);
        if ((unsigned)x < (unsigned)J.cw && (unsigned)y < (unsigned)J.ch) {
          const size_t a = (size_t)y * J.cw + x;
          const uint32_t cm = MAPS ? __ldg(&J.corpus8[a].x) : rs_corpus4<SMEMC>(J, cs, (uint32_t)a);
          if (cm >= 0xFF000000u) c = (uint32_t)x | ((uint32_t)y << 16);
        }
      }
    }
    mycand[rnd] = c;
    if (on) S.q[k] = c;
  }
  G::sync();
  bool pskip[2] = {false, false};
  bool look = true;  // this group (re)reads the stamps in this attempt
  for (int attempt = 0; attempt < 2; attempt++) {
    bool any = false;
#pragma unroll
    for (int rnd = 0; rnd < NR; rnd++) {
      const uint32_t c = mycand[rnd];
      if (look) pskip[rnd] = false;
      if (look && c != RS_NO_SRC) {
        const size_t a = (size_t)(c >> 16) * J.cw + (c & 0xFFFFu);
        uint32_t newest = 0u;
        const ulonglong2 e01 = rs_ld_state2(J.prober + 4 * a);  // the three arrays' words of one corpus pixel share a sector
        const unsigned long long e2 = rs_ld_state(J.prober + 4 * a + 2);
#pragma unroll
        for (int t = 0; t < 3; t++) {
          const unsigned long long e = t == 0 ? e01.x : (t == 1 ? e01.y : e2);
          const uint32_t h = (uint32_t)(e >> 32);
          newest = max(newest, (h >= hide_base) ? (uint32_t)e : h);
        }
        pskip[rnd] = newest != 0u && (newest & RS_IDX_MASK) == v;
        any |= pskip[rnd];
      }
    }
    look = attempt == 0 && hide_from != 0u && G::any(any);  // a skip verdict that the epochs still running could overturn
    if (!G::any_group(look)) break;
    if (look && lane == 0) rs_wait_epochs(J, ctrl, epoch_idx);
    G::sync();
  }
  uint32_t nHeur = 0, nSkips = 0;
#pragma unroll
  for (int rnd = 0; rnd < NR; rnd++) {
    const uint32_t k = lane + (unsigned)LW * rnd;
    const uint32_t c = mycand[rnd];
    bool valid = (c != RS_NO_SRC);
    if (rnd == 0) {  // an earlier neighbour proposing the same point: lanes are neighbours here, one MATCH finds them
      const unsigned same = G::match_any(c);
      if (pskip[0] || (same & lt)) valid = false;
    } else if (valid) {
      bool skip = pskip[rnd];
      for (uint32_t k2 = 0; k2 < k && !skip; k2++) skip = (S.q[k2] == c);
      if (skip) valid = false;
    }
    const unsigned b = G::ballot(valid);
    nSkips += __popc(G::ballot(c != RS_NO_SRC)) - __popc(b);
    if (valid) S.aux[nHeur + __popc(b & lt)] = c;  // sources in S.aux are no longer needed (mycand holds mine)
    nHeur += __popc(b);
    G::sync();
  }
  V.v = v; V.K = K; V.nHeur = nHeur;
  if (lane == 0 && on) {
    S.vis.epoch_idx = epoch_idx; S.vis.hide_from = hide_from;
    S.vis.my_base = tag | epoch0;
    S.st.visits++;
    S.st.skips += nSkips;
  }
  G::sync();
}

// One warp: merge the stamps of visit v into the recentProber array of its epoch (c0/c1: this lane's candidates
// lane and lane + 32 of the list; the first stampEnd of the list are stamped), then publish the visit as complete.
__device__ __forceinline__ void rs_visit_stamps(const RsDev &J, RsCtrl *ctrl, const uint32_t v, const uint32_t epoch_idx,
                                                const uint32_t my_base, const uint32_t hide_from, const uint32_t stampEnd,
                                                const uint32_t c0, const uint32_t c1) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t pass = J.pass, stp = ((pass + 1u) << 29) | v;
  if (stampEnd > 0u && hide_from > 0u) {  // one writing epoch per array: epochs <= e-2 must be complete
    if (lane == 0) rs_wait_epochs(J, ctrl, epoch_idx);
    __syncwarp();
  }
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    if (lane + 32u * rnd < stampEnd) {
      const uint32_t c = rnd ? c1 : c0;
      unsigned long long *pp = J.prober + 4 * ((size_t)(c >> 16) * J.cw + (c & 0xFFFFu)) + epoch_idx % 3u;
      unsigned long long old = rs_ld_state(pp);
      while (true) {
        const uint32_t hi = (uint32_t)(old >> 32), lo = (uint32_t)old;
        const unsigned long long nw = (hi >= my_base) ? (((unsigned long long)max(hi, stp) << 32) | lo)
                                                        : (((unsigned long long)stp << 32) | hi);
        if (nw == old) break;
        const unsigned long long prev = atomicCAS(pp, old, nw);
        if (prev == old) break;
        old = prev;
      }
    }
  }
  __syncwarp();
  // publish: this visit is complete (its stamps were merged by CAS operations that have returned)
  if (lane == 0)
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(&ctrl->epoch_done[pass][epoch_idx].v) : "memory");
}

// One warp: commit the winner (lib/synthesize.h:620-639), merge the heuristic-2 stamps, publish completion.
// hcol[i] = colour of heuristic candidate i (fetched with its first chunk).  For a winning probe: win_pt = its corpus
// point if the distance phase tracked it (else RS_NO_SRC: looked up from the probe's index), win_col = its colour if
// have_col (else fetched here).
template <bool STAMPS = true, bool SMEMC = false, int LW = 32, bool MAPS, int NB>
__device__ __forceinline__ void rs_visit_finish(const RsDev &J, RsCtrl *ctrl, WarpScratch<MAPS, NB> &S, const Visit &V,
                                                uint32_t bestSum, int bestIdx, uint32_t win_pt, uint32_t win_col,
                                                bool have_col, const CorpusSmem &cs = CorpusSmem(), const bool on = true) {
  typedef Grp<LW> G;
  const unsigned lane = G::lane();
  const uint32_t pass = J.pass, v = V.v, nHeur = on ? V.nHeur : 0u;
  const uint32_t *candlist = S.aux, *hcol = S.q;
  const uint32_t epoch_idx = S.vis.epoch_idx, my_base = S.vis.my_base;
  const bool bettered = bestIdx != 0x7FFFFFFF;
  const uint32_t total = nHeur + J.probes;
  const uint32_t seq_evals = !bettered ? 0u : (bestSum == 0u ? (uint32_t)bestIdx + 1u : total);
  if (lane == 0 && on) {  // new colour + source only if the source changed; the new version is always published
    const unsigned long long selfw = S.vis.selfw;
    uint32_t colour = (uint32_t)selfw & 0xFFFFFFu, src = (uint32_t)(selfw >> 32);
    if (bettered) {
      uint32_t bp, bcol = 0;
      const bool heur = (uint32_t)bestIdx < nHeur;
      if (heur) {
        bp = candlist[bestIdx];
        bcol = hcol[bestIdx];
      } else if (win_pt != RS_NO_SRC) {
        bp = win_pt;
        bcol = win_col;
      } else {
        bp = rs_corpus_point(J, J.ctrl->n_corpus, rs_range(rs_probe_hash(J.seed, pass, v, (uint32_t)bestIdx - nHeur), J.ctrl->n_corpus));
      }
      if (bp != src) {
        if (!heur && !(have_col && win_pt != RS_NO_SRC)) {
          const size_t a = (size_t)(bp >> 16) * J.cw + (bp & 0xFFFFu);
          bcol = (MAPS ? __ldg(&J.corpus8[a].x) : rs_corpus4<SMEMC>(J, cs, (uint32_t)a)) & 0xFFFFFFu;
        }
        colour = bcol;
        src = bp;
        S.st.betters++;
      }
      S.st.sumbest += bestSum;
    }
    rs_st_state(J.W + 2 * (size_t)S.vis.selfq + ((pass + 1u) & 1u),
                ((unsigned long long)src << 32) | ((unsigned long long)(pass + 1u) << 24) | colour);
    S.st.evals += seq_evals;
    S.st.heur += min(nHeur, seq_evals);
    S.st.perfect += (bettered && bestSum == 0u) ? 1u : 0u;
  }
  if (STAMPS) {  // (the throughput kernel: this warp does it all; same merge as rs_visit_stamps, list read in place)
    // ---- heuristic 2 bookkeeping: stamp the evaluated heuristic candidates before the perfect one, if any
    const uint32_t tag = (pass + 1u) << 29;
    const uint32_t stampEnd = (bettered && bestSum == 0u && (uint32_t)bestIdx < nHeur) ? (uint32_t)bestIdx : nHeur;
    const bool wait = on && stampEnd > 0u && S.vis.hide_from > 0u;  // one writing epoch per array: epochs <= e-2 must be complete
    if (G::any_group(wait)) {
      if (wait && lane == 0) rs_wait_epochs(J, ctrl, epoch_idx);
      G::sync();
    }
    for (uint32_t i = lane; i < stampEnd; i += LW) {
      const uint32_t c = candlist[i];
      unsigned long long *pp = J.prober + 4 * ((size_t)(c >> 16) * J.cw + (c & 0xFFFFu)) + epoch_idx % 3u;
      const uint32_t stp = tag | v;
      unsigned long long old = rs_ld_state(pp);
      while (true) {
        const uint32_t hi = (uint32_t)(old >> 32), lo = (uint32_t)old;
        const unsigned long long nw = (hi >= my_base) ? (((unsigned long long)max(hi, stp) << 32) | lo)
                                                        : (((unsigned long long)stp << 32) | hi);
        if (nw == old) break;
        const unsigned long long prev = atomicCAS(pp, old, nw);
        if (prev == old) break;
        old = prev;
      }
    }
    G::sync();
    // publish: this visit is complete (its stamps were merged by CAS operations that have returned)
    if (lane == 0 && on)
      asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(&ctrl->epoch_done[pass][epoch_idx].v) : "memory");
  }
}

// One lane: the counters a scratch has collected, into the job's.
__device__ __forceinline__ void rs_flush_warp_stats(const RsDev &J, RsCtrl *ctrl, const WarpStats *ws) {
  if (!ws->visits) return;
  const uint32_t pass = J.pass;
  atomicAdd(&ctrl->visits, (unsigned long long)ws->visits);
  atomicAdd(&ctrl->pass_visits[pass], (unsigned long long)ws->visits);
  atomicAdd(&ctrl->evals, ws->evals);
  atomicAdd(&ctrl->offset_scans, (unsigned long long)ws->scans);
  atomicAdd(&ctrl->heur_evals, (unsigned long long)ws->heur);
  atomicAdd(&ctrl->heur_skips, (unsigned long long)ws->skips);
  atomicAdd(&ctrl->perfect, (unsigned long long)ws->perfect);
  atomicAdd(&ctrl->sum_best[pass], ws->sumbest);
  atomicAdd(&ctrl->betters[pass], ws->betters);
}
// Whole CTA, at kernel end: flush the per-warp counters; the last CTA out decides whether later passes run
// (lib/refiner.h:111): (float)betters/n < 0.1.
__device__ __forceinline__ void rs_pass_epilogue(const RsDev &J, RsCtrl *ctrl, LaneStats &ls, const WarpStats *ws) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t pass = J.pass;
  ls.compares = __reduce_add_sync(RS_FULL, ls.compares);
  ls.issued = __reduce_add_sync(RS_FULL, ls.issued);
  if (lane == 0 && (ls.compares | ls.issued)) {
    atomicAdd(&ctrl->evals_issued, (unsigned long long)ls.issued);
    atomicAdd(&ctrl->compares, (unsigned long long)ls.compares);
  }
  if (lane == 0 && ws != nullptr) rs_flush_warp_stats(J, ctrl, ws);  // ws: the counters of the warp that prepared and committed visits
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&ctrl->done_ctas[J.slot].v, 1u) + 1u;
    if (done == gridDim.x && J.last_seg) {
      __threadfence();
      const unsigned b = atomicAdd(&ctrl->betters[pass], 0u);
      ctrl->passes_run = pass + 1u;
      ctrl->pass_end_ns[pass] = rs_globaltimer();
      if ((double)((float)b / (float)J.nT) < J.terminate_fraction) atomicExch(&ctrl->stop, 1u);
    }
  }
}

// One (heuristic candidate, chunk j) pair of the patch distance; chunk 0 also carries the target point's own terms.
template <bool MAPS, int CH, int NB, bool SMEMC = false>
__device__ __forceinline__ uint32_t rs_heur_pair(const RsDev &J, unsigned lutc, unsigned lutm, const WarpScratch<MAPS, NB> &S,
                                                 uint32_t *hcol, uint32_t K, uint32_t ci, uint32_t j, LaneStats &st,
                                                 const CorpusSmem &cs = CorpusSmem()) {
  const uint32_t c = S.aux[ci];
  const int cx = (int)(c & 0xFFFFu);
  const uint32_t clin = (c >> 16) * (uint32_t)J.cw + (uint32_t)cx, k0 = 1u + j * CH;
  uint32_t own_x = 0, own_y = 0;  // the candidate's own pixel: its colour is what a win commits (synthesize.h:403-419)
  if (j == 0u) {
    if (MAPS) { const uint2 t = __ldg(J.corpus8 + clin); own_x = t.x; own_y = t.y; }
    else own_x = rs_corpus4<SMEMC>(J, cs, clin);
  }
  uint32_t part = rs_chunk_sum<MAPS, CH, SMEMC>(J, lutc, lutm, S.nb, S.map, cx, clin, k0, cs);
  if (j == 0u) {
    if (MAPS) part += rs_lut3(lutm, __vabsdiffu4(own_y, S.map[0]));
    hcol[ci] = own_x & 0xFFFFFFu;
    st.issued++;
    st.compares++;
  }
  st.compares += (k0 < K) ? min((uint32_t)CH, K - k0) : 0u;
  return part;
}

struct TeamShared {  // per team of the latency kernel (k_synth_pass_team, below)
  unsigned long long best;
  uint32_t hsum[RS_MAX_NB];
  uint32_t hcnt[RS_MAX_NB];  // chunks of each heuristic candidate added so far
  uint32_t v, K, nHeur, alive;
  uint32_t win_pt, win_col;  // corpus point and colour of the winning probe, written by the lane that evaluated it
};
// Dynamic shared memory of a pass kernel with `scratch_slots` warp scratches: tables, scratches, two mbarriers, team blocks.
template <bool MAPS, int NB>
__host__ __device__ constexpr unsigned pass_smem_bytes(int scratch_slots) {
  return (MAPS ? 2u : 1u) * RS_LUT_WORDS * 4u + (unsigned)sizeof(WarpScratch<MAPS, NB>) * (unsigned)scratch_slots + 16u +
         (unsigned)sizeof(TeamShared) * RS_TEAM_SLOTS;
}
struct PassSmem {  // carve-up of the dynamic shared memory of the pass kernels
  unsigned lutc, lutm;  // shared-space addresses of this lane's table columns
  void *scratch;
  uint64_t *bar;
};
template <bool MAPS, int NB, int NSCRATCH>
__device__ __forceinline__ PassSmem rs_pass_smem(const RsDev &J, unsigned char *smem_raw) {
  uint32_t *lutc = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *lutm = lutc + RS_LUT_WORDS;  // only staged when MAPS
  const unsigned lut_bytes = (MAPS ? 2u : 1u) * RS_LUT_WORDS * 4u;
  PassSmem P;
  P.scratch = smem_raw + lut_bytes;
  P.bar = reinterpret_cast<uint64_t *>(smem_raw + lut_bytes + sizeof(WarpScratch<MAPS, NB>) * NSCRATCH);
  rs_stage_tables<MAPS>(J, lutc, lutm, P.bar);
  P.lutc = (unsigned)__cvta_generic_to_shared(lutc) + (threadIdx.x & (RS_LUT_REP - 1u)) * 4u;
  P.lutm = (unsigned)__cvta_generic_to_shared(lutm) + (threadIdx.x & (RS_LUT_REP - 1u)) * 4u;
  return P;
}

// ---- throughput mode: one warp per visit -------------------------------------------------------------------
// SMEMC: the corpus (no map channels) lives in the shared memory of the CTA, or of the two CTAs of its cluster; the launch
// passes the slice length in J.sc_slice (pixels per CTA, a multiple of 4) and sizes the dynamic shared memory for it.
// LW: lanes that prepare and commit a visit.  32 = one visit per warp.  16 (patches of at most 16 neighbours) = TWO visits per
// warp, consecutive ones, side by side in its halves through geometry, values, candidates and commit -- the steps that
// leave most lanes of a warp idle when a patch has 9 neighbours -- and one after the other, with all 32 lanes, through
// the distance loop.  Should the second visit's patch hold the first visit's pixel it runs after it instead of beside it.
template <bool MAPS, int CH, int NB, bool SMEMC, int LW>
__global__ void __launch_bounds__(TpShape<MAPS, NB, SMEMC, LW>::warps * 32, TpShape<MAPS, NB, SMEMC, LW>::ctas) k_synth_pass(const RsDev J) {
  constexpr int NV = 32 / LW;  // visits a warp works on at once
  constexpr int TPW = TpShape<MAPS, NB, SMEMC, LW>::warps;  // warps of this CTA
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RsCtrl *ctrl = J.ctrl;
  const bool stopped = rs_ld_u32_relaxed(&ctrl->stop) != 0u;
  if (!SMEMC && stopped) return;  // (a cluster leaves together: both CTAs go through the two cluster barriers below)
  const PassSmem P = rs_pass_smem<MAPS, NB, TPW * NV>(J, smem_raw);
  const unsigned lutc = P.lutc, lutm = P.lutm;
  WarpScratch<MAPS, NB> *Sw = reinterpret_cast<WarpScratch<MAPS, NB> *>(P.scratch) + (threadIdx.x >> 5) * NV;  // this warp's scratches
  const unsigned lane = threadIdx.x & 31u, grp = NV == 1 ? 0u : lane / (unsigned)LW;
  WarpScratch<MAPS, NB> &Sg = Sw[grp];  // my group's
  CorpusSmem cs;
  if (SMEMC) {
    // this CTA's slice of the canonical corpus (sentinel pixel included) -> shared memory, by TMA bulk copies
    uint32_t *slice = reinterpret_cast<uint32_t *>(smem_raw + ((pass_smem_bytes<MAPS, NB>(TPW * NV) + 127u) & ~127u));
    const unsigned rank = rs_cluster_ctarank(), nr = rs_cluster_nctarank();
    const uint32_t first = rank * J.sc_slice, total = J.cn + 1u;
    const uint32_t count = first < total ? min(J.sc_slice, total - first) : 0u;
    const uint32_t bytes = ((count * 4u) + 15u) & ~15u;  // (the corpus buffer is allocated with slack beyond the sentinel)
    uint64_t *cbar = P.bar + 1;
    if (threadIdx.x == 0) {
      rs_mbar_init(cbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      rs_mbar_expect_tx(cbar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u)
        rs_tma_load_1d(reinterpret_cast<unsigned char *>(slice) + off, reinterpret_cast<const unsigned char *>(J.corpus4 + first) + off,
                       min(32768u, bytes - off), cbar);
    }
    __syncthreads();
    rs_mbar_wait(cbar, 0);
    rs_cluster_sync();  // the peer's slice is complete before anyone reads it
    const unsigned base = (unsigned)__cvta_generic_to_shared(slice);
    cs.lo = rs_mapa(base, 0u);
    cs.hi = rs_mapa(base, nr - 1u) - (nr - 1u) * J.sc_slice * 4u;
    cs.split = nr > 1u ? J.sc_slice : 0xFFFFFFFFu;
  }
  LaneStats st;
  if (Grp<LW>::lane() == 0) Sg.st = WarpStats{0ull, 0ull, 0u, 0u, 0u, 0u, 0u, 0u};
  const bool regular = J.regular_r != 0u && ctrl->n_ctx.v + J.nT == (uint32_t)J.tw * (uint32_t)J.th;  // no unusable pixel anywhere
  const uint32_t pass = J.pass, seed = J.seed, nC = ctrl->n_corpus;
  uint32_t v0 = stopped ? J.seg_end : rs_claim_resolve<NV>(J, ctrl, rs_claim_issue<NV>(J, ctrl));
  while (v0 < J.seg_end) {
    const uint32_t vg = v0 + grp;  // my group's visit
    const bool valid = vg < J.seg_end;
    const uint32_t Kg = rs_visit_geometry<CH, LW>(J, Sg, vg, valid ? __ldg(J.targets + vg) : 0u, regular, valid);
    bool dep = false;
    if (NV > 1) {  // does the second visit read the pixel of the first?  (then it has to see its new value)
      __syncwarp();
      const unsigned k = Grp<LW>::lane();
      dep = __any_sync(RS_FULL, grp == 1u && valid && k >= 1u && k < Kg && Sg.q[k] == Sw[0].vis.selfq);
    }
    for (unsigned round = 0; round < (dep ? 2u : 1u); round++) {
      const bool act = valid && (NV == 1 || !dep || grp == round);
      Visit V;
      V.v = vg; V.K = Kg; V.nHeur = 0u;
      rs_visit_values<LW>(J, Sg, vg, act ? Kg : 0u);
      rs_visit_candidates<SMEMC, LW>(J, ctrl, Sg, V, vg, act ? Kg : 0u, cs, act);
      V.K = Kg;
      // ---- evaluate, all 32 lanes on one visit at a time: heuristic candidates first, then the random probes
      //      (lib/synthesize.h:583-604)
      uint32_t gSum = 0xFFFFFFFFu, gWin = RS_NO_SRC;
      int gIdx = 0x7FFFFFFF;
#pragma unroll 1
      for (unsigned x = 0; x < (unsigned)NV; x++) {
        if (NV > 1 && !__shfl_sync(RS_FULL, (int)act, (int)(x * LW))) continue;
        WarpScratch<MAPS, NB> &S = Sw[x];
        const uint32_t v = v0 + x;
        const uint32_t K = NV == 1 ? Kg : __shfl_sync(RS_FULL, Kg, (int)(x * LW));
        const uint32_t nHeur = NV == 1 ? V.nHeur : __shfl_sync(RS_FULL, V.nHeur, (int)(x * LW));
        uint32_t bestSum = 0xFFFFFFFFu, bestLin = RS_NO_SRC;
        int bestIdx = 0x7FFFFFFF, bestCx = 0;
        uint32_t *hsum = S.off;  // offsets are dead by now (WarpScratch)
        uint32_t *hcol = S.q;    // so are the per-neighbour candidates: colour of each heuristic candidate
        const uint32_t hv = rs_probe_hash_visit(seed, pass, v);  // the two visit-constant rounds of rs_probe_hash
        // Heuristic candidates (few, and the likely winners): every (candidate, chunk) pair gets a lane, so all lanes
        // work instead of nHeur of them; full sums, then "first candidate with the minimum sum" as ever.
        if (nHeur) {
          const uint32_t nchr = (K + CH - 2u) / CH, nch = nchr ? nchr : 1u, inv = c_inv32[nch];
          hsum[lane] = 0u;        // (nHeur <= NB <= 64: two slots per lane cover every candidate)
          if (NB > 32) hsum[lane + 32u] = 0u;
          __syncwarp();
          for (uint32_t t = lane; t < nHeur * nch; t += 32) {
            const uint32_t ci = nch > 1u ? __umulhi(t, inv) : t, j = t - ci * nch;  // t / nch, exact while t * nch < 2^32
            atomicAdd(&hsum[ci], rs_heur_pair<MAPS, CH, NB, SMEMC>(J, lutc, lutm, S, hcol, K, ci, j, st, cs));
          }
          __syncwarp();
          const uint32_t h0 = lane < nHeur ? hsum[lane] : 0xFFFFFFFFu, h1 = (NB > 32 && lane + 32u < nHeur) ? hsum[lane + 32u] : 0xFFFFFFFFu;
          const uint32_t msum = __reduce_min_sync(RS_FULL, min(h0, h1));
          const int midx = __reduce_min_sync(RS_FULL, (h0 == msum) ? (int)lane : ((h1 == msum) ? (int)lane + 32 : 0x7FFFFFFF));
          bestSum = msum;
          bestIdx = midx;
        }
        if (bestSum != 0u)
          rs_eval_range<MAPS, CH, SMEMC>(J, lutc, lutm, S.nb, S.map, K, (int)nHeur, (int)(nHeur + J.probes),
                              [&](int i) { return rs_corpus_point(J, nC, rs_range(rs_mix32(hv + ((uint32_t)i - nHeur) * 0xC2B2AE35u), nC)); },
                              bestSum, bestIdx, bestLin, bestCx, st.compares, st.issued, cs);
        if (NV == 1 || grp == x) {
          gSum = bestSum;
          gIdx = bestIdx;
          gWin = bestLin != RS_NO_SRC ? ((uint32_t)bestCx | (((bestLin - (uint32_t)bestCx) / (uint32_t)J.cw) << 16)) : RS_NO_SRC;
        }
      }
      rs_visit_finish<true, SMEMC, LW>(J, ctrl, Sg, V, gSum, gIdx, gWin, 0u, false, cs, act);
      if (NV > 1) __syncwarp();
    }
    v0 = rs_claim_resolve<NV>(J, ctrl, rs_claim_issue<NV>(J, ctrl));
  }
  __syncwarp();
  if (NV > 1 && (!SMEMC || !stopped) && lane == 0) rs_flush_warp_stats(J, ctrl, &Sw[1].st);
  if (!SMEMC || !stopped) rs_pass_epilogue(J, ctrl, st, &Sw[0].st);
  if (SMEMC) rs_cluster_sync();  // nobody leaves while its peer may still read its slice
}

// ---- latency mode: a team of W warps per visit ---------------------------------------------------------------
// Dependency-bound phases (pass 0, small holes) are limited by depth x per-visit latency, not by throughput.
// Warp 0 of a team prepares the visit as above; then all W*32 lanes evaluate: (A) every (heuristic candidate,
// chunk) pair in one gather round, sums merged by shared-memory atomics; (B) the random probes, one
// per lane per round, early-out against a team-shared packed best (sum << 32 | index, atomicMin) -- the same
// "first candidate with the minimum full sum" rule, hence the same bits as the warp kernel.

__device__ __forceinline__ void rs_team_sync(unsigned id, unsigned nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One visit of a team, in the order things become possible (the critical path is what latency mode is about):
//   warp 0                                      the other warps (one probe per lane)
//   claim, target point, patch GEOMETRY
//   ---------------------------------------------------------------- barrier A
//   wait for the neighbours, read VALUES        probe's corpus point, gathers of its first chunk + its own pixel
//   ---------------------------------------------------------------- barrier B
//   heuristic CANDIDATES                        table lookups of the first chunk -> partial sum
//   ---------------------------------------------------------------- barrier C
//   all warps: (heuristic candidate, chunk) pairs; a candidate's last chunk enters it in the shared best -> barrier D
//   all warps: probes continue from their second chunk, early-out against the shared best; rest of the probes
//   ---------------------------------------------------------------- barrier F
//   the lane that owns the winning probe publishes its point + colour -> barrier G -> warp 0 commits
// So the gathers of every probe's first chunk, and the colour a winning probe commits, are fetched while the visit
// is still waiting for its dependencies; none of them is on the critical path.
template <bool MAPS, int CH, int NB>
__global__ void __launch_bounds__(RS_TEAM_WARPS * 32, 2) k_synth_pass_team(const RsDev J, const unsigned W) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RsCtrl *ctrl = J.ctrl;
  if (rs_ld_u32_relaxed(&ctrl->stop)) return;
  const PassSmem P = rs_pass_smem<MAPS, NB, RS_TEAM_SLOTS>(J, smem_raw);
  const unsigned lutc = P.lutc, lutm = P.lutm;
  TeamShared *tshared = reinterpret_cast<TeamShared *>(P.bar + 2);

  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned team = warp / W, wt = warp % W, T = W * 32u, tid = wt * 32u + lane;
  const unsigned bar_id = 1u + team;
  WarpScratch<MAPS, NB> &S = reinterpret_cast<WarpScratch<MAPS, NB> *>(P.scratch)[team];
  TeamShared &TS = tshared[team];
  const uint32_t pass = J.pass, seed = J.seed, nC = ctrl->n_corpus;
  const uint32_t nPre = min(J.probes, T - 32u);  // probes fetched ahead, one per lane of warps 1..W-1
  const bool regular = J.regular_r != 0u && ctrl->n_ctx.v + J.nT == (uint32_t)J.tw * (uint32_t)J.th;  // no unusable pixel anywhere
  LaneStats st;
  Visit V;
  if (wt == 0 && lane == 0) S.st = WarpStats{0ull, 0ull, 0u, 0u, 0u, 0u, 0u, 0u};
  while (true) {
    if (wt == 0) {
      const uint32_t vc = rs_claim_resolve(J, ctrl, rs_claim_issue(J, ctrl));
      const bool ok = vc < J.seg_end;
      uint32_t Kc = 0;
      if (ok) Kc = rs_visit_geometry<CH>(J, S, vc, __ldg(J.targets + vc), regular);
      if (lane == 0) { TS.alive = ok ? 1u : 0u; TS.v = vc; TS.K = Kc; TS.nHeur = 0u; TS.best = ~0ull; TS.win_pt = RS_NO_SRC; }
      for (uint32_t i = lane; i < RS_MAX_NB; i += 32) { TS.hsum[i] = 0u; TS.hcnt[i] = 0u; }
    }
    rs_team_sync(bar_id, T);  // A
    if (!TS.alive) break;
    const uint32_t v = TS.v, K = TS.K;
    const uint32_t hv = rs_probe_hash_visit(seed, pass, v);
    // ---- my probe (warps 1..W-1): point, first-chunk gathers, own pixel -- all independent of other visits
    const bool pre = wt != 0 && (tid - 32u) < nPre;
    uint32_t pc = 0, pclin = 0, pown_x = 0, pown_y = 0, ppartial = 0, cp[CH], cm[CH];
    int pcx = 0;
    if (wt == 0) {
      rs_visit_values(J, S, v, K);
    } else if (pre) {
      pc = rs_corpus_point(J, nC, rs_range(rs_mix32(hv + (tid - 32u) * 0xC2B2AE35u), nC));
      pcx = (int)(pc & 0xFFFFu);
      pclin = (pc >> 16) * (uint32_t)J.cw + (uint32_t)pcx;
      if (MAPS) { const uint2 t = __ldg(J.corpus8 + pclin); pown_x = t.x; pown_y = t.y; }
      else pown_x = __ldg(J.corpus4 + pclin);
      rs_chunk_gather<MAPS, CH>(J, S.nb, pcx, pclin, 1u, cp, cm);
    }
    rs_team_sync(bar_id, T);  // B
    if (wt == 0) {
      rs_visit_candidates(J, ctrl, S, V, v, K);
      if (lane == 0) TS.nHeur = V.nHeur;
    } else if (pre) {
      ppartial = rs_chunk_reduce<MAPS, CH>(lutc, lutm, S.nb, S.map, 1u, cp, cm);
      if (MAPS) ppartial += rs_lut3(lutm, __vabsdiffu4(pown_y, S.map[0]));
    }
    rs_team_sync(bar_id, T);  // C
    const uint32_t nHeur = TS.nHeur;
    // ---- heuristic candidates: all (candidate, chunk) pairs at once
    const uint32_t nchr = (K + CH - 2u) / CH, nch = nchr ? nchr : 1u;
    if (nHeur) {
      for (uint32_t t = tid; t < nHeur * nch; t += T) {
        const uint32_t ci = nch > 1u ? __umulhi(t, c_inv32[nch]) : t, j = t - ci * nch;
        atomicAdd(&TS.hsum[ci], rs_heur_pair<MAPS, CH, NB>(J, lutc, lutm, S, S.q, K, ci, j, st));
        __threadfence_block();
        // the lane that adds a candidate's last chunk holds its full sum: it enters "first candidate with the minimum
        // full sum" directly (packed sum << 32 | index, atomicMin) -- no barrier and no scan by warp 0 in between
        if (atomicAdd(&TS.hcnt[ci], 1u) + 1u == nch)
          atomicMin(&TS.best, ((unsigned long long)atomicAdd(&TS.hsum[ci], 0u) << 32) | ci);
      }
      rs_team_sync(bar_id, T);  // D
    }
    // ---- random probes, early-out against the shared best (sum << 32 | index, atomicMin): the fetched-ahead ones
    //      resume after their first chunk, the rest (probes >= nPre) run from the start
    volatile unsigned long long *vbest = &TS.best;
    const uint32_t selfmap = MAPS ? S.map[0] : 0u;
    unsigned long long mykey = ~0ull;  // key of a probe this lane took to the end
    uint32_t mypt = 0, mycol = 0;
    if ((uint32_t)(*vbest >> 32) != 0u) {
      if (pre) {
        const unsigned long long idx = (unsigned long long)(nHeur + (tid - 32u));
        uint32_t partial = ppartial, k0 = 1u + CH;
        bool alive = (((unsigned long long)partial << 32) | idx) < *vbest;
        st.issued++;
        while (alive && k0 < K) {
          partial += rs_chunk_sum<MAPS, RS_CHUNK_CONT>(J, lutc, lutm, S.nb, S.map, pcx, pclin, k0);
          k0 += RS_CHUNK_CONT;
          alive = !((((unsigned long long)partial << 32) | idx) > *vbest);
        }
        st.compares += min(k0, K);
        if (alive) { mykey = ((unsigned long long)partial << 32) | idx; mypt = pc; mycol = pown_x & 0xFFFFFFu; atomicMin(&TS.best, mykey); }
      }
      for (uint32_t j = nPre + tid; j < J.probes; j += T) {
        if ((uint32_t)(*vbest >> 32) == 0u) break;  // perfect match: nothing later is evaluated (synthesize.h:599)
        const uint32_t c = rs_corpus_point(J, nC, rs_range(rs_mix32(hv + j * 0xC2B2AE35u), nC));
        const int cx = (int)(c & 0xFFFFu);
        const uint32_t clin = (c >> 16) * (uint32_t)J.cw + (uint32_t)cx;
        const unsigned long long idx = (unsigned long long)(nHeur + j);
        uint32_t partial = 0, k0 = 1, own_x, own_y = 0;
        if (MAPS) { const uint2 t = __ldg(J.corpus8 + clin); own_x = t.x; own_y = t.y; }
        else own_x = __ldg(J.corpus4 + clin);
        if (MAPS) partial = rs_lut3(lutm, __vabsdiffu4(own_y, selfmap));
        bool alive = true;
        st.issued++;
        do {
          partial += rs_chunk_sum<MAPS, CH>(J, lutc, lutm, S.nb, S.map, cx, clin, k0);
          k0 += CH;
          if ((((unsigned long long)partial << 32) | idx) > *vbest) { alive = false; break; }
        } while (k0 < K);
        st.compares += min(k0, K);
        if (alive) {
          const unsigned long long key = ((unsigned long long)partial << 32) | idx;
          if (key < mykey) { mykey = key; mypt = c; mycol = own_x & 0xFFFFFFu; }
          atomicMin(&TS.best, key);
        }
      }
    }
    rs_team_sync(bar_id, T);  // F
    const unsigned long long key = TS.best;  // final; warp 0 resets it for the next visit only after barrier G
    if (mykey == key && mykey != ~0ull) { TS.win_pt = mypt; TS.win_col = mycol; }
    // warp 1 merges the visit's stamps while warp 0 commits and claims the next visit: what it needs leaves the
    // scratch (which warp 0 is about to reuse) before the barrier
    uint32_t sc0 = 0, sc1 = 0;
    if (wt == 1) {
      sc0 = lane < nHeur ? S.aux[lane] : 0u;
      sc1 = (NB > 32 && lane + 32u < nHeur) ? S.aux[lane + 32u] : 0u;
    }
    rs_team_sync(bar_id, T);  // G
    const uint32_t bestSum = (key == ~0ull) ? 0xFFFFFFFFu : (uint32_t)(key >> 32);
    const int bestIdx = (key == ~0ull) ? 0x7FFFFFFF : (int)(uint32_t)key;
    if (wt == 0) {
      rs_visit_finish<false>(J, ctrl, S, V, bestSum, bestIdx, TS.win_pt, TS.win_col, true);
    } else if (wt == 1) {
      const uint32_t stampEnd = (bestSum == 0u && (uint32_t)bestIdx < nHeur) ? (uint32_t)bestIdx : nHeur;
      const uint32_t epoch_idx = rs_epoch_of(J, v), epoch0 = epoch_idx * J.epoch_len;  // as in rs_visit_candidates
      rs_visit_stamps(J, ctrl, v, epoch_idx, ((pass + 1u) << 29) | epoch0, epoch_idx ? epoch0 - J.epoch_len : 0u, stampEnd, sc0, sc1);
    }
  }
  __syncwarp();
  rs_pass_epilogue(J, ctrl, st, wt == 0 ? &S.st : nullptr);
}

// ------------------------------------------------------------------------------ standalone best-fit kernel
template <bool MAPS>
__global__ void __launch_bounds__(RS_BF_WARPS * 32, 2)
    k_bestfit_batch(const RsDev J, uint32_t n_visits, const uint32_t *__restrict__ nb_begin,
                    const uint32_t *__restrict__ nb_offsets, const uint8_t *__restrict__ nb_pixels, int n_color,
                    int n_map, int map_bip, const uint32_t *__restrict__ cand_begin, const uint32_t *__restrict__ cands,
                    uint32_t *__restrict__ best_sum, int32_t *__restrict__ best_index) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const PassSmem P = rs_pass_smem<MAPS, RS_NB_FULL, RS_BF_WARPS>(J, smem_raw);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  WarpScratch<MAPS, RS_NB_FULL> &S = reinterpret_cast<WarpScratch<MAPS, RS_NB_FULL> *>(P.scratch)[warp];
  for (uint32_t v = blockIdx.x * RS_BF_WARPS + warp; v < n_visits; v += gridDim.x * RS_BF_WARPS) {
    const uint32_t nb0 = nb_begin[v], K = min(nb_begin[v + 1] - nb0, (uint32_t)RS_MAX_NB);
    const uint32_t nchr = (K + RS_CHUNK_SMALL - 2u) / RS_CHUNK_SMALL, kpad = 1u + (nchr ? nchr : 1u) * RS_CHUNK_SMALL;
    for (uint32_t k = lane; k < kpad; k += 32) {
      RsNb r;
      uint32_t mp = 0;
      if (k < K) {
        const uint32_t o = nb_offsets[nb0 + k];
        const uint8_t *p = nb_pixels + (size_t)(nb0 + k) * 8;
        uint32_t col = 0;
        for (int c = 0; c < n_color; c++) col |= (uint32_t)p[1 + c] << (8 * c);
        for (int c = 0; c < n_map; c++) mp |= (uint32_t)p[map_bip + c] << (8 * c);
        r.dx = rs_off_x(o); r.lin = rs_off_y(o) * J.cw + r.dx; r.pix = col; r.pen = J.penalty;
      } else {
        r.lin = 0; r.dx = RS_PAD_DX; r.pix = 0u; r.pen = 0u;
      }
      S.nb[k] = r;
      if (MAPS) S.map[k] = mp;
    }
    __syncwarp();
    const uint32_t c0 = cand_begin[v], nc = cand_begin[v + 1] - c0;
    uint32_t bestSum = 0xFFFFFFFFu, cmp = 0, iss = 0, blin = 0;
    int bestIdx = 0x7FFFFFFF, bcx = 0;
    if (K)
      rs_eval_range<MAPS, RS_CHUNK_SMALL>(J, P.lutc, P.lutm, S.nb, S.map, K, 0, (int)nc,
                          [&](int i) { return __ldg(cands + c0 + i); }, bestSum, bestIdx, blin, bcx, cmp, iss);
    else if (nc) { bestSum = 0u; bestIdx = 0; }  // an empty patch matches anything perfectly
    if (lane == 0) {
      best_sum[v] = bestSum;
      best_index[v] = (bestIdx == 0x7FFFFFFF) ? -1 : bestIdx;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------- workspaces
// Device buffers, pinned staging, stream and events are expensive to create (cudaFree synchronises the
// device); they live in pooled workspaces that outlive jobs and only ever grow.  A workspace also keeps the
// sorted neighbour-offset table of the last image size it served, which batches of equal-sized jobs reuse.
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};
struct PassVariant {
  void (*tp)(const RsDev) = nullptr;                    // k_synth_pass<maps, chunk, nb>
  void (*team)(const RsDev, const unsigned) = nullptr;  // k_synth_pass_team<maps, chunk, nb>
  size_t smem_tp = 0, smem_team = 0;
  int tp_threads = RS_TP_WARPS * 32;                    // threads of a CTA of `tp` (TpShape)
  int grid = 0, grid_team = 0;                          // persistent grids: resident CTAs per SM x SMs
  int sms = 0;
  void (*tp_smemc)(const RsDev) = nullptr;              // k_synth_pass<false, chunk, nb, true>: corpus in shared memory
  size_t smemc_base = 0;                                // dynamic shared memory before the corpus slice
  uint32_t smemc_slice_max = 0;                         // pixels a CTA's slice can hold
  // two visits per warp (patches of at most RS_NB_SMALL neighbours): k_synth_pass<..., 16>
  void (*tp_pair)(const RsDev) = nullptr;
  void (*tp_pair_smemc)(const RsDev) = nullptr;
  size_t smem_tp_pair = 0, smemc_base_pair = 0;
  uint32_t smemc_slice_max_pair = 0;
};
struct Workspace {
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: work that may run beside the main stream's
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evLater = nullptr;
  cudaEvent_t ev0 = nullptr, evG = nullptr, ev1 = nullptr, evDone = nullptr;
  DevBuf raw_t, raw_c, corpus, W, meta, tmaps, targets, cpts, offsets, lut256, lut_rep, prober, colours,
      sources, ctrl, sort_keys_in, sort_keys_out, sort_vals_in, sort_tmp, nb_lists, nb_counts, nb_later, nb_later_counts, simg, smask, smask2,
      ord_keys_in, ord_keys_out, ord_vals_in, ord_vals_out, ord_tmp, ord_first, ord_points, ord_flags, ord_raw,
      cbits, ccounts, cbefore, csamples, ctx_blocks;
  void *pin = nullptr;  // pinned staging (H2D inputs, D2H results)
  size_t pin_cap = 0;
  void *pin_order = nullptr;  // pinned staging of a visit order
  size_t pin_order_cap = 0;
  void *pin_sort = nullptr;   // pinned staging of rs_job_sort_pairs
  size_t pin_sort_cap = 0;
  void *pin_raw = nullptr;    // pinned buffer the host's PRNG producer fills (rs_job_raw_buffer)
  size_t pin_raw_cap = 0;
  RsTargetDigest *h_digest = nullptr;  // pinned
  cudaEvent_t evDigest = nullptr;
  cudaEvent_t evSel = nullptr, evOrder = nullptr;  // the selection (mask plane / target pixmap) is on the device; the visit order is made
  cudaEvent_t evAux = nullptr;                     // input-independent setup queued on the side stream (prober memset) is done
  unsigned int *h_ticks = nullptr;
  int *h_cancel = nullptr;
  RsCtrl *h_ctrl = nullptr;
  int off_w = 0, off_h = 0;  // dimensions the resident offsets table was built for
  uint32_t off_n = 0;
  PassVariant variant[8];  // the pass-kernel instantiations and their persistent grids, index = rs_variant()
};
static std::atomic<int> g_job_slots{1};
extern "C" void rs_cuda_set_job_slots(int slots) { g_job_slots.store(slots < 1 ? 1 : slots); }
static std::mutex g_pool_mutex;
static std::vector<Workspace *> g_pool;

static int ws_ensure(DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  RS_CHECK(cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}
static int ws_ensure_pinned(Workspace *w, size_t bytes) {
  if (bytes <= w->pin_cap) return 0;
  if (w->pin) cudaFreeHost(w->pin);
  w->pin = nullptr;
  w->pin_cap = 0;
  const size_t want = bytes + bytes / 8 + 4096;
  RS_CHECK(cudaHostAlloc(&w->pin, want, cudaHostAllocDefault));
  w->pin_cap = want;
  return 0;
}
static void ws_free(Workspace *w) {
  DeviceGuard guard(w->device);
  DevBuf *all[] = {&w->raw_t, &w->raw_c, &w->corpus, &w->W, &w->meta, &w->tmaps, &w->targets, &w->cpts, &w->offsets,
                   &w->lut256, &w->lut_rep, &w->prober, &w->colours, &w->sources, &w->ctrl,
                   &w->sort_keys_in, &w->sort_keys_out, &w->sort_vals_in, &w->sort_tmp, &w->nb_lists, &w->nb_counts, &w->nb_later, &w->nb_later_counts,
                   &w->simg, &w->smask, &w->smask2, &w->ord_keys_in, &w->ord_keys_out, &w->ord_vals_in, &w->ord_vals_out,
                   &w->ord_tmp, &w->ord_first, &w->ord_points, &w->ord_flags, &w->ord_raw, &w->cbits, &w->ccounts, &w->cbefore,
                   &w->csamples, &w->ctx_blocks};
  for (DevBuf *b : all) if (b->p) cudaFree(b->p);
  if (w->pin) cudaFreeHost(w->pin);
  if (w->pin_order) cudaFreeHost(w->pin_order);
  if (w->pin_sort) cudaFreeHost(w->pin_sort);
  if (w->pin_raw) cudaFreeHost(w->pin_raw);
  if (w->h_digest) cudaFreeHost(w->h_digest);
  if (w->evDigest) cudaEventDestroy(w->evDigest);
  if (w->evSel) cudaEventDestroy(w->evSel);
  if (w->evOrder) cudaEventDestroy(w->evOrder);
  if (w->evAux) cudaEventDestroy(w->evAux);
  if (w->h_ticks) cudaFreeHost(w->h_ticks);
  if (w->h_cancel) cudaFreeHost(w->h_cancel);
  if (w->h_ctrl) cudaFreeHost(w->h_ctrl);
  if (w->ev0) cudaEventDestroy(w->ev0);
  if (w->evG) cudaEventDestroy(w->evG);
  if (w->ev1) cudaEventDestroy(w->ev1);
  if (w->evDone) cudaEventDestroy(w->evDone);
  if (w->evFork) cudaEventDestroy(w->evFork);
  if (w->evJoin) cudaEventDestroy(w->evJoin);
  if (w->evLater) cudaEventDestroy(w->evLater);
  if (w->stream2) cudaStreamDestroy(w->stream2);
  if (w->stream) cudaStreamDestroy(w->stream);
  delete w;
}

static size_t pass_smem(bool maps, int scratch_slots, bool nb_full = true) {
  return maps ? (nb_full ? pass_smem_bytes<true, RS_NB_FULL>(scratch_slots) : pass_smem_bytes<true, RS_NB_SMALL>(scratch_slots))
              : (nb_full ? pass_smem_bytes<false, RS_NB_FULL>(scratch_slots) : pass_smem_bytes<false, RS_NB_SMALL>(scratch_slots));
}
// The instantiations of the two pass kernels: map channels x chunk size x scratch size (index = rs_variant()).
static int rs_variant(bool maps, bool chunk_large, bool nb_full) { return (maps ? 4 : 0) + (chunk_large ? 2 : 0) + (nb_full ? 1 : 0); }
template <int CH, int NB>
static int configure_pair_smemc(PassVariant &V, size_t smem_max) {
  V.tp_pair_smemc = k_synth_pass<false, CH, NB, true, 16>;
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<false, CH, NB, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  return 0;
}
// The two-visits-per-warp instantiations exist for the small-patch scratch only (a half warp covers 16 neighbours).
template <bool MAPS, int CH, int NB>
static typename std::enable_if<(NB > 16), int>::type configure_pair_kernel(Workspace *, PassVariant &, size_t) { return 0; }
template <bool MAPS, int CH, int NB>
static typename std::enable_if<(NB <= 16), int>::type configure_pair_kernel(Workspace *w, PassVariant &V, size_t sm_max) {
  const size_t smem = pass_smem(MAPS, RS_TP_WARPS * 2, false);
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<MAPS, CH, NB, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int carve = (int)std::min<size_t>(100, ((smem + 1024) * 100 + sm_max - 1) / sm_max);
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<MAPS, CH, NB, false, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  V.tp_pair = k_synth_pass<MAPS, CH, NB, false, 16>;
  V.smem_tp_pair = smem;
  if (!MAPS) {
    const size_t smem_max = 232448;  // 227 KB per CTA on sm_100
    const size_t base = (smem + 127) & ~(size_t)127;
    configure_pair_smemc<CH, NB>(V, smem_max);
    V.smemc_base_pair = base;
    V.smemc_slice_max_pair = (uint32_t)(((smem_max - base) / 4) & ~(size_t)3);
  }
  (void)w;
  return 0;
}
template <bool MAPS, int CH, int NB>
static int configure_pass_kernel(Workspace *w) {
  const bool full = NB == RS_NB_FULL;
  constexpr int tp_warps = TpShape<MAPS, NB, false, 32>::warps;
  const size_t smem_tp = pass_smem(MAPS, tp_warps, full), smem_team = pass_smem(MAPS, RS_TEAM_SLOTS, full);
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<MAPS, CH, NB, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tp));
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass_team<MAPS, CH, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_team));
  int per_sm = 0, per_sm_team = 0, sms = 0;
  RS_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_synth_pass<MAPS, CH, NB, false, 32>, tp_warps * 32, smem_tp));
  RS_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_team, k_synth_pass_team<MAPS, CH, NB>, RS_TEAM_WARPS * 32, smem_team));
  RS_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device));
  if (per_sm < 1 || per_sm_team < 1) { g_err = "the pass kernels do not fit on an SM"; return 100; }
  // what shared memory does not take stays L1: ask for no more than the resident CTAs need
  // (percent of the 228 KB an SM can give to shared memory; every resident CTA also takes 1 KB of system space)
  const size_t sm_max = 233472;
  const int carve_tp = (int)std::min<size_t>(100, ((smem_tp + 1024) * per_sm * 100 + sm_max - 1) / sm_max);
  const int carve_team = (int)std::min<size_t>(100, ((smem_team + 1024) * per_sm_team * 100 + sm_max - 1) / sm_max);
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass<MAPS, CH, NB, false, 32>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_tp));
  RS_CHECK(cudaFuncSetAttribute(k_synth_pass_team<MAPS, CH, NB>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_team));
  PassVariant &V = w->variant[rs_variant(MAPS, CH == RS_CHUNK_LARGE, full)];
  V.tp = k_synth_pass<MAPS, CH, NB, false, 32>;
  V.team = k_synth_pass_team<MAPS, CH, NB>;
  V.smem_tp = smem_tp; V.smem_team = smem_team;
  V.tp_threads = tp_warps * 32;
  V.grid = per_sm * sms; V.grid_team = per_sm_team * sms;
  V.sms = sms;
  if (!MAPS) {  // the corpus-in-shared-memory instantiation: one CTA per SM, everything the SM has left goes to the slice
    const size_t smem_max = 232448;  // 227 KB per CTA on sm_100
    const size_t base = (pass_smem(false, RS_TP_WARPS, full) + 127) & ~(size_t)127;  // (this instantiation: RS_TP_WARPS warps, one CTA)
    V.tp_smemc = k_synth_pass<false, CH, NB, true, 32>;
    V.smemc_base = base;
    V.smemc_slice_max = (uint32_t)(((smem_max - base) / 4) & ~(size_t)3);
    RS_CHECK(cudaFuncSetAttribute(k_synth_pass<false, CH, NB, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  }
  if (int rc = configure_pair_kernel<MAPS, CH, NB>(w, V, sm_max)) return rc;
  return 0;
}

static int ws_acquire(Workspace **out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_err = "no CUDA device"; return 100; }
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    // most recently released first: a caller that runs one job after another keeps getting the workspace whose
    // buffers, pinned staging and offsets table already fit its jobs (oldest-first rotated through every workspace a
    // batch had left in the pool, growing each in turn: 2x the call time of the first jobs after a batch)
    for (size_t i = g_pool.size(); i-- > 0;)
      if (g_pool[i]->device == dev) {
        *out = g_pool[i];
        g_pool.erase(g_pool.begin() + i);
        return 0;
      }
  }
  Workspace *w = new Workspace();
  w->device = dev;
  int rc = configure_pass_kernel<false, RS_CHUNK_SMALL, RS_NB_SMALL>(w);
  if (!rc) rc = configure_pass_kernel<false, RS_CHUNK_SMALL, RS_NB_FULL>(w);
  if (!rc) rc = configure_pass_kernel<false, RS_CHUNK_LARGE, RS_NB_SMALL>(w);
  if (!rc) rc = configure_pass_kernel<false, RS_CHUNK_LARGE, RS_NB_FULL>(w);
  if (!rc) rc = configure_pass_kernel<true, RS_CHUNK_SMALL, RS_NB_SMALL>(w);
  if (!rc) rc = configure_pass_kernel<true, RS_CHUNK_SMALL, RS_NB_FULL>(w);
  if (!rc) rc = configure_pass_kernel<true, RS_CHUNK_LARGE, RS_NB_SMALL>(w);
  if (!rc) rc = configure_pass_kernel<true, RS_CHUNK_LARGE, RS_NB_FULL>(w);
  if (rc) { delete w; return rc; }
#define WCHK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); ws_free(w); return 100; } } while (0)
  WCHK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  WCHK(cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking));
  WCHK(cudaEventCreateWithFlags(&w->evFork, cudaEventDisableTiming));
  WCHK(cudaEventCreateWithFlags(&w->evJoin, cudaEventDisableTiming));
  WCHK(cudaEventCreateWithFlags(&w->evLater, cudaEventDisableTiming));
  WCHK(cudaEventCreate(&w->ev0));
  WCHK(cudaEventCreate(&w->evG));
  WCHK(cudaEventCreate(&w->ev1));
  WCHK(cudaEventCreateWithFlags(&w->evDone, cudaEventDisableTiming | cudaEventBlockingSync));
  WCHK(cudaEventCreateWithFlags(&w->evDigest, cudaEventDisableTiming));
  WCHK(cudaEventCreateWithFlags(&w->evSel, cudaEventDisableTiming));
  WCHK(cudaEventCreateWithFlags(&w->evOrder, cudaEventDisableTiming));
  WCHK(cudaEventCreateWithFlags(&w->evAux, cudaEventDisableTiming));
  WCHK(cudaHostAlloc(&w->h_digest, sizeof(RsTargetDigest), cudaHostAllocDefault));
  WCHK(cudaHostAlloc(&w->h_ticks, 6 * sizeof(unsigned int), cudaHostAllocMapped));
  WCHK(cudaHostAlloc(&w->h_cancel, sizeof(int), cudaHostAllocMapped));
  WCHK(cudaHostAlloc(&w->h_ctrl, RS_CTRL_COPY_BYTES, cudaHostAllocDefault));
#undef WCHK
  *out = w;
  return 0;
}
static void ws_release(Workspace *w) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  g_pool.push_back(w);
}
extern "C" void rs_cuda_release_cached(void) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  for (Workspace *w : g_pool) ws_free(w);
  g_pool.clear();
}

// ---- sorted neighbour offsets on the device (replaces prepareSortedOffsets, lib/engine.c:465-497) ----
// All (x,y), |x|<w, |y|<h, ascending x^2+y^2, equal distances in reverse row-major order (what glibc's merge
// sort makes of the reference's never-equal comparator).  Generated in reverse row-major order and sorted
// by distance with a stable radix sort.
__global__ void k_gen_offsets(int w, int h, uint32_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t row = i / (uint32_t)(2 * w - 1), col = i % (uint32_t)(2 * w - 1);
  const int y = (h - 1) - (int)row, x = (w - 1) - (int)col;
  keys[i] = (uint32_t)(x * x + y * y);
  vals[i] = ((uint32_t)x & 0xFFFFu) | ((uint32_t)y << 16);
}

// ------------------------------------------------------------------------------------------------ the job
// ---- the reference's shuffle on the device, exactly (lib/orderTarget.h:38-53, modes 0 and 1) ----
// The reference runs  for i in [0,n): swap(a[i], a[j_i])  with j_i drawn over the WHOLE vector -- a sequential chain of
// transpositions, not a Fisher-Yates shuffle.  What ends up at position p can still be found independently for every p
// by walking the chain BACKWARDS: at time t the content of position x was last moved by the latest swap before t that
// touches x, which is swap x itself (if x < t) or a swap i with j_i == x.  Follow it to where the content came from
// and repeat until no earlier swap touches the position; the content is then the initial a[x].  Chains are short (2
// hops on average, < 16 for a million points).  The swaps that target a position are found through the pairs (j_i, i)
// sorted by j_i (stable radix sort: ascending i within a target).
__global__ void k_iota(uint32_t *__restrict__ v, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void k_run_heads(const uint32_t *__restrict__ sorted_j, uint32_t n, uint32_t *__restrict__ first) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && (k == 0 || sorted_j[k] != sorted_j[k - 1])) first[sorted_j[k]] = k;
}
__global__ void k_shuffle_trace(const uint32_t *__restrict__ draws, const uint32_t *__restrict__ sorted_j,
                                const uint32_t *__restrict__ sorted_i, const uint32_t *__restrict__ first,
                                const uint32_t *__restrict__ points, uint32_t n, uint32_t *__restrict__ ordered) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t x = p, t = n;
  while (true) {
    // latest swap before t that targets x (j_i == x): its run in the sorted pairs is ascending in i
    uint32_t best = 0xFFFFFFFFu;
    const uint32_t k0 = first[x];
    if (k0 != 0xFFFFFFFFu)
      for (uint32_t k = k0; k < n && sorted_j[k] == x; k++) {
        const uint32_t i = sorted_i[k];
        if (i >= t) break;
        best = i;
      }
    const bool own = x < t;  // swap x itself touches position x
    if (best == 0xFFFFFFFFu && !own) break;
    if (own && (best == 0xFFFFFFFFu || x >= best)) {  // the own swap is the latest (x == best: a self swap, no move)
      t = x;
      x = draws[x];
    } else {
      t = best;
      x = best;
    }
  }
  ordered[p] = points[x];
}
__global__ void k_target_flags(const uint8_t *__restrict__ raw, uint32_t n_px, int bpp, uint8_t *__restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_px) flags[i] = raw[(size_t)i * bpp] != 0 ? 1 : 0;
}

// Visit orders kept on the device, keyed by what they are a function of.  A batch of jobs with the same selection
// (frames of a video, a set of equally sized images with the same hole) orders its target points once.
struct OrderEntry {
  RsOrderKey key;
  int device = 0;
  uint32_t *dev = nullptr;
  uint32_t n = 0;
  unsigned long long stamp = 0;
  cudaEvent_t ready = nullptr;  // recorded behind the upload on the creating job's stream: other streams wait on it
  ~OrderEntry() {
    DeviceGuard guard(device);
    if (ready) cudaEventDestroy(ready);
    if (dev) cudaFree(dev);
  }
};
// Where a job's corpus lives on the device: in its workspace, or in an entry shared by the jobs of a batch that
// synthesise from the SAME corpus (one texture or style source, many targets: SURVEY.md section 8 f4).  The entry
// holds everything that is a function of the corpus alone -- canonical pixels, point list, bitmap + samples, point
// count -- built once per device; a second device of the batch copies it from the first over NVLink (peer copy)
// instead of staging it from the host again.
struct CorpusBufs {
  void *corpus = nullptr;
  uint32_t *cpts = nullptr, *cbits = nullptr;
  uint2 *csamples = nullptr;
};
struct SharedCorpus {
  unsigned long long batch = 0;
  const void *host_key = nullptr;  // the caller's corpus pixmap (identity within one batch call)
  int device = 0;
  int cw = 0, ch = 0, bpp = 0, n_color = 0, n_map = 0, map_bip = 0, alpha_bip = 0, alpha_source = 0;
  DevBuf corpus, cpts, cbits, csamples, n_corpus;
  size_t corpus_bytes = 0, cpts_bytes = 0, cbits_bytes = 0, csamples_bytes = 0;
  bool peer_copied = false;
  cudaEvent_t ready = nullptr;  // recorded behind the build (or the peer copy) on the creating job's stream
  ~SharedCorpus() {
    DeviceGuard guard(device);
    if (ready) cudaEventDestroy(ready);
    for (DevBuf *b : {&corpus, &cpts, &cbits, &csamples, &n_corpus}) if (b->p) cudaFree(b->p);
  }
};
static std::mutex g_corpus_mutex;
static std::vector<std::shared_ptr<SharedCorpus>> g_corpora;
static std::atomic<unsigned long long> g_corpus_peer_copies{0}, g_corpus_hits{0}, g_corpus_builds{0};
extern "C" void rs_cuda_drop_shared_corpora(unsigned long long batch) {
  std::vector<std::shared_ptr<SharedCorpus>> dropped;
  {
    std::lock_guard<std::mutex> lk(g_corpus_mutex);
    for (size_t i = 0; i < g_corpora.size();)
      if (g_corpora[i]->batch == batch) { dropped.push_back(g_corpora[i]); g_corpora.erase(g_corpora.begin() + i); } else i++;
  }
}
extern "C" void rs_cuda_shared_corpus_stats(unsigned long long *builds, unsigned long long *hits, unsigned long long *peer_copies) {
  if (builds) *builds = g_corpus_builds.load();
  if (hits) *hits = g_corpus_hits.load();
  if (peer_copies) *peer_copies = g_corpus_peer_copies.load();
}
struct RsJob {
  CorpusBufs cb;
  std::shared_ptr<SharedCorpus> shared;   // the batch's corpus entry this job reads, if any
  unsigned long long share_batch = 0;     // != 0: the corpus may be shared with other jobs of this batch (rs_job_share_corpus)
  RsJobDesc d;
  Workspace *ws = nullptr;
  bool maps = false;
  uint32_t nT = 0, nC = 0, nOff = 0, penalty = 0;
  uint32_t y_min = 0, y_max = 0;  // rows of the target image that hold target points
  uint32_t x_min = 0, x_max = 0;  // columns likewise (the whole width unless the selection digest narrowed them)
  size_t out_bytes = 0;           // pinned bytes the results need (rows with target points + sources)
  std::shared_ptr<OrderEntry> order;     // cached visit order this job reads, if any
  const uint32_t *targets_dev = nullptr;  // the visit order on the device (cache entry or the workspace's buffer)
  bool want_sources = false;
  float ms_passes = 0.f;
  bool later_lists = false;       // the patches of the passes >= 1 were gathered up front (k_gather_later)
  bool simple = false;            // staged by rs_job_stage_simple: results go back in the caller's image layout
  bool corpus_bits = false;       // the device built the corpus-point bitmap + samples (rs_corpus_point's select path)
  uint32_t launches = 0;          // pass-kernel launches of the last run
  uint32_t pass_launches[6] = {0, 0, 0, 0, 0, 0};
  uint32_t upload_launches = 0;   // kernels launched by the upload (init, offsets, compaction)
  float ms_synth = 0.f;           // CUDA-event time of the pass kernels alone (after the pass-0 patch gather)
  int off_w = 0, off_h = 0;       // dimensions of the full offsets table this job reads (0: a caller's partial table)
  bool result_direct = false;     // the caller's result buffer is page-locked: the rows go there straight from the device (download)
  bool ctx_counted = false;       // k_ctx_blocks ran for this job: RsCtrl::n_ctx holds the number of usable context pixels
};

extern "C" void rs_job_destroy(RsJob *j) {
  if (!j) return;
  if (j->ws) {
    cudaSetDevice(j->ws->device);
    cudaStreamSynchronize(j->ws->stream);
    if (j->ws->stream2) cudaStreamSynchronize(j->ws->stream2);  // (joined into the main stream on the good path; an error may leave it running)
    ws_release(j->ws);
  }
  delete j;
}

// The jobs of batch `batch` (any non-zero id, unique per batch call) that pass the same corpus pixmap to rs_job_stage
// share one device-resident corpus per device; rs_cuda_drop_shared_corpora(batch) releases the entries.
extern "C" void rs_job_share_corpus(RsJob *j, unsigned long long batch) { j->share_batch = batch; }
extern "C" int rs_job_create(const RsJobDesc *desc, RsJob **out) {
  *out = nullptr;
  if (desc->tw <= 0 || desc->th <= 0 || desc->cw <= 0 || desc->ch <= 0 || desc->tw > 32767 || desc->th > 32767 ||
      desc->cw > 32767 || desc->ch > 32767 || desc->bpp < 2 || desc->bpp > 8 || desc->n_color < 1 || desc->n_color > 3 ||
      desc->n_map < 0 || desc->n_map > 3 || desc->n_passes < 1 || desc->n_passes > 6) {
    g_err = "rs_job_create: descriptor out of range";
    return 100;
  }
  Workspace *w = nullptr;
  if (int rc = ws_acquire(&w)) return rc;
  RsJob *j = new RsJob();
  j->d = *desc;
  j->maps = desc->n_map > 0;
  j->ws = w;
  *out = j;
  return 0;
}

static int build_offsets_on_device(Workspace *w, int ow, int oh, uint32_t n) {
  cudaStream_t s = w->stream;
  if (int rc = ws_ensure(w->sort_keys_in, (size_t)n * 4)) return rc;
  if (int rc = ws_ensure(w->sort_keys_out, (size_t)n * 4)) return rc;
  if (int rc = ws_ensure(w->sort_vals_in, (size_t)n * 4)) return rc;
  if (int rc = ws_ensure(w->offsets, (size_t)n * 4)) return rc;
  k_gen_offsets<<<(n + 255) / 256, 256, 0, s>>>(ow, oh, n, (uint32_t *)w->sort_keys_in.p, (uint32_t *)w->sort_vals_in.p);
  const uint32_t maxd = (uint32_t)((ow - 1) * (ow - 1) + (oh - 1) * (oh - 1));
  int bits = 1;
  while (bits < 32 && (maxd >> bits)) bits++;
  size_t tmp = 0;
  RS_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t *)w->sort_keys_in.p, (uint32_t *)w->sort_keys_out.p,
                                           (const uint32_t *)w->sort_vals_in.p, (uint32_t *)w->offsets.p, (int)n, 0, bits, s));
  if (int rc = ws_ensure(w->sort_tmp, tmp)) return rc;
  RS_CHECK(cub::DeviceRadixSort::SortPairs(w->sort_tmp.p, tmp, (const uint32_t *)w->sort_keys_in.p,
                                           (uint32_t *)w->sort_keys_out.p, (const uint32_t *)w->sort_vals_in.p,
                                           (uint32_t *)w->offsets.p, (int)n, 0, bits, s));
  w->off_w = ow; w->off_h = oh; w->off_n = n;
  return 0;
}

// ---- corpus points on the device (replaces prepareCorpusPoints, lib/engine.c:400-431) ----
// Row-major list of corpus pixels that are fully selected (mask 0xFF) and not totally transparent, as packed
// x | y << 16: flag kernel + stable stream compaction + pack.  The count stays on the device (RsCtrl.n_corpus).
__global__ void k_corpus_flags(const uint8_t *__restrict__ raw, uint32_t n_px, int bpp, int alpha_bip, int alpha_source,
                               uint8_t *__restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const uint8_t *p = raw + (size_t)i * bpp;
  flags[i] = (p[0] == 0xFFu && (!alpha_source || p[alpha_bip] != 0)) ? 1 : 0;
}
// The same list without the list: a bitmap of the usable corpus pixels (row-major) and, for every 16th usable pixel, its
// linear index p together with the bitmap window of pixels p .. p + 31.  rs_corpus_point() finds point idx in the window
// of sample idx / 16 -- ONE 8-byte load from a table that stays in L2 wherever at least half of the pixels are usable,
// a short bitmap scan elsewhere -- instead of one 4-byte load from a table of tens of megabytes that never stays
// (4096^2 inpaint: 46 MB of points, one DRAM sector per probe: 45 GB of DRAM reads per pass where 13 GB remain).
__global__ void k_corpus_bits(const uint8_t *__restrict__ flags, uint32_t n_px, uint32_t *__restrict__ bits,
                              uint32_t *__restrict__ counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // launched over n_px rounded up to whole words
  const uint32_t wd = __ballot_sync(RS_FULL, i < n_px && flags[i] != 0);
  if ((threadIdx.x & 31u) == 0u) { bits[i >> 5] = wd; counts[i >> 5] = __popc(wd); }
}
__global__ void k_corpus_samples(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ before, uint32_t n_words,
                                 uint2 *__restrict__ samples) {
  const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= n_words) return;
  const uint32_t wd = bits[wi], b = before[wi], c = __popc(wd);
  // the multiples of 16 among this word's points [b, b + c): at most two
  for (uint32_t first = (b + 15u) & ~15u; c && first < b + c; first += 16u) {
    const uint32_t s = rs_nth_set_bit(wd, first - b), p = (wi << 5) + s;
    const uint32_t win = (wd >> s) | (s ? (bits[wi + 1] << (32u - s)) : 0u);  // usable-pixel bits of pixels p .. p + 31
    samples[first >> 4] = make_uint2(p, win);
  }
}
__global__ void k_pack_points(uint32_t *__restrict__ pts, const unsigned int *__restrict__ n_ptr, int w) {
  const uint32_t n = *n_ptr;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t idx = pts[i];
    pts[i] = (idx % (uint32_t)w) | ((idx / (uint32_t)w) << 16);
  }
}

// ---- target points and their digest on the device ----
// The selection (mask byte != 0) of the target image, 32 pixels per word: count, row range and a 128-bit digest
// (sum over words of two different mixes of (word index, word), so the order of accumulation does not matter).
// The digest keys the visit-order cache: the order is a pure function of (selection, image size, mode, seed).
__device__ __forceinline__ unsigned long long rs_mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x;
}
__global__ void __launch_bounds__(256) k_target_digest(const uint8_t *__restrict__ raw, uint32_t n_px, int bpp, int tw,
                                                       RsCtrl *__restrict__ ctrl) {
  unsigned long long h1 = 0, h2 = 0;
  uint32_t cnt = 0, ymin = 0xFFFFFFFFu, ymax = 0, xmin = 0xFFFFFFFFu, xmax = 0;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_round = (n_px + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool sel = i < n_px && raw[(size_t)i * bpp] != 0;
    const uint32_t w = __ballot_sync(RS_FULL, sel);
    if (lane == 0 && w) {
      const uint32_t wi = i >> 5;
      h1 += rs_mix64(((unsigned long long)wi << 32) | w);
      h2 += rs_mix64((((unsigned long long)w << 32) | wi) ^ 0x9E3779B97F4A7C15ull);
      cnt += __popc(w);
      const uint32_t first = i + (__ffs(w) - 1), last = i + (31 - __clz(w));
      const uint32_t y0 = first / (uint32_t)tw, y1 = last / (uint32_t)tw;
      ymin = min(ymin, y0);
      ymax = max(ymax, y1);
      // columns: exact where the 32 pixels of the word lie in one row, the whole width where they straddle rows
      xmin = min(xmin, y0 == y1 ? first - y0 * (uint32_t)tw : 0u);
      xmax = max(xmax, y0 == y1 ? last - y1 * (uint32_t)tw : (uint32_t)tw - 1u);
    }
  }
  if (lane == 0 && cnt) {
    atomicAdd(&ctrl->dg_h1, h1);
    atomicAdd(&ctrl->dg_h2, h2);
    atomicAdd(&ctrl->dg_n, cnt);
    atomicMin(&ctrl->dg_ymin, ymin);
    atomicMax(&ctrl->dg_ymax, ymax);
    atomicMin(&ctrl->dg_xmin, xmin);
    atomicMax(&ctrl->dg_xmax, xmax);
  }
}
static std::mutex g_order_mutex;
static std::vector<std::shared_ptr<OrderEntry>> g_orders;
static unsigned long long g_order_clock = 0;
static std::atomic<int> g_order_cache_on{1};
extern "C" void rs_cuda_order_cache(int enabled) {
  g_order_cache_on.store(enabled ? 1 : 0);
  if (!enabled) {  // entries are released (cudaFree: a device-wide sync) after the mutex is dropped
    std::vector<std::shared_ptr<OrderEntry>> dropped;
    { std::lock_guard<std::mutex> lk(g_order_mutex); dropped.swap(g_orders); }
  }
}
// A new cache entry for n points on `device`, or nullptr when the device has no room for it even after the cached
// orders of that device were dropped: the job then keeps its order in the workspace, uncached.
static std::shared_ptr<OrderEntry> order_entry_alloc(const RsOrderKey &key, int device, uint32_t n) {
  auto entry = std::make_shared<OrderEntry>();
  entry->key = key; entry->device = device; entry->n = n;
  const size_t bytes = (size_t)n * 4;
  if (cudaMalloc(&entry->dev, bytes) != cudaSuccess) {
    cudaGetLastError();
    std::vector<std::shared_ptr<OrderEntry>> dropped;
    {
      std::lock_guard<std::mutex> lk(g_order_mutex);
      for (size_t i = 0; i < g_orders.size();)
        if (g_orders[i]->device == device) { dropped.push_back(g_orders[i]); g_orders.erase(g_orders.begin() + i); } else i++;
    }
    dropped.clear();
    entry->dev = nullptr;
    if (cudaMalloc(&entry->dev, bytes) != cudaSuccess) { cudaGetLastError(); entry->dev = nullptr; return nullptr; }
  }
  if (cudaEventCreateWithFlags(&entry->ready, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return entry;
}
// Makes the entry visible to other jobs (its `ready` event has been recorded) and evicts the least recently used
// ones: at most 16 entries / 1 GiB per device.  The evicted entries are freed after the mutex is dropped.
static void order_cache_publish(const std::shared_ptr<OrderEntry> &entry) {
  std::vector<std::shared_ptr<OrderEntry>> evicted;
  {
    std::lock_guard<std::mutex> lk(g_order_mutex);
    entry->stamp = ++g_order_clock;
    g_orders.push_back(entry);
    while (true) {
      size_t count = 0, total = 0, lru = g_orders.size();
      for (size_t i = 0; i < g_orders.size(); i++) {
        if (g_orders[i]->device != entry->device) continue;
        count++; total += (size_t)g_orders[i]->n * 4;
        if (g_orders[i] != entry && (lru == g_orders.size() || g_orders[i]->stamp < g_orders[lru]->stamp)) lru = i;
      }
      if (lru == g_orders.size() || !(count > 16 || total > ((size_t)1 << 30))) break;
      evicted.push_back(g_orders[lru]);
      g_orders.erase(g_orders.begin() + lru);
    }
  }
}
static bool key_equal(const RsOrderKey &a, const RsOrderKey &b) {
  return a.h1 == b.h1 && a.h2 == b.h2 && a.n == b.n && a.tw == b.tw && a.th == b.th && a.mode == b.mode && a.seed == b.seed;
}

// ---- simple API (imageSynth): the caller's image + mask planes go up as they are; the internal pixmaps
// [mask][channels] of target and corpus (lib/imageSynth.c:61-137, lib/adaptSimple.h) are built here, and only the
// image rows that hold target points come back, in the caller's layout.
__global__ void k_build_simple(const uint8_t *__restrict__ img, const uint8_t *__restrict__ mask,
                               const uint8_t *__restrict__ mask2, uint32_t n_px, int nc, uint8_t *__restrict__ raw_t,
                               uint8_t *__restrict__ raw_c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const int bpp = nc + 1;
  const uint8_t m = mask[i];
  uint8_t *t = raw_t + (size_t)i * bpp, *c = raw_c + (size_t)i * bpp;
  t[0] = m;
  c[0] = mask2 ? mask2[i] : (uint8_t)~m;  // the corpus is what is NOT selected, or an explicit second mask (imageSynth2)
  for (int k = 0; k < nc; k++) { const uint8_t v = img[(size_t)i * nc + k]; t[1 + k] = v; c[1 + k] = v; }
}
__global__ void k_extract_simple(const uint8_t *__restrict__ raw_t, uint32_t first_px, uint32_t n_px, int nc,
                                 uint8_t *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const uint8_t *t = raw_t + (size_t)(first_px + i) * (nc + 1);
  for (int k = 0; k < nc; k++) out[(size_t)i * nc + k] = t[1 + k];
}

// Host copy into pinned staging followed by the H2D copy, in pieces: a piece goes to the device while the next one is
// being copied, and large images are copied by several cores (a single core moves ~9 GB/s, PCIe 5 takes 25+).
// A caller's buffer that is page-locked (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor): the copy engine
// reads and writes it directly, no staging copy through the workspace's pinned memory.
extern "C" int rs_cuda_host_is_pinned(const void *p) {
  if (p == nullptr || getenv("RS_NO_DIRECT_COPY")) return 0;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}
static int stage_to_device(void *dev, uint8_t *pin, const uint8_t *src, size_t bytes, cudaStream_t s) {
  if (rs_cuda_host_is_pinned(src)) {
    RS_CHECK(cudaMemcpyAsync(dev, src, bytes, cudaMemcpyHostToDevice, s));
    return 0;
  }
  const size_t PIECE = (size_t)4 << 20;
  if (bytes <= PIECE) {
    memcpy(pin, src, bytes);
    RS_CHECK(cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, s));
    return 0;
  }
  unsigned hw = rs_host_cores();
  const size_t nt = std::min<size_t>(rs_copy_threads_max(), hw > 2 ? hw - 1 : 1);
  for (size_t off = 0; off < bytes; off += PIECE * nt) {
    const size_t len = std::min(bytes - off, PIECE * nt);
    if (nt > 1 && len > PIECE) {
      std::vector<std::thread> th;
      const size_t per = (len + nt - 1) / nt;
      for (size_t t = 1; t < nt; t++) {
        const size_t b = std::min(len, t * per), e = std::min(len, (t + 1) * per);
        if (e > b) th.emplace_back([=]() { memcpy(pin + off + b, src + off + b, e - b); });
      }
      memcpy(pin + off, src + off, std::min(len, per));
      for (auto &x : th) x.join();
    } else {
      memcpy(pin + off, src + off, len);
    }
    RS_CHECK(cudaMemcpyAsync((uint8_t *)dev + off, pin + off, len, cudaMemcpyHostToDevice, s));
  }
  return 0;
}

// The same for `rows` rows of row_len bytes that sit src_stride apart in the caller's buffer (ImageBuffer.rowBytes).
static int stage_rows_to_device(void *dev, uint8_t *pin, const uint8_t *src, size_t rows, size_t row_len, size_t src_stride,
                                cudaStream_t s) {
  if (src_stride == row_len) return stage_to_device(dev, pin, src, rows * row_len, s);
  if (rs_cuda_host_is_pinned(src)) {
    RS_CHECK(cudaMemcpy2DAsync(dev, row_len, src, src_stride, row_len, rows, cudaMemcpyHostToDevice, s));
    return 0;
  }
  unsigned hw = rs_host_cores();
  const size_t nt = rows * row_len < ((size_t)4 << 20) ? 1 : std::min<size_t>(rs_copy_threads_max(), hw > 2 ? hw - 1 : 1);
  auto band = [=](size_t t) {
    const size_t per = (rows + nt - 1) / nt, b = std::min(rows, t * per), e = std::min(rows, (t + 1) * per);
    for (size_t y = b; y < e; y++) memcpy(pin + y * row_len, src + y * src_stride, row_len);
  };
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; t++) th.emplace_back(band, t);
  band(0);
  for (auto &x : th) x.join();
  RS_CHECK(cudaMemcpyAsync(dev, pin, rows * row_len, cudaMemcpyHostToDevice, s));
  return 0;
}

struct SimpleSource {  // imageSynth's inputs, as the caller holds them
  const uint8_t *img, *mask, *mask2;
  size_t img_rb, mask_rb, mask2_rb;
  int nc;
};

// Stage 1 of an upload: everything that does not depend on the target points.  Asynchronous on the job's stream.
static int stage_images(RsJob *j, const uint8_t *target_raw, const uint8_t *corpus_raw, const uint32_t *corpus_points,
                        uint32_t n_corpus, const uint32_t *offsets, uint32_t n_offsets, const uint32_t *color_lut256,
                        const uint32_t *map_lut256, uint32_t map_lut_max, bool digest, const SimpleSource *simple = nullptr) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  const RsJobDesc &d = j->d;
  j->simple = simple != nullptr;
  if (corpus_points && n_corpus == 0) { g_err = "rs_job upload: empty corpus point list"; return 100; }
  const size_t tn = (size_t)d.tw * d.th, cn = (size_t)d.cw * d.ch;
  cudaStream_t s = w->stream;
  j->nC = corpus_points ? n_corpus : 0;
  j->corpus_bits = false;
  j->penalty = 65535u * (uint32_t)d.n_color + map_lut_max * (uint32_t)d.n_map;
  const size_t cap_cpts = corpus_points ? (size_t)n_corpus : cn;
  int rc = 0;
  if ((rc = ws_ensure(w->raw_t, tn * d.bpp)) || (rc = ws_ensure(w->raw_c, cn * d.bpp)) ||
      (rc = ws_ensure(w->corpus, (cn + 1) * (j->maps ? 8 : 4))) || (rc = ws_ensure(w->W, tn * 16)) ||
      (rc = ws_ensure(w->meta, tn * 4)) || (rc = ws_ensure(w->tmaps, j->maps ? tn * 4 : 4)) ||
      (rc = ws_ensure(w->cpts, cap_cpts * 4)) ||
      (rc = ws_ensure(w->lut256, 512 * 4)) || (rc = ws_ensure(w->lut_rep, 2 * RS_LUT_WORDS * 4)) ||
      (rc = ws_ensure(w->prober, cn * 32)) ||
      (rc = ws_ensure(w->ctrl, sizeof(RsCtrl))))
    return rc;
  // neighbour offsets: caller-provided table, or built (and kept) on the device
  const int ow = d.tw < d.cw ? d.tw : d.cw, oh = d.th < d.ch ? d.th : d.ch;
  const uint32_t full_n = (uint32_t)(2 * ow - 1) * (uint32_t)(2 * oh - 1);
  // stage all host inputs through pinned memory so the copies are truly asynchronous
  const size_t sz_t = tn * d.bpp, sz_c = cn * d.bpp, sz_cp = corpus_points ? (size_t)n_corpus * 4 : 0,
               sz_off = offsets ? (size_t)n_offsets * 4 : 0, sz_lut = 512 * 4;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_t = 0, o_c = o_t + up(sz_t), o_cp = o_c + up(sz_c), o_off = o_cp + up(sz_cp),
               o_lut = o_off + up(sz_off), total = o_lut + up(sz_lut);
  const size_t need_pin = total > j->out_bytes ? total : j->out_bytes;
  if ((rc = ws_ensure_pinned(w, need_pin))) return rc;
  uint8_t *pin = (uint8_t *)w->pin;
  RS_CHECK(cudaMemsetAsync(w->ctrl.p, 0, sizeof(RsCtrl), s));
  // the recentProber words depend on no input: cleared on the side stream while the images are on their way (32 bytes per
  // corpus pixel: half a gigabyte for a 4096x4096 image), joined below
  RS_CHECK(cudaMemsetAsync(w->prober.p, 0, cn * 32, w->stream2));
  RS_CHECK(cudaEventRecord(w->evAux, w->stream2));
  const int T = 256;
  j->upload_launches = 0;
  auto enqueue_digest = [&](const uint8_t *mask_bytes, int stride) -> int {
    // count, row range and digest of the selection: queued as early as the mask is on its way, so that the result is
    // back (rs_job_digest) while the host is still copying the rest of the job into the staging buffer
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    const uint32_t ymin_init = 0xFFFFFFFFu;
    RS_CHECK(cudaMemcpyAsync(&((RsCtrl *)w->ctrl.p)->dg_ymin, &ymin_init, 4, cudaMemcpyHostToDevice, s));
    RS_CHECK(cudaMemcpyAsync(&((RsCtrl *)w->ctrl.p)->dg_xmin, &ymin_init, 4, cudaMemcpyHostToDevice, s));
    k_target_digest<<<sms * 8, 256, 0, s>>>(mask_bytes, (uint32_t)tn, stride, d.tw, (RsCtrl *)w->ctrl.p);
    RS_CHECK(cudaMemcpyAsync(w->h_digest, &((RsCtrl *)w->ctrl.p)->dg_h1, sizeof(RsTargetDigest), cudaMemcpyDeviceToHost, s));
    RS_CHECK(cudaEventRecord(w->evDigest, s));
    j->upload_launches += 1u;
    return 0;
  };
  if (simple) {  // image + mask planes up (they fit the staging regions of the two pixmaps), pixmaps built on the device
    const size_t sz_img = tn * simple->nc;
    if ((rc = ws_ensure(w->simg, sz_img)) || (rc = ws_ensure(w->smask, tn)) || (rc = ws_ensure(w->smask2, simple->mask2 ? tn : 4)))
      return rc;
    if ((rc = stage_rows_to_device(w->smask.p, pin + o_t + sz_img, simple->mask, d.th, d.tw, simple->mask_rb, s))) return rc;
    RS_CHECK(cudaEventRecord(w->evSel, s));  // what the order pipeline reads (shuffle_order_impl, on the side stream) is up
    if (digest && (rc = enqueue_digest((const uint8_t *)w->smask.p, 1))) return rc;  // the selection IS the mask plane
    if ((rc = stage_rows_to_device(w->simg.p, pin + o_t, simple->img, d.th, (size_t)d.tw * simple->nc, simple->img_rb, s))) return rc;
    if (simple->mask2 && (rc = stage_rows_to_device(w->smask2.p, pin + o_c, simple->mask2, d.th, d.tw, simple->mask2_rb, s))) return rc;
    k_build_simple<<<(unsigned)((tn + T - 1) / T), T, 0, s>>>((const uint8_t *)w->simg.p, (const uint8_t *)w->smask.p,
                                                            simple->mask2 ? (const uint8_t *)w->smask2.p : nullptr, (uint32_t)tn,
                                                            simple->nc, (uint8_t *)w->raw_t.p, (uint8_t *)w->raw_c.p);
    j->upload_launches += 1u;
  } else {
    if ((rc = stage_to_device(w->raw_t.p, pin + o_t, target_raw, sz_t, s))) return rc;
    RS_CHECK(cudaEventRecord(w->evSel, s));
    if (digest && (rc = enqueue_digest((const uint8_t *)w->raw_t.p, d.bpp))) return rc;
  }
  memcpy(pin + o_lut, color_lut256, 256 * 4);
  memcpy(pin + o_lut + 256 * 4, map_lut256, 256 * 4);
  RS_CHECK(cudaMemcpyAsync(w->lut256.p, pin + o_lut, sz_lut, cudaMemcpyHostToDevice, s));
  bool built_offsets = false;
  int offset_sort_bits = 1;
  unsigned int *d_ncorpus = &((RsCtrl *)w->ctrl.p)->n_corpus;
  j->cb = CorpusBufs{w->corpus.p, (uint32_t *)w->cpts.p, nullptr, nullptr};
  j->shared.reset();
  bool corpus_ready = false;  // canonical pixels + points already on the device (a shared entry)
  if (corpus_points) {
    memcpy(pin + o_cp, corpus_points, sz_cp);
    RS_CHECK(cudaMemcpyAsync(w->cpts.p, pin + o_cp, sz_cp, cudaMemcpyHostToDevice, s));
    RS_CHECK(cudaMemcpyAsync(d_ncorpus, &j->nC, 4, cudaMemcpyHostToDevice, s));
    if (!simple && (rc = stage_to_device(w->raw_c.p, pin + o_c, corpus_raw, sz_c, s))) return rc;
  } else {
    const uint32_t n_words = (uint32_t)((cn + 31) / 32);
    std::unique_lock<std::mutex> share_lock;
    std::shared_ptr<SharedCorpus> mine, peer;
    if (j->share_batch && !simple) {  // jobs of a batch that name the same corpus pixmap: built once per device
      share_lock = std::unique_lock<std::mutex>(g_corpus_mutex);
      for (auto &e : g_corpora) {
        if (e->batch != j->share_batch || e->host_key != (const void *)corpus_raw || e->cw != d.cw || e->ch != d.ch || e->bpp != d.bpp ||
            e->n_color != d.n_color || e->n_map != d.n_map || e->map_bip != d.map_bip || e->alpha_bip != d.alpha_bip ||
            e->alpha_source != d.alpha_source)
          continue;
        if (e->device == w->device) { mine = e; break; }
        if (!peer) peer = e;
      }
      if (mine) {  // hit: nothing of the corpus is staged or built again
        RS_CHECK(cudaStreamWaitEvent(s, mine->ready, 0));
        RS_CHECK(cudaMemcpyAsync(d_ncorpus, mine->n_corpus.p, 4, cudaMemcpyDeviceToDevice, s));
        g_corpus_hits.fetch_add(1);
        corpus_ready = true;
      } else {
        mine = std::make_shared<SharedCorpus>();
        mine->batch = j->share_batch; mine->host_key = corpus_raw; mine->device = w->device;
        mine->cw = d.cw; mine->ch = d.ch; mine->bpp = d.bpp; mine->n_color = d.n_color; mine->n_map = d.n_map;
        mine->map_bip = d.map_bip; mine->alpha_bip = d.alpha_bip; mine->alpha_source = d.alpha_source;
        mine->corpus_bytes = (cn + 1) * (j->maps ? 8 : 4); mine->cpts_bytes = cn * 4;
        mine->cbits_bytes = (size_t)(n_words + 2) * 4; mine->csamples_bytes = (size_t)(2 * n_words + 2) * 8;
        if ((rc = ws_ensure(mine->corpus, mine->corpus_bytes)) || (rc = ws_ensure(mine->cpts, mine->cpts_bytes)) ||
            (rc = ws_ensure(mine->cbits, mine->cbits_bytes)) || (rc = ws_ensure(mine->csamples, mine->csamples_bytes)) ||
            (rc = ws_ensure(mine->n_corpus, 4)))
          return rc;
        RS_CHECK(cudaEventCreateWithFlags(&mine->ready, cudaEventDisableTiming));
        if (peer) {  // another device of the batch has it: device-to-device over NVLink, no host staging
          RS_CHECK(cudaStreamWaitEvent(s, peer->ready, 0));
          RS_CHECK(cudaMemcpyPeerAsync(mine->corpus.p, w->device, peer->corpus.p, peer->device, mine->corpus_bytes, s));
          RS_CHECK(cudaMemcpyPeerAsync(mine->cpts.p, w->device, peer->cpts.p, peer->device, mine->cpts_bytes, s));
          RS_CHECK(cudaMemcpyPeerAsync(mine->cbits.p, w->device, peer->cbits.p, peer->device, mine->cbits_bytes, s));
          RS_CHECK(cudaMemcpyPeerAsync(mine->csamples.p, w->device, peer->csamples.p, peer->device, mine->csamples_bytes, s));
          RS_CHECK(cudaMemcpyPeerAsync(mine->n_corpus.p, w->device, peer->n_corpus.p, peer->device, 4, s));
          RS_CHECK(cudaMemcpyAsync(d_ncorpus, mine->n_corpus.p, 4, cudaMemcpyDeviceToDevice, s));
          mine->peer_copied = true;
          g_corpus_peer_copies.fetch_add(1);
          corpus_ready = true;
        }
      }
      j->shared = mine;
      j->cb = CorpusBufs{mine->corpus.p, (uint32_t *)mine->cpts.p, (uint32_t *)mine->cbits.p, (uint2 *)mine->csamples.p};
    } else {
      if ((rc = ws_ensure(w->cbits, (size_t)(n_words + 2) * 4)) || (rc = ws_ensure(w->csamples, (size_t)(2 * n_words + 2) * 8))) return rc;
      j->cb.cbits = (uint32_t *)w->cbits.p; j->cb.csamples = (uint2 *)w->csamples.p;
    }
    if (!corpus_ready) {
      if (!simple && (rc = stage_to_device(w->raw_c.p, pin + o_c, corpus_raw, sz_c, s))) return rc;
      if ((rc = ws_ensure(w->sort_keys_in, cn))) return rc;  // flags
      k_corpus_flags<<<(unsigned)((cn + T - 1) / T), T, 0, s>>>((const uint8_t *)w->raw_c.p, (uint32_t)cn, d.bpp, d.alpha_bip,
                                                              d.alpha_source, (uint8_t *)w->sort_keys_in.p);
      thrust::counting_iterator<uint32_t> idx(0);
      size_t tmp = 0, tmp_scan = 0;
      RS_CHECK(cub::DeviceSelect::Flagged(nullptr, tmp, idx, (const uint8_t *)w->sort_keys_in.p, j->cb.cpts, d_ncorpus, (int)cn, s));
      RS_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (const uint32_t *)w->ccounts.p, (uint32_t *)w->cbefore.p, (int)n_words, s));
      if ((rc = ws_ensure(w->sort_tmp, tmp_scan > tmp ? tmp_scan : tmp)) || (rc = ws_ensure(w->ccounts, (size_t)n_words * 4)) ||
          (rc = ws_ensure(w->cbefore, (size_t)n_words * 4)))
        return rc;
      RS_CHECK(cub::DeviceSelect::Flagged(w->sort_tmp.p, tmp, idx, (const uint8_t *)w->sort_keys_in.p, j->cb.cpts, d_ncorpus, (int)cn, s));
      k_pack_points<<<592, T, 0, s>>>(j->cb.cpts, d_ncorpus, d.cw);
      // bitmap + every-32nd-point samples of the same selection (rs_corpus_point picks by density at run time)
      k_corpus_bits<<<(n_words * 32u + T - 1) / T, T, 0, s>>>((const uint8_t *)w->sort_keys_in.p, (uint32_t)cn, j->cb.cbits,
                                                            (uint32_t *)w->ccounts.p);
      RS_CHECK(cudaMemsetAsync(j->cb.cbits + n_words, 0, 8, s));  // the scan may peek one word past the end
      RS_CHECK(cub::DeviceScan::ExclusiveSum(w->sort_tmp.p, tmp_scan, (const uint32_t *)w->ccounts.p, (uint32_t *)w->cbefore.p, (int)n_words, s));
      k_corpus_samples<<<(n_words + T - 1) / T, T, 0, s>>>(j->cb.cbits, (const uint32_t *)w->cbefore.p, n_words, j->cb.csamples);
      if (mine) RS_CHECK(cudaMemcpyAsync(mine->n_corpus.p, d_ncorpus, 4, cudaMemcpyDeviceToDevice, s));
    }
    j->corpus_bits = true;
    if (mine && !corpus_ready) {  // canonical pixels belong to the entry too; then it becomes visible to the batch
      k_canon_corpus<<<(unsigned)((cn + T) / T), T, 0, s>>>((const uint8_t *)w->raw_c.p, (int)cn, d.bpp, d.n_color, d.n_map, d.map_bip,
                                                              j->maps ? nullptr : (uint32_t *)j->cb.corpus,
                                                              j->maps ? (uint2 *)j->cb.corpus : nullptr);
      g_corpus_builds.fetch_add(1);
      corpus_ready = true;
    }
    if (mine && !g_corpora.empty() && std::find(g_corpora.begin(), g_corpora.end(), mine) != g_corpora.end()) {
      // (an entry found in the list: nothing to publish)
    } else if (mine) {
      RS_CHECK(cudaEventRecord(mine->ready, s));
      g_corpora.push_back(mine);
    }
  }
  if (offsets) {
    if ((rc = ws_ensure(w->offsets, sz_off))) return rc;
    memcpy(pin + o_off, offsets, sz_off);
    RS_CHECK(cudaMemcpyAsync(w->offsets.p, pin + o_off, sz_off, cudaMemcpyHostToDevice, s));
    w->off_w = w->off_h = 0; w->off_n = 0;  // not a cached full table
    j->nOff = n_offsets;
    j->off_w = j->off_h = 0;
  } else {
    if (!(w->off_w == ow && w->off_h == oh && w->off_n == full_n)) {
      if ((rc = build_offsets_on_device(w, ow, oh, full_n))) return rc;
      built_offsets = true;
      const uint32_t maxd = (uint32_t)((ow - 1) * (ow - 1) + (oh - 1) * (oh - 1));
      while (offset_sort_bits < 32 && (maxd >> offset_sort_bits)) offset_sort_bits++;
    }
    j->nOff = full_n;
    j->off_w = ow; j->off_h = oh;
  }
  if ((rc = ws_ensure(w->ctx_blocks, (size_t)((d.tw + 31) / 32) * (size_t)((d.th + 31) / 32) * 4))) return rc;
  RS_CHECK(cudaStreamWaitEvent(s, w->evAux, 0));
  if (!corpus_ready)
    k_canon_corpus<<<(unsigned)((cn + T) / T), T, 0, s>>>((const uint8_t *)w->raw_c.p, (int)cn, d.bpp, d.n_color, d.n_map,
                                                            d.map_bip, j->maps ? nullptr : (uint32_t *)j->cb.corpus,
                                                            j->maps ? (uint2 *)j->cb.corpus : nullptr);
  k_init_target<<<(unsigned)((tn + T - 1) / T), T, 0, s>>>((const uint8_t *)w->raw_t.p, (int)tn, d.bpp, d.n_color, d.n_map,
                                                         d.map_bip, d.alpha_bip, d.alpha_target, d.use_context,
                                                         (unsigned long long *)w->W.p, (uint32_t *)w->meta.p,
                                                         j->maps ? (uint32_t *)w->tmaps.p : nullptr);
  k_replicate_lut<<<(RS_LUT_WORDS + T - 1) / T, T, 0, s>>>((const uint32_t *)w->lut256.p, (const uint32_t *)w->lut256.p + 256,
                                                         (uint32_t *)w->lut_rep.p);
  RS_CHECK(cudaGetLastError());
  j->upload_launches += 3u + (corpus_points ? 0u : 7u) + (built_offsets ? 1u + (uint32_t)((offset_sort_bits + 7) / 8) + 2u : 0u);
  for (int p = 0; p < 6; p++) w->h_ticks[p] = 0;
  *w->h_cancel = 0;
  return 0;
}

// Everything whose size follows the number of target points.  `idle`: nothing of this job is in flight on the stream
// (pinned staging may be reallocated).
static int set_targets(RsJob *j, uint32_t n_targets, uint32_t y_min, uint32_t y_max, bool idle) {
  Workspace *w = j->ws;
  const RsJobDesc &d = j->d;
  if (n_targets == 0 || n_targets >= RS_IDX_MASK || y_max < y_min || y_max >= (uint32_t)d.th) {
    g_err = "rs_job upload: empty or oversized target point list";
    return 100;
  }
  j->nT = n_targets; j->y_min = y_min; j->y_max = y_max;
  j->x_min = 0; j->x_max = (uint32_t)d.tw - 1u;
  uint32_t kmax = d.patch_size < 2 ? 2 : d.patch_size;
  if (kmax > RS_MAX_NB) kmax = RS_MAX_NB;
  int rc = 0;
  if ((rc = ws_ensure(w->nb_lists, (size_t)n_targets * (kmax - 1) * sizeof(uint2))) || (rc = ws_ensure(w->nb_counts, n_targets)) ||
      ((n_targets >= (1u << 21) || getenv("RS_LATER_LISTS_MIN")) &&
       ((rc = ws_ensure(w->nb_later, (size_t)n_targets * (kmax - 1) * sizeof(uint2))) || (rc = ws_ensure(w->nb_later_counts, n_targets)))) ||
      (rc = ws_ensure(w->sources, (size_t)n_targets * 4)))
    return rc;
  j->out_bytes = (size_t)(y_max - y_min + 1) * d.tw * d.bpp + (size_t)n_targets * 4 + 512;
  if (idle && (rc = ws_ensure_pinned(w, j->out_bytes))) return rc;
  return 0;
}

// Phase 1 of the upload with the target points known to the caller (count and row range): everything that does not
// depend on the visit ORDER.  Asynchronous on the job's stream, so the host can order the points meanwhile.
extern "C" int rs_job_upload_images(RsJob *j, const uint8_t *target_raw, const uint8_t *corpus_raw, uint32_t n_targets,
                                    uint32_t y_min, uint32_t y_max, const uint32_t *corpus_points, uint32_t n_corpus,
                                    const uint32_t *offsets, uint32_t n_offsets, const uint32_t *color_lut256,
                                    const uint32_t *map_lut256, uint32_t map_lut_max) {
  if (int rc = set_targets(j, n_targets, y_min, y_max, true)) return rc;
  return stage_images(j, target_raw, corpus_raw, corpus_points, n_corpus, offsets, n_offsets, color_lut256, map_lut256,
                      map_lut_max, false);
}

// Phase 2 of the upload: the visit order (n_targets points as given to rs_job_upload_images).
static int upload_order_impl(RsJob *j, const uint32_t *targets, const RsOrderKey *key) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  const size_t bytes = (size_t)j->nT * 4;
  if (bytes > w->pin_order_cap) {
    if (w->pin_order) cudaFreeHost(w->pin_order);
    w->pin_order = nullptr; w->pin_order_cap = 0;
    RS_CHECK(cudaHostAlloc(&w->pin_order, bytes + bytes / 8 + 4096, cudaHostAllocDefault));
    w->pin_order_cap = bytes + bytes / 8 + 4096;
  }
  uint32_t *dst = nullptr;
  j->order.reset();
  std::shared_ptr<OrderEntry> entry;
  if (key && g_order_cache_on.load()) entry = order_entry_alloc(*key, w->device, j->nT);  // the uploaded order becomes a cache entry
  if (entry) {
    dst = entry->dev;
    j->order = entry;
  } else {
    if (int rc = ws_ensure(w->targets, bytes)) return rc;
    dst = (uint32_t *)w->targets.p;
  }
  j->targets_dev = dst;
  memcpy(w->pin_order, targets, bytes);
  RS_CHECK(cudaMemcpyAsync(dst, w->pin_order, bytes, cudaMemcpyHostToDevice, w->stream));
  if (entry) {
    // visible to other jobs only once the event that follows its upload exists: a job on another stream that hits this
    // entry makes its stream wait for that event before reading the order
    RS_CHECK(cudaEventRecord(entry->ready, w->stream));
    order_cache_publish(entry);
  }
  j->upload_launches += 1u;
  k_scatter_order<<<(j->nT + 255) / 256, 256, 0, w->stream>>>(dst, j->nT, j->d.tw, (uint32_t *)w->meta.p);
  RS_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int rs_job_upload_order(RsJob *j, const uint32_t *targets) { return upload_order_impl(j, targets, nullptr); }

// ---- the same upload with the target points found on the device ----
extern "C" int rs_job_stage(RsJob *j, const uint8_t *target_raw, const uint8_t *corpus_raw, const uint32_t *color_lut256,
                            const uint32_t *map_lut256, uint32_t map_lut_max) {
  j->out_bytes = 0;
  return stage_images(j, target_raw, corpus_raw, nullptr, 0, nullptr, 0, color_lut256, map_lut256, map_lut_max, true);
}
extern "C" int rs_job_stage_simple(RsJob *j, const uint8_t *img, size_t img_row_bytes, const uint8_t *mask,
                                   size_t mask_row_bytes, const uint8_t *mask2, size_t mask2_row_bytes,
                                   const uint32_t *color_lut256, const uint32_t *map_lut256, uint32_t map_lut_max) {
  const RsJobDesc &d = j->d;
  if (d.tw != d.cw || d.th != d.ch || d.n_map != 0 || d.bpp < 2 || d.bpp > 5) {
    g_err = "rs_job_stage_simple: the simple API has one image (target = corpus), 1-4 channels and no maps";
    return 100;
  }
  const SimpleSource src{img, mask, mask2, img_row_bytes, mask_row_bytes, mask2_row_bytes, d.bpp - 1};
  j->out_bytes = 0;
  return stage_images(j, nullptr, nullptr, nullptr, 0, nullptr, 0, color_lut256, map_lut256, map_lut_max, true, &src);
}
// The rows of the image that hold target points, back into the caller's buffer (all channels: alpha is unchanged).
extern "C" int rs_job_download_simple(RsJob *j, uint8_t *img, size_t img_row_bytes) {
  const Workspace *w = j->ws;
  if (!j->simple) { g_err = "rs_job_download_simple: the job was not staged by rs_job_stage_simple"; return 100; }
  const size_t row_len = (size_t)j->d.tw * (j->d.bpp - 1), rows = j->y_max - j->y_min + 1;
  const uint8_t *src = (const uint8_t *)w->pin;
  const uint32_t y0 = j->y_min;
  if (j->result_direct) {  // device -> the caller's page-locked image, in place: the box that holds the target points
    RS_CHECK(cudaSetDevice(w->device));
    const size_t nc = (size_t)j->d.bpp - 1, x0 = (size_t)j->x_min * nc, box_len = ((size_t)j->x_max - j->x_min + 1) * nc;
    RS_CHECK(cudaMemcpy2DAsync(img + (size_t)y0 * img_row_bytes + x0, img_row_bytes, (const uint8_t *)w->simg.p + x0, row_len, box_len, rows,
                               cudaMemcpyDeviceToHost, w->stream));
    RS_CHECK(cudaStreamSynchronize(w->stream));
    return 0;
  }
  unsigned hw = rs_host_cores();
  const size_t nt = rows * row_len < ((size_t)4 << 20) ? 1 : std::min<size_t>(rs_copy_threads_max(), hw > 2 ? hw - 1 : 1);
  auto band = [=](size_t t) {
    const size_t per = (rows + nt - 1) / nt, b = std::min(rows, t * per), e = std::min(rows, (t + 1) * per);
    for (size_t r = b; r < e; r++) memcpy(img + (size_t)(y0 + r) * img_row_bytes, src + r * row_len, row_len);
  };
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; t++) th.emplace_back(band, t);
  band(0);
  for (auto &x : th) x.join();
  return 0;
}
extern "C" int rs_job_digest(RsJob *j, RsTargetDigest *out) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  RS_CHECK(cudaEventSynchronize(w->evDigest));
  *out = *w->h_digest;
  return 0;
}
// Binds the cached visit order for `key` to the job.  Returns 1 on a hit, 0 on a miss (the caller orders the points
// and calls rs_job_set_order), 100 on error.  `dg` = what rs_job_digest returned.
extern "C" int rs_job_bind_order(RsJob *j, const RsTargetDigest *dg, const RsOrderKey *key) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  // the staging copies are not all done yet: the pinned buffer must not move now (idle = false); it already holds the
  // whole input, which is at least as large as the rows that come back
  if (int rc = set_targets(j, dg->n, dg->ymin, dg->ymax, false)) return rc;
  if (dg->xmin <= dg->xmax && dg->xmax < (uint32_t)j->d.tw) { j->x_min = dg->xmin; j->x_max = dg->xmax; }
  if (j->out_bytes > w->pin_cap) {  // (cannot happen for same-sized in/out images; be safe)
    RS_CHECK(cudaStreamSynchronize(w->stream));
    if (int rc = ws_ensure_pinned(w, j->out_bytes)) return rc;
  }
  if (!key || !g_order_cache_on.load()) return 0;
  std::shared_ptr<OrderEntry> hit;
  {
    std::lock_guard<std::mutex> lk(g_order_mutex);
    for (auto &e : g_orders)
      if (e->device == w->device && key_equal(e->key, *key)) { e->stamp = ++g_order_clock; hit = e; break; }
  }
  if (!hit) return 0;
  j->order = hit;
  j->targets_dev = hit->dev;
  RS_CHECK(cudaStreamWaitEvent(w->stream, hit->ready, 0));  // the entry may still be on its way up on another stream
  j->upload_launches += 1u;
  k_scatter_order<<<(j->nT + 255) / 256, 256, 0, w->stream>>>(hit->dev, j->nT, j->d.tw, (uint32_t *)w->meta.p);
  RS_CHECK(cudaGetLastError());
  return 1;
}
// Stable ascending radix sort of n (key, value) pairs, host to host, on the job's SIDE stream: it runs beside the staging
// of the images.  Used for the sort step of the target orderings 2-8 (lib/orderTarget.h:154-262); the keys are computed
// on the host because the brushfire ray index needs libm's atan2 bit for bit.
extern "C" int rs_job_sort_pairs(RsJob *j, uint32_t *keys, uint32_t *vals, uint32_t n, int key_bits) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  if (n == 0) return 0;
  const size_t bytes = (size_t)n * 4;
  int rc = 0;
  if ((rc = ws_ensure(w->ord_keys_in, bytes)) || (rc = ws_ensure(w->ord_keys_out, bytes)) ||
      (rc = ws_ensure(w->ord_vals_in, bytes)) || (rc = ws_ensure(w->ord_vals_out, bytes)))
    return rc;
  if (2 * bytes > w->pin_sort_cap) {
    if (w->pin_sort) cudaFreeHost(w->pin_sort);
    w->pin_sort = nullptr; w->pin_sort_cap = 0;
    RS_CHECK(cudaHostAlloc(&w->pin_sort, 2 * bytes + bytes / 4 + 4096, cudaHostAllocDefault));
    w->pin_sort_cap = 2 * bytes + bytes / 4 + 4096;
  }
  uint8_t *pin = (uint8_t *)w->pin_sort;
  cudaStream_t s2 = w->stream2;
  memcpy(pin, keys, bytes);
  memcpy(pin + bytes, vals, bytes);
  RS_CHECK(cudaMemcpyAsync(w->ord_keys_in.p, pin, bytes, cudaMemcpyHostToDevice, s2));
  RS_CHECK(cudaMemcpyAsync(w->ord_vals_in.p, pin + bytes, bytes, cudaMemcpyHostToDevice, s2));
  size_t tmp = 0;
  if (key_bits < 1 || key_bits > 32) key_bits = 32;
  RS_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t *)w->ord_keys_in.p, (uint32_t *)w->ord_keys_out.p,
                                           (const uint32_t *)w->ord_vals_in.p, (uint32_t *)w->ord_vals_out.p, (int)n, 0, key_bits, s2));
  if ((rc = ws_ensure(w->ord_tmp, tmp))) return rc;
  RS_CHECK(cub::DeviceRadixSort::SortPairs(w->ord_tmp.p, tmp, (const uint32_t *)w->ord_keys_in.p, (uint32_t *)w->ord_keys_out.p,
                                           (const uint32_t *)w->ord_vals_in.p, (uint32_t *)w->ord_vals_out.p, (int)n, 0, key_bits, s2));
  RS_CHECK(cudaMemcpyAsync(pin, w->ord_keys_out.p, bytes, cudaMemcpyDeviceToHost, s2));
  RS_CHECK(cudaMemcpyAsync(pin + bytes, w->ord_vals_out.p, bytes, cudaMemcpyDeviceToHost, s2));
  RS_CHECK(cudaStreamSynchronize(s2));
  memcpy(keys, pin, bytes);
  memcpy(vals, pin + bytes, bytes);
  j->upload_launches += 2u + (uint32_t)((key_bits + 7) / 8);
  return 0;
}

// Visit order of the shuffling modes (matchContextType 0, 1) built on the device from the host's draws j_i (the
// reference's PRNG stream, n of them): target points compacted from the staged image, pairs sorted, chains traced.
// With a key the order becomes a cache entry.  ordered_out (optional): the order back on the host.
// Range reduction of the raw PRNG words on the device: g_rand_int_range(0, n) accepts a word v iff v <= maxvalue (the
// rejection rule of GLib's g_rand_int_range) and returns v % n; the accepted words keep their order.
struct RsAccept {
  uint32_t maxvalue;
  __host__ __device__ bool operator()(const uint32_t &v) const { return v <= maxvalue; }
};
__global__ void k_reduce_draws(uint32_t *__restrict__ draws, const unsigned int *__restrict__ n_accepted, uint32_t n, RsCtrl *ctrl) {
  // too few words sent: fail loudly (rs_job_run reports the fault) -- but the kernels queued behind this one still
  // index with the draws, so they are reduced into range regardless
  if (*n_accepted < n && blockIdx.x == 0 && threadIdx.x == 0) atomicExch(&ctrl->fault, 1u);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) draws[i] %= n;
}

// The PRNG stream itself on the device: the first n_words raw 32-bit words of GLib's GRand seeded with `seed` (MT19937 with
// init_genrand seeding -- g_rand_new_with_seed / g_rand_int, what the reference draws its visit order from,
// lib/engine.c:643, lib/orderTarget.h:38-53).
//
// One CTA (rs_mt_rounds).  Word i of the stream is x[i] = x[i-227] ^ f(x[i-624], x[i-623]): every f of a round of 623
// consecutive words reads words older than the round, and the x[i-227] chain is one XOR per word that stays inside a lane
// -- lane t of the 227 MAKER threads makes words t, t + 227 and (t < 169) t + 454 of the round.  The state lives in a
// ring in shared memory, one barrier per round; warps 8-19 temper the words of the round before and write them out, so
// that a makers' round is the dependency chain and nothing else (loads, f, XORs, stores, barrier).  Shapes measured on
// B200 with tools/microbench/mt_bench.cu (4.39 M words, cfg3): this one 1.54 ms; rounds of 454 words 1.71 (makers store) /
// 1.95 (256 writers); fewer, wider lanes 2.4-10 ms -- a round is latency (shared-memory round trip + barrier with its
// store drain), so lanes beat words per lane.
//
// Many CTAs (k_mt19937_raw, grid > 1).  The generator is linear over GF(2), so CTA q can start RS_MT_JUMP * q words into
// the stream without making them: with g_q(z) = z^(q * RS_MT_JUMP) mod the characteristic polynomial (host_prep.cpp:
// mt_jump_poly, set-bit positions in a device table), its state is the XOR over the set bits k of g_q of the windows
// x[k .. k + 624) of the untempered stream -- 19936 words that every CTA makes itself from the seed (32 rounds, 7 us),
// combined by one thread per state word (~10 k shared-memory loads each; the index list staged in shared memory), after
// which it runs its own 2^18 words.  cfg3's 4.39 M words: 17 CTAs on 17 SMs, 0.26 ms beside the image upload; the host's AVX2 producer takes 3-4 ms of a
// core for the same words, and the words another 17 MB of PCIe traffic.
#define RS_MT_WRITERS 384
#define RS_MT_THREADS (256 + RS_MT_WRITERS)
#define RS_MT_JUMP (1u << 18)        // words per CTA of a multi-CTA launch
#define RS_MT_MAX_CTAS 512u          // above that many (134 M words) one CTA makes the whole stream
#define RS_MT_IDX_STRIDE 19968u      // entries per polynomial in the device table (at most 19937 set bits)
#define RS_MT_X_PAD (624u + 19936u)  // index of 624 zero words behind the stream: what the padding of an index list selects
#define RS_MT_X_WORDS (RS_MT_X_PAD + 624u)
#define RS_MT_SMEM_BYTES (RS_MT_X_WORDS * 4u + RS_MT_IDX_STRIDE * 2u)  // the words, then the CTA's index list
__device__ __forceinline__ uint32_t rs_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void rs_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// n_words words from the state in ring[0..623] (= words -624..-1), tempered into out[0..n_words).  TO_X: the untempered
// words go to X[624 + i] instead and nothing is written out.  Every thread of the CTA calls it (one barrier per round,
// the same instruction for makers and writers).
template <bool TO_X>
__device__ __forceinline__ void rs_mt_rounds(const uint32_t sb, const uint32_t t, const uint32_t n_words, uint32_t *__restrict__ out,
                                             uint32_t *__restrict__ X) {
  const uint32_t rounds = (n_words + 622u) / 623u;
  auto f = [](uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  };
  const bool maker = t < 227u, third = t < 169u, writer = t >= 256u;  // (lanes 227..255 only keep the barrier count)
  const uint32_t u = t - 256u;
  constexpr int K = (623 + RS_MT_WRITERS - 1) / RS_MT_WRITERS;
  uint32_t pb = t * 4u;  // makers: byte offset in the ring of x[i0 - 624], i0 = 623 r + t
  for (uint32_t r = 0; r <= rounds; r++) {
    if (maker) {
      if (r < rounds) {
        const uint32_t c = rs_lds32(sb + ((pb + 1588u) & 8188u));  // x[i0 - 227]
        const uint32_t a0 = rs_lds32(sb + pb), b0 = rs_lds32(sb + ((pb + 4u) & 8188u));
        const uint32_t a1 = rs_lds32(sb + ((pb + 908u) & 8188u)), b1 = rs_lds32(sb + ((pb + 912u) & 8188u));
        uint32_t a2 = 0, b2 = 0;
        if (third) { a2 = rs_lds32(sb + ((pb + 1816u) & 8188u)); b2 = rs_lds32(sb + ((pb + 1820u) & 8188u)); }
        const uint32_t v0 = c ^ f(a0, b0), v1 = v0 ^ f(a1, b1), v2 = v1 ^ f(a2, b2);
        rs_sts32(sb + ((pb + 2496u) & 8188u), v0);
        rs_sts32(sb + ((pb + 3404u) & 8188u), v1);
        if (third) rs_sts32(sb + ((pb + 4312u) & 8188u), v2);
        if (TO_X) {
          const uint32_t i0 = 624u + r * 623u + t;
          X[i0] = v0;
          X[i0 + 227u] = v1;
          if (third) X[i0 + 454u] = v2;
        }
        pb = (pb + 2492u) & 8188u;
      }
    } else if (writer && !TO_X && r > 0) {
      // the 623 words of round r - 1 (still in the ring: a position is reused 2048 words later), tempered
      const uint32_t base = (r - 1u) * 623u;
      uint32_t y[K];
#pragma unroll
      for (int k = 0; k < K; k++) { const uint32_t o = u + k * RS_MT_WRITERS; y[k] = o < 623u ? rs_lds32(sb + (((base + o + 624u) & 2047u) << 2)) : 0u; }
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t o = u + k * RS_MT_WRITERS, i = base + o;
        if (o < 623u && i < n_words) {
          uint32_t z = y[k];
          z ^= z >> 11;
          z ^= (z << 7) & 0x9d2c5680u;
          z ^= (z << 15) & 0xefc60000u;
          out[i] = z ^ (z >> 18);
        }
      }
    }
    __syncthreads();
  }
}
// CTA q makes words [q * jump, min((q + 1) * jump, n_words)).  grid 1: jump >= n_words, no table, no dynamic shared memory.
__global__ void __launch_bounds__(RS_MT_THREADS, 1) k_mt19937_raw(uint32_t seed, uint32_t n_words, uint32_t jump, uint32_t *__restrict__ out,
                                                                  const uint16_t *__restrict__ jump_idx, const uint32_t *__restrict__ jump_cnt) {
  extern __shared__ __align__(16) uint32_t mt_x[];  // CTAs q > 0: the untempered words -624 .. 19935 of the stream, then the index list
  __shared__ __align__(16) uint32_t ring[2048];    // word i of the stream lives at ring[(i + 624) & 2047]; words -624..-1 = the state
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t t = threadIdx.x, q = blockIdx.x;
  const unsigned long long first = (unsigned long long)q * jump;
  if (first >= n_words) return;
  const uint32_t count = (uint32_t)min((unsigned long long)jump, (unsigned long long)n_words - first);
  if (t == 0) {
    uint32_t x = seed;
    ring[0] = x;
    for (uint32_t i = 1; i < 624; i++) { x = 1812433253u * (x ^ (x >> 30)) + i; ring[i] = x; }
  }
  __syncthreads();
  if (q > 0) {
    if (t < 624u) { mt_x[t] = ring[t]; mt_x[RS_MT_X_PAD + t] = 0u; }
    // the CTA's index list into shared memory (coalesced, once), eight positions per 16-byte word; the list is padded to
    // a multiple of eight with RS_MT_X_PAD.  (Read from global memory inside the loop, every second iteration waited for
    // a new sector: 0.35 instead of 0.26 ms for the kernel.)
    const uint32_t cnt8 = (__ldg(jump_cnt + (q - 1u)) + 7u) / 8u;
    uint4 *idx_s = reinterpret_cast<uint4 *>(mt_x + RS_MT_X_WORDS);
    {
      const uint4 *__restrict__ idx_g = reinterpret_cast<const uint4 *>(jump_idx + (size_t)(q - 1u) * RS_MT_IDX_STRIDE);
      for (uint32_t e = t; e < cnt8; e += RS_MT_THREADS) idx_s[e] = __ldg(idx_g + e);
    }
    rs_mt_rounds<true>(sb, t, 19936u, nullptr, mt_x);
    __syncthreads();
    // the loop is bound by the shared-memory loads of the windows: 624 x ~10 k words
    uint32_t acc = 0;
    if (t < 624u) {
      const uint32_t *__restrict__ xt = mt_x + t;
#pragma unroll 2
      for (uint32_t e = 0; e < cnt8; e++) {
        const uint4 p = idx_s[e];
        acc ^= xt[p.x & 0xFFFFu] ^ xt[p.x >> 16] ^ xt[p.y & 0xFFFFu] ^ xt[p.y >> 16] ^ xt[p.z & 0xFFFFu] ^ xt[p.z >> 16] ^
               xt[p.w & 0xFFFFu] ^ xt[p.w >> 16];
      }
    }
    __syncthreads();
    if (t < 624u) ring[t] = acc;  // (of word -624 only the top bit is state, and only that bit is right)
    __syncthreads();
  }
  rs_mt_rounds<false>(sb, t, count, out + first, nullptr);
}
// The jump polynomials of a device: built on the host on demand (mt_jump_poly), kept for the life of the process.
struct MtJumpDev {
  uint16_t *idx = nullptr;
  uint32_t *cnt = nullptr;
  uint32_t have = 0, cap = 0;
};
static std::mutex g_mtj_mutex;
static MtJumpDev g_mtj[64];
// Launches the stream kernel on `s` of the current device (the launch happens under the table's mutex: a table that
// grows is freed only after the kernels that read it have run).
static int mt19937_launch(uint32_t seed, uint32_t n_words, uint32_t *out, cudaStream_t s) {
  if (n_words == 0) return 0;
  uint32_t ctas = (n_words + RS_MT_JUMP - 1u) / RS_MT_JUMP;
  if (ctas <= 1u || ctas > RS_MT_MAX_CTAS || getenv("RS_MT_ONE_CTA")) {
    k_mt19937_raw<<<1, RS_MT_THREADS, 0, s>>>(seed, n_words, 0xFFFFFFFFu, out, nullptr, nullptr);
    RS_CHECK(cudaGetLastError());
    return 0;
  }
  int device = 0;
  RS_CHECK(cudaGetDevice(&device));
  if (device < 0 || device >= 64) { g_err = "mt19937_launch: device ordinal out of range"; return 100; }
  std::lock_guard<std::mutex> lk(g_mtj_mutex);
  MtJumpDev &T = g_mtj[device];
  const uint32_t need = ctas - 1u;
  if (need > T.cap) {
    const uint32_t cap = std::max(need, std::max(2u * T.cap, 32u));
    uint16_t *idx = nullptr;
    uint32_t *cnt = nullptr;
    RS_CHECK(cudaMalloc(&idx, (size_t)cap * RS_MT_IDX_STRIDE * sizeof(uint16_t)));
    RS_CHECK(cudaMalloc(&cnt, (size_t)cap * sizeof(uint32_t)));
    if (T.have) {
      RS_CHECK(cudaMemcpy(idx, T.idx, (size_t)T.have * RS_MT_IDX_STRIDE * sizeof(uint16_t), cudaMemcpyDeviceToDevice));
      RS_CHECK(cudaMemcpy(cnt, T.cnt, (size_t)T.have * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
    }
    if (T.idx) cudaFree(T.idx);  // (cudaFree waits for the kernels in flight)
    if (T.cnt) cudaFree(T.cnt);
    T.idx = idx; T.cnt = cnt; T.cap = cap;
  }
  for (uint32_t q = T.have + 1u; q <= need; q++) {
    const std::vector<uint16_t> &v = rs::mt_jump_poly(q, RS_MT_JUMP);
    const uint32_t c = (uint32_t)v.size();
    if (c == 0 || c + 8u > RS_MT_IDX_STRIDE) { g_err = "mt19937_launch: bad jump polynomial"; return 100; }
    std::vector<uint16_t> padded(v);
    while (padded.size() % 8u) padded.push_back((uint16_t)RS_MT_X_PAD);
    RS_CHECK(cudaMemcpy(T.idx + (size_t)(q - 1u) * RS_MT_IDX_STRIDE, padded.data(), padded.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    RS_CHECK(cudaMemcpy(T.cnt + (q - 1u), &c, sizeof(uint32_t), cudaMemcpyHostToDevice));
    T.have = q;
  }
  RS_CHECK(cudaFuncSetAttribute(k_mt19937_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_MT_SMEM_BYTES));
  k_mt19937_raw<<<ctas, RS_MT_THREADS, RS_MT_SMEM_BYTES, s>>>(seed, n_words, RS_MT_JUMP, out, T.idx, T.cnt);
  RS_CHECK(cudaGetLastError());
  return 0;
}
// Test entry: the first n_words words of the stream of `seed`, made on the current device, into host memory.
extern "C" int rs_cuda_mt19937_raw(uint32_t seed, uint32_t n_words, uint32_t *out_host) {
  if (n_words == 0) return 0;
  uint32_t *d = nullptr;
  RS_CHECK(cudaMalloc(&d, (size_t)n_words * 4));
  int rc = mt19937_launch(seed, n_words, d, nullptr);
  cudaError_t e = rc ? cudaSuccess : cudaMemcpy(out_host, d, (size_t)n_words * 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (rc) return rc;
  RS_CHECK(e);
  return 0;
}

static int shuffle_order_impl(RsJob *j, const uint32_t *draws, const uint32_t *raw_pinned, uint32_t n_raw, const RsOrderKey *key,
                              uint32_t *ordered_out, bool raw_on_device = false, uint32_t seed = 0);
extern "C" int rs_job_shuffle_order(RsJob *j, const uint32_t *draws, const RsOrderKey *key, uint32_t *ordered_out) {
  return shuffle_order_impl(j, draws, nullptr, 0, key, ordered_out);
}
// The same with the PRNG stream made on the device too (k_mt19937_raw): nothing of the order comes from the host.
extern "C" int rs_job_shuffle_order_seed(RsJob *j, uint32_t seed, const RsOrderKey *key, uint32_t *ordered_out) {
  const uint32_t n = j->nT;
  const uint64_t n_raw = (uint64_t)n + n / 32u + 65536u;  // room for the rejections (probability n / 2^32 per word)
  if (n_raw > 0xFFFFFFFFull) { g_err = "rs_job_shuffle_order_seed: too many target points"; return 100; }
  return shuffle_order_impl(j, nullptr, nullptr, (uint32_t)n_raw, key, ordered_out, true, seed);
}
// The same from RAW words of the PRNG stream (the first n_raw of them, in the pinned buffer rs_job_raw_buffer returned):
// the device applies the rejection rule and the modulo itself.  n_raw must leave room for the rejections (probability
// n / 2^32 per word); if fewer than n words survive the job is flagged faulty and rs_job_run returns an error.
extern "C" int rs_job_shuffle_order_raw(RsJob *j, uint32_t n_raw, const RsOrderKey *key, uint32_t *ordered_out) {
  if (!j->ws->pin_raw || (size_t)n_raw * 4 > j->ws->pin_raw_cap) { g_err = "rs_job_shuffle_order_raw: no raw buffer of that size"; return 100; }
  return shuffle_order_impl(j, nullptr, (const uint32_t *)j->ws->pin_raw, n_raw, key, ordered_out);
}
// Pinned buffer for n_words raw PRNG words (the caller's producer thread fills it while the images are staged).
extern "C" uint32_t *rs_job_raw_buffer(RsJob *j, size_t n_words) {
  Workspace *w = j->ws;
  if (cudaSetDevice(w->device) != cudaSuccess) return nullptr;
  if (n_words * 4 > w->pin_raw_cap) {
    if (w->pin_raw) cudaFreeHost(w->pin_raw);
    w->pin_raw = nullptr; w->pin_raw_cap = 0;
    if (cudaHostAlloc(&w->pin_raw, n_words * 4 + 4096, cudaHostAllocDefault) != cudaSuccess) { w->pin_raw = nullptr; return nullptr; }
    w->pin_raw_cap = n_words * 4 + 4096;
  }
  return (uint32_t *)w->pin_raw;
}
static int shuffle_order_impl(RsJob *j, const uint32_t *draws, const uint32_t *raw_pinned, uint32_t n_raw, const RsOrderKey *key,
                              uint32_t *ordered_out, bool raw_on_device, uint32_t seed) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  const RsJobDesc &d = j->d;
  const uint32_t n = j->nT;
  const size_t bytes = (size_t)n * 4, tn = (size_t)d.tw * d.th;
  // The whole pipeline reads the selection only (the mask plane of the simple API, the target pixmap of the full one):
  // it runs on the SIDE stream from the moment that is on the device (evSel), beside the upload of the image and the
  // kernels that build the job's state on the main stream, which joins it (evOrder) before the order is scattered.
  cudaStream_t s = w->stream2, s_main = w->stream;
  RS_CHECK(cudaStreamWaitEvent(s, w->evSel, 0));
  const bool raw_words = raw_pinned || raw_on_device;
  int rc = 0;
  if (raw_words && n_raw < n) { g_err = "rs_job_shuffle_order_raw: fewer raw words than target points"; return 100; }
  if (raw_words && (rc = ws_ensure(w->ord_raw, (size_t)n_raw * 4))) return rc;
  if ((rc = ws_ensure(w->ord_keys_in, bytes)) || (rc = ws_ensure(w->ord_keys_out, bytes)) || (rc = ws_ensure(w->ord_vals_in, bytes)) ||
      (rc = ws_ensure(w->ord_vals_out, bytes)) || (rc = ws_ensure(w->ord_first, bytes)) || (rc = ws_ensure(w->ord_points, bytes + 4)) ||
      (rc = ws_ensure(w->ord_flags, tn)))
    return rc;
  if (bytes > w->pin_order_cap) {
    if (w->pin_order) cudaFreeHost(w->pin_order);
    w->pin_order = nullptr; w->pin_order_cap = 0;
    RS_CHECK(cudaHostAlloc(&w->pin_order, bytes + bytes / 8 + 4096, cudaHostAllocDefault));
    w->pin_order_cap = bytes + bytes / 8 + 4096;
  }
  uint32_t *dst = nullptr;
  j->order.reset();
  std::shared_ptr<OrderEntry> entry;
  if (key && g_order_cache_on.load()) entry = order_entry_alloc(*key, w->device, n);
  if (entry) {
    dst = entry->dev;
    j->order = entry;
  } else {
    if ((rc = ws_ensure(w->targets, bytes))) return rc;
    dst = (uint32_t *)w->targets.p;
  }
  j->targets_dev = dst;
  const int T = 256;
  size_t tmp3 = 0;
  if (raw_words) {  // raw words made here or sent up (already pinned), accepted ones compacted into the draw array, then reduced mod n
    uint32_t leftover = (0x80000000u % n) * 2u;
    if (leftover >= n) leftover -= n;
    const RsAccept acc{(n <= 0x80000000u) ? 0xffffffffu - leftover : n - 1u};
    // (the draw array takes up to n_raw words here; only the first n are used afterwards)
    if ((rc = ws_ensure(w->ord_keys_in, (size_t)n_raw * 4))) return rc;
    if (raw_on_device) {
      if ((rc = mt19937_launch(seed, n_raw, (uint32_t *)w->ord_raw.p, s))) return rc;
      j->upload_launches += 1u;
    } else {
      RS_CHECK(cudaMemcpyAsync(w->ord_raw.p, raw_pinned, (size_t)n_raw * 4, cudaMemcpyHostToDevice, s));
    }
    unsigned int *d_acc = &((RsCtrl *)w->ctrl.p)->dg_acc;
    RS_CHECK(cub::DeviceSelect::If(nullptr, tmp3, (const uint32_t *)w->ord_raw.p, (uint32_t *)w->ord_keys_in.p, d_acc, (int)n_raw, acc, s));
    if ((rc = ws_ensure(w->ord_tmp, tmp3))) return rc;
    RS_CHECK(cub::DeviceSelect::If(w->ord_tmp.p, tmp3, (const uint32_t *)w->ord_raw.p, (uint32_t *)w->ord_keys_in.p, d_acc, (int)n_raw, acc, s));
    k_reduce_draws<<<(n + T - 1) / T, T, 0, s>>>((uint32_t *)w->ord_keys_in.p, d_acc, n, (RsCtrl *)w->ctrl.p);
    j->upload_launches += 3u;
  } else {  // the draws go up while the device compacts the target points
    memcpy(w->pin_order, draws, bytes);
    RS_CHECK(cudaMemcpyAsync(w->ord_keys_in.p, w->pin_order, bytes, cudaMemcpyHostToDevice, s));
  }
  // (a job of the simple API: the target pixmap is still being built on the main stream; its mask plane is the selection)
  k_target_flags<<<(unsigned)((tn + T - 1) / T), T, 0, s>>>(j->simple ? (const uint8_t *)w->smask.p : (const uint8_t *)w->raw_t.p, (uint32_t)tn,
                                                          j->simple ? 1 : d.bpp, (uint8_t *)w->ord_flags.p);
  thrust::counting_iterator<uint32_t> idx(0);
  unsigned int *d_cnt = &((RsCtrl *)w->ctrl.p)->dg_sel;
  size_t tmp = 0, tmp2 = 0;
  RS_CHECK(cub::DeviceSelect::Flagged(nullptr, tmp, idx, (const uint8_t *)w->ord_flags.p, (uint32_t *)w->ord_points.p, d_cnt, (int)tn, s));
  int bits = 1;
  while (bits < 32 && ((n - 1) >> bits)) bits++;
  RS_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, (const uint32_t *)w->ord_keys_in.p, (uint32_t *)w->ord_keys_out.p,
                                           (const uint32_t *)w->ord_vals_in.p, (uint32_t *)w->ord_vals_out.p, (int)n, 0, bits, s));
  if ((rc = ws_ensure(w->ord_tmp, tmp > tmp2 ? tmp : tmp2))) return rc;
  RS_CHECK(cub::DeviceSelect::Flagged(w->ord_tmp.p, tmp, idx, (const uint8_t *)w->ord_flags.p, (uint32_t *)w->ord_points.p, d_cnt, (int)tn, s));
  k_pack_points<<<592, T, 0, s>>>((uint32_t *)w->ord_points.p, d_cnt, d.tw);
  k_iota<<<(n + T - 1) / T, T, 0, s>>>((uint32_t *)w->ord_vals_in.p, n);
  RS_CHECK(cub::DeviceRadixSort::SortPairs(w->ord_tmp.p, tmp2, (const uint32_t *)w->ord_keys_in.p, (uint32_t *)w->ord_keys_out.p,
                                           (const uint32_t *)w->ord_vals_in.p, (uint32_t *)w->ord_vals_out.p, (int)n, 0, bits, s));
  RS_CHECK(cudaMemsetAsync(w->ord_first.p, 0xFF, bytes, s));
  k_run_heads<<<(n + T - 1) / T, T, 0, s>>>((const uint32_t *)w->ord_keys_out.p, n, (uint32_t *)w->ord_first.p);
  k_shuffle_trace<<<(n + T - 1) / T, T, 0, s>>>((const uint32_t *)w->ord_keys_in.p, (const uint32_t *)w->ord_keys_out.p,
                                               (const uint32_t *)w->ord_vals_out.p, (const uint32_t *)w->ord_first.p,
                                               (const uint32_t *)w->ord_points.p, n, dst);
  RS_CHECK(cudaGetLastError());
  j->upload_launches += 8u + (uint32_t)((bits + 7) / 8);
  if (entry) {
    RS_CHECK(cudaEventRecord(entry->ready, s));
    order_cache_publish(entry);
  }
  RS_CHECK(cudaEventRecord(w->evOrder, s));
  RS_CHECK(cudaStreamWaitEvent(s_main, w->evOrder, 0));  // the join: meta[] (k_init_target, main stream) takes the visit indices
  j->upload_launches += 1u;
  k_scatter_order<<<(n + 255) / 256, 256, 0, s_main>>>(dst, n, d.tw, (uint32_t *)w->meta.p);
  RS_CHECK(cudaGetLastError());
  if (ordered_out) {
    RS_CHECK(cudaMemcpyAsync(w->pin_order, dst, bytes, cudaMemcpyDeviceToHost, s));
    RS_CHECK(cudaStreamSynchronize(s));
    memcpy(ordered_out, w->pin_order, bytes);
  }
  return 0;
}

extern "C" void rs_job_set_passes(RsJob *j, const uint32_t *pass_end, uint32_t n_passes) {
  for (uint32_t p = 0; p < 6; p++) j->d.pass_end[p] = p < n_passes ? pass_end[p] : 0u;
  j->d.n_passes = n_passes > 6 ? 6 : n_passes;
}
extern "C" int rs_job_set_order(RsJob *j, const uint32_t *targets, const RsOrderKey *key) {
  return upload_order_impl(j, targets, key);
}

extern "C" int rs_job_upload(RsJob *j, const uint8_t *target_raw, const uint8_t *corpus_raw, const uint32_t *targets,
                             uint32_t n_targets, const uint32_t *corpus_points, uint32_t n_corpus,
                             const uint32_t *offsets, uint32_t n_offsets, const uint32_t *color_lut256,
                             const uint32_t *map_lut256, uint32_t map_lut_max) {
  if (n_targets == 0) { g_err = "rs_job_upload: no target points"; return 100; }
  uint32_t lo = 0xFFFFu, hi = 0;
  for (uint32_t i = 0; i < n_targets; i++) { const uint32_t y = targets[i] >> 16; lo = y < lo ? y : lo; hi = y > hi ? y : hi; }
  if (int rc = rs_job_upload_images(j, target_raw, corpus_raw, n_targets, lo, hi, corpus_points, n_corpus, offsets,
                                    n_offsets, color_lut256, map_lut256, map_lut_max))
    return rc;
  return rs_job_upload_order(j, targets);
}

static RsDev make_dev(const RsJob *j, uint32_t pass) {
  RsDev D;
  memset(&D, 0, sizeof D);
  const RsJobDesc &d = j->d;
  const Workspace *w = j->ws;
  D.corpus4 = j->maps ? nullptr : (const uint32_t *)j->cb.corpus;
  D.corpus8 = j->maps ? (const uint2 *)j->cb.corpus : nullptr;
  D.W = (unsigned long long *)w->W.p; D.meta = (const uint32_t *)w->meta.p;
  D.tmaps = j->maps ? (const uint32_t *)w->tmaps.p : nullptr;
  D.targets = j->targets_dev; D.corpus_pts = j->cb.cpts;
  D.cbits = j->corpus_bits && !getenv("RS_NO_CORPUS_BITS") ? j->cb.cbits : nullptr;
  D.csamples = j->cb.csamples;
  D.offsets = (const uint32_t *)w->offsets.p; D.lut_rep = (const uint32_t *)w->lut_rep.p;
  D.prober = (unsigned long long *)w->prober.p;
  { const uint32_t e = (j->nT + 31u) / 32u; D.epoch_len = e < 64u ? 64u : e; D.epoch_inv = (uint32_t)(0x100000000ull / D.epoch_len); }
  D.nb_lists = (const uint2 *)w->nb_lists.p; D.nb_counts = (const uint8_t *)w->nb_counts.p;
  D.nb_later = j->later_lists ? (const uint2 *)w->nb_later.p : nullptr;
  D.nb_later_counts = (const uint8_t *)w->nb_later_counts.p;
  D.ctrl = (RsCtrl *)w->ctrl.p; D.host_ticks = w->h_ticks; D.host_cancel = w->h_cancel;
  D.tw = d.tw; D.th = d.th; D.cw = d.cw; D.ch = d.ch; D.cn = (uint32_t)d.cw * (uint32_t)d.ch;
  D.ow = j->off_w; D.oh = j->off_h; D.gw = (d.tw + 31) / 32; D.gh = (d.th + 31) / 32;
  D.regular_r = 0;
  if (pass != 0 && j->ctx_counted && j->off_w > 0 && !(getenv("RS_REGULAR") && atoi(getenv("RS_REGULAR")) == 0)) {
    // largest |component| among the first kmax-1 offsets of the table (ascending distance): floor(sqrt(their largest d^2))
    uint32_t kmax = d.patch_size < 2 ? 2 : d.patch_size;
    if (kmax > RS_MAX_NB) kmax = RS_MAX_NB;
    std::vector<int> d2;
    for (int y = -9; y <= 9; y++) for (int x = -9; x <= 9; x++) if (x || y) d2.push_back(x * x + y * y);
    std::sort(d2.begin(), d2.end());
    int r = 0;
    while ((r + 1) * (r + 1) <= d2[kmax - 2]) r++;
    if (j->off_w > r && j->off_h > r && d.tw > 2 * r + 1 && d.th > 2 * r + 1 && j->nOff >= kmax) D.regular_r = (uint32_t)r;
  }
  D.ctx_blocks = (const uint32_t *)w->ctx_blocks.p;
  D.nT = j->nT; D.nOff = j->nOff;
  uint32_t kmax = d.patch_size < 2 ? 2 : d.patch_size;  // the size test follows the append (synthesize.h:222-224)
  D.kmax = kmax > RS_MAX_NB ? RS_MAX_NB : kmax;
  D.probes = d.max_probes; D.seed = d.seed; D.penalty = j->penalty;
  D.pass = pass; D.pass_end = d.pass_end[pass];
  for (int p = 0; p < 6; p++) D.ends[p] = d.pass_end[p];
  D.htile = d.htile; D.vtile = d.vtile; D.terminate_fraction = d.terminate_fraction;
  D.select_min = RS_SELECT_MIN_POINTS;
  if (const char *e = getenv("RS_SELECT_MIN")) D.select_min = (uint32_t)strtoul(e, nullptr, 10);  // tests, sweeps
  D.cw_inv = d.cw > 1 ? (uint32_t)(0x100000000ull / (uint32_t)d.cw) : 0xFFFFFFFFu;
  return D;
}

// Warps per visit: 1 = throughput kernel, 2/4/8 = team kernel (latency mode).  A stretch of a pass is bound either
// by visit throughput (4736 resident warps / W visits in flight) or by the depth of its dependency chains times the
// latency of one visit.  Small jobs are the latter throughout; so is the BEGINNING of pass 0 of any job: while few
// target points have a value, the nearest valued pixels of a visit are the visits just before it, wherever they
// are, and the pass runs almost serially (profiles/timeline_*.txt).  A pass is therefore cut into segments, one
// launch each, whose width shrinks as the pass fills in.  Thresholds from sweeps on B200 (profiles/); RS_TEAM_P0 /
// RS_TEAM_PN force one width for a whole pass, RS_SEG_P0="end:width,end:width,..." forces the pass-0 plan.
struct Segment { uint32_t end; unsigned width; };
static unsigned pass_width(uint32_t n, uint32_t p, int slots = 1) {
  if (n <= 32768u) return 8;
  // (four or more such jobs side by side on SM shares, a batch: what counts is the visits in flight on the whole device, and
  //  teams half as wide keep twice as many -- 64 heal jobs 2048^2 / 256^2 hole: 4.17 -> 3.42 ms per job, tools/batch_width_sweep.py)
  if (n <= 200000u) return slots >= 4 ? (p == 0 ? 4 : 2) : (p == 0 ? 8 : 4);
  if (n <= 600000u) return p == 0 ? 4 : 2;
  return 1;
}
static int plan_segments(uint32_t n_targets, uint32_t end, int ordered_visits, int patch_size, uint32_t p, Segment *out /*[4]*/) {
  const char *e = getenv(p == 0 ? "RS_TEAM_P0" : "RS_TEAM_PN");
  if (e) { const int w = atoi(e); if (w == 1 || w == 2 || w == 4 || w == 8) { out[0] = {end, (unsigned)w}; return 1; } }
  const int slots = g_job_slots.load();
  const unsigned base = pass_width(n_targets, p, slots);
  if (p != 0 || (slots >= 4 && n_targets > 32768u && n_targets <= 200000u)) { out[0] = {end, base}; return 1; }
  if (ordered_visits && !getenv("RS_SEG_P0")) {
    // A spatially sorted order (inwards, outwards, by rows...) keeps pass 0 a narrow dependency front from its first visit
    // to its last: latency mode throughout (1 Mi targets: 11.8 ms at 4 warps per visit, 20.0 with the shuffle's plan).
    out[0] = {end, base > 4u ? base : 4u};
    return 1;
  }
  Segment plan[4] = {{16384u, 8u}, {65536u, 4u}, {262144u, 2u}, {0xFFFFFFFFu, 1u}};
  if (patch_size < RS_CHUNK_SWITCH_K) {  // small patches depend on fewer earlier visits: the pass widens sooner
    plan[0].end = 8192u; plan[1].end = 24576u; plan[2].end = 65536u;  // (cfg2 4.33 -> 4.08 ms, cfg4 86.8 -> 85.4)
  }
  if (const char *sp = getenv("RS_SEG_P0")) {  // "end:width,..." ; the last entry runs to the end of the pass
    int k = 0;
    while (*sp && k < 4) {
      char *q = nullptr;
      const unsigned long en = strtoul(sp, &q, 10);
      if (q == sp || *q != ':') break;
      const unsigned long wd = strtoul(q + 1, &q, 10);
      plan[k++] = {(uint32_t)en, (unsigned)((wd == 2 || wd == 4 || wd == 8) ? wd : 1)};
      sp = (*q == ',') ? q + 1 : q;
    }
    if (k) { plan[k - 1].end = 0xFFFFFFFFu; for (; k < 4; k++) plan[k] = {0xFFFFFFFFu, 1u}; }
  }
  int n = 0;
  uint32_t begin = 0;
  for (int k = 0; k < 4 && begin < end; k++) {
    const unsigned w = plan[k].width > base ? plan[k].width : base;
    const uint32_t se = plan[k].end < end ? plan[k].end : end;
    if (se <= begin) continue;
    if (n && out[n - 1].width == w) out[n - 1].end = se;  // same width as the previous segment: one launch
    else out[n++] = {se, w};
    begin = se;
  }
  if (n == 0) out[n++] = {end, base};
  out[n - 1].end = end;
  return n;
}

// The launch plan of one pass, for tests and tools (no device needed): segment ends and warps per visit (1 = throughput kernel).
extern "C" int rs_cuda_plan_pass(uint32_t n_targets, uint32_t pass_end, int ordered_visits, int patch_size, uint32_t pass,
                                 uint32_t *ends4, uint32_t *widths4) {
  Segment seg[4];
  const int n = plan_segments(n_targets, pass_end, ordered_visits, patch_size, pass, seg);
  for (int k = 0; k < n; k++) { ends4[k] = seg[k].end; widths4[k] = seg[k].width; }
  return n;
}

extern "C" int rs_job_run(RsJob *j, RsTickFn tick, void *tick_ctx) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  cudaStream_t s = w->stream;
  uint32_t kmax_run = j->d.patch_size < 2 ? 2 : j->d.patch_size;
  bool large = kmax_run >= RS_CHUNK_SWITCH_K;
  if (const char *e = getenv("RS_CHUNK")) large = atoi(e) >= RS_CHUNK_LARGE;
  // patches of at most RS_NB_SMALL neighbours run the small-scratch instantiation (more of the SM's memory stays L1)
  bool nb_full = kmax_run > RS_NB_SMALL;
  if (const char *e = getenv("RS_NB_FULL")) nb_full = nb_full || atoi(e) != 0;
  const PassVariant &PV = w->variant[rs_variant(j->maps, large, nb_full)];
  int grid = PV.grid, grid_team = PV.grid_team;
  {  // several jobs sharing the device: each persistent grid takes its share of the SMs (rs_cuda_set_job_slots)
    const int slots = g_job_slots.load();
    if (slots > 1) {
      grid = (grid + slots - 1) / slots; if (grid < 4) grid = 4;
      grid_team = (grid_team + slots - 1) / slots; if (grid_team < 8) grid_team = 8;
    }
    if (const char *e = getenv("RS_GRID_CAP")) { const int c = atoi(e); if (c > 0 && c < grid) grid = c; if (c > 0 && c < grid_team) grid_team = c; }
  }
  // Patches of at most RS_NB_SMALL neighbours in a shuffled order: two visits per warp (k_synth_pass<..., 16>).  A spatially
  // sorted order puts consecutive visits next to each other -- the second of a pair would wait for the first every time.
  bool pair = !nb_full && PV.tp_pair != nullptr && !j->d.ordered_visits;
  if (const char *e = getenv("RS_PAIR")) {  // 0: one visit per warp; 2: pairs for sorted orders too (tests: every pair is dependent)
    const int f = atoi(e);
    pair = f == 2 ? (!nb_full && PV.tp_pair != nullptr) : (pair && f != 0);
  }
  uint32_t pair_from = 262144u;
  if (const char *e = getenv("RS_PAIR_FROM")) pair_from = (uint32_t)strtoul(e, nullptr, 10);  // sweeps
  const size_t smem_team = PV.smem_team;
  // Corpus on chip (throughput kernel, no map channels): whole in one CTA's shared memory when it fits, else split over the
  // two CTAs of a cluster.  RS_SMEM_CORPUS=0 keeps it in L2; =1/2 forces the cluster size (tests).
  int smemc_ctas = 0, grid_smemc = grid < PV.sms ? grid : PV.sms;
  uint32_t smemc_slice = 0;
  if (!j->maps && PV.tp_smemc != nullptr) {
    const uint32_t total = (uint32_t)j->d.cw * (uint32_t)j->d.ch + 1u;  // pixels + the sentinel
    const uint32_t slice_max = pair ? PV.smemc_slice_max_pair : PV.smemc_slice_max;
    int want = total <= slice_max ? 1 : (total <= 2u * slice_max ? 2 : 0);
    if (const char *e = getenv("RS_SMEM_CORPUS")) {
      const int f = atoi(e);
      if (f == 0) want = 0;
      else if (f == 2 && want == 1) want = 2;
    }
    if (want == 2 && grid_smemc < 2) want = 0;
    smemc_ctas = want;
    if (want) smemc_slice = (uint32_t)((((total + (uint32_t)want - 1u) / (uint32_t)want) + 3u) & ~3u);
  }
  RS_CHECK(cudaEventRecord(w->ev0, s));
  {  // all pass-0 patches, dependency-free
    RsDev D0 = make_dev(j, 0);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    RsCtrl *dctrl = (RsCtrl *)w->ctrl.p;
    RsLine *claims = dctrl->claims;  // [0] search kernel, [1] cooperative scan, [2] the later visits' scans
    // The first visits: the valued pixels are few and far away.  Up to 8192 of them, fewer in small jobs: a search step costs
    // about a tenth of a scan step, and visit v scans ~ K n / v table entries where the search looks at v points.
    uint32_t v1 = j->nT < 8192u ? j->nT : 8192u;
    // Their patches come from a search over the earlier target points and the context blocks (k_gather_pass0_sparse) when
    // the job reads the full offsets table; from the cooperative scan otherwise, and for tiled jobs with context.
    bool sparse = j->off_w > 0 && w->ctx_blocks.p != nullptr;
    if (const char *e = getenv("RS_SPARSE_GATHER")) sparse = sparse && atoi(e) != 0;
    if (sparse) {
      const uint32_t vs = (uint32_t)sqrt(72.0 * (double)j->nT);
      if (vs < v1) v1 = vs < 256u ? (j->nT < 256u ? j->nT : 256u) : vs;
    }
    if (const char *e = getenv("RS_GATHER_FIRST")) { const uint32_t f = (uint32_t)strtoul(e, nullptr, 10); v1 = f < j->nT ? f : j->nT; }  // sweeps
    j->ctx_counted = w->ctx_blocks.p != nullptr;
    if (j->ctx_counted) {
      RS_CHECK(cudaMemsetAsync(&dctrl->n_ctx.v, 0, 4, s));
      k_ctx_blocks<<<sms * 4, 256, 0, s>>>(D0.meta, D0.tw, D0.th, D0.gw, D0.gh, (uint32_t *)w->ctx_blocks.p, &dctrl->n_ctx.v);
    }
    if (!sparse) D0.ctx_blocks = nullptr;
    // The gather kernels write disjoint visits.  The first visits' search goes to the device first (it is short and takes
    // half of every SM's threads); the later visits' scans fill the other half from a second stream and spread over the
    // whole SM as the search kernel's CTAs leave.
    RS_CHECK(cudaEventRecord(w->evFork, s));
    if (sparse) {
      uint32_t probe = RS_SPARSE_PROBE;
      if (const char *e = getenv("RS_SPARSE_PROBE")) probe = (uint32_t)strtoul(e, nullptr, 10);  // tests
      k_gather_pass0_sparse<<<sms * 4, RS_SPARSE_WARPS * 32, 0, s>>>(D0, (uint2 *)w->nb_lists.p, (uint8_t *)w->nb_counts.p, v1,
                                                                      &claims[0].v, probe);
    }
    if (!sparse || j->d.htile || j->d.vtile)
      k_gather_pass0_coop<<<sms * 2, RS_COOP_THREADS, 0, s>>>(D0, (uint2 *)w->nb_lists.p, (uint8_t *)w->nb_counts.p, v1, &claims[1].v,
                                                               sparse ? 1 : 0);
    RS_CHECK(cudaStreamWaitEvent(w->stream2, w->evFork, 0));
    if (v1 < j->nT)
      k_gather_pass0<<<sms * 8, 256, 0, w->stream2>>>(D0, (uint2 *)w->nb_lists.p, (uint8_t *)w->nb_counts.p, v1, j->nT, &claims[2].v);
    RS_CHECK(cudaEventRecord(w->evJoin, w->stream2));
    // (pays off when the meta words of the scan no longer sit in L1/L2 next to everything else: cfg4 98 -> 94 ms, cfg3
    //  59.3 -> 58.2; a 1 Mi-target job loses 3 % to the extra kernel and the streamed lists, so small jobs keep scanning)
    uint32_t later_min = 1u << 21;
    if (const char *e = getenv("RS_LATER_LISTS_MIN")) later_min = (uint32_t)strtoul(e, nullptr, 10);  // tests, sweeps
    j->later_lists = j->d.n_passes > 1 && j->nT >= later_min;
    // (a target that is the whole image has no unusable pixel: its later passes take their patches off the head of the
    //  offsets table, rs_visit_geometry; for patches below 16 neighbours that is as fast as a list -- cfg4 50.8 -> 50.4 ms)
    if (j->nT == (uint32_t)j->d.tw * (uint32_t)j->d.th && j->d.patch_size < RS_CHUNK_SWITCH_K && j->off_w > 0 &&
        !getenv("RS_LATER_LISTS_MIN") && !(getenv("RS_REGULAR") && atoi(getenv("RS_REGULAR")) == 0))
      j->later_lists = false;
    if (j->later_lists) {  // beside pass 0, needed from pass 1 on
      k_gather_later<<<sms * 8, 256, 0, w->stream2>>>(D0, (uint2 *)w->nb_later.p, (uint8_t *)w->nb_later_counts.p);
      RS_CHECK(cudaEventRecord(w->evLater, w->stream2));
    }
    RS_CHECK(cudaStreamWaitEvent(s, w->evJoin, 0));
  }
  RS_CHECK(cudaEventRecord(w->evG, s));
  uint32_t slot = 0;
  for (uint32_t p = 0; p < j->d.n_passes; p++) {
    if (p == 1 && j->later_lists) RS_CHECK(cudaStreamWaitEvent(s, w->evLater, 0));
    RsDev D = make_dev(j, p);
    Segment seg[4];
    const int nseg = plan_segments(j->nT, j->d.pass_end[p], j->d.ordered_visits, j->d.patch_size, p, seg);
    uint32_t begin = 0, n_launched = 0;
    // one launch: the visits [b, e) of the pass at W warps per visit (W <= 1: the throughput kernel, one visit per warp or two)
    auto launch = [&](uint32_t b, uint32_t e, unsigned W, bool two, bool last) -> int {
      D.seg_begin = b; D.seg_end = e; D.slot = slot++; D.last_seg = last ? 1u : 0u;
      D.chunk = large ? RS_CHUNK_LARGE : RS_CHUNK_SMALL;
      n_launched++;
      if (W <= 1 && smemc_ctas) {  // corpus staged into the shared memory of a CTA or of a 2-CTA cluster
        D.sc_slice = smemc_slice;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(grid_smemc / smemc_ctas * smemc_ctas), 1, 1);
        cfg.blockDim = dim3(RS_TP_WARPS * 32, 1, 1);
        cfg.dynamicSmemBytes = (two ? PV.smemc_base_pair : PV.smemc_base) + (size_t)smemc_slice * 4 + 16;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)smemc_ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        RS_CHECK(cudaLaunchKernelEx(&cfg, two ? PV.tp_pair_smemc : PV.tp_smemc, D));
      } else if (W <= 1) {
        if (two) PV.tp_pair<<<grid, RS_TP_WARPS * 32, PV.smem_tp_pair, s>>>(D);
        else PV.tp<<<grid, PV.tp_threads, PV.smem_tp, s>>>(D);
      } else {
        PV.team<<<grid_team, RS_TEAM_WARPS * 32, smem_team, s>>>(D, W);
      }
      return 0;
    };
    for (int k = 0; k < nseg; k++) {
      const unsigned W = seg[k].width;
      const bool last = k == nseg - 1;
      // Two visits per warp double the visits in flight; while pass 0 is young a visit's neighbours are often among them,
      // and a warp then waits with both of its visits (cfg2: pass 0 2.00 -> 2.13 ms with pairs throughout).  Pairs from
      // visit pair_from on, where a neighbour in flight has become rare.
      if (W <= 1 && pair && p == 0 && begin < pair_from) {
        const uint32_t mid = seg[k].end < pair_from ? seg[k].end : pair_from;
        if (int rc = launch(begin, mid, W, false, last && mid == seg[k].end)) return rc;
        if (mid < seg[k].end) { if (int rc = launch(mid, seg[k].end, W, true, last)) return rc; }
      } else {
        if (int rc = launch(begin, seg[k].end, W, pair, last)) return rc;
      }
      begin = seg[k].end;
    }
    j->pass_launches[p] = n_launched;
  }
  j->launches = slot;
  RS_CHECK(cudaGetLastError());
  RS_CHECK(cudaEventRecord(w->ev1, s));
  k_writeback<<<(j->nT + 255) / 256, 256, 0, s>>>((const unsigned long long *)w->W.p, j->targets_dev, j->nT,
                                                 j->d.tw, j->d.bpp, j->d.n_color, (uint8_t *)w->raw_t.p,
                                                 j->want_sources ? (uint32_t *)w->sources.p : nullptr);
  // the rows that contain target points land in pinned memory behind the same event
  const size_t row_bytes = (size_t)j->d.tw * j->d.bpp, rows_bytes = (size_t)(j->y_max - j->y_min + 1) * row_bytes;
  if (j->simple) {  // the caller's layout: channels only, no mask byte
    const int nc = j->d.bpp - 1;
    const uint32_t npx = (j->y_max - j->y_min + 1) * (uint32_t)j->d.tw;
    k_extract_simple<<<(npx + 255) / 256, 256, 0, s>>>((const uint8_t *)w->raw_t.p, j->y_min * (uint32_t)j->d.tw, npx, nc,
                                                       (uint8_t *)w->simg.p);
    if (!j->result_direct) RS_CHECK(cudaMemcpyAsync(w->pin, w->simg.p, (size_t)npx * nc, cudaMemcpyDeviceToHost, s));
  } else if (!j->result_direct) {
    RS_CHECK(cudaMemcpyAsync(w->pin, (const uint8_t *)w->raw_t.p + (size_t)j->y_min * row_bytes, rows_bytes,
                             cudaMemcpyDeviceToHost, s));
  }
  if (j->want_sources)
    RS_CHECK(cudaMemcpyAsync((uint8_t *)w->pin + ((rows_bytes + 255) & ~(size_t)255), w->sources.p, (size_t)j->nT * 4,
                             cudaMemcpyDeviceToHost, s));
  RS_CHECK(cudaMemcpyAsync(w->h_ctrl, w->ctrl.p, RS_CTRL_COPY_BYTES, cudaMemcpyDeviceToHost, s));
  RS_CHECK(cudaEventRecord(w->evDone, s));
  // Host side of the progress/cancel contract: replay ticks in order while the device runs.
  uint32_t emitted[6] = {0, 0, 0, 0, 0, 0};  // ticks already forwarded per pass
  bool cancelled = false;
  uint32_t cancel_pass = 0, cancel_index = 0;
  unsigned idle_spins = 0;
  auto emit_upto = [&](uint32_t p, uint32_t upto) {
    while (emitted[p] < upto) {
      const uint32_t idx = emitted[p] * 4096u;
      emitted[p]++;
      if (tick && !cancelled && tick(tick_ctx, p, idx)) {
        cancelled = true;
        cancel_pass = p;
        cancel_index = idx;
        *(volatile int *)w->h_cancel = 1;
      }
    }
  };
  while (tick != nullptr) {  // (no progress callback to serve -- a batch job: wait below without polling the driver)
    cudaError_t q = cudaEventQuery(w->evDone);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) { g_err = std::string("rs_job_run: ") + cudaGetErrorString(q); return 100; }
    bool progressed = false;
    for (uint32_t p = 0; p < j->d.n_passes; p++) {
      const unsigned int seen = ((volatile unsigned int *)w->h_ticks)[p];  // a started tick index + 1 (not monotone)
      if (seen == 0) break;
      const uint32_t upto = (seen - 1u) / 4096u + 1u;
      progressed |= emitted[p] < upto;
      emit_upto(p, upto);
    }
    if (!progressed) {  // nothing new: do not burn the core (many jobs may be waiting like this one)
      if (++idle_spins > 200) { struct timespec ts = {0, 20000}; nanosleep(&ts, nullptr); }
    } else {
      idle_spins = 0;
    }
  }
  if (tick == nullptr) RS_CHECK(cudaEventSynchronize(w->evDone));  // blocking: the thread sleeps, the driver is left to the others
  RS_CHECK(cudaStreamSynchronize(s));
  // final, exact replay from the device's own counters: visits [0, pass_visits) of each pass were started
  for (uint32_t p = 0; p < j->d.n_passes; p++) {
    const unsigned long long started = w->h_ctrl->pass_visits[p];
    if (started) emit_upto(p, (uint32_t)((started - 1ull) / 4096ull + 1ull));
  }
  if (cancelled && tick && cancel_pass + 1u < j->d.n_passes) {
    // What the reference does after a cancel (lib/synthesize.h:493-497, lib/refiner.h:75-121): the cancelled synthesize()
    // call returns the betters it had; unless that is below the stop fraction the next pass starts, ticks at its index
    // 0, sees the flag and returns no betters -- which ends the pass loop.  So exactly one more tick, or none.
    // (A cancel at index 0 means no visit of that pass counted; otherwise the device's count stands in for the
    //  reference's -- it includes the visits that were already in flight beyond the cancelled tick.)
    const float frac = cancel_index == 0u ? 0.f : (float)w->h_ctrl->betters[cancel_pass] / (float)j->nT;
    if (!((double)frac < j->d.terminate_fraction)) tick(tick_ctx, cancel_pass + 1u, 0u);
  }
  if (w->h_ctrl->fault) {
    g_err = "rs_job_run: a visit waited more than 10 s for another one (inconsistent inputs); the result is invalid";
    return 100;
  }
  RS_CHECK(cudaEventElapsedTime(&j->ms_passes, w->ev0, w->ev1));
  RS_CHECK(cudaEventElapsedTime(&j->ms_synth, w->evG, w->ev1));
  return 0;
}

extern "C" void rs_job_want_sources(RsJob *j, int yes) { j->want_sources = yes != 0; }
extern "C" void rs_job_result_direct(RsJob *j, int yes) { j->result_direct = yes != 0; }

extern "C" int rs_job_download(RsJob *j, uint8_t *target_raw_out, uint32_t *sources_out) {
  const Workspace *w = j->ws;  // rs_job_run left the rows (and sources) in pinned memory
  const size_t row_bytes = (size_t)j->d.tw * j->d.bpp, rows_bytes = (size_t)(j->y_max - j->y_min + 1) * row_bytes;
  if (target_raw_out) {
    if (j->simple) { g_err = "rs_job_download: a job staged by rs_job_stage_simple returns its rows through rs_job_download_simple"; return 100; }
    uint8_t *dst = target_raw_out + (size_t)j->y_min * row_bytes;
    if (j->result_direct) {  // device -> the caller's page-locked pixmap: the box that holds the target points
      RS_CHECK(cudaSetDevice(w->device));
      const size_t x0 = (size_t)j->x_min * j->d.bpp, box_len = ((size_t)j->x_max - j->x_min + 1) * j->d.bpp;
      RS_CHECK(cudaMemcpy2DAsync(dst + x0, row_bytes, (const uint8_t *)w->raw_t.p + (size_t)j->y_min * row_bytes + x0, row_bytes, box_len,
                                 (size_t)(j->y_max - j->y_min + 1), cudaMemcpyDeviceToHost, w->stream));
      RS_CHECK(cudaStreamSynchronize(w->stream));
    } else {
      const uint8_t *src = (const uint8_t *)w->pin;
      unsigned hw = rs_host_cores();
      const size_t nt = rows_bytes < ((size_t)4 << 20) ? 1 : std::min<size_t>(rs_copy_threads_max(), hw > 2 ? hw - 1 : 1);
      const size_t per = (rows_bytes + nt - 1) / nt;
      std::vector<std::thread> th;
      for (size_t t = 1; t < nt; t++) {
        const size_t b = std::min(rows_bytes, t * per), e = std::min(rows_bytes, (t + 1) * per);
        if (e > b) th.emplace_back([=]() { memcpy(dst + b, src + b, e - b); });
      }
      memcpy(dst, src, std::min(rows_bytes, per));
      for (auto &x : th) x.join();
    }
  }
  if (sources_out) {
    if (!j->want_sources) { g_err = "rs_job_download: sources were not requested before rs_job_run"; return 100; }
    memcpy(sources_out, (const uint8_t *)w->pin + ((rows_bytes + 255) & ~(size_t)255), (size_t)j->nT * 4);
  }
  return 0;
}

extern "C" int rs_job_counters(RsJob *j, RsJobCounters *out) {
  const RsCtrl &c = *j->ws->h_ctrl;
  memset(out, 0, sizeof *out);
  out->visits = c.visits; out->evals = c.evals; out->evals_issued = c.evals_issued; out->compares = c.compares;
  out->offset_scans = c.offset_scans; out->heur_evals = c.heur_evals; out->heur_skips = c.heur_skips;
  out->perfect = c.perfect;
  for (int p = 0; p < 6; p++) { out->betters[p] = c.betters[p]; out->pass_visits[p] = c.pass_visits[p]; out->sum_best[p] = c.sum_best[p]; }
  out->passes_run = c.passes_run;
  out->n_corpus = c.n_corpus;
  out->ms_passes = j->ms_passes;
  out->ms_synth = j->ms_synth;
  out->kernel_launches = j->upload_launches + 2u + (j->later_lists ? 1u : 0u) + j->launches + 1u;  // + gathers, passes, write-back
  for (uint32_t p = 0; p < c.passes_run && p < 6; p++) out->synth_launches_run += j->pass_launches[p];
  for (uint32_t p = 0; p < c.passes_run && p < 6; p++)
    out->ms_pass[p] = c.pass_end_ns[p] > c.tick_ns[p][0] ? (float)((c.pass_end_ns[p] - c.tick_ns[p][0]) * 1e-6) : 0.f;
  return 0;
}

// Start time (ns since the pass began) of every 4096th visit of a pass of the last run: the throughput profile of
// the pass.  Returns the number of entries written.
extern "C" uint32_t rs_job_timeline(RsJob *j, uint32_t pass, uint64_t *out_ns, uint32_t cap) {
  const RsCtrl &c = *j->ws->h_ctrl;
  if (pass >= 6 || pass >= c.passes_run) return 0;
  const unsigned long long started = c.pass_visits[pass];
  uint32_t n = started ? (uint32_t)((started - 1ull) / 4096ull + 1ull) : 0u;
  if (n > RS_TIMELINE) n = RS_TIMELINE;
  if (n > cap) n = cap;
  for (uint32_t i = 0; i < n; i++) out_ns[i] = c.tick_ns[pass][i] - c.tick_ns[pass][0];
  return n;
}

// Device-built offsets table of a job, for parity tests against the host/oracle table.
extern "C" int rs_job_read_offsets(RsJob *j, uint32_t *out, uint32_t cap) {
  Workspace *w = j->ws;
  RS_CHECK(cudaSetDevice(w->device));
  const uint32_t n = j->nOff < cap ? j->nOff : cap;
  RS_CHECK(cudaStreamSynchronize(w->stream));
  RS_CHECK(cudaMemcpy(out, w->offsets.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------- gather-rate micro-benchmark
// What the distance loop is made of: independent, uniformly random aligned loads of one corpus pixel (4 or 8 bytes)
// from a buffer the size of a corpus.  Gives the rate this GPU sustains for that access pattern (L1/L2 sector
// gathers), the practical ceiling of neighbour compares per second.
// Pinned down so that two runs agree: the launch shape and the shared-memory carve-out of the throughput kernel (one
// 1024-thread CTA per SM, the same dynamic shared memory, hence the same L1 size -- a 256 KB corpus is L1-resident or
// not depending on exactly that), launches of >= 20 ms each, the median of `repeats` (>= 5) of them after a warm-up.
template <typename T>
__global__ void __launch_bounds__(1024, 1) k_gather_rate(const T *__restrict__ buf, uint32_t n_elems, uint32_t iters,
                                                         unsigned long long *__restrict__ sink) {
  extern __shared__ __align__(16) unsigned char gr_smem[];
  uint32_t h = rs_mix32(blockIdx.x * 1024u + threadIdx.x + 0x9E3779B9u);
  unsigned long long acc = 0;
  for (uint32_t i = 0; i < iters; i++) {
    uint32_t idx[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { h = h * 1664525u + 1013904223u; idx[u] = __umulhi(rs_mix32(h), n_elems); }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const T v = __ldg(buf + idx[u]);
      acc += (sizeof(T) == 8) ? (unsigned long long)((const uint32_t *)&v)[0] + ((const uint32_t *)&v)[sizeof(T) / 4 - 1]
                              : (unsigned long long)((const uint32_t *)&v)[0];
    }
  }
  if (acc == 0x123456789ull) { *sink = acc; gr_smem[threadIdx.x] = 1; }  // keep the loads (and the carve-out) alive
}
extern "C" int rs_cuda_gather_rate(size_t buffer_bytes, int elem_bytes, int repeats, double *loads_per_s) {
  if (elem_bytes != 4 && elem_bytes != 8) { g_err = "rs_cuda_gather_rate: element size must be 4 or 8"; return 100; }
  const uint32_t n = (uint32_t)(buffer_bytes / (size_t)elem_bytes);
  void *buf = nullptr;
  unsigned long long *sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int sms = 148, dev = 0, rc = 0;
  uint32_t iters = 256;
  std::vector<float> times;
  if (repeats < 5) repeats = 5;
  const size_t smem = pass_smem(elem_bytes == 8, RS_TP_WARPS);
#define GCHK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); rc = 100; goto out; } } while (0)
  GCHK(cudaGetDevice(&dev));
  GCHK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GCHK(cudaFuncSetAttribute(k_gather_rate<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GCHK(cudaFuncSetAttribute(k_gather_rate<uint2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GCHK(cudaMalloc(&buf, (size_t)n * elem_bytes));
  GCHK(cudaMalloc(&sink, 8));
  GCHK(cudaMemset(buf, 1, (size_t)n * elem_bytes));
  GCHK(cudaEventCreate(&e0));
  GCHK(cudaEventCreate(&e1));
  for (int r = -2; r < repeats; r++) {  // r = -2: warm-up; r = -1: calibrates the launch length to >= 20 ms
    GCHK(cudaEventRecord(e0));
    if (elem_bytes == 4) k_gather_rate<uint32_t><<<sms, 1024, smem>>>((const uint32_t *)buf, n, iters, sink);
    else k_gather_rate<uint2><<<sms, 1024, smem>>>((const uint2 *)buf, n, iters, sink);
    GCHK(cudaEventRecord(e1));
    GCHK(cudaEventSynchronize(e1));
    float ms = 0.f;
    GCHK(cudaEventElapsedTime(&ms, e0, e1));
    if (r == -1) {
      const double scale = ms > 0.f ? 20.0 / ms : 1.0;
      if (scale > 1.0) iters = (uint32_t)std::min(1.0e7, iters * scale + 1.0);
    } else if (r >= 0) {
      times.push_back(ms / (float)iters);
    }
  }
  std::sort(times.begin(), times.end());
  *loads_per_s = (double)sms * 1024.0 * 8.0 / ((double)times[times.size() / 2] * 1e-3);
out:
#undef GCHK
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  return rc;
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
  if (maps) BCHK(cudaMalloc(&d_c8, (cn + 1) * 8)); else BCHK(cudaMalloc(&d_c4, (cn + 1) * 4));
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
    k_canon_corpus<<<(unsigned)((cn + 256) / 256), 256>>>(d_raw, (int)cn, desc->bpp, desc->n_color, desc->n_map,
                                                        desc->map_bip, d_c4, d_c8);
    k_replicate_lut<<<(RS_LUT_WORDS + 255) / 256, 256>>>(d_lut, d_lut + 256, d_rep);
    RsDev D;
    memset(&D, 0, sizeof D);
    D.corpus4 = d_c4; D.corpus8 = d_c8; D.lut_rep = d_rep; D.cw = desc->cw; D.ch = desc->ch; D.cn = (uint32_t)cn;
    D.penalty = 65535u * (uint32_t)desc->n_color + map_lut_max * (uint32_t)desc->n_map;
    D.chunk = RS_CHUNK_SMALL;
    const size_t smem = pass_smem(maps, RS_BF_WARPS);
    const unsigned grid = (n_visits + RS_BF_WARPS - 1) / RS_BF_WARPS;
    if (maps) {
      BCHK(cudaFuncSetAttribute(k_bestfit_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_bestfit_batch<true><<<grid > 1184 ? 1184 : grid, RS_BF_WARPS * 32, smem>>>(D, n_visits, d_nbb, d_nbo, d_nbpix, desc->n_color, desc->n_map,
                                                       desc->map_bip, d_cb, d_c, d_bs, d_bi);
    } else {
      BCHK(cudaFuncSetAttribute(k_bestfit_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_bestfit_batch<false><<<grid > 1184 ? 1184 : grid, RS_BF_WARPS * 32, smem>>>(D, n_visits, d_nbb, d_nbo, d_nbpix, desc->n_color, desc->n_map,
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
