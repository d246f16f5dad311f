// Host-side preparation of one synthesis job: everything lib/engine.c:539-655 does before it calls
// refiner(), restated over flat vectors.  The outputs are the kernel's schedule (visit order), its
// neighbour-offset table and its two metric tables, so they must equal the reference's arrays exactly;
// tests/test_host_prep.py compares them with the oracle.
#pragma once
#include <cstdint>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/resynthesizer.h"

extern "C" unsigned rs_host_cores(void);  // cores this process may use for helper threads (csrc/rs_kernels.cu)

namespace rs {

// Points are packed x | y << 16 (coordinates < 32768) all the way from the scan to the device.
inline uint32_t pack_xy(int x, int y) { return (uint32_t)x | ((uint32_t)y << 16); }
inline int unpack_x(uint32_t p) { return (int)(p & 0xFFFFu); }
inline int unpack_y(uint32_t p) { return (int)(p >> 16); }

// GLib GRand (MT19937) as the reference's product build draws it (lib/engine.c:643, lib/orderTarget.h:44,76).
class GRandMT {
 public:
  explicit GRandMT(uint32_t seed);
  uint32_t next32();
  uint32_t int_range(uint32_t n);  // g_rand_int_range(0, n)
  void fill_int_range(uint32_t n, uint32_t *out, size_t count);  // `count` successive int_range(n) draws
  void fill_raw(uint32_t *out, size_t count);                    // `count` successive next32() words
 private:
  void refill();
  uint32_t mt_[624];
  uint32_t out_[624];  // tempered outputs of the current block
  int mti_;
};

// The raw 32-bit words of a GRand stream do not depend on the range they are later reduced to.  RawStream starts producing
// them on its own thread as soon as the seed is known (before the number of target points is); reduce() then turns the
// first words into `count` g_rand_int_range(0, n) draws exactly as GRandMT::fill_int_range would (same rejections).
class RawStream {
 public:
  // external: caller-owned buffer of max_words words (e.g. pinned memory), else an internal per-thread buffer
  RawStream(uint32_t seed, size_t max_words, uint32_t *external = nullptr);
  ~RawStream();
  void reduce(uint32_t n, uint32_t *out, size_t count);
  void wait_ready(size_t words);  // until the first min(words, max_words) words are in the buffer
 private:
  uint32_t *buf_;
  std::atomic<size_t> ready_{0};
  std::atomic<bool> stop_{false};
  size_t cap_;
  uint32_t seed_;
  std::thread th_;
};

// Jump-ahead of MT19937, so that the device makes the PRNG stream of a large job on many SMs at once (k_mt19937_raw in
// csrc/rs_kernels.cu).  The generator is linear over GF(2): with phi(z) its characteristic polynomial (degree 19937) and
// g(z) = z^J mod phi(z), the untempered word stream x obeys x[J + j] = XOR over the set bits k of g of x[k + j].
// mt_jump_poly(q, jump) = the positions of the set bits of z^(q * jump) mod phi, ascending: the state after q * jump words
// is that combination of the first 19937 + 624 words.  phi comes from Berlekamp-Massey on one output bit (computed once
// per process, ~20 ms); the polynomials of q = 1, 2, ... are built on demand, each from the one before (~10 ms), and kept.
// Thread-safe; the reference returned stays valid.
constexpr uint32_t kMtDegree = 19937;
const std::vector<uint16_t> &mt_jump_poly(uint32_t q, uint32_t jump_words);

// lib/matchWeighting.h:142-204.  Tables over the signed difference, index 256+d, like the reference.
void build_metric_tables(double sensitivity, double map_weight, uint16_t color512[512], uint32_t map512[512]);

// lib/engine.c:465-497 with glibc's merge-sort tie order (ascending x^2+y^2, ties in reverse row-major order).
// Packed int16 pairs (x | y << 16).
void build_sorted_offsets(int target_w, int target_h, int corpus_w, int corpus_h, std::vector<uint32_t> &out);

// lib/engine.c:338-431.  Points in row-major scan order.
void collect_target_points(const uint8_t *pix, int w, int h, int bpp, std::vector<uint32_t> &out);
void collect_corpus_points(const uint8_t *pix, int w, int h, int bpp, const TFormatIndices &fi, std::vector<uint32_t> &out);

// Is there any corpus point at all (mask 0xFF and not transparent)?  Stops at the first one (lib/engine.c:620-627).
bool has_target_point(const uint8_t *pix, int w, int h, int bpp);
bool has_corpus_point(const uint8_t *pix, int w, int h, int bpp, const TFormatIndices &fi);

// lib/orderTarget.h:268-343 (+ brushfire.h, engineTypes.h).  Returns 0 or IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE.
// Stable ascending sort of n (key, value) pairs in place; false = not done (the host's own radix sort runs instead).
using PairSorter = std::function<bool(uint32_t *keys, uint32_t *vals, size_t n)>;
int order_target_points(int match_context_type, std::vector<uint32_t> &pts, GRandMT &prng, const PairSorter *sorter = nullptr);
// Both steps for a pixmap with n_known selected pixels, overlapped where the mode allows; -1 if the count was wrong.
int collect_and_order(int match_context_type, const uint8_t *mask0, int w, int h, size_t pixel_stride, size_t row_stride,
                      size_t n_known, uint32_t seed, std::vector<uint32_t> &pts, const PairSorter *sorter = nullptr);
void collect_target_points_strided(const uint8_t *mask0, int w, int h, size_t pixel_stride, size_t row_stride,
                                   std::vector<uint32_t> &out);

// lib/passes.h:67-93.  Returns the estimated total visit count.
uint32_t pass_schedule(uint32_t n_targets, uint32_t ends[6]);

}  // namespace rs
