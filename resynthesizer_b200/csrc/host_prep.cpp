#include "host_prep.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <memory>
#include <climits>
#include <cmath>
#include <cstring>

namespace rs {

// ------------------------------------------------------------------------------------------- GRand
GRandMT::GRandMT(uint32_t seed) {
  mt_[0] = seed;
  for (int i = 1; i < 624; i++) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
  mti_ = 624;
}

// Regenerates the 624-word state.  Three dependency-free segments (each reads only words that are already final),
// so the compiler can vectorise them; then the whole block is tempered in place into out_[].
__attribute__((target_clones("avx2", "default"))) void GRandMT::refill() {
  uint32_t *mt = mt_;
  auto twist = [](uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  };
#pragma GCC ivdep
  for (int k = 0; k < 227; k++) mt[k] = twist(mt[k], mt[k + 1], mt[k + 397]);
#pragma GCC ivdep
  for (int k = 227; k < 454; k++) mt[k] = twist(mt[k], mt[k + 1], mt[k - 227]);
#pragma GCC ivdep
  for (int k = 454; k < 623; k++) mt[k] = twist(mt[k], mt[k + 1], mt[k - 227]);
  mt[623] = twist(mt[623], mt[0], mt[396]);
#pragma GCC ivdep
  for (int k = 0; k < 624; k++) {
    uint32_t y = mt[k];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    out_[k] = y;
  }
  mti_ = 0;
}

uint32_t GRandMT::next32() {
  if (mti_ >= 624) refill();
  return out_[mti_++];
}

uint32_t GRandMT::int_range(uint32_t n) {
  if (n == 0) return 0;
  uint32_t v;
  if (n <= 0x80000000u) {
    uint32_t leftover = (0x80000000u % n) * 2u;
    if (leftover >= n) leftover -= n;
    const uint32_t maxvalue = 0xffffffffu - leftover;
    do v = next32(); while (v > maxvalue);
  } else {
    do v = next32(); while (v >= n);
  }
  return v % n;
}

// ----------------------------------------------------------------------------------- metric tables
void build_metric_tables(double sensitivity, double map_weight, uint16_t color512[512], uint32_t map512[512]) {
  const float cauchy = (float)sensitivity, mapw = (float)map_weight;  // narrowed first (matchWeighting.h:194-197)
  const double scale = (double)(cauchy * 256);
  const double full = std::log((256.0 / scale) * (256.0 / scale) + 1.0);
  for (int d = -256; d < 256; d++) {
    const double r = (double)d / scale;
    const double v = std::log(r * r + 1.0) / full * (float)65535;
    color512[256 + d] = (uint16_t)v;
    map512[256 + d] = (uint32_t)(d * d * mapw * 4.0);
  }
}

// --------------------------------------------------------------------------------- neighbour offsets
void build_sorted_offsets(int tw, int th, int cw, int ch, std::vector<uint32_t> &out) {
  const int w = std::min(tw, cw), h = std::min(th, ch);
  const size_t n = (size_t)(2 * w - 1) * (size_t)(2 * h - 1);
  const uint32_t maxd = (uint32_t)((w - 1) * (w - 1) + (h - 1) * (h - 1));
  std::vector<uint32_t> start((size_t)maxd + 2, 0u);
  for (int y = -h + 1; y < h; y++) {
    const uint32_t yy = (uint32_t)(y * y);
    for (int x = -w + 1; x < w; x++) start[yy + (uint32_t)(x * x) + 1]++;
  }
  for (uint32_t d = 0; d <= maxd; d++) start[d + 1] += start[d];
  out.resize(n);
  // stable counting sort fed in reverse row-major order == what the never-equal comparator yields
  for (int y = h - 1; y > -h; y--) {
    const uint32_t yy = (uint32_t)(y * y);
    for (int x = w - 1; x > -w; x--)
      out[start[yy + (uint32_t)(x * x)]++] = ((uint32_t)x & 0xFFFFu) | ((uint32_t)y << 16);
  }
}

// ------------------------------------------------------------------------------------------ points
void collect_target_points(const uint8_t *pix, int w, int h, int bpp, std::vector<uint32_t> &out) {
  collect_target_points_strided(pix, w, h, (size_t)bpp, (size_t)w * bpp, out);
}

// mask0 = the mask byte of pixel (0,0); pixel (x,y) has its mask byte at mask0 + y * row_stride + x * pixel_stride.
void collect_target_points_strided(const uint8_t *mask0, int w, int h, size_t pixel_stride, size_t row_stride,
                                   std::vector<uint32_t> &out) {
  const size_t npx = (size_t)w * h;
  unsigned hw = rs_host_cores();
  const int nt = npx < ((size_t)1 << 19) ? 1 : (int)std::min<unsigned>(npx < ((size_t)1 << 21) ? 4u : 8u, hw ? hw : 1u);
  if (nt <= 1) {
    out.clear();
    for (int y = 0; y < h; y++) {
      const uint8_t *row = mask0 + (size_t)y * row_stride;
      for (int x = 0; x < w; x++)
        if (row[(size_t)x * pixel_stride] != 0) out.push_back(pack_xy(x, y));
    }
    return;
  }
  // large images: row bands on several cores (count, then fill at the band's offset: row-major order is kept)
  std::vector<size_t> cnt((size_t)nt + 1, 0);
  const int per = (h + nt - 1) / nt;
  auto band = [&](int t, bool fill) {
    const int y0 = std::min(h, t * per), y1 = std::min(h, (t + 1) * per);
    size_t c = 0;
    uint32_t *dst = fill ? out.data() + cnt[t] : nullptr;
    for (int y = y0; y < y1; y++) {
      const uint8_t *row = mask0 + (size_t)y * row_stride;
      for (int x = 0; x < w; x++)
        if (row[(size_t)x * pixel_stride] != 0) { if (fill) dst[c] = pack_xy(x, y); c++; }
    }
    if (!fill) cnt[t + 1] = c;
  };
  for (int phase = 0; phase < 2; phase++) {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(band, t, phase == 1);
    band(0, phase == 1);
    for (auto &x : th) x.join();
    if (phase == 0) {
      for (int t = 0; t < nt; t++) cnt[t + 1] += cnt[t];
      out.resize(cnt[nt]);
    }
  }
}

void collect_corpus_points(const uint8_t *pix, int w, int h, int bpp, const TFormatIndices &fi, std::vector<uint32_t> &out) {
  out.clear();
  const bool alpha = fi.isAlphaSource != 0;
  const int ab = fi.alpha_bip;
  for (int y = 0; y < h; y++) {
    const uint8_t *row = pix + (size_t)y * w * bpp;
    for (int x = 0; x < w; x++) {
      const uint8_t *p = row + (size_t)x * bpp;
      if (p[0] == 0xFF && (!alpha || p[ab] != 0)) out.push_back(pack_xy(x, y));
    }
  }
}

// Is any pixel selected?  Sparse samples first (a selection is rarely a single pixel), then every pixel.
bool has_target_point(const uint8_t *pix, int w, int h, int bpp) {
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; i += 61) if (pix[i * bpp] != 0) return true;
  for (size_t i = 0; i < n; i++) if (pix[i * bpp] != 0) return true;
  return false;
}

bool has_corpus_point(const uint8_t *pix, int w, int h, int bpp, const TFormatIndices &fi) {
  const bool alpha = fi.isAlphaSource != 0;
  const int ab = fi.alpha_bip;
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; i++) {
    const uint8_t *p = pix + i * bpp;
    if (p[0] == 0xFF && (!alpha || p[ab] != 0)) return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------- ordering
// v % n for a fixed n without a division per draw (Lemire, "Faster remainder by direct computation", 2019)
struct FastMod {
  uint64_t M;
  uint32_t n;
  explicit FastMod(uint32_t n_) : M(~0ull / n_ + 1ull), n(n_) {}
  uint32_t mod(uint32_t v) const {
    const uint64_t low = M * v;
    return (uint32_t)(((__uint128_t)low * n) >> 64);
  }
};

void GRandMT::fill_int_range(uint32_t n, uint32_t *out, size_t count) {
  if (n == 0) { for (size_t i = 0; i < count; i++) out[i] = 0; return; }
  uint32_t leftover = (0x80000000u % n) * 2u;
  if (leftover >= n) leftover -= n;
  const uint32_t maxvalue = (n <= 0x80000000u) ? 0xffffffffu - leftover : n - 1u;
  const FastMod fm(n);
  size_t i = 0;
  while (i < count) {
    if (mti_ >= 624) refill();
    const int avail = 624 - mti_;
    const size_t want = count - i;
    const int take = (size_t)avail < want ? avail : (int)want;
    const uint32_t *src = out_ + mti_;
    int used = 0;
    for (; used < take && i < count; used++) {  // rejection is rare (< 2^-12 per draw for n <= 2^20)
      const uint32_t v = src[used];
      if (v <= maxvalue) out[i++] = fm.mod(v);
    }
    mti_ += used;
  }
}

// for i in [0, n): swap(a[i], a[j_i]) where the draws j_i come from `fill(offset, count, out)` in order.  The draws
// (MT19937 blocks + range reduction) and the swaps (random access) cost about the same and the draws do not depend on
// the swaps, so for large vectors a second thread produces them a block ahead of the swap loop.
template <class Fill>
static void swaps_from_draws(uint32_t *a, size_t n, Fill fill) {
  static thread_local std::vector<uint32_t> js;  // reused: no page faults per job
  js.resize(n);
  uint32_t *jp = js.data();
  constexpr size_t AHEAD = 64, BLOCK = 1u << 16;
  if (n < 4 * BLOCK) {
    fill((size_t)0, n, jp);
    for (size_t i = 0; i < n; i++) {
      if (i + AHEAD < n) __builtin_prefetch(a + jp[i + AHEAD], 1);
      std::swap(a[i], a[jp[i]]);
    }
    return;
  }
  std::atomic<size_t> ready{0};
  std::thread producer([&]() {
    for (size_t off = 0; off < n; off += BLOCK) {
      const size_t len = n - off < BLOCK ? n - off : BLOCK;
      fill(off, len, jp + off);
      ready.store(off + len, std::memory_order_release);
    }
  });
  for (size_t off = 0; off < n; off += BLOCK) {
    const size_t end = n - off < BLOCK ? n : off + BLOCK;
    while (ready.load(std::memory_order_acquire) < end) { /* spin: the producer is at most a block away */ }
    for (size_t i = off; i < end; i++) {
      if (i + AHEAD < end) __builtin_prefetch(a + jp[i + AHEAD], 1);
      std::swap(a[i], a[jp[i]]);
    }
  }
  producer.join();
}

// lib/orderTarget.h:57-80: every i swaps with a draw from the band [i - half, i + half) clipped to the vector.
void GRandMT::fill_raw(uint32_t *out, size_t count) {
  size_t i = 0;
  while (i < count) {
    if (mti_ >= 624) refill();
    const size_t take = std::min<size_t>((size_t)(624 - mti_), count - i);
    std::memcpy(out + i, out_ + mti_, take * 4);
    mti_ += (int)take;
    i += take;
  }
}

// ------------------------------------------------------------------------------------------- MT19937 jump-ahead
namespace {
constexpr size_t kPolyWords = (kMtDegree + 64) / 64;  // bits 0..19937 (phi has bit 19937 set)
using Poly = std::vector<uint64_t>;                    // bit k = coefficient of z^k

inline bool poly_bit(const Poly &p, size_t k) { return (p[k >> 6] >> (k & 63)) & 1u; }
// dst ^= src << shift (whole words of src; dst is long enough)
void xor_shl(uint64_t *dst, const uint64_t *src, size_t src_words, size_t shift) {
  const size_t ws = shift >> 6, bs = shift & 63;
  if (bs == 0) {
    for (size_t i = 0; i < src_words; i++) dst[i + ws] ^= src[i];
  } else {
    uint64_t carry = 0;
    for (size_t i = 0; i < src_words; i++) {
      dst[i + ws] ^= (src[i] << bs) | carry;
      carry = src[i] >> (64 - bs);
    }
    dst[src_words + ws] ^= carry;
  }
}

// The characteristic polynomial: Berlekamp-Massey over GF(2) on 2 * 19937 successive values of one state bit (bit 0 of
// the untempered words; phi is irreducible, so every non-zero bit stream of the generator has it as minimal polynomial).
Poly mt_characteristic_polynomial() {
  const size_t N = kMtDegree, T = 2 * N;
  std::vector<uint32_t> x(624 + T);
  x[0] = 5489u;  // any non-zero state does
  for (size_t i = 1; i < 624; i++) x[i] = 1812433253u * (x[i - 1] ^ (x[i - 1] >> 30)) + (uint32_t)i;
  for (size_t i = 624; i < x.size(); i++) {
    const uint32_t y = (x[i - 624] & 0x80000000u) | (x[i - 623] & 0x7fffffffu);
    x[i] = x[i - 227] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  }
  const size_t W = kPolyWords + 1;
  Poly C(W, 0), B(W, 0), R(W, 0), Tmp(W, 0);
  C[0] = 1; B[0] = 1;
  size_t L = 0, m = 1;
  for (size_t i = 0; i < T; i++) {
    // R: bit j = s[i - j]
    uint64_t carry = x[624 + i] & 1u;
    for (size_t w = 0; w < W; w++) { const uint64_t nc = R[w] >> 63; R[w] = (R[w] << 1) | carry; carry = nc; }
    uint64_t acc = 0;
    for (size_t w = 0; w <= (L >> 6); w++) acc ^= C[w] & R[w];
    if (__builtin_parityll(acc)) {
      const bool grow = 2 * L <= i;
      if (grow) Tmp = C;
      // C ^= z^m B: deg(z^m B) <= the new L <= N, so the words of B whose image would leave C are zero
      const size_t ws = m >> 6;
      if (ws + 1 < W) xor_shl(C.data(), B.data(), W - ws - 1, m);
      if (grow) { L = i + 1 - L; B.swap(Tmp); m = 1; } else { m++; }
    } else {
      m++;
    }
  }
  Poly phi(kPolyWords, 0);
  if (L != N) return phi;  // (cannot happen)
  for (size_t k = 0; k <= N; k++)
    if (poly_bit(C, N - k)) phi[k >> 6] |= (uint64_t)1 << (k & 63);
  return phi;
}

// a * b mod phi; a, b of degree < 19937
Poly poly_mulmod(const Poly &a, const Poly &b, const Poly &phi) {
  std::vector<uint64_t> acc(2 * kPolyWords + 2, 0);
  for (size_t w = 0; w < kPolyWords; w++) {
    uint64_t bits = b[w];
    while (bits) {
      const int k = __builtin_ctzll(bits);
      bits &= bits - 1;
      xor_shl(acc.data(), a.data(), kPolyWords, w * 64 + (size_t)k);
    }
  }
  for (size_t i = 2 * (size_t)kMtDegree; i-- > kMtDegree;)
    if ((acc[i >> 6] >> (i & 63)) & 1u) xor_shl(acc.data(), phi.data(), kPolyWords, i - kMtDegree);
  Poly r(acc.begin(), acc.begin() + kPolyWords);
  r[kPolyWords - 1] &= ((uint64_t)1 << (kMtDegree & 63)) - 1u;  // (bit 19937 and above are zero after the reduction)
  return r;
}

struct JumpTable {
  std::mutex mu;
  Poly phi;
  uint32_t jump = 0;
  Poly g1;                                                  // z^jump mod phi
  Poly last;                                                // z^(q * jump) mod phi of the newest entry
  std::vector<std::unique_ptr<std::vector<uint16_t>>> idx;  // idx[q - 1]
};
JumpTable g_jump;
}  // namespace

const std::vector<uint16_t> &mt_jump_poly(uint32_t q, uint32_t jump_words) {
  static const std::vector<uint16_t> none;
  if (q == 0 || jump_words == 0) return none;
  std::lock_guard<std::mutex> lk(g_jump.mu);
  if (g_jump.phi.empty()) g_jump.phi = mt_characteristic_polynomial();
  if (g_jump.jump != jump_words) {  // z^jump by square and multiply
    g_jump.idx.clear();
    Poly r(kPolyWords, 0), base(kPolyWords, 0);
    r[0] = 1; base[0] = 2;
    for (uint32_t e = jump_words; e; e >>= 1) {
      if (e & 1u) r = poly_mulmod(r, base, g_jump.phi);
      if (e > 1u) base = poly_mulmod(base, base, g_jump.phi);
    }
    g_jump.g1 = r;
    g_jump.jump = jump_words;
  }
  while (g_jump.idx.size() < q) {
    g_jump.last = g_jump.idx.empty() ? g_jump.g1 : poly_mulmod(g_jump.last, g_jump.g1, g_jump.phi);
    std::unique_ptr<std::vector<uint16_t>> v(new std::vector<uint16_t>());
    for (size_t k = 0; k < kMtDegree; k++)
      if (poly_bit(g_jump.last, k)) v->push_back((uint16_t)k);
    g_jump.idx.push_back(std::move(v));
  }
  return *g_jump.idx[q - 1];
}

RawStream::RawStream(uint32_t seed, size_t max_words, uint32_t *external) : cap_(max_words), seed_(seed) {
  if (external) {
    buf_ = external;
  } else {
    static thread_local std::vector<uint32_t> storage;  // reused by the calling thread's next job
    storage.resize(cap_);
    buf_ = storage.data();
  }
  uint32_t *b = buf_;
  th_ = std::thread([this, b]() {
    GRandMT g(seed_);
    constexpr size_t BLOCK = 624 * 64;
    for (size_t off = 0; off < cap_ && !stop_.load(std::memory_order_relaxed); off += BLOCK) {
      const size_t len = std::min(BLOCK, cap_ - off);
      g.fill_raw(b + off, len);
      ready_.store(off + len, std::memory_order_release);
    }
  });
}
RawStream::~RawStream() {
  stop_.store(true);
  if (th_.joinable()) th_.join();
}
void RawStream::wait_ready(size_t words) {
  const size_t want = std::min(words, cap_);
  while (ready_.load(std::memory_order_acquire) < want) { /* spin: the producer has been running since the job began */ }
}
void RawStream::reduce(uint32_t n, uint32_t *out, size_t count) {
  if (n == 0) { for (size_t i = 0; i < count; i++) out[i] = 0; return; }
  uint32_t leftover = (0x80000000u % n) * 2u;
  if (leftover >= n) leftover -= n;
  const uint32_t maxvalue = (n <= 0x80000000u) ? 0xffffffffu - leftover : n - 1u;
  const FastMod fm(n);
  size_t src = 0, i = 0;
  const uint32_t *b = buf_;
  while (i < count) {
    size_t avail = ready_.load(std::memory_order_acquire);
    if (avail <= src) {
      if (src >= cap_) {  // out of prepared words (cannot happen with the caller's slack): finish from a fresh stream
        GRandMT g(seed_);
        std::vector<uint32_t> skip(4096);
        for (size_t done = 0; done < cap_; done += skip.size()) g.fill_raw(skip.data(), std::min(skip.size(), cap_ - done));
        g.fill_int_range(n, out + i, count - i);
        return;
      }
      continue;
    }
    for (; src < avail && i < count; src++) {
      const uint32_t v = b[src];
      if (v <= maxvalue) out[i++] = fm.mod(v);  // rejection is rare (< 2^-12 per draw for n <= 2^20)
    }
  }
}

static void shuffle_bands(std::vector<uint32_t> &p, GRandMT &prng) {
  const int last = (int)p.size() - 1;
  const int half = (int)(p.size() * 0.1);  // IMAGE_SYNTH_BAND_FRACTION
  // the band is 2*half wide except within `half` of either end: the full-band run is drawn in bulk
  swaps_from_draws(p.data(), p.size(), [&](size_t off, size_t count, uint32_t *out) {
    size_t k = 0;
    while (k < count) {
      const int i = (int)(off + k);
      const int lo = std::max(i - half, 0), hi = std::min(i + half, last);
      if (i >= half && i + half <= last && hi - lo == 2 * half) {
        const int run_end = std::min(last - half, (int)(off + count) - 1);  // inclusive: last i of this call with a full band
        const size_t cnt = (size_t)(run_end - i + 1);
        prng.fill_int_range((uint32_t)(2 * half), out + k, cnt);
        for (size_t t = 0; t < cnt; t++) out[k + t] += (uint32_t)(i + (int)t - half);
        k += cnt;
      } else {
        out[k] = (uint32_t)lo + prng.int_range((uint32_t)(hi - lo));
        k++;
      }
    }
  });
}

static unsigned ray_index(int x, int y) {
  return (unsigned)(std::atan2((double)(float)y, (double)(float)x) * 200 /
                        3.1415926535897932384626433832795028841971693993751 + 200);
}

// Stable LSD radix sort of (key, payload) pairs by 32-bit key, ascending: 11-bit digits (3 passes), passes whose digit
// is the same for every key skipped, buffers reused across calls.
static void radix_sort_pairs(std::vector<uint32_t> &keys, std::vector<uint32_t> &vals) {
  const size_t n = keys.size();
  static thread_local std::vector<uint32_t> k2, v2;
  k2.resize(n); v2.resize(n);
  constexpr int BITS = 11, BUCKETS = 1 << BITS;
  std::vector<size_t> hist((size_t)3 * BUCKETS, 0);
  for (size_t i = 0; i < n; i++) {
    const uint32_t k = keys[i];
    hist[k & (BUCKETS - 1)]++;
    hist[BUCKETS + ((k >> BITS) & (BUCKETS - 1))]++;
    hist[2 * BUCKETS + (k >> (2 * BITS))]++;
  }
  for (int pass = 0; pass < 3; pass++) {
    size_t *h = hist.data() + (size_t)pass * BUCKETS;
    const int shift = pass * BITS;
    bool trivial = false;
    size_t sum = 0;
    for (int d = 0; d < BUCKETS; d++) { if (h[d] == n) trivial = true; const size_t c = h[d]; h[d] = sum; sum += c; }
    if (trivial) continue;  // every key has the same digit: the pass is the identity
    const uint32_t *ks = keys.data(), *vs = vals.data();
    uint32_t *kd = k2.data(), *vd = v2.data();
    for (size_t i = 0; i < n; i++) {
      const size_t pos = h[(ks[i] >> shift) & (BUCKETS - 1)]++;
      kd[pos] = ks[i];
      vd[pos] = vs[i];
    }
    keys.swap(k2);
    vals.swap(v2);
  }
}

// Runs body(begin, end) over [0, n) on up to 8 threads (one for small n).
template <class Body>
static void parallel_ranges(size_t n, Body body) {
  unsigned hw = rs_host_cores();
  size_t nt = n < ((size_t)1 << 17) ? 1 : std::min<size_t>(8, hw ? hw : 1);
  if (nt <= 1) { body((size_t)0, n, (size_t)0); return; }
  std::vector<std::thread> th;
  const size_t per = (n + nt - 1) / nt;
  for (size_t t = 1; t < nt; t++) th.emplace_back([=, &body]() { body(std::min(n, t * per), std::min(n, (t + 1) * per), t); });
  body((size_t)0, std::min(n, per), (size_t)0);
  for (auto &x : th) x.join();
}

// collect_target_points + order_target_points for a pixmap whose number of selected pixels is already known (the device
// counted them).  For the shuffling modes the draws depend on that number and the seed only, so their producer thread
// starts before the points are collected and the two overlap.
int collect_and_order(int mode, const uint8_t *mask0, int w, int h, size_t pixel_stride, size_t row_stride, size_t n_known,
                      uint32_t seed, std::vector<uint32_t> &pts, const PairSorter *sorter) {
  if (mode < 0 || mode > 8) return IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE;
  constexpr size_t BLOCK = 1u << 16;
  if (mode > 1 || n_known < 4 * BLOCK) {
    collect_target_points_strided(mask0, w, h, pixel_stride, row_stride, pts);
    GRandMT prng(seed);
    return order_target_points(mode, pts, prng, sorter);
  }
  static thread_local std::vector<uint32_t> js;
  js.resize(n_known);
  uint32_t *jp = js.data();
  const size_t n = n_known;
  std::atomic<size_t> ready{0};
  std::thread producer([&]() {
    GRandMT prng(seed);
    for (size_t off = 0; off < n; off += BLOCK) {
      const size_t len = n - off < BLOCK ? n - off : BLOCK;
      prng.fill_int_range((uint32_t)n, jp + off, len);
      ready.store(off + len, std::memory_order_release);
    }
  });
  collect_target_points_strided(mask0, w, h, pixel_stride, row_stride, pts);
  if (pts.size() != n) { producer.join(); return -1; }  // the caller's count was wrong: nothing was swapped yet
  uint32_t *a = pts.data();
  constexpr size_t AHEAD = 64;
  for (size_t off = 0; off < n; off += BLOCK) {
    const size_t end = n - off < BLOCK ? n : off + BLOCK;
    while (ready.load(std::memory_order_acquire) < end) { /* spin: the producer is ahead after the collect */ }
    for (size_t i = off; i < end; i++) {
      if (i + AHEAD < end) __builtin_prefetch(a + jp[i + AHEAD], 1);
      std::swap(a[i], a[jp[i]]);
    }
  }
  producer.join();
  return 0;
}

int order_target_points(int mode, std::vector<uint32_t> &pts, GRandMT &prng, const PairSorter *sorter) {
  const size_t n = pts.size();
  if (mode < 0 || mode > 8) return IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE;
  if (mode <= 1) {  // not Fisher-Yates: every i swaps with a draw over the whole vector
    swaps_from_draws(pts.data(), n, [&](size_t, size_t count, uint32_t *out) { prng.fill_int_range((uint32_t)n, out, count); });
    return 0;
  }
  // centre of the bounding box; the upper bounds start at 0 as in the reference (engineTypes.h:192-226)
  int ulx = INT_MAX, uly = INT_MAX, lrx = 0, lry = 0;
  for (uint32_t p : pts) {
    const int x = unpack_x(p), y = unpack_y(p);
    ulx = std::min(ulx, x); uly = std::min(uly, y);
    lrx = std::max(lrx, x); lry = std::max(lry, y);
  }
  const int cx = (lrx - ulx) / 2 + ulx, cy = (lry - uly) / 2 + uly;
  // Sort keys as order-preserving 32-bit patterns (non-negative floats compare like their bit patterns).
  std::vector<uint32_t> keys(n), idx(n);
  const bool brush = (mode == 2 || mode == 5 || mode == 8);  // 8 ("squeeze") nets out to mode 2's sort
  bool descending;
  if (brush) {
    // the ray index needs libm's atan2 bit for bit (brushfire.h); it is the bulk of the work, so it runs on several cores
    unsigned maxray_t[8][401];
    std::memset(maxray_t, 0, sizeof maxray_t);
    std::vector<uint16_t> ray(n);
    parallel_ranges(n, [&](size_t b, size_t e, size_t t) {
      unsigned *mr = maxray_t[t];
      for (size_t i = b; i < e; i++) {
        const int ox = unpack_x(pts[i]) - cx, oy = unpack_y(pts[i]) - cy;
        const unsigned g = ray_index(ox, oy);
        ray[i] = (uint16_t)g;
        mr[g] = std::max(mr[g], (unsigned)(ox * ox + oy * oy));
      }
    });
    unsigned maxray[401];
    for (int g = 0; g < 401; g++) { unsigned m = 0; for (int t = 0; t < 8; t++) m = std::max(m, maxray_t[t][g]); maxray[g] = m; }
    parallel_ranges(n, [&](size_t b, size_t e, size_t) {
      for (size_t i = b; i < e; i++) {
        const int ox = unpack_x(pts[i]) - cx, oy = unpack_y(pts[i]) - cy;
        const float k = (float)(oy * oy + ox * ox) / maxray[ray[i]];  // NaN only when n == 1
        uint32_t bits;
        std::memcpy(&bits, &k, 4);
        keys[i] = bits;
      }
    });
    descending = (mode != 5);
  } else {
    const bool by_y = (mode == 4 || mode == 7);
    for (size_t i = 0; i < n; i++) {
      const int ox = unpack_x(pts[i]) - cx, oy = unpack_y(pts[i]) - cy;
      keys[i] = (uint32_t)(by_y ? oy * oy : ox * ox);
    }
    descending = (mode == 3 || mode == 4);
  }
  // glibc merge sort under comparators that never answer "equal" (engineTypes.h:51-56):
  // "less" kinds  -> ascending, equal keys in REVERSED input order; "more" kinds -> descending, input order kept.
  if (n > 1) {
    if (descending) {
      for (size_t i = 0; i < n; i++) { keys[i] = ~keys[i]; idx[i] = (uint32_t)i; }
    } else {
      std::reverse(keys.begin(), keys.end());
      for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)(n - 1 - i);
    }
    bool done = false;
    if (sorter && *sorter) {
      // the stable sort of (key, point) pairs is handed to the caller's sorter (the device, csrc/rs_kernels.cu): the
      // points travel as the payload, in the same (possibly reversed) input order as the keys
      std::vector<uint32_t> vals(n);
      for (size_t i = 0; i < n; i++) vals[i] = pts[idx[i]];
      if ((*sorter)(keys.data(), vals.data(), n)) { pts.swap(vals); done = true; }  // (keys and vals are untouched on failure)
    }
    if (!done) {
      radix_sort_pairs(keys, idx);
      std::vector<uint32_t> sorted(n);
      for (size_t i = 0; i < n; i++) sorted[i] = pts[idx[i]];
      pts.swap(sorted);
    }
  }
  shuffle_bands(pts, prng);
  return 0;
}

uint32_t pass_schedule(uint32_t n, uint32_t ends[6]) {
  uint32_t total = n;
  ends[0] = n;
  for (int p = 1; p < 6; p++) { ends[p] = n; total += n; n = n * 3 / 4; }
  return total;
}

}  // namespace rs
