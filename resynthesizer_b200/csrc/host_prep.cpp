#include "host_prep.h"

#include <algorithm>
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

uint32_t GRandMT::next32() {
  if (mti_ >= 624) {
    for (int k = 0; k < 624; k++) {
      const uint32_t y = (mt_[k] & 0x80000000u) | (mt_[(k + 1) % 624] & 0x7fffffffu);
      mt_[k] = mt_[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    mti_ = 0;
  }
  uint32_t y = mt_[mti_++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
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
void collect_target_points(const uint8_t *pix, int w, int h, int bpp, std::vector<Point> &out) {
  out.clear();
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; i++)
    if (pix[i * bpp] != 0) out.push_back(Point{(int)(i % w), (int)(i / w)});
}

void collect_corpus_points(const uint8_t *pix, int w, int h, int bpp, const TFormatIndices &fi, std::vector<Point> &out) {
  out.clear();
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; i++) {
    const uint8_t *p = pix + i * bpp;
    if (p[0] == 0xFF && (fi.isAlphaSource ? p[fi.alpha_bip] != 0 : true)) out.push_back(Point{(int)(i % w), (int)(i / w)});
  }
}

// ---------------------------------------------------------------------------------------- ordering
static void shuffle_bands(std::vector<Point> &p, GRandMT &prng) {
  const int last = (int)p.size() - 1;
  const int half = (int)(p.size() * 0.1);  // IMAGE_SYNTH_BAND_FRACTION
  for (int i = 0; i <= last; i++) {
    const int lo = std::max(i - half, 0), hi = std::min(i + half, last);
    const int j = lo + (int)prng.int_range((uint32_t)(hi - lo));
    std::swap(p[i], p[j]);
  }
}

static unsigned ray_index(const Point &a) {
  return (unsigned)(std::atan2((double)(float)a.y, (double)(float)a.x) * 200 /
                        3.1415926535897932384626433832795028841971693993751 + 200);
}

int order_target_points(int mode, std::vector<Point> &pts, GRandMT &prng) {
  const size_t n = pts.size();
  if (mode < 0 || mode > 8) return IMAGE_SYNTH_ERROR_MATCH_CONTEXT_TYPE_RANGE;
  if (mode <= 1) {  // not Fisher-Yates: every i swaps with a draw over the whole vector
    for (size_t i = 0; i < n; i++) std::swap(pts[i], pts[prng.int_range((uint32_t)n)]);
    return 0;
  }
  // centre of the bounding box; the upper bounds start at 0 as in the reference (engineTypes.h:192-226)
  int ulx = INT_MAX, uly = INT_MAX, lrx = 0, lry = 0;
  for (const Point &p : pts) {
    ulx = std::min(ulx, p.x); uly = std::min(uly, p.y);
    lrx = std::max(lrx, p.x); lry = std::max(lry, p.y);
  }
  const Point c{(lrx - ulx) / 2 + ulx, (lry - uly) / 2 + uly};
  struct Keyed { float fkey; int ikey; uint32_t orig; };
  std::vector<Keyed> keys(n);
  std::vector<Point> off(n);
  for (size_t i = 0; i < n; i++) { off[i] = Point{pts[i].x - c.x, pts[i].y - c.y}; keys[i].orig = (uint32_t)i; keys[i].fkey = 0.f; keys[i].ikey = 0; }
  const bool brush = (mode == 2 || mode == 5 || mode == 8);  // 8 ("squeeze") nets out to mode 2's sort
  bool descending;
  if (brush) {
    unsigned maxray[401];
    std::memset(maxray, 0, sizeof maxray);
    for (size_t i = 0; i < n; i++) {
      const unsigned d = (unsigned)(off[i].x * off[i].x + off[i].y * off[i].y), g = ray_index(off[i]);
      maxray[g] = std::max(maxray[g], d);
    }
    for (size_t i = 0; i < n; i++)
      keys[i].fkey = (float)(off[i].y * off[i].y + off[i].x * off[i].x) / maxray[ray_index(off[i])];
    descending = (mode != 5);
  } else {
    const bool by_y = (mode == 4 || mode == 7);
    for (size_t i = 0; i < n; i++) keys[i].ikey = by_y ? off[i].y * off[i].y : off[i].x * off[i].x;
    descending = (mode == 3 || mode == 4);
  }
  // glibc merge sort under comparators that never answer "equal" (engineTypes.h:51-56):
  // "less" kinds  -> ascending, equal keys in REVERSED input order; "more" kinds -> descending, input order kept.
  if (n > 1) {
    if (descending) {
      if (brush) std::stable_sort(keys.begin(), keys.end(), [](const Keyed &a, const Keyed &b) { return a.fkey > b.fkey; });
      else std::stable_sort(keys.begin(), keys.end(), [](const Keyed &a, const Keyed &b) { return a.ikey > b.ikey; });
    } else {
      std::reverse(keys.begin(), keys.end());
      if (brush) std::stable_sort(keys.begin(), keys.end(), [](const Keyed &a, const Keyed &b) { return a.fkey < b.fkey; });
      else std::stable_sort(keys.begin(), keys.end(), [](const Keyed &a, const Keyed &b) { return a.ikey < b.ikey; });
    }
  }
  for (size_t i = 0; i < n; i++) pts[i] = Point{off[keys[i].orig].x + c.x, off[keys[i].orig].y + c.y};
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
