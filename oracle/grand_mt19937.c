/*
 * TEST INFRASTRUCTURE (oracle/). Not part of the product path.
 *
 * GLib GRand stand-in for the compiled reference (oracle/_ref/libref_mt_1t.so).
 *
 * The reference's real product build draws from GLib's GRand
 * (engine.c:441, engine.c:643, orderTarget.h:44,76 under SYNTH_USE_GLIB);
 * GLib is a third-party dependency that is absent from /root/reference and
 * from this image (unpinned: "gimp-2.0 >= 2.2.0", configure.ac:57-61).
 * Its published algorithm is MT19937 (Matsumoto & Nishimura 1998) with the
 * GLib >= 2.2 seeding and g_rand_int_range rules restated below.  Parity is
 * anchored on the reference's own golden images (Test/reference_out_images),
 * which this generator reproduces bit-exactly (tests/test_oracle_goldens.py).
 *
 * The standalone proxy (glibProxy.c:29-49) is compiled with its two PRNG
 * functions renamed away (see build_ref.sh) and these definitions linked in.
 */
#include <stdint.h>
#include <stdlib.h>

#define MT_N 624
#define MT_M 397

typedef struct {
  uint32_t mt[MT_N];
  int mti;
} RefGRand;

/* 0 = use the seed the engine passes (1198472, engine.c:643). */
unsigned int ref_seed_override = 0;
/* number of raw 32-bit draws since the last seeding (test instrumentation) */
unsigned long long ref_draw_count = 0;

static RefGRand the_rand;

static void mt_seed(RefGRand *r, uint32_t s)
{
  r->mt[0] = s;
  for (int i = 1; i < MT_N; i++)
    r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = MT_N;
}

static uint32_t mt_next(RefGRand *r)
{
  static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
  uint32_t y;
  if (r->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (r->mt[MT_N - 1] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
    r->mt[MT_N - 1] = r->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    r->mti = 0;
  }
  y = r->mt[r->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  ref_draw_count++;
  return y;
}

/* replaces s_rand_new_with_seed (glibProxy.c:29-34) */
void *s_rand_new_with_seed(unsigned int seed)
{
  mt_seed(&the_rand, ref_seed_override ? ref_seed_override : seed);
  ref_draw_count = 0;
  return (void *)&the_rand;
}

/* replaces s_rand_int_range (glibProxy.c:36-49) with g_rand_int_range rules */
unsigned int s_rand_int_range(void *prng, unsigned int begin, unsigned int end)
{
  RefGRand *r = prng ? (RefGRand *)prng : &the_rand;
  uint32_t dist = end - begin;
  uint32_t v;
  if (dist == 0) return begin;
  if (dist <= 0x80000000u) {
    uint32_t leftover = (0x80000000u % dist) * 2u;
    if (leftover >= dist) leftover -= dist;
    uint32_t maxvalue = 0xffffffffu - leftover;
    do { v = mt_next(r); } while (v > maxvalue);
  } else {
    do { v = mt_next(r); } while (v >= dist); /* not reachable from the engine */
  }
  return begin + v % dist;
}
