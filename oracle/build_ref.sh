#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle/).  Compiles the UNMODIFIED reference library
# from the sources where they lie under /root/reference into oracle/_ref/*.so.
#
# No reference source is copied into the repo: a scratch directory of symlinks
# is made under $TMPDIR, with exactly one generated file -- buildSwitches.h,
# the compile-switch header the reference tells its users to edit
# (lib/buildSwitches.h:34-38, lib/imageSynth.c:17-19) -- produced by sed.
#
#   libref_mt_1t.so    unthreaded, GLib-GRand-compatible MT19937 (grand_mt19937.c)
#                      -> the golden-pinned oracle
#   libref_rand_1t.so  unthreaded, libc rand() proxy exactly as shipped
#   libref_rand_8t.so  the reference's threaded refiner (THREAD_LIMIT 8, pthreads)
#
# Flags follow the reference's autotools default (-O2, no -ffast-math).
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
HERE=$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/lib" ]; then
  echo "build_ref.sh: $REF/lib not present; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT

mk_tree() { # $1 = dir, $2 = threaded? (0/1)
  mkdir -p "$1"
  for f in "$REF"/lib/*.c "$REF"/lib/*.h; do
    ln -s "$f" "$1/$(basename "$f")"
  done
  rm "$1/buildSwitches.h"
  if [ "$2" = 1 ]; then
    sed -e 's|^#define SYNTH_USE_GLIB_THREADS|// &|' "$REF/lib/buildSwitches.h" > "$1/buildSwitches.h"
  else
    sed -e 's|^#define SYNTH_USE_GLIB_THREADS|// &|' \
        -e 's|^#define SYNTH_THREADED TRUE|// &|' "$REF/lib/buildSwitches.h" > "$1/buildSwitches.h"
  fi
}
SRCS="imageSynth.c engine.c glibProxy.c engineParams.c imageFormat.c progress.c"
CFLAGS="-DSYNTH_LIB_ALONE -O2 -std=gnu99 -fPIC -w"

mk_tree "$TMP/t1" 0
mk_tree "$TMP/t8" 1

( cd "$TMP/t1" && gcc $CFLAGS -shared -o "$OUT/libref_rand_1t.so" $SRCS -lm )
( cd "$TMP/t8" && gcc $CFLAGS -pthread -shared -o "$OUT/libref_rand_8t.so" $SRCS -lm -lpthread )
# MT variant: rename the proxy's two PRNG functions away, link the GRand shim.
( cd "$TMP/t1" && for s in $SRCS; do
    extra=""
    [ "$s" = glibProxy.c ] && extra="-Ds_rand_new_with_seed=unused_proxy_rand_new -Ds_rand_int_range=unused_proxy_rand_int_range"
    gcc $CFLAGS $extra -c "$s" -o "${s%.c}.o"
  done
  gcc $CFLAGS -c "$HERE/grand_mt19937.c" -o grand_mt19937.o
  gcc -shared -o "$OUT/libref_mt_1t.so" *.o -lm )
# The reference's own test harness (src/testSynth.c, minus its duplicate include at line 9), compiled unchanged
# and linked against THIS repo's library instead of the reference objects: the drop-in check of INTEGRATION.md 3.1.
PROD="$HERE/../resynthesizer_b200/lib"
if [ -f "$PROD/libresynthesizer_b200.so" ]; then
  sed '9d' "$REF/src/testSynth.c" > "$TMP/testSynth.c"
  gcc -DSYNTH_LIB_ALONE -w -I "$TMP/t1" "$TMP/testSynth.c" -o "$OUT/testSynth_b200" \
      -L "$PROD" -lresynthesizer_b200 -Wl,-rpath,'$ORIGIN/../../resynthesizer_b200/lib'
  ( cd "$TMP/t1" && gcc $CFLAGS -I "$TMP/t1" -o "$OUT/testSynth_ref" "$TMP/testSynth.c" $SRCS -lm )
fi
echo "built: $(ls "$OUT")"
