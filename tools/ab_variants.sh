#!/bin/bash
# A/B of library variants in ONE call (same box, same clocks): bash tools/ab_variants.sh <tag> <variant suffixes...>
tag=$1; shift; out=gpurun_out; mkdir -p $out
for rep in 1 2; do for v in "$@"; do
  [ "$v" = "-" ] && v=""
  echo "== variant '$v' rep $rep"
  RS_LIB_VARIANT=$v timeout 300 python tools/quick.py cfg3 cfg4 cfg2 cfg1 cfg5 heal:1024:512 2>&1 | cut -c1-105
done; done > $out/${tag}_ab.txt 2>&1
cat $out/${tag}_ab.txt
