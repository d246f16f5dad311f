#!/bin/bash
# Sweep of pass-0 segment plans (RS_SEG_P0 = "end:width,..."): kernel ms and pass-0 ms per plan and workload.
# Usage: bash tools/seg_sweep.sh "<workloads>" "<plan>" "<plan>" ...
wl=$1; shift
for plan in "$@"; do
  echo "== $plan"; RS_SEG_P0=$plan timeout 90 python tools/quick.py $wl 2>&1 | awk '{print $1,$3,$6}'
done
