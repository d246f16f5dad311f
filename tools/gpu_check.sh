#!/bin/bash
# One short GPU call after a kernel change: smoke, the whole GPU suite, kernel time of every BASELINE configuration.
# Usage: bash tools/gpu_check.sh <tag> [ENV=VALUE for a second timing run ...]   -> gpurun_out/<tag>_*
tag=${1:-check}; shift; out=gpurun_out
mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
timeout 300 python tools/quick.py cfg3 cfg4 cfg2 cfg1 cfg5 heal:1024:512 > $out/${tag}_quick.txt 2>&1
cat $out/${tag}_quick.txt
for kv in "$@"; do
  echo "== $kv"
  env $kv timeout 300 python tools/quick.py cfg3 cfg4 cfg2 cfg1 cfg5 heal:1024:512 2>&1 | tee $out/${tag}_quick_${kv//[^A-Za-z0-9]/_}.txt
done
