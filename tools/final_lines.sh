#!/bin/bash
# Bench line + GPU test tail at the final commit of a round (kernels unchanged since the last round_profile.sh): <tag>
tag=${1:-final}; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/${tag}_pytest_gpu.txt; cat $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err
cut -c1-700 $out/${tag}_bench.json
