#!/bin/bash
# Round profile artifacts, one gpurun call: bench lines of every BASELINE config, pass timelines, ncu launch list.
# Usage (from the repo root, on a GPU box):  bash tools/round_profile.sh <tag>     -> gpurun_out/<tag>_*
# Every step runs under its own timeout: a step that hangs costs its limit, not the whole call.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 300 python bench.py > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err
for w in cfg1 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 2 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
timeout 200 python bench.py --workload cfg5 --batch 64 --slots 8 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_cfg5_batch64.json 2> $out/${tag}_bench_cfg5_batch64.err
timeout 200 python bench.py --workload cfg1 --batch 64 --slots 8 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_cfg1_batch64.json 2> $out/${tag}_bench_cfg1_batch64.err
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python tools/pass_timeline.py --workload $w --out $out/${tag}_timeline_$w.txt > /dev/null 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
for p in 50 100 200 500 1000; do timeout 150 python bench.py --workload cfg5 --probes $p --steps 5 --warmup 2 --no-cpu-baseline > $out/${tag}_bench_cfg5_probes$p.json 2>/dev/null; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -c 300 $out/${tag}_bench_*.err
