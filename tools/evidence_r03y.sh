#!/bin/bash
# What changed with the device-side PRNG stream / side-stream order pipeline / direct copies: sanitizer on the ordering
# paths, launch list of the headline, call phases, timeline.  bash tools/evidence_r03y.sh <tag>
tag=${1:-r03y}; out=gpurun_out; mkdir -p $out
( timeout 200 compute-sanitizer --tool memcheck python tools/sanitizer_jobs.py 2>&1 | tail -6; timeout 200 compute-sanitizer --tool racecheck python tools/sanitizer_jobs.py 2>&1 | tail -6 ) > $out/${tag}_sanitizer.txt 2>&1
tail -12 $out/${tag}_sanitizer.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
grep -c k_ $out/${tag}_launches_cfg3.csv
timeout 120 python tools/phase_times.py cfg3 cfg2 cfg5 > $out/${tag}_call_phases.txt 2>&1; cat $out/${tag}_call_phases.txt
timeout 120 python tools/pass_timeline.py --workload cfg3 --out $out/${tag}_timeline_cfg3.txt > /dev/null 2>&1
