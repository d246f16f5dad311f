#!/bin/bash
# Evidence at the final commit of the round: smoke, GPU suite, bench line, launch list of the headline, call phases,
# timeline.  bash tools/final_evidence.sh <tag>  -> gpurun_out/<tag>_*
tag=${1:-final}; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
tail -1 $out/${tag}_smoke.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/${tag}_pytest_gpu.txt; cat $out/${tag}_pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err
python tools/bench_table.py $out/${tag}_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 120 python tools/phase_times.py cfg3 cfg2 cfg5 > $out/${tag}_call_phases.txt 2>&1
timeout 120 python tools/pass_timeline.py --workload cfg3 --out $out/${tag}_timeline_cfg3.txt > /dev/null 2>&1
ls $out | grep -c $tag
