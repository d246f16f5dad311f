#!/bin/bash
# Bench line + GPU test tail at the final commit of a round, plus what changed since the last round_profile.sh for the
# headline kernel (ncu counters, DRAM traffic and launch list of cfg3): <tag>
tag=${1:-final}; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/${tag}_pytest_gpu.txt; cat $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err
cut -c1-700 $out/${tag}_bench.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 200 ncu --metrics $M --clock-control none -k regex:'k_synth_pass|k_gather|k_ctx' --csv --log-file $out/${tag}_traffic_cfg3.csv python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_traffic_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg3_pass1 python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_ncu_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
for w in cfg3 cfg2; do timeout 120 python tools/pass_timeline.py --workload $w --out $out/${tag}_timeline_$w.txt > /dev/null 2>&1; done
ls $out | grep $tag | wc -l
