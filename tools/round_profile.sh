#!/bin/bash
# Round profile artifacts, one gpurun call (1 GPU): the bench line (all BASELINE configurations ride along as sub-records),
# the reference arm, DRAM traffic of every configuration, ncu counters of the dominant kernels, launch list, timelines,
# sanitizer, GPU test tail.  Usage (repo root, on a GPU box):  bash tools/round_profile.sh <tag>   -> gpurun_out/<tag>_*
# Every step runs under its own timeout: a step that hangs costs its limit, not the whole call.
tag=${1:-r03m}
out=gpurun_out
mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/${tag}_pytest_gpu.txt
cat $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for w in cfg3 cfg2 cfg5 cfg4 cfg1; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:'k_synth_pass|k_gather|k_ctx' --csv --log-file $out/${tag}_traffic_$w.csv python tools/ncu_job.py --workload $w --jobs 2 > $out/${tag}_traffic_$w.log 2>&1
done
NCU="ncu --set full --clock-control none --import-source on"
# k_synth_pass launches per job: cfg3 6 (one per pass); cfg2 and cfg4 7 (pass 0 = one visit per warp up to visit 262144, two from
# there on; then one launch per pass): -s skips to pass 1 of the second job
timeout 300 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg3_pass1 python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_ncu_cfg3.log 2>&1
timeout 200 $NCU -k k_synth_pass -s 9 -c 1 -f -o $out/${tag}_cfg2_pass1 python tools/ncu_job.py --workload cfg2 --jobs 2 > $out/${tag}_ncu_cfg2.log 2>&1
timeout 300 $NCU -k k_synth_pass -s 9 -c 1 -f -o $out/${tag}_cfg4_pass1 python tools/ncu_job.py --workload cfg4 --jobs 2 > $out/${tag}_ncu_cfg4.log 2>&1
timeout 200 $NCU -k k_synth_pass_team -s 6 -c 2 -f -o $out/${tag}_cfg5_team python tools/ncu_job.py --workload cfg5 --jobs 2 > $out/${tag}_ncu_cfg5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python tools/pass_timeline.py --workload $w --out $out/${tag}_timeline_$w.txt > /dev/null 2>&1; done
timeout 300 python tools/quality_table.py --out $out/${tag}_quality > /dev/null 2> $out/${tag}_quality.err
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitizer_jobs.py 2>&1 | tail -8; RS_TEAM_P0=1 RS_TEAM_PN=1 RS_SMEM_CORPUS=2 RS_SPARSE_PROBE=0 timeout 400 compute-sanitizer --tool memcheck python tools/sanitizer_jobs.py 2>&1 | tail -6; RS_TEAM_P0=1 RS_TEAM_PN=1 RS_SELECT_MIN=0 RS_SPARSE_PROBE=0 timeout 400 compute-sanitizer --tool racecheck python tools/sanitizer_jobs.py 2>&1 | tail -6; RS_TEAM_P0=1 RS_TEAM_PN=1 RS_PAIR=2 timeout 300 compute-sanitizer --tool synccheck python tools/sanitizer_jobs.py 2>&1 | tail -4 ) > $out/${tag}_sanitizer.txt 2>&1
tail -20 $out/${tag}_sanitizer.txt
ls $out | grep $tag | wc -l
