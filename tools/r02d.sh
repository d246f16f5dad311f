#!/bin/bash
# Round 2, call D: full GPU suite, bench line, DRAM traffic of every BASELINE configuration, ncu captures of the dominant kernels.
out=gpurun_out; tag=r02d
mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
tail -1 $out/${tag}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
timeout 120 python tools/quick.py cfg2 heal:1024:512 > $out/${tag}_quick_smemc_on.txt 2>&1
RS_SMEM_CORPUS=0 timeout 120 python tools/quick.py cfg2 heal:1024:512 > $out/${tag}_quick_smemc_off.txt 2>&1
cat $out/${tag}_quick_smemc_on.txt $out/${tag}_quick_smemc_off.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 400 $out/${tag}_bench.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for w in cfg3 cfg2 cfg5 cfg4 cfg1; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:'k_synth_pass|k_gather' --csv --log-file $out/${tag}_traffic_$w.csv python tools/ncu_job.py --workload $w --jobs 2 > $out/${tag}_traffic_$w.log 2>&1
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg3_pass1 python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_ncu_cfg3.log 2>&1
timeout 200 $NCU -k k_synth_pass -s 3 -c 1 -f -o $out/${tag}_cfg2_pass1 python tools/ncu_job.py --workload cfg2 --jobs 2 > $out/${tag}_ncu_cfg2.log 2>&1
timeout 200 $NCU -k k_synth_pass_team -s 6 -c 2 -f -o $out/${tag}_cfg5_team python tools/ncu_job.py --workload cfg5 --jobs 2 > $out/${tag}_ncu_cfg5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
ls -la $out | grep $tag | head -40
