#!/bin/bash
# Round 2, call A: the widened GPU parity suite on the round-1 kernels + ncu captures of both pass kernels (baseline).
out=gpurun_out; tag=r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt
nproc >> $out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
cat $out/${tag}_pytest_gpu.txt | tail -5
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg2_pass1 python tools/ncu_job.py --workload cfg2 --jobs 2 > $out/${tag}_ncu_cfg2.log 2>&1
timeout 300 $NCU -k k_synth_pass_team -s 6 -c 2 -f -o $out/${tag}_cfg5_team python tools/ncu_job.py --workload cfg5 --jobs 2 > $out/${tag}_ncu_cfg5.log 2>&1
timeout 400 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg3_pass1 python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_ncu_cfg3.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python tools/pass_timeline.py --workload $w --out $out/${tag}_timeline_$w.txt > /dev/null 2>&1; done
tail -3 $out/${tag}_ncu_*.log
ls -la $out | tail -20
