#!/bin/bash
# Round 2, call F: windowed corpus-point samples (suite + timings), L2 traffic of cfg2 with the corpus on chip vs in L2.
out=gpurun_out; tag=r02f
mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
timeout 240 python tools/quick.py cfg3 cfg5 cfg1 heal:1024:512 cfg4 > $out/${tag}_quick_default.txt 2>&1
RS_SELECT_MIN=0 timeout 240 python tools/quick.py cfg5 cfg1 heal:1024:512 > $out/${tag}_quick_select0.txt 2>&1
cat $out/${tag}_quick_default.txt $out/${tag}_quick_select0.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg2_pass1_smemc python tools/ncu_job.py --workload cfg2 --jobs 2 > $out/${tag}_ncu_cfg2a.log 2>&1
RS_SMEM_CORPUS=0 timeout 200 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg2_pass1_l2 python tools/ncu_job.py --workload cfg2 --jobs 2 > $out/${tag}_ncu_cfg2b.log 2>&1
timeout 300 $NCU -k k_synth_pass -s 7 -c 1 -f -o $out/${tag}_cfg3_pass1 python tools/ncu_job.py --workload cfg3 --jobs 2 > $out/${tag}_ncu_cfg3.log 2>&1
ls -la $out | grep $tag
