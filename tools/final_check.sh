out=gpurun_out; tag=r01j
timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > $out/${tag}_pytest_gpu.txt
timeout 80 python bench.py > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 30 python tools/pass_timeline.py --workload cfg2 --out $out/${tag}_timeline_cfg2.txt > /dev/null 2>&1
timeout 40 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg4.json 2>/dev/null
cat $out/${tag}_pytest_gpu.txt; cut -c1-400 $out/${tag}_bench_cfg2.json
