#!/bin/bash
# Round 2, call B: new bench contract (cfg3 headline, sub-records, cfg5 batch), reference arm, boundary tests.
out=gpurun_out; tag=r02b
mkdir -p $out
nproc > $out/${tag}_nproc.txt
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_full_size.py tests/test_gpu_order_cache.py -m gpu -q -x 2>&1 | tail -15 > $out/${tag}_pytest.txt
tail -3 $out/${tag}_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -c 300 $out/${tag}_bench_reference.err
for s in 1 2 3 4; do
  timeout 300 python bench.py --quick --steps 1 --warmup 0 --no-cpu-baseline --slots $s > /dev/null 2>&1
done
ls -la $out | tail -8
