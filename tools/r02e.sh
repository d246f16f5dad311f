#!/bin/bash
# Round 2, call E (2 GPUs): multi-device dealer tests, bench under torchrun at N=2, in-process dealer timing.
out=gpurun_out; tag=r02e
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_smi.txt
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 400 python -m pytest tests/test_gpu_batch.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 | tail -15 > $out/${tag}_pytest.txt
tail -3 $out/${tag}_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err
tail -c 500 $out/${tag}_bench_2gpu.err
timeout 300 python tools/batch_multi.py > $out/${tag}_batch_multi.txt 2>&1
cat $out/${tag}_batch_multi.txt
