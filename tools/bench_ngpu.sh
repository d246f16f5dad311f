#!/bin/bash
# bench.py under torchrun on all GPUs of the box: bash tools/bench_ngpu.sh <tag> <n>
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l > $out/${tag}_ngpus.txt
nproc >> $out/${tag}_ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
tail -c 300 $out/${tag}_bench_${n}gpu.err; tail -c 1500 $out/${tag}_bench_${n}gpu.json
