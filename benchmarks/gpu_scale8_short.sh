#!/bin/bash
# 8-GPU lines of configs 4 and 5 with the fused exchange (run under `gpurun --gpus 8`)
T=${1:-r2_n8b}
O=gpurun_out
mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29521 bench.py --gpus 8 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet18_fused.json 2> $O/${T}_resnet18_fused.err; echo "r18 rc $?"
timeout 300 $R --master-port 29523 bench.py --gpus 8 --workload resnet50 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet50_g16.json 2> $O/${T}_resnet50_g16.err; echo "r50 rc $?"
for f in $O/${T}_bench_*.json; do cut -c1-260 $f; echo; done
for f in $O/${T}_*.err; do tail -n 2 $f; done
