#!/bin/bash
# 8-GPU evidence (run under `gpurun --gpus 8`): config 4 (ResNet-18 b256/GPU) with the fused NVLink exchange and with NCCL,
# config 5 (ResNet-50, 16-bit gradients, b128/GPU), and the real-peer correctness test of the exchange.
T=${1:-r2_n8}
O=gpurun_out
mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -x -q > $O/${T}_dp_tests.log 2>&1; echo "dp tests rc $?"; tail -2 $O/${T}_dp_tests.log
timeout 400 $R --master-port 29511 bench.py --gpus 8 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet18_fused.json 2> $O/${T}_resnet18_fused.err; echo "r18 fused rc $?"
LBT_DP=nccl timeout 400 $R --master-port 29512 bench.py --gpus 8 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet18_nccl.json 2> $O/${T}_resnet18_nccl.err; echo "r18 nccl rc $?"
timeout 400 $R --master-port 29513 bench.py --gpus 8 --workload resnet50 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet50_g16.json 2> $O/${T}_resnet50_g16.err; echo "r50 rc $?"
for f in $O/${T}_bench_*.json; do cut -c1-260 $f; echo; done
for f in $O/${T}_*.err; do tail -n 2 $f; done
