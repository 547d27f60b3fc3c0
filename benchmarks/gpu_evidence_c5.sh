#!/bin/bash
# Config 5 (ResNet-50, 16-bit gradients): `ncu --set full` of the dual-accumulator kernels inside ONE step of the bench command.
T=${1:-r2_c5}
O=gpurun_out
mkdir -p $O
export LBT_PROFILE_REGION=1
B="python bench.py --workload resnet50 --steps 1 --warmup 3 --no-micro --no-cpu-baseline"
cap() {  # name, regex, skip, count
  timeout 420 ncu --profile-from-start off --set full --clock-control none -k "regex:$2" -s $3 -c $4 -o /tmp/${T}_$1 -f $B > $O/${T}_ncu_$1.log 2>&1
  echo "ncu $1 rc $?"
  ncu -i /tmp/${T}_$1.ncu-rep --page raw --csv > $O/${T}_ncu_$1_raw.csv 2>/dev/null
}
cap dual "conv_wgrad_kernel|conv_fprop_kernel|conv_halo_kernel|gemm_i8_kernel" 120 16
du -sh $O
