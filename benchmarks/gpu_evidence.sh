#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, the default bench line, the other workloads, an ncu launch list
# and targeted `ncu --set full` captures of kernels of ONE step of the default bench command (never a bench value).
# Reports stay in /tmp on the box (gpurun_out/ is capped at 64 MiB): only their CSV pages come back.
T=${1:-r2_s4}
O=gpurun_out
mkdir -p $O
if [ "$2" != "notests" ]; then
python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?"; tail -2 $O/${T}_tests.log
fi
python bench.py > $O/${T}_bench_resnet18.json 2> $O/${T}_bench_resnet18.err; echo "bench rc $?"
python bench.py --workload resnet20 --no-micro > $O/${T}_bench_resnet20.json 2> $O/${T}_bench_resnet20.err
python bench.py --workload resnet50 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet50_g16.json 2> $O/${T}_bench_resnet50_g16.err
python bench.py --workload resnet50_g8 --no-micro --no-cpu-baseline > $O/${T}_bench_resnet50_g8.json 2> $O/${T}_bench_resnet50_g8.err
python bench.py --no-micro --no-cpu-baseline --dump-launches $O/${T}_launch_times_resnet18.csv > /dev/null 2>&1
export LBT_PROFILE_REGION=1
B="python bench.py --steps 1 --warmup 3 --no-micro --no-cpu-baseline"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${T}_launches_resnet18.csv $B > $O/${T}_ncu_list.log 2>&1
echo "ncu list rc $?"
cap() {  # name, regex, skip, count
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 \
      -o /tmp/${T}_$1 -f $B > $O/${T}_ncu_$1.log 2>&1
  echo "ncu $1 rc $?"
  ncu -i /tmp/${T}_$1.ncu-rep --page raw --csv > $O/${T}_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${T}_ncu_$1_source.csv.gz
  ls -la /tmp/${T}_$1.ncu-rep
}
cap fprop "conv_halo|conv_ldg_kernel|conv_fprop" 0 45
cap bnfwd "bn_fwd2|bn_bwd2|bn_bwd1" 3 6
cap bnbwd "bn_fwd2|bn_bwd2|bn_bwd1" 48 12
cap wgrad "wgrad" 14 6
cap misc "maxpool|param_prep|xent_fwd|finalize|dp_step|quantize_|stem_pack8" 0 14
du -sh $O
