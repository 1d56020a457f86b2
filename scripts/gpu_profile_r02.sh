#!/usr/bin/env bash
# Round-2 profile visit: launch lists (gpu__time_duration) and `ncu --set full` captures of the kernels that dominate
# each benchmarked configuration.  Every ncu command runs only after the same command line exited 0 without ncu.
# Output: gpurun_out/r02_*.  usage: scripts/gpu_profile_r02.sh [what ...]   (what: m3 m4 m1; default all)
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
WHAT="${*:-m3 m4 m1}"
raw() {  # rep -> raw csv; drop big reports (gpurun_out/ is capped at 64 MiB)
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  sz=$(stat -c %s $OUT/$1.ncu-rep 2>/dev/null || echo 0)
  [ "$sz" -gt 12000000 ] && rm -f $OUT/$1.ncu-rep
}
launches() {  # tag cmd...
  tag=$1; shift
  "$@" > $OUT/r02_plain_$tag.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r02_launches_$tag.csv \
      "$@" > $OUT/r02_ncu_launch_$tag.log 2>&1
  echo "launch list $tag rc=$?"
}
full() {  # tag regex skip count cmd...
  tag=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $OUT/r02_plain_full_$tag.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt \
      -f -o $OUT/r02_full_$tag "$@" > $OUT/r02_ncu_full_$tag.log 2>&1
  echo "full $tag rc=$?"
  raw r02_full_$tag
}
for w in $WHAT; do
  case $w in
    m3)
      M3="python bench.py --steps 1 --warmup 3 --batch 8880 --profile-only"
      launches m3_b8880_bf16x2 $M3
      # the 4th (timed) step: 3 fwd + 3 bwd + 3 wgrad_rows ConvLSTM kernels, then the dense conv / wgrad launches
      full m3_convlstm 'tc_wgrad_rows|convlstm_seq' 27 9 $M3
      full m3_dense 'tc_conv_kernel|tc_wgrad_kernel' 36 12 $M3
      full m3_lstm 'lstm_' 6 2 $M3
      ;;
    m4)
      launches m4_b32_bf16_train python scripts/profile_step.py m4 32 bf16 train 2
      launches m4_b32_bf16_infer python scripts/profile_step.py m4 32 bf16 infer 2
      # second step of the training run: head convs (fwd / bwd-data), TMA-fed weight gradients, general weight gradients
      full m4_conv 'tc_conv_kernel' 244 12 python scripts/profile_step.py m4 32 bf16 train 2
      full m4_wgrad_planes 'wgrad_planes|planes_kernel' 30 9 python scripts/profile_step.py m4 32 bf16 train 2
      ;;
    m4conv)
      full m4_conv 'tc_conv_kernel' 244 12 python scripts/profile_step.py m4 32 bf16 train 2
      ;;
    m1)
      launches m1_b8192_bf16x2_train python scripts/profile_step.py m1 8192 bf16x2 train 3
      full m1_xproj_lstm 'xproj|lstm_tc_fwd' 4 2 python scripts/profile_step.py m1 8192 bf16x2 train 3
      ;;
  esac
done
du -sh $OUT
