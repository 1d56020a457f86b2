#!/usr/bin/env bash
# Profile visit: launch list of one bench step + full ncu captures of the dominant kernels.
# usage: scripts/gpu_profile.sh <tag> [batch]
# The .ncu-rep files are exported to CSV on the box and deleted when large (gpurun_out/ is capped at 64 MiB).
TAG="${1:-r}"
B="${2:-4096}"
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
BENCH="python bench.py --steps 1 --warmup 3 --batch $B --profile-only"
timeout 300 python bench.py --steps 3 --warmup 3 --batch $B --profile-only > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_$TAG.log; exit 1; }
tail -1 $OUT/plain_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH > $OUT/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
# full captures: every weight-gradient launch of the 4th (timed) step, then a sample of conv / ConvLSTM-step launches
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 -c $4 \
      -f -o $OUT/prof_${TAG}_$1 $BENCH > $OUT/ncu_full_${TAG}_$1.log 2>&1; echo "full $1 rc=$?"
  ncu -i $OUT/prof_${TAG}_$1.ncu-rep --page raw --csv > $OUT/prof_${TAG}_$1_raw.csv 2>/dev/null
  ls -la $OUT/prof_${TAG}_$1.ncu-rep
  sz=$(stat -c %s $OUT/prof_${TAG}_$1.ncu-rep)
  [ "$sz" -gt 16000000 ] && rm -f $OUT/prof_${TAG}_$1.ncu-rep
}
# one timed step (the 4th) of: the three persistent / fused ConvLSTM kernels per layer, the dense GEMMs, the fc-LSTM
cap convlstm 'tc_wgrad_rows|convlstm_seq' ${CL_SKIP:-27} ${CL_N:-9}
[ -n "$ONLY_CONVLSTM" ] || cap conv 'tc_conv_kernel|tc_wgrad_kernel' ${CV_SKIP:-36} ${CV_N:-12}
[ -n "$ONLY_CONVLSTM" ] || cap lstm 'lstm_' ${LS_SKIP:-6} 2
du -sh $OUT
