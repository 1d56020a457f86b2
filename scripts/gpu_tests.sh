#!/usr/bin/env bash
# Run the GPU parity suite group by group (a faulting kernel poisons its own process only).
# usage: scripts/gpu_tests.sh [outdir]
OUT="${1:-gpurun_out}"
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
rc=0
for grp in featuriser "window or whole_span or one_hot or hit_rate" "lstm or encoder_decoder" conv2d convlstm "losses or adam" m3 m4 fit; do
  tag=$(echo "$grp" | tr ' ' '_')
  CUDA_LAUNCH_BLOCKING=${CUDA_LAUNCH_BLOCKING:-0} timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "$grp" \
      > "$OUT/test_$tag.log" 2>&1
  r=$?
  echo "group [$grp] exit $r: $(tail -1 "$OUT/test_$tag.log")" | tee -a "$OUT/summary.txt"
  [ $r -ne 0 ] && rc=1
done
exit $rc
