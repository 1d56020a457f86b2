#!/usr/bin/env bash
# One GPU visit: selftest (quick), parity suite, bench at two sizes, launch list under ncu.
# usage: scripts/gpu_round.sh <tag>
TAG="${1:-r}"
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
timeout 200 longterm360fov_b200/csrc/build/tc_selftest quick > $OUT/selftest_$TAG.log 2>&1; echo "selftest rc=$?"
rm -f $OUT/summary.txt
bash scripts/gpu_tests.sh $OUT; echo "tests rc=$?"
cat $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --batch 4096 > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -c 3000 $OUT/bench_$TAG.log
