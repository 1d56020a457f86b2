#!/usr/bin/env bash
# Final visit of a round: the whole GPU suite + smoke(), the default bench line, and the launch list of the headline step
# (ncu only after the same command exited 0 without it).  Output: gpurun_out/final_*.
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > $OUT/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/final_smoke.log
python bench.py > $OUT/final_bench_n1.json 2> $OUT/final_bench_n1.err; echo "bench rc=$?"
M3="python bench.py --steps 1 --warmup 3 --batch 8880 --profile-only"
$M3 > $OUT/final_plain_m3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/final_launches_m3_b8880_bf16x2.csv \
    $M3 > $OUT/final_ncu_launch_m3.log 2>&1
echo "launch list rc=$?"
tail -2 $OUT/final_pytest_gpu.log; tail -1 $OUT/final_smoke.log; head -c 300 $OUT/final_bench_n1.json
