#!/bin/bash
# usage: gpu_bench.sh [batch] [steps]  -- bench + ncu launch list on one GPU; logs in gpurun_out/
B=${1:-4096}; K=${2:-20}
mkdir -p gpurun_out
timeout 900 python bench.py --steps $K --warmup 3 --batch $B > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
# every launch of two denoising steps at a reduced batch (ncu serialises and replays; shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --batch 256 --max-rows 512 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu exit $?"; tail -n 3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
