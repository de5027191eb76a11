#!/bin/bash
# ab_step.sh A B ...: per-family ncu step breakdown of several library builds on the SAME box
for v in "$@"; do
  export VDT_LIB=$PWD/v-diffusion-torch_b200/lib/libvdt_b200_$v.so
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_$v.csv python bench.py --steps 1 --warmup 3 --batch 512 --max-rows 1024 --no-cpu-baseline > gpurun_out/ncu_$v.log 2>&1
  echo "== variant $v"; python scripts/step_breakdown.py gpurun_out/launches_$v.csv
done
