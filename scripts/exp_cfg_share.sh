#!/bin/bash
# first five conv launches of a step (in_conv, block 0 conv1 / conv2, block 1 conv1 / conv2): time and DRAM reads under
# the CFG shared prefix with / without paired tiles, and row by row
mkdir -p gpurun_out
for cfg in "VDT_PAIR_TILES=1" "VDT_PAIR_TILES=0" "VDT_NO_CFG_SHARE=1"; do
  echo "== $cfg"
  env $cfg timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:conv_gemm -c 5 --csv \
      --log-file gpurun_out/exp_$cfg.csv python bench.py --steps 1 --warmup 3 --batch 512 --max-rows 1024 --no-cpu-baseline > /dev/null 2>&1
  python - "gpurun_out/exp_$cfg.csv" <<'PY'
import csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
by = {}
for r in rows:
    by.setdefault(int(r["ID"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
for i in sorted(by):
    print(i, f'{by[i]["gpu__time_duration.sum"] / 1e3:9.1f} us  {by[i]["dram__bytes_read.sum"] / 1e6:9.1f} MB read')
PY
done
