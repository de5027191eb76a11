#!/bin/bash
# next-tile A prefetch of the conv kernel on / off: first five conv launches under ncu, then the bench on the same box
mkdir -p gpurun_out
for cfg in "VDT_CONV_PREFETCH=1" "VDT_CONV_PREFETCH=0"; do
  echo "== $cfg"
  env $cfg timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_gemm -c 5 --csv \
      --log-file gpurun_out/exp_$cfg.csv python bench.py --steps 1 --warmup 3 --batch 512 --max-rows 1024 --no-cpu-baseline > /dev/null 2>&1
  python - "gpurun_out/exp_$cfg.csv" <<'PY'
import csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
by = {}
for r in rows:
    by.setdefault(int(r["ID"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
for i in sorted(by):
    d = by[i]
    print(i, f'{d["gpu__time_duration.sum"] / 1e3:9.1f} us  {d["dram__bytes_read.sum"] / 1e6:9.1f} MB read  tensor {d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]:.1f} %')
PY
done
for r in 1 2; do for v in 0 1; do
  VDT_CONV_PREFETCH=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefetch=$v', d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['family_ms_per_step']['conv'])"
done; done
